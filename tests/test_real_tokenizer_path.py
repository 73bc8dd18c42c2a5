"""The user-supplied-tokenizer path: when the checkpoint directory holds a `tokenizer.json`, `AutoProcessor` /
`from_pretrained` hand out a `transformers` fast tokenizer instead of the synthetic stand-in. The Llama-2 vocabulary is gated and
absent here, so this test builds a tokenizer with the same SHAPE with the `tokenizers` library — 32000 entries, `<unk>/<s>/</s>` =
0/1/2, the SentencePiece dummy prefix `▁` at id 29871, the 256 action bins on the last 256 ids (31744..31999), Metaspace
pre-tokenisation with a prepended `▁`, BOS added by the post-processor — and drives the host pipeline through it:
processor call (processing_prismatic.py:187-216), ActionTokenizer round trip (action_tokenizer.py:28-68) and
Solver.extract_action_policies (solver.py:108-137: re-tokenise each `;` piece, drop the leading 29871, de-tokenise 7 values)."""

import json
import os

import numpy as np
import pytest
import torch

tokenizers = pytest.importorskip("tokenizers")


from _llama_shaped_tokenizer import build_llama_shaped_tokenizer as _build_llama_shaped_tokenizer  # noqa: E402


@pytest.fixture(scope="module")
def tok_dir(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("llama_shaped_tok"))
    _build_llama_shaped_tokenizer(d)
    return d


def test_loader_returns_the_directory_tokenizer(tok_dir):
    from emmax_b200.tokenization import SyntheticLlamaTokenizer, load_tokenizer

    tok = load_tokenizer(tok_dir)
    assert not isinstance(tok, SyntheticLlamaTokenizer) and tok.vocab_size == 32000
    ids = tok("In: pick up\nOut:").input_ids
    assert ids[0] == 1 and ids[1] == 29871, "BOS then the SentencePiece dummy prefix, as Llama-2"
    assert tok("x", add_special_tokens=False).input_ids[0] == 29871
    assert isinstance(load_tokenizer(None), SyntheticLlamaTokenizer)


def test_processor_and_solver_through_a_real_fast_tokenizer(tok_dir):
    from PIL import Image

    from emmax_b200 import AutoProcessor
    from emmax_b200.action_tokenizer import ActionTokenizer
    from emmax_b200.solver import Solver

    proc = AutoProcessor.from_pretrained(tok_dir, trust_remote_code=True)
    image = Image.fromarray(np.random.default_rng(0).integers(0, 256, (224, 224, 3), dtype=np.uint8))
    prompt, image = proc.get_prompt("put carrot in pot", image)
    batch = proc(prompt, image)
    assert batch["input_ids"].dtype == torch.long and batch["input_ids"][0, 0] == 1
    assert batch["attention_mask"].shape == batch["input_ids"].shape and batch["pixel_values"].shape == (1, 6, 224, 224)

    at = ActionTokenizer(proc.tokenizer)
    actions = np.array([[-0.93, 0.41, 0.0, 0.77, -0.12, 0.3, 1.0], [0.05, -0.6, 0.25, -1.0, 0.9, 0.0, -0.33]])
    pieces = at(actions)  # two strings of 7 action characters each
    assert all(len(p) == 7 for p in pieces)
    text = "the arm is left of the pot.\nPOLICIES:\n" + ";".join(pieces) + "\n"
    policies, reasoning = Solver(at, verbose=False).extract_action_policies(text)
    assert reasoning.strip() == "the arm is left of the pot."
    assert len(policies) == 2 and all(len(p) == 7 for p in policies)
    want = at.decode_token_ids_to_actions(at.tokenizer.vocab_size - np.digitize(np.clip(actions, -1, 1), at.bins))
    assert np.array_equal(np.asarray(policies), want), "text round trip must return the bin centres of the encoded actions"
    assert np.abs(np.asarray(policies) - np.clip(actions, -1, 1)).max() <= 1.0 / 255 + 1e-12
    # decode() of generated ids with specials skipped, as generate_actions does
    ids = [1] + proc.tokenizer("ok", add_special_tokens=False).input_ids + [2]
    assert proc.tokenizer.decode(ids, skip_special_tokens=True).strip() == "ok"


def test_processor_call_matches_the_reference_class(tok_dir, golden_dir):
    """`PrismaticProcessor.__call__` vs the reference's own class (processing_prismatic.py:187-216, run by oracle/gen_golden_processor_call.py with
    the same Llama-shaped tokenizer): key order, dtypes, shapes, ids, mask, pixel bytes (sha256), `model_input_names`, and the malformed-batch error."""
    import hashlib

    from PIL import Image

    from emmax_b200 import AutoProcessor

    g = json.load(open(os.path.join(golden_dir, "processor_call_golden.json")))
    proc = AutoProcessor.from_pretrained(tok_dir, trust_remote_code=True)
    for case in g["cases"]:
        h, w = case["h"], case["w"]
        img = Image.fromarray(np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8))
        out = proc(case["prompt"], img)
        assert list(out.keys()) == case["keys"]
        assert {k: str(v.dtype) for k, v in out.items()} == case["dtypes"] and {k: list(v.shape) for k, v in out.items()} == case["shapes"]
        assert out["input_ids"].tolist() == case["input_ids"] and out["attention_mask"].tolist() == case["attention_mask"]
        assert hashlib.sha256(out["pixel_values"].float().contiguous().numpy().tobytes()).hexdigest() == case["pixel_sha256"]
    assert proc.model_input_names == g["model_input_names"]
    with pytest.raises(ValueError) as e:
        proc([g["cases"][1]["prompt"]] * 2, Image.new("RGB", (224, 224)))
    assert str(e.value) == g["batch_error"] == "Batch is malformed; expected same number of images and text inputs!"
