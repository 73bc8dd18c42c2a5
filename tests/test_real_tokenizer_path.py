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


def _build_llama_shaped_tokenizer(path: str) -> None:
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, processors

    vocab = {"<unk>": 0, "<s>": 1, "</s>": 2}
    text_chars = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789:;,._-?!\n")
    for i, ch in enumerate(text_chars):
        vocab[ch] = 3 + i
    vocab["▁"] = 29871
    action_chars = [chr(0x4E00 + i) for i in range(256)]  # 256 distinct single code points (CJK block), like Llama-2's tail
    for i, ch in enumerate(action_chars):
        vocab[ch] = 31744 + i
    used = set(vocab.values())
    for i in range(32000):  # fill the remaining ids so that vocab_size == 32000
        if i not in used:
            vocab[f"<filler_{i}>"] = i
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    # "▁" is prepended to the text and every character is its own token (a WordLevel stand-in for SentencePiece pieces)
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.Metaspace(replacement="▁", prepend_scheme="first", split=False),
                                                 pre_tokenizers.Split("", behavior="isolated")])  # fmt: skip
    tok.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    tok.decoder = decoders.Sequence([decoders.Replace("▁", " "), decoders.Fuse(), decoders.Strip(" ", 1, 0)])
    os.makedirs(path, exist_ok=True)
    tok.save(os.path.join(path, "tokenizer.json"))
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"tokenizer_class": "PreTrainedTokenizerFast", "bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>",
                   "model_max_length": 2048, "clean_up_tokenization_spaces": False}, f)  # fmt: skip


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
