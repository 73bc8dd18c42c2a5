"""Host-side de-tokeniser / Solver / prompt builder against fixtures frozen from the REFERENCE'S OWN code
(tests/golden/detok_golden.json, made by oracle/gen_golden.py from /root/reference/prismatic/vla/{action_tokenizer,solver}.py)."""

import json
import os

import numpy as np
import pytest

from emmax_b200.action_tokenizer import ActionTokenizer
from emmax_b200.prompting import PurePromptBuilder, emma_x_prompt, openvla_prompt
from emmax_b200.solver import Solver, unnormalize
from emmax_b200.tokenization import SyntheticLlamaTokenizer
from oracle import detok as oracle_detok


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "detok_golden.json")) as f:
        return json.load(f)


def unhex(xs):
    return np.array([float.fromhex(x) for x in xs], dtype=np.float64)


def test_decode_ids_bit_exact(golden):
    tok = SyntheticLlamaTokenizer()
    at = ActionTokenizer(tok)
    ids = np.array(golden["ids"])
    want = unhex(golden["decoded"])
    assert np.array_equal(at.decode_token_ids_to_actions(ids), want)
    assert np.array_equal(oracle_detok.decode_token_ids_to_actions(ids, tok.vocab_size), want)
    assert at.action_token_begin_idx == golden["action_token_begin_idx"]
    assert np.array_equal(at.bin_centers, unhex(golden["bin_centers_hex"]))


def test_known_answers():
    # SURVEY.md §8c known-answer vectors: centres[k] = -1 + (2k+1)/255, k = clip(32000 - id - 1, 0, 254)
    at = ActionTokenizer(SyntheticLlamaTokenizer())
    got = at.decode_token_ids_to_actions(np.array([31999, 31872, 31745, 31744, 32000, 32063, 5]))
    c = lambda k: (np.linspace(-1, 1, 256)[k] + np.linspace(-1, 1, 256)[k + 1]) / 2  # noqa: E731
    assert got[0] == c(0) and abs(got[0] - (-1 + 1 / 255)) < 1e-15
    assert got[1] == 0.0 or abs(got[1]) < 1e-16
    assert got[2] == c(254) and got[3] == c(254)  # documented clip case (action_tokenizer.py:60-64)
    assert got[4] == c(0) and got[5] == c(0)  # ids >= vocab clip to bin 0
    assert got[6] == c(254)  # ordinary text ids are not rejected by the reference


def test_encode_text_matches_reference(golden):
    at = ActionTokenizer(SyntheticLlamaTokenizer())
    acts = np.array(golden["encode_actions"])
    assert [at(a) for a in acts] == golden["encode_text"]
    assert at(acts) == golden["encode_batch_text"]


def test_solver_cases(golden):
    solver = Solver(ActionTokenizer(SyntheticLlamaTokenizer()), verbose=False)
    for name, case in golden["solver_cases"].items():
        pol, remain = solver.extract_action_policies(case["text"])
        want = [unhex(p) for p in case["policies_hex"]]
        assert len(pol) == len(want), name
        for g, w in zip(pol, want):
            assert np.array_equal(np.asarray(g, dtype=np.float64), w), name
        assert remain == case["remain"], name
        req, mov = solver.extract_movement_plan(case["text"])
        assert req == case["require_unorm"], name
        assert np.array_equal(np.asarray(mov, dtype=np.float64), unhex(case["movement_hex"])), name


def test_unnormalize(golden):
    u = golden["unnorm"]
    at = ActionTokenizer(SyntheticLlamaTokenizer())
    normalized = at.decode_token_ids_to_actions(np.array(u["ids"]))
    assert np.array_equal(normalized, unhex(u["normalized_hex"]))
    assert np.array_equal(unnormalize(normalized, u["stats"]), unhex(u["actions_hex"]))
    assert np.array_equal(oracle_detok.unnormalize_actions(normalized, u["stats"]), unhex(u["actions_hex"]))
    # DummyDataset stats (datasets.py:200-204): q01=0, q99=1, all-true mask -> 0.5*(a+1)
    dummy = {"q01": [0.0] * 7, "q99": [1.0] * 7}
    assert np.array_equal(unnormalize(normalized, dummy), 0.5 * (normalized + 1))


def test_prompt_builder():
    # base_prompter.py:36,44,73 — "In: {msg}\nOut: " then rstrip
    pb = PurePromptBuilder("prismatic")
    pb.add_turn("human", "  hello <image> world ")
    assert pb.get_prompt() == "In: hello  world\nOut:"
    assert pb.get_potential_prompt("x") == "In: hello  world\nOut: In: x\nOut:"
    with pytest.raises(AssertionError):
        PurePromptBuilder().add_turn("gpt", "first turn must be human")
    assert emma_x_prompt("put carrot in pot") == (
        "In: What action should the robot take to achieve the instruction\nINSTRUCTION: \nput carrot in pot\nOut:"
    )
    assert openvla_prompt("Put Carrot") == "In: What action should the robot take to put carrot?\nOut:"


def test_tokenizer_roundtrip_properties():
    tok = SyntheticLlamaTokenizer()
    ids = tok("POLICIES:\nabc", add_special_tokens=False).input_ids
    assert ids[0] == 29871  # SentencePiece dummy prefix: why solver.py:125-126 drops the first value
    assert tok("x").input_ids[0] == tok.bos_token_id
    acts = list(range(31744, 32000))
    text = tok.decode(acts)
    assert len(text) == 256 and tok(text, add_special_tokens=False).input_ids == [29871] + acts
    assert tok.decode([1, 29871, 300, 2], skip_special_tokens=True) == "POLICIES:"
    enc = tok(["ab", "a"], return_tensors="pt", padding=True)
    assert enc.input_ids.shape == (2, 4) and enc.attention_mask[1].tolist() == [1, 1, 1, 0]


def test_pil_bicubic_restatement_is_bit_exact_with_pillow():
    """The integer resample the GPU kernel implements (emx_resize_preprocess_u8) is pinned here against Pillow itself — the code
    torchvision's `resize(PIL image, BICUBIC, antialias=True)` runs in the reference's processor (processing_prismatic.py:133)."""
    from PIL import Image

    from emmax_b200.processing import pil_bicubic_resize_reference

    rng = np.random.default_rng(0)
    for h, w in [(256, 256), (480, 640), (224, 224), (200, 300), (97, 131), (720, 1280)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((224, 224), Image.BICUBIC))
        assert np.array_equal(pil_bicubic_resize_reference(img, 224, 224), want), (h, w)
    flat = np.full((300, 300, 3), 255, dtype=np.uint8)  # saturation: overshoot of the negative lobes must clamp, not wrap
    assert np.array_equal(pil_bicubic_resize_reference(flat, 224, 224), np.asarray(Image.fromarray(flat).resize((224, 224), Image.BICUBIC)))


def test_prompt_builder_matches_the_reference_class(golden_dir):
    """PurePromptBuilder vs the reference's own class (base_prompter.py:28-73, loaded by path and frozen by oracle/gen_golden_prompts.py):
    per-turn wrapped strings, `get_prompt`, `get_potential_prompt` (which does not consume a turn), `<image>` stripping, empty gpt turns."""
    g = json.load(open(os.path.join(golden_dir, "prompt_golden.json")))
    for case in g["cases"]:
        pb = PurePromptBuilder("prismatic")
        wrapped = [pb.add_turn(role, msg) for role, msg in case["turns"]]
        assert wrapped == case["wrapped"] and pb.get_prompt() == case["prompt"]
        assert pb.get_potential_prompt("next question") == case["potential"] and pb.turn_count == case["turn_count"]
    with pytest.raises(AssertionError):
        PurePromptBuilder("prismatic").add_turn("gpt", "out of turn")
    assert emma_x_prompt("put carrot in pot") == g["cases"][0]["prompt"]


def test_image_processor_matches_the_reference_class(golden_dir):
    """`PrismaticImageProcessor.apply_transform` vs the reference's own class (processing_prismatic.py:128-145 executed by
    oracle/gen_golden_processor.py on Pillow + torchvision): all three resize strategies, square / landscape / portrait inputs, the
    bf16-rounded OpenVLA means/stds. Bit-exact: sha256 of the float32 output bytes."""
    import hashlib

    import torch
    from PIL import Image

    from emmax_b200 import PrismaticImageProcessor

    g = json.load(open(os.path.join(golden_dir, "processor_golden.json")))
    for case in g["cases"]:
        h, w = case["h"], case["w"]
        img = Image.fromarray(np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8))
        proc = PrismaticImageProcessor(use_fused_vision_backbone=True, image_resize_strategy=case["strategy"], input_sizes=[(3, 224, 224)] * 2,
                                       interpolations=["bicubic"] * 2, means=[tuple(m) for m in g["means"]], stds=[tuple(s) for s in g["stds"]])  # fmt: skip
        t = proc.apply_transform(img.convert("RGB")).float().contiguous()
        assert list(t.shape) == case["shape"], case
        flat = t.flatten()
        for i, v in case["probes"].items():
            assert float(flat[int(i)]) == v, (case["strategy"], h, w, i)
        assert hashlib.sha256(t.numpy().tobytes()).hexdigest() == case["sha256"], (case["strategy"], h, w)
    # the defaults of this package are the OpenVLA export values used above
    d = PrismaticImageProcessor()
    assert [list(m) for m in d.means] == g["means"] and [list(s) for s in d.stds] == g["stds"] and d.image_resize_strategy == "resize-naive"
    assert isinstance(d.apply_transform(Image.new("RGB", (300, 200))), torch.Tensor)
