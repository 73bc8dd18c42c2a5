"""The oracle restatement against its frozen fixture (tests/golden/tiny_vla_golden.npz) and the image processor
against the reference transform arithmetic — CPU only."""

import os

import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "tiny_vla_golden.npz"))


def test_image_processor_matches_reference_arithmetic(g):
    from PIL import Image

    from emmax_b200 import PrismaticImageProcessor
    from emmax_b200.processing import OPENVLA_MEANS, OPENVLA_STDS

    img = Image.fromarray(g["image"])
    pv = PrismaticImageProcessor()(img, return_tensors="pt")["pixel_values"]
    assert pv.shape == (1, 6, 224, 224) and np.array_equal(pv.numpy(), g["pixel_values"])
    # 224x224 input + resize-naive = identity resize: to_tensor -> (x - mean) / std per backbone, DINO first
    x = torch.from_numpy(g["image"]).permute(2, 0, 1).float().div(255)
    for i in range(2):
        m, s = torch.tensor(OPENVLA_MEANS[i])[:, None, None], torch.tensor(OPENVLA_STDS[i])[:, None, None]
        assert torch.allclose(pv[0, 3 * i : 3 * i + 3], (x - m) / s, atol=1e-6)
    # letterbox pads to a square with the (last) backbone mean before resizing (processing_prismatic.py:24-30,117-118)
    wide = Image.fromarray(np.zeros((100, 200, 3), dtype=np.uint8))
    lb = PrismaticImageProcessor(image_resize_strategy="letterbox")(wide, return_tensors="pt")["pixel_values"]
    assert lb.shape == (1, 6, 224, 224)
    with pytest.raises(ValueError):
        PrismaticImageProcessor(image_resize_strategy="stretch")


def test_processor_batch_validation():
    from PIL import Image

    from emmax_b200 import AutoProcessor

    proc = AutoProcessor.from_pretrained(None)
    img = Image.fromarray(np.zeros((224, 224, 3), dtype=np.uint8))
    out = proc("In: x\nOut:", img)
    assert set(out.keys()) == {"input_ids", "attention_mask", "pixel_values"} and proc.model_input_names == ["input_ids", "attention_mask", "pixel_values"]
    cast = out.to("cpu", dtype=torch.bfloat16)
    assert cast["pixel_values"].dtype == torch.bfloat16 and cast["input_ids"].dtype == torch.long
    with pytest.raises(ValueError):
        proc(["a", "b"], img, padding=True)  # 2 texts, 1 image (processing_prismatic.py:213-214)
    prompt, image = proc.get_prompt("put carrot in pot", img)
    assert prompt.endswith("\nOut:") and "INSTRUCTION: \nput carrot in pot" in prompt


def test_oracle_reproduces_fixture(g):
    from emmax_b200 import SyntheticLlamaTokenizer, tiny_config
    from emmax_b200.synthetic import make_state_dict
    from oracle.model import OracleVLA

    cfg = tiny_config()
    input_ids = torch.from_numpy(g["input_ids"])
    script = [int(x) for x in g["script"]]
    sd = make_state_dict(cfg, seed=0, script=script, script_prev=int(input_ids[0, -1]))
    m = OracleVLA.from_state_dict(cfg, sd, dtype=torch.bfloat16)
    pv = torch.from_numpy(g["pixel_values"]).to(torch.bfloat16)
    ids, logits = m.generate(input_ids, pv, len(script), return_logits=True)
    assert ids.numpy().tolist() == g["generated_ids"].tolist()
    assert ids[0, input_ids.shape[1] :].tolist() == script  # the planted known answer
    assert np.allclose(logits.float().numpy(), g["step_logits"], atol=5e-2)
    assert np.array_equal(m.predict_action(input_ids, pv), g["action_pred"])
    tok = SyntheticLlamaTokenizer()
    text = tok.decode(script, skip_special_tokens=True)
    assert "MOVEMENT:\n" in text and "POLICIES:\n" in text and text.count(";") == 1


def test_oracle_skipping_last_vit_block_is_exact():
    """`get_intermediate_layers(n={depth-2})` taps block depth-2: the final block cannot change the output
    (the engine never uploads it)."""
    from emmax_b200 import tiny_config
    from emmax_b200.synthetic import make_state_dict
    from oracle.model import OracleVLA

    cfg = tiny_config()
    m = OracleVLA.from_state_dict(cfg, make_state_dict(cfg, seed=1), dtype=torch.float32)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.inference_mode():
        for vit in (m.vision_backbone.featurizer, m.vision_backbone.fused_featurizer):
            assert torch.equal(vit(x), vit(x, run_all_blocks=True))
            assert vit(x).shape == (2, 256, vit.v.embed_dim)


def test_oracle_wiring_matches_the_reference_hf_model_class(golden_dir):
    """tests/golden/hf_model_golden.npz holds outputs of the reference's OWN `OpenVLAForActionPrediction` (modeling_prismatic.py executed on the
    CPU with toy widths by oracle/gen_golden_hf_model.py; only timm's ViT internals and two transformers-5.x shims are stand-ins). The oracle,
    built from the same seeded state dict, must reproduce them: state-dict key set, vision features, projector output, multimodal prefill
    logits at every position, the cached single-token step, greedy continuation ids (4.40.1-style loop over the reference's forward) and
    `predict_action` on planted action tokens. fp32 CPU; logits to 1e-4
    (thread-count-dependent summation order), integer and fp64 results exactly."""
    from emmax_b200 import tiny_config
    from emmax_b200.synthetic import make_state_dict
    from oracle.model import OracleVLA

    g = np.load(os.path.join(golden_dir, "hf_model_golden.npz"))
    cfg = tiny_config()
    script = [int(x) for x in g["action_script"]]  # 7 planted action tokens after the id 29871, then EOS
    sd = make_state_dict(cfg, seed=int(g["pixel_seed"]), device="cpu", script=script, script_prev=29871)
    assert sorted(sd.keys()) == list(g["state_dict_keys"]), "state-dict names accepted by the reference class (0 missing / 0 unexpected)"
    oracle = OracleVLA.from_state_dict(cfg, sd, device="cpu", dtype=torch.float32, attn_implementation="eager")
    ids = torch.from_numpy(g["input_ids"])
    pv = torch.randn(1, 6, 224, 224, generator=torch.Generator().manual_seed(int(g["pixel_seed"])))
    with torch.no_grad():
        feats = oracle.vision_backbone(pv)
        proj = oracle.projector(feats)
    assert np.allclose(feats[0, ::37, ::53].numpy(), g["features_probe"], atol=1e-5) and np.allclose(proj[0, ::37, ::29].numpy(), g["projected_probe"], atol=1e-5)
    logits, past = oracle.prefill(ids, pv)
    assert logits.shape[1] == ids.shape[1] + cfg.num_patches
    assert np.array_equal(logits.argmax(-1).numpy(), g["prefill_argmax"]), "argmax at every one of the 277 positions"
    assert np.allclose(logits[0, -1].numpy(), g["prefill_last_logits"], atol=1e-4)
    step_logits, _ = oracle.step(torch.from_numpy(g["step_token"]), past)
    assert np.allclose(step_logits[0, -1].numpy(), g["step_logits"], atol=1e-4)
    n_new = g["generated_ids"].shape[1] - ids.shape[1]
    assert np.array_equal(oracle.generate(ids, pv, n_new, eos_token_id=None).numpy(), g["generated_ids"])
    # predict_action: the reference appends 29871 itself (:512-515), generates action_dim tokens, de-tokenises, un-normalises
    ids29871 = torch.cat([ids, torch.tensor([[29871]])], dim=1)
    assert oracle.generate(ids29871, pv, 7, eos_token_id=None)[0, -7:].tolist() == g["action_ids"].tolist() == script[:7]
    assert np.array_equal(np.asarray(oracle.predict_action(ids, pv), dtype=np.float64), g["predict_action"])
    assert len(set(np.round(g["predict_action"], 6))) == 7, "seven distinct action values: the de-tokeniser saw real action tokens, not the clip value"
