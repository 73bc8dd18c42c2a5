"""Robot-loop shims: host twins of the TensorFlow image steps (known answers; TF itself is absent, so these are
restatements of its published kernels) and the two evaluation entry points over stand-in model / processor objects."""

import numpy as np
import pytest
import torch
from PIL import Image


def test_center_crop_known_answers():
    from emmax_b200 import robot_utils as R

    rng = np.random.default_rng(0)
    y1, x1, y2, x2 = R.center_crop_box(0.9)
    assert y1 == x1 and y2 == x2 and abs(float(y2 - y1) - 0.9**0.5) < 1e-6 and abs(float(y1) - (1 - 0.9**0.5) / 2) < 1e-6
    # crop_scale = 1 on a 224 x 224 frame samples every pixel exactly: the whole uint8 -> float -> uint8 chain is the identity
    img = rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)
    assert np.array_equal(R.center_crop_frame(img, 1.0), img)
    # a constant frame stays constant; output is always 224 x 224 (openvla_utils.py:117)
    c = np.full((300, 260, 3), 200, np.uint8)
    out = R.center_crop_frame(c)
    assert out.shape == (224, 224, 3) and np.unique(out).tolist() == [200]
    # a horizontal ramp: the 0.9 crop keeps the centre, so the output range is the input range shrunk by sqrt(0.9) around the middle
    ramp = np.tile(np.arange(256, dtype=np.uint8)[None, :, None], (256, 1, 3))
    out = R.center_crop_frame(ramp)
    lo, hi = float(x1) * 255, float(x2) * 255
    assert abs(int(out[0, 0, 0]) - lo) <= 1 and abs(int(out[0, -1, 0]) - hi) <= 1 and np.all(np.diff(out[5, :, 1].astype(int)) >= 0)
    # float API keeps the reference's shapes ([H, W, C] and [B, H, W, C])
    x = rng.random((2, 64, 48, 3), dtype=np.float32)
    assert R.crop_and_resize(x, 0.9, 2).shape == (2, 224, 224, 3) and R.crop_and_resize(x[0], 0.9, 1).shape == (224, 224, 3)


def test_lanczos3_resize_known_answers():
    from emmax_b200 import robot_utils as R

    rng = np.random.default_rng(1)
    starts, w = R.lanczos3_spans(256, 224)
    assert w.shape == (224, 9) and np.allclose(w.sum(1), 1.0, atol=1e-6) and starts.min() == 0 and starts.max() + 9 >= 256
    _, w_id = R.lanczos3_spans(224, 224)  # same size: the kernel degenerates to the identity tap
    assert np.allclose(np.sort(w_id, axis=1)[:, -1], 1.0, atol=1e-6)
    c = np.full((256, 256, 3), 137, np.uint8)
    assert np.unique(R.lanczos3_resize(c, (224, 224))).tolist() == [137]
    img = rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)
    assert np.array_equal(R.lanczos3_resize(img, (224, 224)), img)
    # independent cross-check (not TF, but the same windowed-sinc definition): Pillow's LANCZOS resample agrees to rounding
    # (on a smooth image: Pillow clips its uint8 intermediate between the two passes, which only shows where the lobes overshoot 0 / 255)
    yy, xx = np.mgrid[0:256, 0:256].astype(np.float32)
    smooth = np.stack([125 + 90 * np.sin(xx / 17 + c) * np.cos(yy / 23 - c) for c in range(3)], axis=-1).round().astype(np.uint8)
    ours = R.lanczos3_resize(smooth, (224, 224)).astype(int)
    pil = np.asarray(Image.fromarray(smooth).resize((224, 224), Image.LANCZOS)).astype(int)
    assert np.abs(ours - pil).max() <= 1 and np.abs(ours - pil).mean() < 0.3
    img = rng.integers(0, 256, (256, 256, 3), dtype=np.uint8)
    out = R.resize_image(img, (224, 224))  # JPEG round trip + resize
    assert out.shape == (224, 224, 3) and out.dtype == np.uint8
    obs = {"full_image": img}
    assert R.get_preprocessed_image(obs, 224).shape == (224, 224, 3) and obs["full_image"].shape == (224, 224, 3)


class _FakeInputs(dict):
    def to(self, device, dtype=None):
        self["moved"] = (str(device), dtype)
        return self


def test_entry_points_call_the_model_as_the_reference_does():
    """openvla_utils.py:127-218: prompt text, processor call, `predict_action(**inputs, unnorm_key=..., do_sample=False)` and
    `generate_actions(image=..., prompt_text=..., type=..., temperature=0.0, max_new_tokens=512, min_length=1, do_sample=False)`."""
    from emmax_b200 import PurePromptBuilder
    from emmax_b200 import robot_utils as R

    calls = {}

    class FakeVLA:
        device = torch.device("cpu")

        def predict_action(self, **kw):
            calls["predict"] = kw
            return np.zeros(7)

        def get_prompt_builder(self):
            return PurePromptBuilder("prismatic")

        def generate_actions(self, **kw):
            calls["generate"] = kw
            return [np.ones(7)], "text"

    def processor(prompt, image):
        calls["processor"] = (prompt, image.size, image.mode)
        return _FakeInputs(input_ids=torch.ones(1, 3, dtype=torch.long))

    obs = {"full_image": np.random.default_rng(0).integers(0, 256, (256, 256, 3), dtype=np.uint8)}
    a = R.get_vla_action(FakeVLA(), processor, "openvla-7b", obs, "Put Carrot In Pot", "bridge_orig", center_crop=True)
    assert a.shape == (7,)
    assert calls["processor"] == ("In: What action should the robot take to put carrot in pot?\nOut:", (224, 224), "RGB")
    assert calls["predict"]["unnorm_key"] == "bridge_orig" and calls["predict"]["do_sample"] is False and calls["predict"]["moved"][1] == torch.bfloat16
    R.get_vla_action(FakeVLA(), processor, "openvla-v01-7b", obs, "Lift", None)
    assert calls["processor"][0].startswith(R.OPENVLA_V01_SYSTEM_PROMPT + " USER: What action should the robot take to lift? ASSISTANT:")
    assert calls["processor"][1] == (256, 256)  # no centre crop: the frame goes through unchanged
    acts, text = R.get_seq_action(FakeVLA(), processor, "emma-x", obs, "put carrot in pot", None, type="pos")
    g = calls["generate"]
    assert g["prompt_text"] == "In: put carrot in pot\nOut:" and g["type"] == "pos" and g["max_new_tokens"] == 512 and g["min_length"] == 1
    assert g["temperature"] == 0.0 and g["do_sample"] is False and g["image"].size == (256, 256) and text == "text"

    class Cfg:
        model_family, pretrained_checkpoint, unnorm_key, center_crop = "pred-all", "x", None, False

    assert R.get_action(Cfg, FakeVLA(), obs, "t", processor, type="act")[1] == "text"
    Cfg.model_family = "openvla"
    assert R.get_action(Cfg, FakeVLA(), obs, "t", processor)[1] is None
    Cfg.model_family = "other"
    with pytest.raises(ValueError):
        R.get_action(Cfg, FakeVLA(), obs, "t", processor)
