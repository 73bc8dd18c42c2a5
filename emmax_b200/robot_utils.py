"""
Caller-side pipeline of the robot loop: the functions `experiments/robot` wraps around the model, over this package's classes.

    get_vla_action / get_seq_action   /root/reference/experiments/robot/openvla_utils.py:127-170, :173-218
    crop_and_resize                   openvla_utils.py:81-124     (tf.image.crop_and_resize, bilinear, to 224 x 224)
    resize_image                      experiments/robot/bridge/bridgev2_utils.py:152-166 (JPEG round trip + tf.image.resize lanczos3, antialias)
    get_model / get_action            experiments/robot/robot_utils.py:34-82

The reference runs the image steps in TensorFlow on the host; TensorFlow does not exist in this image, so their arithmetic is
RESTATED here from TensorFlow's published kernels (tensorflow/core/kernels/image/crop_and_resize_op.cc, scale_and_translate_op.cc,
image_ops_impl.py: convert_image_dtype) in numpy float32 — "parity unpinned" against TF itself — and the GPU twins (emx_crop_resize_u8,
emx_lanczos3_resize_u8) are pinned bit-exactly against these host twins. The JPEG round trip of `resize_image` uses Pillow's libjpeg
(quality 95, 4:2:0) and is not claimed to reproduce TensorFlow's encoder byte for byte.
"""

from __future__ import annotations

import io
from typing import Any, Optional, Tuple

import numpy as np
import torch
from PIL import Image

ACTION_DIM = 7
OPENVLA_V01_SYSTEM_PROMPT = (
    "A chat between a curious user and an artificial intelligence assistant. "
    "The assistant gives helpful, detailed, and polite answers to the user's questions."
)
F32 = np.float32


# ---------------------------------------------------------------------------------------------------------------------
# crop_and_resize (openvla_utils.py:81-124)
# ---------------------------------------------------------------------------------------------------------------------
def center_crop_box(crop_scale: float) -> Tuple[np.float32, np.float32, np.float32, np.float32]:
    """(y1, x1, y2, x2) of the centred box with area `crop_scale`, in float32 as TF computes it (openvla_utils.py:103-117)."""
    side = np.clip(np.sqrt(F32(crop_scale), dtype=F32), F32(0), F32(1))
    off = (F32(1) - side) / F32(2)
    return off, off, off + side, off + side


def crop_resize_taps(in_size: int, out_size: int, lo: np.float32, hi: np.float32):
    """Per output index: (first source index, second source index, lerp weight, inside?) of tf.image.crop_and_resize's bilinear sampling
    (crop_and_resize_op.cc: in = lo * (in_size - 1) + i * scale, scale = (hi - lo) * (in_size - 1) / (out_size - 1)), all float32."""
    i = np.arange(out_size, dtype=F32)
    if out_size > 1:
        scale = (hi - lo) * F32(in_size - 1) / F32(out_size - 1)
        pos = lo * F32(in_size - 1) + i * scale
    else:
        pos = np.full(1, F32(0.5) * (lo + hi) * F32(in_size - 1), dtype=F32)
    inside = (pos >= 0) & (pos <= F32(in_size - 1))
    top = np.floor(pos)
    bot = np.ceil(pos)
    lerp = (pos - top).astype(F32)
    return top.astype(np.int32), bot.astype(np.int32), lerp, inside


def crop_and_resize(image: np.ndarray, crop_scale: float, batch_size: int = 1, out_size: Tuple[int, int] = (224, 224)) -> np.ndarray:
    """float32 [H, W, C] or [B, H, W, C] in [0, 1] -> centre crop of area `crop_scale` resized (bilinear) to 224 x 224, float32
    (extrapolation value 0 outside the image, which a centre crop never reaches)."""
    x = np.asarray(image, dtype=F32)
    squeeze = x.ndim == 3
    if squeeze:
        x = x[None]
    assert x.shape[0] == batch_size
    y1, x1, y2, x2 = center_crop_box(crop_scale)
    H, W = x.shape[1:3]
    ty, by, ly, iy = crop_resize_taps(H, out_size[0], y1, y2)
    tx, bx, lx, ix = crop_resize_taps(W, out_size[1], x1, x2)
    ty_c, by_c, tx_c, bx_c = (np.clip(a, 0, n - 1) for a, n in ((ty, H), (by, H), (tx, W), (bx, W)))
    tl, tr = x[:, ty_c][:, :, tx_c], x[:, ty_c][:, :, bx_c]
    bl, br = x[:, by_c][:, :, tx_c], x[:, by_c][:, :, bx_c]
    lxb, lyb = lx[None, None, :, None], ly[None, :, None, None]
    top = tl + (tr - tl) * lxb
    bottom = bl + (br - bl) * lxb
    out = (top + (bottom - top) * lyb).astype(F32)
    out = np.where((iy[None, :, None, None]) & (ix[None, None, :, None]), out, F32(0))
    return out[0] if squeeze else out


def center_crop_frame(frame_u8: np.ndarray, crop_scale: float = 0.9) -> np.ndarray:
    """The whole `center_crop` branch of get_vla_action on a uint8 frame (openvla_utils.py:136-156): uint8 -> float32 / 255 ->
    crop_and_resize -> clip [0, 1] -> * 255.5, saturate, truncate to uint8 (tf.image.convert_image_dtype both ways)."""
    x = frame_u8.astype(F32) * (F32(1) / F32(255))
    y = np.clip(crop_and_resize(x, crop_scale, 1), F32(0), F32(1))
    return np.clip(y * F32(255.5), F32(0), F32(255)).astype(np.uint8)


def center_crop_frame_device(frame_u8: torch.Tensor, crop_scale: float = 0.9) -> torch.Tensor:
    """GPU twin of `center_crop_frame` (emx_crop_resize_u8): uint8 CUDA [H, W, 3] or [B, H, W, 3] -> uint8 [.., 224, 224, 3], bit-exact."""
    from ._lib import call, ptr, stream

    squeeze = frame_u8.dim() == 3
    x = (frame_u8[None] if squeeze else frame_u8).contiguous()
    if x.dtype != torch.uint8 or x.device.type != "cuda" or x.shape[-1] != 3:
        raise ValueError("center_crop_frame_device expects a uint8 CUDA tensor [B, H, W, 3]")
    B, H, W, _ = x.shape
    y1, x1, y2, x2 = center_crop_box(crop_scale)
    out = torch.empty((B, 224, 224, 3), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        call("emx_crop_resize_u8", ptr(x), B, H, W, float(y1), float(x1), float(y2), float(x2), ptr(out), 224, 224, stream())
    return out[0] if squeeze else out


# ---------------------------------------------------------------------------------------------------------------------
# resize_image (bridgev2_utils.py:152-166): tf.image.resize(method="lanczos3", antialias=True) -> round -> clip -> uint8
# ---------------------------------------------------------------------------------------------------------------------
def _lanczos3(x: np.ndarray) -> np.ndarray:
    """TensorFlow's LanczosKernelFunc (scale_and_translate_op.cc), radius 3, float32."""
    x = np.abs(x).astype(F32)
    pi = F32(3.14159265359)
    with np.errstate(divide="ignore", invalid="ignore"):
        v = (F32(3) * np.sin(pi * x, dtype=F32) * np.sin(pi * x / F32(3), dtype=F32) / (pi * pi * x * x)).astype(F32)
    v = np.where(x <= F32(1e-3), F32(1), v)
    return np.where(x > F32(3), F32(0), v).astype(F32)


def lanczos3_spans(in_size: int, out_size: int):
    """ComputeSpansCore of TF's ScaleAndTranslate for scale = out/in, translate 0, antialias: per output index the first source index and
    `span_size` normalised float32 weights (zero-padded)."""
    scale = F32(out_size) / F32(in_size)
    inv_scale = F32(1) / scale
    kernel_scale = max(inv_scale, F32(1))
    radius = F32(3)
    span_size = min(2 * int(np.ceil(radius * kernel_scale)) + 1, in_size)
    starts = np.zeros(out_size, dtype=np.int32)
    weights = np.zeros((out_size, span_size), dtype=F32)
    one_over = F32(1) / kernel_scale
    for x in range(out_size):
        sample_f = (F32(x) + F32(0.5)) * inv_scale
        if sample_f < 0 or sample_f > in_size:
            continue
        span_start = max(int(np.ceil(sample_f - radius * kernel_scale - F32(0.5))), 0)
        span_end = min(int(np.floor(sample_f + radius * kernel_scale - F32(0.5))), in_size - 1) + 1
        span_start = min(span_start, span_end)
        src = np.arange(span_start, span_end, dtype=F32)
        w = _lanczos3((src + F32(0.5) - sample_f) * one_over)
        tot = F32(np.sum(w, dtype=F32))
        if abs(tot) >= F32(1000) * np.finfo(F32).tiny:
            w = (w * (F32(1) / tot)).astype(F32)
        starts[x] = span_start
        weights[x, : span_end - span_start] = w
    return starts, weights


def lanczos3_resize(img_u8: np.ndarray, resize_size: Tuple[int, int]) -> np.ndarray:
    """uint8 [H, W, C] -> uint8 [h, w, C]: rows first (horizontal pass), then columns, float32 accumulation in source order, then
    round-half-even, clip, cast — `tf.cast(tf.clip_by_value(tf.round(tf.image.resize(img, size, "lanczos3", antialias=True)), 0, 255), tf.uint8)`."""
    H, W, C = img_u8.shape
    h, w = resize_size
    sx, wx = lanczos3_spans(W, w)
    sy, wy = lanczos3_spans(H, h)
    x = img_u8.astype(F32)
    tmp = np.zeros((H, w, C), dtype=F32)
    for k in range(wx.shape[1]):
        idx = np.minimum(sx + k, W - 1)
        tmp = (tmp + x[:, idx, :] * wx[None, :, k, None]).astype(F32)
    out = np.zeros((h, w, C), dtype=F32)
    for k in range(wy.shape[1]):
        idx = np.minimum(sy + k, H - 1)
        out = (out + tmp[idx] * wy[:, k, None, None]).astype(F32)
    return np.clip(np.round(out), 0, 255).astype(np.uint8)


def lanczos3_resize_device(frame_u8: torch.Tensor, resize_size: Tuple[int, int]) -> torch.Tensor:
    """GPU twin of `lanczos3_resize` (emx_lanczos3_resize_u8), uint8 CUDA [H, W, 3] -> uint8 [h, w, 3], bit-exact with the host twin."""
    from ._lib import call, ptr, stream

    if frame_u8.dtype != torch.uint8 or frame_u8.device.type != "cuda" or frame_u8.dim() != 3 or frame_u8.shape[-1] != 3:
        raise ValueError("lanczos3_resize_device expects a uint8 CUDA tensor [H, W, 3]")
    H, W, _ = frame_u8.shape
    h, w = resize_size
    dev = frame_u8.device
    sx, wx = lanczos3_spans(W, w)
    sy, wy = lanczos3_spans(H, h)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_sx, d_wx, d_sy, d_wy = t(sx), t(wx), t(sy), t(wy)
    tmp = torch.empty((H, w, 3), dtype=torch.float32, device=dev)
    out = torch.empty((h, w, 3), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        call("emx_lanczos3_resize_u8", ptr(frame_u8.contiguous()), H, W, h, w, ptr(d_sx), ptr(d_wx), wx.shape[1], ptr(d_sy), ptr(d_wy), wy.shape[1],
             ptr(tmp), ptr(out), stream())  # fmt: skip
    return out


def resize_image(img: np.ndarray, resize_size: Tuple[int, int]) -> np.ndarray:
    """bridgev2_utils.py:152-166: JPEG encode / decode (as the RLDS builder stored the training frames), then the lanczos3 resize."""
    assert isinstance(resize_size, tuple)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=95)
    img = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
    return lanczos3_resize(img, resize_size)


def get_preprocessed_image(obs: dict, resize_size) -> np.ndarray:
    """bridgev2_utils.py:169-175"""
    assert isinstance(resize_size, int) or isinstance(resize_size, tuple)
    if isinstance(resize_size, int):
        resize_size = (resize_size, resize_size)
    obs["full_image"] = resize_image(obs["full_image"], resize_size)
    return obs["full_image"]


# ---------------------------------------------------------------------------------------------------------------------
# the two entry points of the evaluation scripts
# ---------------------------------------------------------------------------------------------------------------------
def _prepare_image(obs: dict, center_crop: bool, device: Optional[torch.device]) -> Image.Image:
    image = Image.fromarray(obs["full_image"]).convert("RGB")
    if center_crop:
        frame = np.asarray(image)
        if device is not None and device.type == "cuda":  # overlap-friendly: the crop runs on the GPU, only the 150 KB result comes back
            frame = center_crop_frame_device(torch.from_numpy(frame.copy()).to(device)).cpu().numpy()
        else:
            frame = center_crop_frame(frame)
        image = Image.fromarray(frame).convert("RGB")
    return image


def get_vla_action(vla: Any, processor: Any, base_vla_name: str, obs: dict, task_label: str, unnorm_key: Optional[str], center_crop: bool = False) -> np.ndarray:
    """openvla_utils.py:127-170: frame (+ optional 0.9 centre crop) -> OpenVLA prompt -> processor -> predict_action."""
    image = _prepare_image(obs, center_crop, getattr(vla, "device", None))
    if "openvla-v01" in base_vla_name:  # OpenVLA v0.1
        prompt = f"{OPENVLA_V01_SYSTEM_PROMPT} USER: What action should the robot take to {task_label.lower()}? ASSISTANT:"
    else:  # OpenVLA
        prompt = f"In: What action should the robot take to {task_label.lower()}?\nOut:"
    inputs = processor(prompt, image).to(vla.device, dtype=torch.bfloat16)
    return vla.predict_action(**inputs, unnorm_key=unnorm_key, do_sample=False)


def get_seq_action(vla: Any, processor: Any, base_vla_name: str, obs: dict, task_label: str, unnorm_key: Optional[str], type: str,  # noqa: A002
                   center_crop: bool = False):
    """openvla_utils.py:173-218: frame -> PurePromptBuilder turn -> `generate_actions(image, prompt, type, max_new_tokens=512, ...)`."""
    image = _prepare_image(obs, center_crop, getattr(vla, "device", None))
    prompt_builder = vla.get_prompt_builder()
    prompt_builder.add_turn(role="human", message=task_label)
    prompt = prompt_builder.get_prompt()
    return vla.generate_actions(image=image, prompt_text=prompt, type=type, temperature=0.0, max_new_tokens=512, min_length=1, do_sample=False)


def get_action(cfg: Any, model: Any, obs: dict, task_label: str, processor: Any = None, type: str = "act"):  # noqa: A002
    """robot_utils.py:63-82: dispatch on cfg.model_family exactly as the evaluation scripts do."""
    if cfg.model_family == "openvla":
        action = get_vla_action(model, processor, cfg.pretrained_checkpoint, obs, task_label, cfg.unnorm_key, center_crop=cfg.center_crop)
        assert action.shape == (ACTION_DIM,)
        return [action], None
    if cfg.model_family == "pred-all":
        assert type in ["pos", "act"]
        return get_seq_action(model, processor, cfg.pretrained_checkpoint, obs, task_label, cfg.unnorm_key, type=type, center_crop=cfg.center_crop)
    raise ValueError("Unexpected `model_family` found in config.")
