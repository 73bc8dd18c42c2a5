"""
`AutoProcessor` surface: image transform + tokenizer.

Behavioural mirror of /root/reference/prismatic/extern/hf/processing_prismatic.py:
  * `PrismaticImageProcessor.apply_transform` (:128-145): [letterbox ->] resize (bicubic, antialias) -> center-crop
    -> to_tensor -> normalize, once per backbone, channel-stacked to [6, H, W] (DINO-normalised first);
  * `PrismaticProcessor.__call__` (:187-216): tokenizer + image processor, batch-size check (`ValueError` :213-214);
  * `get_prompt(task_label, image)`: hub-only helper of the Emma-X model card (README.md:42-44), prompt per
    prompting.py.
The reference obtains the transform parameters from `timm.data.create_transform` (:71-79); timm is not available, so
the four-stage structure it validates (:82-93) is constructed directly. Means/stds are *inputs* (they come from
`preprocessor_config.json`); the defaults are the bf16-rounded ImageNet values the OpenVLA exporter wrote
 followed by SigLIP's 0.5/0.5.
"""

from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
import torchvision.transforms.functional as TVF
from PIL import Image

from .prompting import emma_x_prompt
from .tokenization import load_tokenizer

_INTERP = {"bicubic": TVF.InterpolationMode.BICUBIC, "bilinear": TVF.InterpolationMode.BILINEAR,
           "nearest": TVF.InterpolationMode.NEAREST, "lanczos": TVF.InterpolationMode.LANCZOS}  # fmt: skip

OPENVLA_MEANS = [(0.484375, 0.455078125, 0.40625), (0.5, 0.5, 0.5)]
OPENVLA_STDS = [(0.228515625, 0.2236328125, 0.224609375), (0.5, 0.5, 0.5)]


class BatchFeature(dict):
    """Dict with attribute access and the HF `.to(device, dtype=...)` rule: only floating tensors change dtype."""

    def __getattr__(self, k: str) -> Any:
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, *args: Any, **kwargs: Any) -> "BatchFeature":
        device, dtype = kwargs.get("device"), kwargs.get("dtype")
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
            else:
                device = a
        out = BatchFeature()
        for k, v in self.items():
            if isinstance(v, torch.Tensor):
                v = v.to(device=device, dtype=dtype) if torch.is_floating_point(v) else v.to(device=device)
            out[k] = v
        return out


def letterbox_pad_transform(image: Image.Image, padding_fill_value: Tuple[int, int, int]) -> Image.Image:
    (w, h), side = image.size, max(image.size)
    pw, ph = int((side - w) / 2), int((side - h) / 2)
    return TVF.pad(image, (pw, ph, pw, ph), fill=padding_fill_value, padding_mode="constant")


def pil_bicubic_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Coefficient table of Pillow's antialiased bicubic resample for one axis (src/libImaging/Resample.c: `precompute_coeffs` with
    `bicubic_filter` (a = -0.5, support 2) + `normalize_coeffs_8bpc`, PRECISION_BITS = 22), which is what torchvision's
    `resize(PIL image, BICUBIC, antialias=True)` executes at processing_prismatic.py:133. Returns (kk int32 [out_size, ksize],
    bounds int32 [out_size, 2] = (first input index, number of taps)). Same double-precision operations in the same order, so
    the integer coefficients are identical to Pillow's (tests/test_detok_host.py pins the resulting resize against PIL itself)."""
    import math

    def filt(x: float) -> float:
        a = -0.5
        x = -x if x < 0.0 else x
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        if x < 2.0:
            return (((x - 5) * x + 8) * x - 4) * a
        return 0.0

    scale = in_size / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in enumerate(w):
            if ww != 0.0:
                v = v / ww
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return kk, bounds


def pil_bicubic_resize_reference(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """numpy twin of the two integer passes (horizontal, then vertical on the uint8 intermediate) — the host-side statement of what
    `emx_resize_preprocess_u8` computes; used by the CPU tests to pin the coefficient builder against Pillow."""

    def axis_pass(x: np.ndarray, kk: np.ndarray, bounds: np.ndarray, axis: int) -> np.ndarray:
        x = np.moveaxis(x, axis, 0).astype(np.int64)
        out = np.empty((kk.shape[0],) + x.shape[1:], dtype=np.uint8)
        for i in range(kk.shape[0]):
            lo, n = bounds[i]
            acc = np.full(x.shape[1:], 1 << 21, dtype=np.int64)
            for j in range(n):
                acc += x[lo + j] * int(kk[i, j])
            out[i] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)

    x = axis_pass(img, *pil_bicubic_coeffs(img.shape[1], out_w), axis=1)
    return axis_pass(x, *pil_bicubic_coeffs(img.shape[0], out_h), axis=0)


class PrismaticImageProcessor:
    model_input_names = ["pixel_values"]

    def __init__(
        self,
        use_fused_vision_backbone: bool = True,
        image_resize_strategy: str = "resize-naive",
        input_sizes: Optional[List[Tuple[int, int, int]]] = None,
        interpolations: Optional[List[str]] = None,
        means: Optional[List[Tuple[float, float, float]]] = None,
        stds: Optional[List[Tuple[float, float, float]]] = None,
        **kwargs: Any,
    ) -> None:
        self.use_fused_vision_backbone = use_fused_vision_backbone
        self.image_resize_strategy = image_resize_strategy
        n = 2 if use_fused_vision_backbone else 1
        self.input_sizes = [tuple(s) for s in (input_sizes or [(3, 224, 224)] * n)]
        self.interpolations = list(interpolations or ["bicubic"] * n)
        self.means = [tuple(m) for m in (means or (OPENVLA_MEANS[:n] if n == 2 else [(0.5, 0.5, 0.5)]))]
        self.stds = [tuple(s) for s in (stds or (OPENVLA_STDS[:n] if n == 2 else [(0.5, 0.5, 0.5)]))]
        if image_resize_strategy not in ("resize-naive", "letterbox", "resize-crop"):
            raise ValueError(f"Image resize strategy `{image_resize_strategy}` is not supported!")
        self.tvf_do_letterbox = image_resize_strategy == "letterbox"
        # the reference overwrites the fill per backbone, so the last backbone's mean wins (:117-118)
        self.tvf_letterbox_fill = tuple(int(x * 255) for x in self.means[-1]) if self.tvf_do_letterbox else None

    def _resize_size(self, idx: int) -> Union[int, Tuple[int, int]]:
        side = self.input_sizes[idx][-1]
        return (side, side) if self.image_resize_strategy == "resize-naive" else side

    def apply_transform(self, img: Image.Image) -> torch.Tensor:
        if self.tvf_do_letterbox:
            img = letterbox_pad_transform(img, self.tvf_letterbox_fill)
        outs = []
        for idx in range(len(self.input_sizes)):
            x = TVF.resize(img, size=self._resize_size(idx), interpolation=_INTERP[self.interpolations[idx]],
                           max_size=None, antialias=True)  # fmt: skip
            x = TVF.center_crop(x, output_size=self.input_sizes[idx][-2:])
            x = TVF.to_tensor(x)
            x = TVF.normalize(x, mean=list(self.means[idx]), std=list(self.stds[idx]), inplace=False)
            outs.append(x)
        return torch.vstack(outs)

    def preprocess_device(self, frames: torch.Tensor) -> torch.Tensor:
        """GPU twin of `apply_transform` for the `resize-naive` (Emma-X: conf/models.py:494) and `letterbox` strategies: uint8 [B, H, W, 3] or [H, W, 3]
        on a CUDA device -> bf16 [B, 6, 224, 224], bit-exact with `preprocess(...)["pixel_values"].to(device, dtype=torch.bfloat16)`.
        Frames already at the input size (the robot loop pre-resizes, bridgev2_utils.py:152-166) take one kernel (`emx_preprocess_u8`);
        other sizes (256x256 sim frames, run_bridgev2_eval.py:161) go through Pillow's antialiased bicubic resample restated in its
        own integer arithmetic (`emx_resize_preprocess_u8`). Replaces PIL -> float -> normalize on the host and uploads the raw frame."""
        from ._lib import call, ptr, stream

        if frames.dim() == 3:
            frames = frames[None]
        if frames.dtype != torch.uint8 or frames.device.type != "cuda" or frames.shape[-1] != 3:
            raise ValueError("preprocess_device expects a uint8 CUDA tensor [B, H, W, 3]")
        B, H, W, _ = frames.shape
        n = len(self.input_sizes)
        Ho, Wo = self.input_sizes[0][-2:]
        if self.image_resize_strategy not in ("resize-naive", "letterbox") or any(tuple(s[-2:]) != (Ho, Wo) for s in self.input_sizes):
            raise ValueError(f"preprocess_device implements the `resize-naive` and `letterbox` strategies with one common input size; got "
                             f"{self.image_resize_strategy} / {self.input_sizes}: use the host transform")  # fmt: skip
        if self.tvf_do_letterbox:
            # letterbox_pad_transform (processing_prismatic.py:23-29): symmetric constant border of int((max - side) / 2) pixels per side,
            # filled with int(255 * mean) of the LAST backbone; pure data movement, then the same resample as `resize-naive`
            hp, vp = int((max(H, W) - W) / 2), int((max(H, W) - H) / 2)
            if hp or vp:
                fill = torch.tensor(self.tvf_letterbox_fill, dtype=torch.uint8, device=frames.device)
                padded = fill.expand(B, H + 2 * vp, W + 2 * hp, 3).contiguous()
                padded[:, vp : vp + H, hp : hp + W] = frames
                frames, H, W = padded, H + 2 * vp, W + 2 * hp
        if (H, W) != (Ho, Wo) and any(i != "bicubic" for i in self.interpolations):
            raise ValueError(f"preprocess_device resizes with Pillow's antialiased bicubic only; got {self.interpolations}")
        frames = frames.contiguous()
        dev = frames.device
        key = (str(dev), n)
        if getattr(self, "_dev_stats", {}).get("key") != key:
            self._dev_stats = {"key": key,
                               "mean": torch.tensor([c for m in self.means for c in m], dtype=torch.float32, device=dev),
                               "std": torch.tensor([c for m in self.stds for c in m], dtype=torch.float32, device=dev)}  # fmt: skip
        out = torch.empty((B, 3 * n, Ho, Wo), dtype=torch.bfloat16, device=dev)
        with torch.cuda.device(dev):
            if (H, W) == (Ho, Wo):
                call("emx_preprocess_u8", ptr(frames), B, H, W, n, ptr(self._dev_stats["mean"]), ptr(self._dev_stats["std"]), ptr(out), stream())
            else:
                ck = (str(dev), H, W, Ho, Wo)
                tabs = getattr(self, "_dev_resample", {})
                if ck not in tabs:  # Pillow's coefficient tables for this geometry, built once on the host
                    kh, bh = pil_bicubic_coeffs(W, Wo)
                    kv, bv = pil_bicubic_coeffs(H, Ho)
                    tabs[ck] = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (kh, bh, kv, bv))
                    self._dev_resample = tabs
                kh, bh, kv, bv = tabs[ck]
                tmp = torch.empty((B, H, Wo, 3), dtype=torch.uint8, device=dev)
                call("emx_resize_preprocess_u8", ptr(frames), B, H, W, Ho, Wo, ptr(kh), ptr(bh), kh.shape[1], ptr(kv), ptr(bv), kv.shape[1],
                     ptr(tmp), n, ptr(self._dev_stats["mean"]), ptr(self._dev_stats["std"]), ptr(out), stream())  # fmt: skip
        return out

    def preprocess(self, images: Union[Image.Image, List[Image.Image]], return_tensors: Optional[str] = None, **_: Any) -> BatchFeature:
        if not isinstance(images, list):
            images = [images]
        pixel_values = torch.stack([self.apply_transform(img.convert("RGB")) for img in images]).float()
        if return_tensors is None:
            return BatchFeature(pixel_values=pixel_values.numpy())
        return BatchFeature(pixel_values=pixel_values)

    def __call__(self, images: Union[Image.Image, List[Image.Image]], **kwargs: Any) -> BatchFeature:
        return self.preprocess(images, **kwargs)

    def to_dict(self) -> Dict[str, Any]:
        return dict(
            use_fused_vision_backbone=self.use_fused_vision_backbone, image_resize_strategy=self.image_resize_strategy,
            input_sizes=self.input_sizes, interpolations=self.interpolations, means=self.means, stds=self.stds,
            image_processor_type="PrismaticImageProcessor", processor_class="PrismaticProcessor",
        )  # fmt: skip

    @classmethod
    def from_pretrained(cls, path: Optional[str] = None, **kwargs: Any) -> "PrismaticImageProcessor":
        cfg: Dict[str, Any] = {}
        if path is not None and os.path.exists(os.path.join(path, "preprocessor_config.json")):
            with open(os.path.join(path, "preprocessor_config.json")) as f:
                cfg = json.load(f)
            for k in ("image_processor_type", "processor_class", "auto_map", "tvf_resize_params", "tvf_crop_params",
                      "tvf_normalize_params", "tvf_do_letterbox", "tvf_letterbox_fill"):  # fmt: skip
                cfg.pop(k, None)
        cfg.update(kwargs)
        return cls(**cfg)


class PrismaticProcessor:
    attributes = ["image_processor", "tokenizer"]

    def __init__(self, image_processor: Optional[PrismaticImageProcessor] = None, tokenizer: Any = None) -> None:
        self.image_processor = image_processor or PrismaticImageProcessor()
        self.tokenizer = tokenizer or load_tokenizer(None)

    @classmethod
    def from_pretrained(cls, path: Optional[str] = None, trust_remote_code: bool = True, **kwargs: Any) -> "PrismaticProcessor":
        return cls(PrismaticImageProcessor.from_pretrained(path), load_tokenizer(path))

    def __call__(
        self,
        text: Union[str, Sequence[str]],
        images: Union[Image.Image, List[Image.Image]],
        padding: Any = False,
        truncation: Any = None,
        max_length: Optional[int] = None,
        return_tensors: Optional[str] = "pt",
    ) -> BatchFeature:
        pixel_values = self.image_processor(images, return_tensors=return_tensors)["pixel_values"]
        text_inputs = self.tokenizer(
            [text] if isinstance(text, str) else list(text),
            return_tensors=return_tensors, padding=padding, truncation=truncation, max_length=max_length,
        )  # fmt: skip
        if pixel_values.shape[0] != text_inputs.input_ids.shape[0]:
            raise ValueError("Batch is malformed; expected same number of images and text inputs!")
        return BatchFeature(**{k: text_inputs[k] for k in text_inputs.keys()}, pixel_values=pixel_values)

    # hub-only helper used by the model card (README.md:42-44)
    def get_prompt(self, task_label: str, image: Image.Image) -> Tuple[str, Image.Image]:
        return emma_x_prompt(task_label), image.convert("RGB")

    def batch_decode(self, sequences: Any, skip_special_tokens: bool = False, **kw: Any) -> List[str]:
        return self.tokenizer.batch_decode(sequences, skip_special_tokens=skip_special_tokens, **kw)

    def decode(self, token_ids: Any, skip_special_tokens: bool = False, **kw: Any) -> str:
        return self.tokenizer.decode(token_ids, skip_special_tokens=skip_special_tokens, **kw)

    @property
    def model_input_names(self) -> List[str]:
        return list(dict.fromkeys(list(self.tokenizer.model_input_names) + self.image_processor.model_input_names))


class AutoProcessor:
    """`AutoProcessor.from_pretrained(path, trust_remote_code=True)` (README.md:42; openvla_utils.py:75-78)."""

    @staticmethod
    def from_pretrained(path: Optional[str] = None, **kwargs: Any) -> PrismaticProcessor:
        return PrismaticProcessor.from_pretrained(path, **kwargs)

    @staticmethod
    def register(*_: Any, **__: Any) -> None:  # openvla_utils.py:38-41 registers classes; nothing to do here
        return None


AutoImageProcessor = AutoProcessor
