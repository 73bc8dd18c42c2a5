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
        """GPU path for frames that already have the model's input size (the robot loop pre-resizes to 224x224,
        experiments/robot/bridge/bridgev2_utils.py:152-166; `resize-naive` is then the identity): uint8 [B, H, W, 3] or [H, W, 3] on a
        CUDA device -> bf16 [B, 6, H, W], bit-exact with `preprocess(...)["pixel_values"].to(device, dtype=torch.bfloat16)`. One kernel
        (`emx_preprocess_u8`) replaces PIL -> float -> normalize on the host and cuts the upload from 602 KB to 150 KB per frame."""
        from ._lib import call, ptr, stream

        if frames.dim() == 3:
            frames = frames[None]
        if frames.dtype != torch.uint8 or frames.device.type != "cuda" or frames.shape[-1] != 3:
            raise ValueError("preprocess_device expects a uint8 CUDA tensor [B, H, W, 3]")
        B, H, W, _ = frames.shape
        if self.image_resize_strategy != "resize-naive" or any(tuple(s[-2:]) != (H, W) for s in self.input_sizes):
            raise ValueError(f"preprocess_device handles frames already at the input size {self.input_sizes}; got {H}x{W} "
                             f"({self.image_resize_strategy}): use the host transform")  # fmt: skip
        frames = frames.contiguous()
        n = len(self.input_sizes)
        key = (str(frames.device), n)
        if getattr(self, "_dev_stats", {}).get("key") != key:
            self._dev_stats = {"key": key,
                               "mean": torch.tensor([c for m in self.means for c in m], dtype=torch.float32, device=frames.device),
                               "std": torch.tensor([c for m in self.stds for c in m], dtype=torch.float32, device=frames.device)}  # fmt: skip
        out = torch.empty((B, 3 * n, H, W), dtype=torch.bfloat16, device=frames.device)
        with torch.cuda.device(frames.device):
            call("emx_preprocess_u8", ptr(frames), B, H, W, n, ptr(self._dev_stats["mean"]), ptr(self._dev_stats["std"]), ptr(out), stream())
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
