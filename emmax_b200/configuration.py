"""
Configuration objects for the B200-native Emma-X / OpenVLA action-generation path.

Mirrors the reference's `PrismaticConfig` / `OpenVLAConfig`
(/root/reference/prismatic/extern/hf/configuration_prismatic.py:72-140) closely enough that a reference
`config.json` round-trips (`from_dict` / `to_dict`), but without depending on the `transformers` config registry
(the reference resolves `text_config` through `CONFIG_MAPPING`, :121-125; here the Llama fields are plain data).

Only the `dinosiglip-vit-so-224px` + `llama2-7b-pure` family (Emma-X, conf/vla.py:302-315) is accelerated; the
vision-backbone tables below carry the timm hyper-parameters the reference gets from `timm.create_model`
(modeling_prismatic.py:78-101), which is not vendored in the reference tree.
"""

from __future__ import annotations

import copy
import json
import os
from dataclasses import asdict, dataclass, field
from typing import Any, Dict, List, Optional


@dataclass
class ViTDims:
    """Hyper-parameters of one timm `VisionTransformer` (restated; timm==0.9.10 is not vendored in the reference)."""

    timm_id: str
    embed_dim: int
    depth: int
    num_heads: int
    mlp_dim: int
    num_prefix_tokens: int  # 1 cls + 4 reg for DINOv2-reg4; 0 for SigLIP (class_token=False)
    layerscale: bool  # DINOv2: init_values=1e-5 -> LayerScale present; SigLIP: none
    patch_size: int = 14
    image_size: int = 224
    ln_eps: float = 1e-6

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def num_patches(self) -> int:
        return (self.image_size // self.patch_size) ** 2

    @property
    def num_tokens(self) -> int:
        return self.num_patches + self.num_prefix_tokens

    @property
    def used_depth(self) -> int:
        # `get_intermediate_layers(n={depth-2})` returns the output of block index depth-2
        # (modeling_prismatic.py:85-87, :99-101), i.e. depth-1 blocks contribute to the result.
        return self.depth - 1


@dataclass
class LlamaDims:
    vocab_size: int = 32064  # 32000 + PAD, padded to a multiple of 64 (llama2.py:74-76)
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: int = 32
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_position_embeddings: int = 4096
    bos_token_id: int = 1
    eos_token_id: int = 2
    pad_token_id: int = 32000

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


# Defaults of `transformers.LlamaConfig()` (the class behind CONFIG_MAPPING["llama"], configuration_prismatic.py:119-123) for the
# fields this package reads; used to complete a `text_config` read from a checkpoint's config.json.
HF_LLAMA_CONFIG_DEFAULTS: Dict[str, Any] = dict(
    vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
    num_key_value_heads=None, rms_norm_eps=1e-6, rope_theta=10000.0, max_position_embeddings=2048, bos_token_id=1,
    eos_token_id=2,
)  # fmt: skip


# timm model id -> dims (values as created by `timm.create_model(id, img_size=224, num_classes=0)`)
TIMM_VIT_DIMS: Dict[str, ViTDims] = {
    "vit_large_patch14_reg4_dinov2.lvd142m": ViTDims(
        "vit_large_patch14_reg4_dinov2.lvd142m", 1024, 24, 16, 4096, num_prefix_tokens=5, layerscale=True
    ),
    "vit_so400m_patch14_siglip_224": ViTDims(
        "vit_so400m_patch14_siglip_224", 1152, 27, 16, 4304, num_prefix_tokens=0, layerscale=False
    ),
}

# Same tables as configuration_prismatic.py:15-46, restricted to the accelerated family.
VISION_BACKBONE_TO_RESOLUTION = {"dinosiglip-vit-so-224px": [224, 224]}
VISION_BACKBONE_TO_TIMM_ID = {
    "dinosiglip-vit-so-224px": ["vit_large_patch14_reg4_dinov2.lvd142m", "vit_so400m_patch14_siglip_224"]
}
TIMM_OVERRIDE_ACT_LAYER = {"dinosiglip-vit-so-224px": [None, None]}
LLM_BACKBONE_TO_HF_PATH = {"llama2-7b-pure": "meta-llama/Llama-2-7b-hf"}
VALID_VISION_BACKBONES = set(VISION_BACKBONE_TO_RESOLUTION)
VALID_LLM_BACKBONES = set(LLM_BACKBONE_TO_HF_PATH)


class PrismaticConfig:
    model_type: str = "prismatic"

    def __init__(
        self,
        vision_backbone_id: str = "dinosiglip-vit-so-224px",
        llm_backbone_id: str = "llama2-7b-pure",
        arch_specifier: str = "no-align+fused-gelu-mlp",
        use_fused_vision_backbone: Optional[bool] = None,
        image_resize_strategy: str = "resize-naive",
        text_config: Optional[Dict[str, Any]] = None,
        llm_max_length: int = 2048,
        pad_token_id: int = 32000,
        pad_to_multiple_of: int = 64,
        output_projector_states: bool = False,
        vision_dims: Optional[List[Dict[str, Any]]] = None,
        **kwargs: Any,
    ) -> None:
        if vision_backbone_id not in VALID_VISION_BACKBONES:
            raise ValueError(f"Vision backbone `{vision_backbone_id}` not in {VALID_VISION_BACKBONES = }")
        if llm_backbone_id not in VALID_LLM_BACKBONES:
            raise ValueError(f"LLM backbone `{llm_backbone_id}` not in {VALID_LLM_BACKBONES = }")

        self.vision_backbone_id = vision_backbone_id
        self.llm_backbone_id = llm_backbone_id
        self.arch_specifier = arch_specifier
        self.output_projector_states = output_projector_states
        self.use_fused_vision_backbone = (
            use_fused_vision_backbone
            if use_fused_vision_backbone is not None
            else any(vision_backbone_id.startswith(v) for v in ["dinoclip", "dinosiglip"])
        )
        self.timm_model_ids = VISION_BACKBONE_TO_TIMM_ID[vision_backbone_id]
        self.timm_override_act_layers = TIMM_OVERRIDE_ACT_LAYER[vision_backbone_id]
        self.image_sizes = VISION_BACKBONE_TO_RESOLUTION[vision_backbone_id]
        self.image_resize_strategy = image_resize_strategy
        self.hf_llm_id = LLM_BACKBONE_TO_HF_PATH[llm_backbone_id]
        self.llm_max_length = llm_max_length
        self.pad_token_id, self.pad_to_multiple_of = pad_token_id, pad_to_multiple_of

        # `text_config`: plain Llama fields (the reference builds a `LlamaConfig`, :121-125)
        known = set(LlamaDims.__dataclass_fields__)
        tc = dict(text_config or {})
        self.text_config = LlamaDims(**{k: v for k, v in tc.items() if k in known})
        self._text_config_extra = {k: v for k, v in tc.items() if k not in known}

        # Vision dims: defaults come from the timm-id table; an explicit list overrides (used by the tiny test config)
        if vision_dims is not None:
            self.vision_dims = [ViTDims(**d) if isinstance(d, dict) else d for d in vision_dims]
        else:
            self.vision_dims = [copy.deepcopy(TIMM_VIT_DIMS[i]) for i in self.timm_model_ids]

        # HF-compat odds and ends the callers touch
        self.output_attentions = kwargs.pop("output_attentions", False)
        self.output_hidden_states = kwargs.pop("output_hidden_states", False)
        self.use_return_dict = kwargs.pop("use_return_dict", True)
        self._attn_implementation = kwargs.pop("attn_implementation", "flash_attention_2")
        self._extra = kwargs

    # --- derived sizes -------------------------------------------------------------------------------------------
    @property
    def vision_embed_dim(self) -> int:
        return sum(v.embed_dim for v in self.vision_dims)

    @property
    def num_patches(self) -> int:
        return self.vision_dims[0].num_patches

    # --- (de)serialisation ----------------------------------------------------------------------------------------
    def to_dict(self) -> Dict[str, Any]:
        d = {
            "model_type": self.model_type,
            "vision_backbone_id": self.vision_backbone_id,
            "llm_backbone_id": self.llm_backbone_id,
            "arch_specifier": self.arch_specifier,
            "use_fused_vision_backbone": self.use_fused_vision_backbone,
            "image_resize_strategy": self.image_resize_strategy,
            "text_config": {**asdict(self.text_config), **self._text_config_extra},
            "llm_max_length": self.llm_max_length,
            "pad_token_id": self.pad_token_id,
            "pad_to_multiple_of": self.pad_to_multiple_of,
            "output_projector_states": self.output_projector_states,
            "timm_model_ids": self.timm_model_ids,
            "timm_override_act_layers": self.timm_override_act_layers,
            "image_sizes": self.image_sizes,
            "hf_llm_id": self.hf_llm_id,
            "vision_dims": [asdict(v) for v in self.vision_dims],
        }
        return d

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "PrismaticConfig":
        d = dict(d)
        # The reference builds `text_config` as `LlamaConfig(**text_config)` (configuration_prismatic.py:119-123), so a key that a
        # checkpoint's config.json omits takes the *transformers* default there (the HF exporter only patches vocab_size, pad_token_id
        # and the dtype, convert_openvla_weights_to_hf.py:175-177) - NOT the defaults of `LlamaDims`, which describe the synthetic /
        # native Llama-2 configuration. Every RMSNorm of the path depends on `rms_norm_eps`, so the fill-in must match.
        d["text_config"] = {**HF_LLAMA_CONFIG_DEFAULTS, **(d.get("text_config") or {})}
        if d["text_config"].get("num_key_value_heads") is None:
            d["text_config"]["num_key_value_heads"] = d["text_config"]["num_attention_heads"]
        for derived in ("model_type", "timm_model_ids", "timm_override_act_layers", "image_sizes", "hf_llm_id"):
            d.pop(derived, None)
        for hf_noise in ("architectures", "auto_map", "torch_dtype", "transformers_version", "_name_or_path"):
            d.pop(hf_noise, None)
        return cls(**d)

    @classmethod
    def from_pretrained(cls, path: str, **kwargs: Any) -> "PrismaticConfig":
        with open(os.path.join(path, "config.json") if os.path.isdir(path) else path) as f:
            d = json.load(f)
        d.update(kwargs)
        return cls.from_dict(d)

    def save_pretrained(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "config.json"), "w") as f:
            json.dump(self.to_dict(), f, indent=2)


class OpenVLAConfig(PrismaticConfig):
    model_type: str = "openvla"

    def __init__(
        self,
        norm_stats: Optional[Dict[str, Dict[str, Dict[str, List[float]]]]] = None,
        n_action_bins: int = 256,
        **kwargs: Any,
    ) -> None:
        self.norm_stats, self.n_action_bins = norm_stats, n_action_bins
        super().__init__(**kwargs)

    def to_dict(self) -> Dict[str, Any]:
        d = super().to_dict()
        d["norm_stats"], d["n_action_bins"] = self.norm_stats, self.n_action_bins
        return d


# === Named configurations ===========================================================================================
def synthetic_norm_stats(seed: int = 0, action_dim: int = 7) -> Dict[str, Any]:
    """Seeded stand-in for `dataset_statistics.json` (bridge-style mask: gripper dim is not un-normalised,
    /root/reference/prismatic/vla/datasets/rlds/oxe/materialize.py:37-39)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    q01 = (-rng.uniform(0.01, 0.06, action_dim)).round(6)
    q99 = (rng.uniform(0.01, 0.06, action_dim)).round(6)
    q01[-1], q99[-1] = 0.0, 1.0
    return {
        "synthetic": {
            "action": {
                "q01": q01.tolist(),
                "q99": q99.tolist(),
                "mask": [True] * (action_dim - 1) + [False],
            }
        }
    }


def emma_x_config(**kwargs: Any) -> OpenVLAConfig:
    """Full Emma-X architecture (conf/vla.py:302-315; conf/models.py:490-497)."""
    kwargs.setdefault("norm_stats", synthetic_norm_stats())
    return OpenVLAConfig(**kwargs)


def tiny_config(**kwargs: Any) -> OpenVLAConfig:
    """Same topology at toy widths: every kernel shape class of the full model appears (head dims 64 / 72 / 128,
    a non-multiple-of-64 MLP width, 5 prefix tokens, LayerScale on one tower) but a CPU oracle runs it in milliseconds."""
    kwargs.setdefault("norm_stats", synthetic_norm_stats())
    vision = [
        ViTDims("tiny_dino", 128, 4, 2, 512, num_prefix_tokens=5, layerscale=True),
        ViTDims("tiny_siglip", 144, 4, 2, 536, num_prefix_tokens=0, layerscale=False),
    ]
    text = dict(hidden_size=256, intermediate_size=688, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=2)
    return OpenVLAConfig(vision_dims=vision, text_config=text, **kwargs)
