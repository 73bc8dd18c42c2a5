"""
Native ("Prismatic run directory") checkpoint loader: `load_vla(".../<RUN_ID>/checkpoints/<step>.pt")`.

Mirror of /root/reference/prismatic/models/load.py:122-228 for the one model family this package accelerates
(`dinosiglip-vit-so-224px` + `llama2-7b-pure`, Emma-X: conf/vla.py:302-315), as `experiments/robot` uses it
(robot_utils.py:37-42: `load_vla(model_id_or_path=cfg.model_pretrained_checkpoint, hf_token=..., proprio_norm_stats=...)`).

A native checkpoint is `torch.load(pt)["model"] = {"vision_backbone": {...}, "projector": {...}, "llm_backbone": {...}}` with the
module-local parameter names of the reference classes (prismatic.py:112-120):
    vision_backbone:  dino_featurizer.* / siglip_featurizer.*   (timm names; LayerScale as `ls{1,2}.gamma`)
    projector:        projector.{0,2,4}.{weight,bias}           (FusedMLPProjector's nn.Sequential)
    llm_backbone:     llm.model.* / llm.lm_head.weight
`remap_native_state_dict` is the name map of vla-scripts/extern/convert_openvla_weights_to_hf.py:74-116 (PROJECTOR_KEY_MAPPING,
`llm.` -> `language_model.`, `dino_featurizer.` -> `vision_backbone.featurizer.`, `.gamma` -> `.scale_factor`,
`siglip_featurizer.` -> `vision_backbone.fused_featurizer.`), after which the engine sees exactly what `from_pretrained` feeds it.
"""

from __future__ import annotations

import json
import os
from pathlib import Path
from typing import Any, Dict, Optional, Union

import torch

from .configuration import OpenVLAConfig
from .modeling import OpenVLAForActionPrediction
from .tokenization import load_tokenizer

# conf/models.py: the registered base VLMs built on the accelerated backbone pair (model_id -> ModelConfig fields)
BASE_VLM_REGISTRY: Dict[str, Dict[str, Any]] = {
    "prism-dinosiglip-224px+7b": dict(vision_backbone_id="dinosiglip-vit-so-224px", llm_backbone_id="llama2-7b-pure",
                                      arch_specifier="no-align+fused-gelu-mlp", image_resize_strategy="resize-naive", llm_max_length=2048),
    "prism-dinosiglip-224px-controlled+7b": dict(vision_backbone_id="dinosiglip-vit-so-224px", llm_backbone_id="llama2-7b-pure",
                                                 arch_specifier="no-align+fused-gelu-mlp", image_resize_strategy="resize-naive", llm_max_length=2048),
}  # fmt: skip

PROJECTOR_KEY_MAPPING = {f"projector.{i}.{p}": f"projector.fc{j}.{p}" for i, j in ((0, 1), (2, 2), (4, 3)) for p in ("weight", "bias")}


def remap_native_state_dict(model_state_dict: Dict[str, Dict[str, torch.Tensor]]) -> Dict[str, torch.Tensor]:
    """{"vision_backbone", "projector", "llm_backbone"} component dicts -> one flat dict in the HF export naming."""
    missing = [k for k in ("vision_backbone", "projector", "llm_backbone") if k not in model_state_dict]
    if missing:
        # the reference keeps timm's pretrained vision weights when the checkpoint holds none (prismatic.py:119-120); there is no timm
        # (and no network) to fetch them from here
        raise ValueError(f"native checkpoint is missing component(s) {missing}; expected keys vision_backbone / projector / llm_backbone")
    out: Dict[str, torch.Tensor] = {}
    for key, value in model_state_dict["projector"].items():
        if key not in PROJECTOR_KEY_MAPPING:
            raise KeyError(f"unexpected projector key {key!r} (fused-gelu-mlp projector expected: {sorted(PROJECTOR_KEY_MAPPING)})")
        out[PROJECTOR_KEY_MAPPING[key]] = value
    for key, value in model_state_dict["llm_backbone"].items():
        out[key.replace("llm.", "language_model.", 1) if key.startswith("llm.") else key] = value
    for key, value in model_state_dict["vision_backbone"].items():
        if key.startswith("dino_featurizer."):
            if key.endswith(".gamma"):  # timm LayerScale parameter, renamed because transformers rewrites `gamma` (convert script :60-71)
                key = key[: -len(".gamma")] + ".scale_factor"
            out["vision_backbone.featurizer." + key[len("dino_featurizer.") :]] = value
        elif key.startswith("siglip_featurizer."):
            out["vision_backbone.fused_featurizer." + key[len("siglip_featurizer.") :]] = value
        else:
            raise KeyError(f"unexpected vision-backbone key {key!r} (dino_featurizer.* / siglip_featurizer.* expected)")
    return out


def to_native_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, Dict[str, torch.Tensor]]:
    """Inverse of `remap_native_state_dict` (used by tests and tools to write a run directory from an HF-named state dict)."""
    inv_proj = {v: k for k, v in PROJECTOR_KEY_MAPPING.items()}
    comp: Dict[str, Dict[str, torch.Tensor]] = {"vision_backbone": {}, "projector": {}, "llm_backbone": {}}
    for key, value in sd.items():
        if key.startswith("projector."):
            comp["projector"][inv_proj[key]] = value
        elif key.startswith("language_model."):
            comp["llm_backbone"]["llm." + key[len("language_model.") :]] = value
        elif key.startswith("vision_backbone.featurizer."):
            k = "dino_featurizer." + key[len("vision_backbone.featurizer.") :]
            comp["vision_backbone"][k[: -len(".scale_factor")] + ".gamma" if k.endswith(".scale_factor") else k] = value
        elif key.startswith("vision_backbone.fused_featurizer."):
            comp["vision_backbone"]["siglip_featurizer." + key[len("vision_backbone.fused_featurizer.") :]] = value
        else:
            raise KeyError(key)
    return comp


def load_vla(model_id_or_path: Union[str, Path], hf_token: Optional[str] = None, cache_dir: Optional[Union[str, Path]] = None,
             load_for_training: bool = False, step_to_load: Optional[int] = None, model_type: str = "pretrained",
             proprio_norm_stats: Optional[dict] = None, config: Optional[OpenVLAConfig] = None, **model_kwargs: Any) -> OpenVLAForActionPrediction:  # fmt: skip
    """Loads a pretrained VLA from a local native checkpoint (load.py:122-228). `model_id_or_path` must be the checkpoint `.pt` FILE
    `<RUN_ID>/checkpoints/<name>.pt`, next to `<RUN_ID>/config.json` ({"vla": {"base_vlm": ...}}) and `<RUN_ID>/dataset_statistics.json`,
    exactly as the reference validates it; hub ids cannot be resolved offline. `config` overrides the registry lookup (toy widths in
    tests). Returns the same class `from_pretrained` does, so `generate_actions(image, prompt, type)` / `predict_action` behave alike."""
    if load_for_training:
        raise NotImplementedError("emmax_b200 is the inference hot path; training is out of scope")
    if not os.path.isfile(model_id_or_path):
        raise ValueError(f"Couldn't find valid HF Hub Path `{model_type}/{model_id_or_path}` (offline: only local checkpoint files can be loaded)")
    checkpoint_pt = Path(model_id_or_path)
    assert (checkpoint_pt.suffix == ".pt") and (checkpoint_pt.parent.name == "checkpoints"), "Invalid checkpoint!"
    run_dir = checkpoint_pt.parents[1]
    config_json, dataset_statistics_json = run_dir / "config.json", run_dir / "dataset_statistics.json"
    assert config_json.exists(), f"Missing `config.json` for `{run_dir = }`"
    assert dataset_statistics_json.exists(), f"Missing `dataset_statistics.json` for `{run_dir = }`"
    with open(config_json, "r") as f:
        vla_cfg = json.load(f)["vla"]
    with open(dataset_statistics_json, "r") as f:
        norm_stats = json.load(f)
    if config is None:
        base_vlm = vla_cfg["base_vlm"]
        if base_vlm not in BASE_VLM_REGISTRY:
            raise ValueError(f"base_vlm `{base_vlm}` is not one of the accelerated model families {sorted(BASE_VLM_REGISTRY)}")
        config = OpenVLAConfig(**BASE_VLM_REGISTRY[base_vlm])
    config.norm_stats = norm_stats
    model_state_dict = torch.load(checkpoint_pt, map_location="cpu")["model"]
    assert ("downsampler" not in model_state_dict) or (len(model_state_dict["downsampler"]) == 0), "Downsampler?"
    sd = remap_native_state_dict(model_state_dict)
    vla = OpenVLAForActionPrediction(config, sd, tokenizer=load_tokenizer(str(run_dir)), **model_kwargs)
    vla.proprio_norm_stats = proprio_norm_stats
    return vla
