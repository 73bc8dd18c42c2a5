"""Freezes outputs of the reference's own `PrismaticImageProcessor.apply_transform` (TEST INFRASTRUCTURE) for tests/test_detok_host.py.

/root/reference/prismatic/extern/hf/processing_prismatic.py is loaded by path. Two of its imports do not resolve in this container and are
replaced by stand-ins BEFORE the import: `timm.data.create_transform` (timm is absent; for `is_training=False, crop_pct=1.0` timm 0.9.10
returns Compose([Resize(size, interpolation), CenterCrop(size), ToTensor(), Normalize(mean, std)]) — exactly the structure the reference
validates at :82-93 — which is what the stand-in builds), and four type aliases of `transformers.tokenization_utils` that moved in
transformers 5.x (annotations only). Everything else — letterbox padding, the per-strategy resize sizes, torchvision's functional
resize / center_crop / to_tensor / normalize calls and the channel stacking — is the reference's code running on Pillow + torchvision.
The fixture stores, per case, the sha256 of the float32 output bytes and a few probe values (a 6x224x224 tensor per case would be 1.2 MB).
Usage (container with /root/reference):  python oracle/gen_golden_processor.py"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch
from PIL import Image

REF = "/root/reference/prismatic/extern/hf/processing_prismatic.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "processor_golden.json")

MEANS = [(0.484375, 0.455078125, 0.40625), (0.5, 0.5, 0.5)]
STDS = [(0.228515625, 0.2236328125, 0.224609375), (0.5, 0.5, 0.5)]
CASES = [(strategy, h, w) for strategy in ("resize-naive", "letterbox", "resize-crop") for (h, w) in ((224, 224), (256, 256), (480, 640), (301, 200))]


def case_image(h: int, w: int) -> Image.Image:
    return Image.fromarray(np.random.default_rng(h * 1000 + w).integers(0, 256, (h, w, 3), dtype=np.uint8))


def load_reference():
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor

    def create_transform(input_size, interpolation, mean, std, crop_pct, crop_mode, is_training):
        assert not is_training and crop_pct == 1.0 and crop_mode == "center" and input_size[-1] == input_size[-2]
        return Compose([Resize(input_size[-1], interpolation=InterpolationMode(interpolation)), CenterCrop(input_size[-1]), ToTensor(),
                        Normalize(mean=torch.tensor(mean), std=torch.tensor(std))])  # fmt: skip

    import transformers.image_processing_utils  # noqa: F401  (resolve transformers' lazy imports BEFORE the timm stand-in exists:
    import transformers.processing_utils  # noqa: F401         its optional-dependency probe must keep seeing "timm not installed")
    import transformers.tokenization_utils as tu
    from importlib.machinery import ModuleSpec

    timm, timm_data = types.ModuleType("timm"), types.ModuleType("timm.data")
    timm.__spec__, timm_data.__spec__ = ModuleSpec("timm", None), ModuleSpec("timm.data", None)
    timm_data.create_transform = create_transform
    timm.data = timm_data
    sys.modules.update({"timm": timm, "timm.data": timm_data})

    for name in ("PaddingStrategy", "PreTokenizedInput", "TextInput", "TruncationStrategy"):
        if not hasattr(tu, name):
            setattr(tu, name, object)
    spec = importlib.util.spec_from_file_location("ref_processing_prismatic", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


if __name__ == "__main__":
    ref = load_reference()
    out = []
    for strategy, h, w in CASES:
        proc = ref.PrismaticImageProcessor(use_fused_vision_backbone=True, image_resize_strategy=strategy, input_sizes=[(3, 224, 224)] * 2,
                                           interpolations=["bicubic"] * 2, means=MEANS, stds=STDS)  # fmt: skip
        t = proc.apply_transform(case_image(h, w).convert("RGB")).float().contiguous()
        flat = t.flatten()
        probes = [0, 1, 224 * 224 - 1, 3 * 224 * 224, flat.numel() // 2 + 17, flat.numel() - 1]
        out.append({"strategy": strategy, "h": h, "w": w, "shape": list(t.shape), "sha256": hashlib.sha256(t.numpy().tobytes()).hexdigest(),
                    "probes": {str(i): float(flat[i]) for i in probes}})
    with open(OUT, "w") as f:
        json.dump({"source": REF + ":128-145", "means": MEANS, "stds": STDS, "cases": out}, f, indent=1)
    print("wrote", OUT, len(out), "cases")
