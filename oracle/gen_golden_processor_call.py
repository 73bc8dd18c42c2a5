"""Freezes outputs of the reference's own `PrismaticProcessor.__call__` (processing_prismatic.py:187-216) for
tests/test_real_tokenizer_path.py (TEST INFRASTRUCTURE). The reference classes are loaded as in gen_golden_processor.py; the tokenizer is the
Llama-2-shaped fast tokenizer of tests/_llama_shaped_tokenizer.py (the real vocabulary is gated). `ProcessorMixin.__init__` of transformers
5.x validates its arguments against auto-class registries the reference never registered with, so the instance is created with `__new__`
and the two attributes `__call__` uses are set directly.
Usage (container with /root/reference):  python oracle/gen_golden_processor_call.py"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
OUT = os.path.join(ROOT, "tests", "golden", "processor_call_golden.json")

PROMPTS = ["In: What action should the robot take to achieve the instruction\nINSTRUCTION: \nput carrot in pot\nOut:", "In: pick up the spoon?\nOut:"]

if __name__ == "__main__":
    from _llama_shaped_tokenizer import build_llama_shaped_tokenizer
    from gen_golden_processor import MEANS, STDS, case_image, load_reference
    from transformers import AutoTokenizer

    ref = load_reference()
    d = tempfile.mkdtemp()
    build_llama_shaped_tokenizer(d)
    tok = AutoTokenizer.from_pretrained(d, model_max_length=2048, padding_side="right")
    ip = ref.PrismaticImageProcessor(use_fused_vision_backbone=True, image_resize_strategy="resize-naive", input_sizes=[(3, 224, 224)] * 2,
                                     interpolations=["bicubic"] * 2, means=MEANS, stds=STDS)  # fmt: skip
    proc = ref.PrismaticProcessor.__new__(ref.PrismaticProcessor)
    proc.image_processor, proc.tokenizer = ip, tok
    cases = []
    for prompt, (h, w) in zip(PROMPTS, [(224, 224), (256, 256)]):
        out = proc(prompt, case_image(h, w))
        cases.append({"prompt": prompt, "h": h, "w": w, "keys": list(out.keys()), "dtypes": {k: str(v.dtype) for k, v in out.items()},
                      "shapes": {k: list(v.shape) for k, v in out.items()}, "input_ids": out["input_ids"].tolist(),
                      "attention_mask": out["attention_mask"].tolist(),
                      "pixel_sha256": hashlib.sha256(out["pixel_values"].float().contiguous().numpy().tobytes()).hexdigest()})  # fmt: skip
    try:
        proc([PROMPTS[1], PROMPTS[1]], case_image(224, 224))  # two texts (same length: no padding needed), one image
        err = None
    except ValueError as e:
        err = str(e)
    with open(OUT, "w") as f:
        json.dump({"source": "/root/reference/prismatic/extern/hf/processing_prismatic.py:187-216", "cases": cases, "batch_error": err,
                   "model_input_names": proc.model_input_names}, f, indent=1)  # fmt: skip
    print("wrote", OUT, "batch_error:", err)
