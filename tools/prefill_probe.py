"""Tensor-core roofline probe of the prefill side (BASELINE.json configs[2]: bs=32, 224x224 images + 40-id prompt, 1 new token) and
of the bs=1 prefill that precedes every decode. Times the three stages of Engine._prefill_body separately with CUDA events
(no graph) and the whole body as one CUDA graph, and reports TFLOP/s against MEASURED_PEAKS.json (bf16 dense, sustained).
Algorithmic FLOPs: SURVEY.md §8(d) — vision+projector 405.2 GFLOP/image (used blocks only), LLM 2*6,476,005,376*S + 2*S^2*4096*32.
Usage (GPU box): python tools/prefill_probe.py [--batches 1,32] [--reps 5]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emmax_b200 import OpenVLAForActionPrediction, emma_x_config
from emmax_b200.synthetic import make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,32")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
batches = [int(b) for b in args.batches.split(",")]

cfg = emma_x_config()
sd = make_state_dict(cfg, seed=0, device="cuda")
model = OpenVLAForActionPrediction(cfg, sd, max_context=320, max_batch=max(batches)).to("cuda")
eng = model.engine
t = cfg.text_config
P, n_ids = cfg.num_patches, 40
S = P + n_ids


def vit_flops(v):
    D, T, M, hd, nh = v.embed_dim, v.num_tokens, v.mlp_dim, v.head_dim, v.num_heads
    per_block = 2 * T * D * 3 * D + 2 * T * D * D + 4 * T * D * M + 4 * T * T * hd * nh
    return v.used_depth * per_block + 2 * P * D * 3 * v.patch_size * v.patch_size


vd, H = cfg.vision_embed_dim, t.hidden_size
flops_vision = sum(vit_flops(v) for v in cfg.vision_dims)
flops_proj = 2 * P * (vd * 4 * vd + 4 * vd * H + H * H)
L, I, V = t.num_hidden_layers, t.intermediate_size, t.vocab_size
flops_llm = 2 * S * L * (4 * H * H + 3 * H * I) + 2 * S * S * H * L + 2 * V * H  # causal QK^T + PV counted as S^2 (half of 2x dense)
peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
peak = json.load(open(peaks))["bf16_tflops_sustained"] if os.path.exists(peaks) else 1390.0


def time_ms(fn, reps):
    fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return float(np.median(out))


for B in batches:
    ids = torch.tensor([[1] + np.random.default_rng(1234).integers(3, 31744, n_ids - 1).tolist()] * B, device="cuda")
    pv = torch.randn(B, 6, 224, 224, device="cuda").to(torch.bfloat16)
    ws = eng._workspace(B, n_ids)
    ws["ids"].copy_(ids)
    ws["pixels"].copy_(pv)
    t_v = time_ms(lambda: eng._vision(ws, B), args.reps)
    t_p = time_ms(lambda: eng._projector(ws), args.reps)
    t_l = time_ms(lambda: eng._llm_prefill(ws, B, n_ids), args.reps)
    t_g = time_ms(lambda: eng.prefill(ids, pv, use_graph=True), args.reps)
    tf = lambda fl, ms: fl * B / (ms * 1e-3) / 1e12  # noqa: E731
    total = flops_vision + flops_proj + flops_llm
    print(json.dumps({
        "batch": B, "prefill_positions": S,
        "vision_ms": round(t_v, 3), "vision_tflops": round(tf(flops_vision, t_v), 1),
        "projector_ms": round(t_p, 3), "projector_tflops": round(tf(flops_proj, t_p), 1),
        "llm_prefill_ms": round(t_l, 3), "llm_prefill_tflops": round(tf(flops_llm, t_l), 1),
        "whole_graph_ms": round(t_g, 3), "whole_tflops": round(tf(total, t_g), 1),
        "peak_tflops_sustained": peak, "frac_of_peak": round(tf(total, t_g) / peak, 3),
        "gflop_per_image": {"vision": round(flops_vision / 1e9, 1), "projector": round(flops_proj / 1e9, 1), "llm": round(flops_llm / 1e9, 1)},
    }), flush=True)  # fmt: skip
