"""emx_attn_fwd (tcgen05 kernel, scores resident in tensor memory; and the mma.sync kernel it replaced) head to head with flash_attn 2.8.3
(`flash_attn_qkvpacked_func`, the kernel behind the reference's attn_implementation="flash_attention_2") on the attention shapes of the
path: SigLIP [B,16,256,72], DINOv2 [B,16,261,64] (non-causal), Llama prefill [B,32,296,128] (causal). CUDA events, median of --reps.
Usage (GPU box): python tools/attn_probe.py [--reps 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emmax_b200._lib import call, ptr, stream

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--warm", type=int, default=3)
args = ap.parse_args()
try:
    from flash_attn import flash_attn_qkvpacked_func
except Exception as e:  # noqa: BLE001
    flash_attn_qkvpacked_func = None
    print(f"# flash_attn not importable: {e}")


def time_ms(fn, reps):
    for _ in range(args.warm):
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


g = torch.Generator(device="cuda").manual_seed(0)
for B, H, T, D, causal in [(32, 16, 256, 72, 0), (32, 16, 261, 64, 0), (1, 16, 256, 72, 0), (1, 16, 261, 64, 0), (1, 32, 296, 128, 1),
                           (8, 32, 296, 128, 1), (32, 32, 296, 128, 1)]:  # fmt: skip
    qkv = torch.randn(B * T, 3 * H * D, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.empty(B * T, H * D, dtype=torch.bfloat16, device="cuda")
    row = {"B": B, "heads": H, "T": T, "head_dim": D, "causal": causal}
    flops = 4 * B * H * T * T * D / (2 if causal else 1)
    for name, env in (("emx_tcgen05", "1"), ("emx_mma_sync", "0")):
        os.environ["EMX_ATTN_TC"] = env
        t = time_ms(lambda: call("emx_attn_fwd", ptr(qkv), ptr(out), B, T, H, D, causal, D ** -0.5, stream()), args.reps)
        row[name + "_us"] = round(t * 1e3, 1)
        row[name + "_tflops"] = round(flops / t / 1e9, 1)
    if flash_attn_qkvpacked_func is not None:
        q5 = qkv.view(B, T, 3, H, D)
        t = time_ms(lambda: flash_attn_qkvpacked_func(q5, causal=bool(causal), softmax_scale=D ** -0.5), args.reps)
        row["flash_attn_2_8_3_us"] = round(t * 1e3, 1)
        row["emx_tcgen05_over_flash_attn"] = round(t * 1e3 / row["emx_tcgen05_us"], 2)
        ref = flash_attn_qkvpacked_func(q5, causal=bool(causal), softmax_scale=D ** -0.5).reshape(B * T, H * D)
        os.environ["EMX_ATTN_TC"] = "1"
        call("emx_attn_fwd", ptr(qkv), ptr(out), B, T, H, D, causal, D ** -0.5, stream())
        torch.cuda.synchronize()
        row["max_abs_diff_vs_flash_attn"] = float((out.float() - ref.float()).abs().max())
    print(json.dumps(row), flush=True)
os.environ.pop("EMX_ATTN_TC", None)
