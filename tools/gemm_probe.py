"""emx_gemm_bf16 head to head with cuBLAS (torch.matmul) on the GEMM shapes of the path: the ViT / projector / Llama-prefill Linears at
bs=32 (BASELINE.json configs[2]) and at bs=1. Times each with CUDA events (median of --reps after a warm-up; operands rotate through
buffers larger than L2? no: the operands of one problem are reused, as they are inside the prefill graph, where A was just written).
EMX_GEMM_PAIR=0|1|2 selects never / heuristic / always the CTA-pair kernel (csrc/gemm_tcgen05.cu).
Usage (GPU box): python tools/gemm_probe.py [--reps 20] [--modes 0,1,2]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emmax_b200._lib import EPI_SWIGLU
from emmax_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--modes", default="0,1,2")
ap.add_argument("--batches", default="32,1")
ap.add_argument("--warm", type=int, default=3)
ap.add_argument("--no-cublas", action="store_true")
args = ap.parse_args()
BF = torch.bfloat16


def shapes(B):
    d, s, l = 261 * B, 256 * B, 296 * B
    return [
        ("dino qkv", d, 3072, 1024, "bias"), ("dino proj", d, 1024, 1024, "bias+ls+res"), ("dino fc1", d, 4096, 1024, "bias+gelu"),
        ("dino fc2", d, 1024, 4096, "bias+ls+res"),
        ("siglip qkv", s, 3456, 1152, "bias"), ("siglip proj", s, 1152, 1152, "bias+res"), ("siglip fc1", s, 4304, 1152, "bias+gelu"),
        ("siglip fc2", s, 1152, 4304, "bias+res"),
        ("proj fc1", s, 8704, 2176, "bias+gelu"), ("proj fc2", s, 4096, 8704, "bias+gelu"), ("proj fc3", s, 4096, 4096, "bias"),
        ("llama qkv", l, 12288, 4096, ""), ("llama o", l, 4096, 4096, "res"), ("llama gateup", l, 22016, 4096, "swiglu"),
        ("llama down", l, 4096, 11008, "res"),
    ]  # fmt: skip


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
for B in [int(b) for b in args.batches.split(",")]:
    for name, M, N, K, epi in shapes(B):
        a = (torch.randn(M, K, generator=g, device="cuda")).to(BF)
        w = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(BF)
        bias = torch.randn(N, generator=g, device="cuda").to(BF) if "bias" in epi else None
        ls = torch.randn(N, generator=g, device="cuda").to(BF) if "ls" in epi else None
        swiglu = "swiglu" in epi
        out = torch.zeros(M, N // 2 if swiglu else N, dtype=BF, device="cuda")
        res = out if "res" in epi else None
        flags = EPI_SWIGLU if swiglu else (1 if "gelu" in epi else 0)
        ref = torch.empty(M, N, dtype=BF, device="cuda")
        t_cublas = time_ms(lambda: torch.matmul(a, w.T, out=ref), args.reps) if not args.no_cublas else float("nan")
        row = {"batch": B, "gemm": name, "M": M, "N": N, "K": K, "epilogue": epi, "cublas_plain_ms": round(t_cublas, 4),
               "cublas_tflops": round(2 * M * N * K / t_cublas / 1e9, 1)}  # fmt: skip
        for mode in args.modes.split(","):
            os.environ["EMX_GEMM_PAIR"] = mode
            t = time_ms(lambda: Engine.gemm(a, w, out, bias=bias, ls=ls, resid=res, flags=flags), args.reps)
            row[f"emx_mode{mode}_ms"] = round(t, 4)
            row[f"emx_mode{mode}_tflops"] = round(2 * M * N * K / t / 1e9, 1)
        # correctness of the last mode run against cuBLAS on the plain product (epilogue-free problems only)
        if epi == "":
            torch.cuda.synchronize()
            row["max_abs_diff_vs_cublas"] = float((out.float() - ref.float()).abs().max())
        print(json.dumps(row), flush=True)
os.environ.pop("EMX_GEMM_PAIR", None)
