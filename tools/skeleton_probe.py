"""Dataflow skeleton of the decode kernel (emx_debug_skeleton): what do grid barriers and consumer stalls cost under HBM load?
Usage (GPU box): python tools/skeleton_probe.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from emmax_b200 import _lib

lib = _lib.load()
GRID = 148
nbytes = 12 * 1024**3
buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
buf.random_(0, 255)
sync = torch.zeros(32 * (2 * GRID + 2), dtype=torch.int32, device="cuda")
out = torch.zeros(32 + 2 * GRID, dtype=torch.int64, device="cuda")
st = _lib.stream()


def run(label, phases, stalls, reps=32, rows=16, seg=4096, stride=8192, stages=3, consume=0, variant=1, n=5, n_prod=2, n_cons=8, dist=False, weight=None, timers=1, quiet=False, pf=0, pf_mode=0, pace=1400, launch=0):
    ph = (C.c_int * len(phases))(*phases)
    sl = (C.c_int * len(stalls))(*stalls)
    region = nbytes // GRID // (rows * stride) * (rows * stride)
    times = []
    for i in range(n + 2):
        sync.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.emx_debug_skeleton(buf.data_ptr(), region, len(phases), ph, sl, reps, rows, seg, stride, stages, consume, variant,
                                          n_prod, n_cons, None if weight is None else weight.data_ptr(), timers, pf, pf_mode, pace, launch, sync.data_ptr(),
                                          out.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            times.append(e0.elapsed_time(e1))
    o = out.cpu().tolist()
    ms = sum(times) / len(times)
    nb = sum(phases) * reps * rows * seg * GRID
    per_phase = " ".join(f"{o[i] / reps / 1e3:.1f}+{o[8 + i] / reps / 1e3:.1f}" for i in range(len(phases)))
    import numpy as np
    t = np.array(o[32 : 32 + GRID]) / 1e6
    sm = np.array(o[32 + GRID : 32 + 2 * GRID])
    if not quiet:
        print(f"{label:58s} {ms:7.3f} ms  {nb / ms / 1e6:7.0f} GB/s  per rep (phase+barrier us): {per_phase}", flush=True)
    if dist:
        order = np.argsort(t)
        print(f"    per-CTA total ms: min {t.min():.3f} p10 {np.percentile(t, 10):.3f} median {np.median(t):.3f} mean {t.mean():.3f} p90 {np.percentile(t, 90):.3f} max {t.max():.3f}")
        print("    fastest (cta:smid:ms):", " ".join(f"{i}:{sm[i]}:{t[i]:.3f}" for i in order[:10]))
        print("    slowest (cta:smid:ms):", " ".join(f"{i}:{sm[i]}:{t[i]:.3f}" for i in order[-10:]))
    return ms, t, sm


LAYER = [11, 0, 4, 19, 9]  # 64 KB stages per CTA: qkv | (attention) | o | gate/up | down  = 2752 KB (real: 2735 KB)
STALL = [8000, 700, 1000, 1000, 1000]  # attention, load attn, rmsnorm2, load h, rmsnorm1
NOST = [0, 0, 0, 0, 0]

import numpy as np

print("--- free-run streaming: producer warp placement (launch bit 1: producers are warps 1,2 instead of 8,9)")
run("1 producer (warp 8), 1 consumer", [43], [0], variant=0, n_prod=1, n_cons=1, timers=0)
run("1 producer (warp 1), 1 consumer", [43], [0], variant=0, n_prod=1, n_cons=1, timers=0, launch=2)
run("2 producers (warps 1,2), 1 consumer", [43], [0], variant=0, n_prod=2, n_cons=1, timers=0, launch=2)
run("2 producers (warps 8,9), 1 consumer", [43], [0], variant=0, n_prod=2, n_cons=1, timers=0)
run("1 producer (warp 1), 1 consumer, 2 x 64 KB", [43], [0], variant=0, n_prod=1, n_cons=1, timers=0, launch=2, stages=2)
