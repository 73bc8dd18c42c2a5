"""Phase-level timing of the persistent decode kernel (CTA 0's view, %globaltimer) + ring wait counters.
Usage (GPU box): python tools/decode_probe.py [--steps 16] [--ctx-extra 0]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emmax_b200 import OpenVLAForActionPrediction, _lib, emma_x_config
from emmax_b200.synthetic import make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--warm", type=int, default=8)
args = ap.parse_args()

cfg = emma_x_config()
sd = make_state_dict(cfg, seed=0, device="cuda")
model = OpenVLAForActionPrediction(cfg, sd).to("cuda")
eng = model.engine
L = cfg.text_config.num_hidden_layers
ids = torch.tensor([[1] + np.random.default_rng(1234).integers(3, 31744, 39).tolist()], device="cuda")
pv = torch.randn(1, 6, 224, 224, device="cuda").to(torch.bfloat16)
eng.generate(ids, pv, 2, eos_token_id=None)
p = eng._decode_params(0)
lib = _lib.load()
dbg = torch.zeros(15 * L + 16 + 2 * 148 + 8, dtype=torch.int64, device="cuda")
skews = []
late = []
names = ["P1 rmsnorm+qkv", "barrier", "attention", "barrier", "load attn", "o_proj", "barrier", "rmsnorm2", "gate/up", "barrier",
         "load h", "down", "barrier"]
acc = np.zeros(13)
tail = np.zeros(3)
tot = []
cw, pw = [], []
for s in range(args.warm + args.steps):
    p.dbg = dbg.data_ptr() if s >= args.warm else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.emx_decode_step(C.byref(p), _lib.stream()))
    e1.record()
    torch.cuda.synchronize()
    if s >= args.warm:
        t = dbg.cpu().numpy()
        marks = t[: 13 * L + 4].astype(np.float64)
        d = np.diff(marks)
        per_layer = d[: 13 * L].reshape(L, 13) if False else None
        # interval k of layer l = marks[13l + k + 1] - marks[13l + k]
        iv = np.array([[marks[13 * l + k + 1] - marks[13 * l + k] for k in range(13)] for l in range(L)])
        acc += iv.mean(0)
        base = 13 * L
        tail += np.array([marks[base + 1] - marks[base], marks[base + 2] - marks[base + 1], marks[base + 3] - marks[base + 2]])
        tot.append(e0.elapsed_time(e1))
        arr = t[15 * L + 16 : 15 * L + 16 + 296].reshape(148, 2).astype(np.float64)
        skews.append((arr[:, 0].max() - arr[:, 0].min(), arr[:, 0].max() - np.median(arr[:, 0]), (arr[:, 1] - arr[:, 0].max()).mean(), (arr[:, 1] - arr[:, 0].max()).max()))
        late.append(arr[:, 0] - np.median(arr[:, 0]))
        cw.append(t[15 * L + 8])
        pw.append(t[15 * L + 9])
n = args.steps
print(f"kernel time: mean {np.mean(tot):.3f} ms  (min {np.min(tot):.3f})")
print("per-layer phase means (us), CTA 0:")
for nm, v in zip(names, acc / n / 1e3):
    print(f"  {nm:18s} {v:8.2f}")
print(f"  layer total        {acc.sum() / n / 1e3:8.2f}   x{L} = {acc.sum() / n / 1e6 * L:.3f} ms")
print("tail (us): final rmsnorm %.2f, lm_head %.2f, final barrier %.2f" % tuple(tail / n / 1e3))
clk = 1.965e9
print(f"consumer warp0 waited on weights: {np.mean(cw) / clk * 1e3:.3f} ms/step; producer waited on free slots: {np.mean(pw) / clk * 1e3:.3f} ms/step")
sk = np.array(skews) / 1e3
print("gate/up phase end across CTAs (layer 1): max-min %.2f us, max-median %.2f us; release after last arrival: mean %.2f us, max %.2f us" % tuple(sk.mean(0)))
late = np.array(late) / 1e3  # [steps, 148] us relative to the median CTA
m, sd = late.mean(0), late.std(0)
order = np.argsort(-m)
print("slowest CTAs (mean lateness us +- std over steps):", ", ".join(f"{i}:{m[i]:.2f}+-{sd[i]:.2f}" for i in order[:12]))
print("fastest CTAs:", ", ".join(f"{i}:{m[i]:.2f}+-{sd[i]:.2f}" for i in order[-8:]))
print("systematic part: std of per-CTA means %.2f us; mean of per-CTA stds %.2f us" % (m.std(), sd.mean()))
np.save("gpurun_out/cta_lateness.npy", late)
