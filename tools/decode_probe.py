"""Phase-level timing of the persistent decode kernel (CTA 0's view, %globaltimer) + ring wait counters, swept over
(L2 look-ahead KiB, debug_flags) configurations with ONE model build.
Usage (GPU box): python tools/decode_probe.py [--steps 16] [--configs 0:0,256:0,256:16] [--ctx 296] [--brief]
debug_flags: 1 do not wait for LL tags, 2 skip attention (with 1), 4 no evict-first hint, 8 L2 prefetch also in catch-up mode, 16 drop LL stores, 64 skip MMAs, bits 8..19 idle prefetch pace, bits 20.. catch-up prefetch pace (10 ns / 64 KB)."""
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
ap.add_argument("--configs", default="0:0")
ap.add_argument("--brief", action="store_true")
ap.add_argument("--advance", type=int, default=0, help="decode this many extra tokens first (longer context)")
args = ap.parse_args()

cfg = emma_x_config()
sd = make_state_dict(cfg, seed=0, device="cuda")
model = OpenVLAForActionPrediction(cfg, sd).to("cuda")
eng = model.engine
L = cfg.text_config.num_hidden_layers
ids = torch.tensor([[1] + np.random.default_rng(1234).integers(3, 31744, 39).tolist()], device="cuda")
pv = torch.randn(1, 6, 224, 224, device="cuda").to(torch.bfloat16)
eng.generate(ids, pv, 2 + args.advance, eos_token_id=None)
lib = _lib.load()
names = ["x in + rmsnorm1", "q rows", "-", "k rows", "-", "v rows", "attn in (waits for attention)", "o_proj", "x in + rmsnorm2", "gate/up", "h in", "down"]
NM = len(names)  # timestamps per layer
clk = 1.965e9

for conf in args.configs.split(","):
    la, flags = (int(v) for v in conf.split(":"))
    p = eng._decode_params(0)
    p.l2_lookahead_kb, p.debug_flags = la, flags
    dbg = torch.zeros(15 * L + 16 + 2 * 148 + 8, dtype=torch.int64, device="cuda")
    skews, late, tot, cw, pw, ts, te = [], [], [], [], [], [], []
    acc, tail = np.zeros(NM), np.zeros(3)
    att = np.zeros(7)
    # the product kernel (no dbg buffer: the un-instrumented twin), timed over args.steps launches with one event pair
    p.dbg = None
    for s in range(args.warm):
        _lib.check(lib.emx_decode_step(C.byref(p), _lib.stream()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        _lib.check(lib.emx_decode_step(C.byref(p), _lib.stream()))
    e1.record()
    torch.cuda.synchronize()
    print(f"-- lookahead {la} KiB, debug_flags {flags}: PRODUCT kernel {e0.elapsed_time(e1) / args.steps:.3f} ms/token (ctx {int(eng.d_state[1].item())})", flush=True)
    for s in range(args.warm + args.steps):
        p.dbg = dbg.data_ptr() if s >= args.warm else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.emx_decode_step(C.byref(p), _lib.stream()))
        e1.record()
        torch.cuda.synchronize()
        if s >= args.warm:
            t = dbg.cpu().numpy()
            marks = t[: NM * L + 4].astype(np.float64)
            iv = np.array([[marks[NM * l + k + 1] - marks[NM * l + k] for k in range(NM)] for l in range(L)])
            acc += iv.mean(0)
            base = NM * L
            att += t[15 * L + 16 + 148 : 15 * L + 16 + 155].astype(np.float64) / L
            tail += np.array([marks[base + 1] - marks[base], marks[base + 2] - marks[base + 1], marks[base + 3] - marks[base + 2]])
            tot.append(e0.elapsed_time(e1))
            arr = t[15 * L + 16 : 15 * L + 16 + 148].astype(np.float64)  # end of layer 1 on every CTA
            skews.append((arr.max() - arr.min(), arr.max() - np.median(arr)))
            late.append(arr - np.median(arr))
            cw.append(t[15 * L + 8])
            pw.append(t[15 * L + 9])
            ts.append(t[15 * L + 11])
            te.append(t[15 * L + 12])
            pfb = t[15 * L + 13]
            gp = t[15 * L + 4 : 15 * L + 8] / clk * 1e6 / (2 * L + 1)
    n = args.steps
    a = acc / n / 1e3
    sk = np.array(skews).mean(0) / 1e3
    print(f"== lookahead {la} KiB, debug_flags {flags}: kernel {np.mean(tot):.3f} ms (min {np.min(tot):.3f}) | per layer {a.sum():.2f} us: "
          f"weights {a[1] + a[3] + a[5] + a[7] + a[9] + a[11]:.2f} (q {a[1]:.2f} k {a[3]:.2f} v {a[5]:.2f} o {a[7]:.2f} gateup {a[9]:.2f} down {a[11]:.2f}) exchanges "
          f"{a[0] + a[2] + a[4] + a[6] + a[8] + a[10]:.2f} (x {a[0]:.2f} attn {a[6]:.2f} xo {a[8]:.2f} h {a[10]:.2f}) | attention warps of CTA 0 (us/layer): wait q {att[0] / n / 1e3:.2f}, cached keys {att[1] / n / 1e3:.2f} (scores done at {att[4] / n / 1e3:.2f}, softmax at {att[5] / n / 1e3:.2f}), new key {att[2] / n / 1e3:.2f}, combine+publish {att[3] / n / 1e3:.2f}, staging next layer into TMEM {att[6] / n / 1e3:.2f} | lm_head {tail[1] / n / 1e3:.1f} us | "
          f"consumer wait {np.mean(cw) / clk * 1e3:.3f} ms, warp0 partial-sync {np.mean(ts) / clk * 1e3:.3f} ms, epilogue {np.mean(te) / clk * 1e3:.3f} ms, CTA0 prefetched {pfb / 1e6:.1f} MB of 89.3 | layer-1 end skew max-min {sk[0]:.2f} us | gather+norm us/call: loads {gp[0]:.2f} ln-wait {gp[1]:.2f} sum {gp[2]:.2f} norm+bar {gp[3]:.2f}", flush=True)  # fmt: skip
    if args.brief:
        continue
    print("per-layer phase means (us), CTA 0:")
    for nm, v in zip(names, a):
        print(f"  {nm:18s} {v:8.2f}")
    print("tail (us): final rmsnorm %.2f, lm_head %.2f, final barrier %.2f" % tuple(tail / n / 1e3))
    print(f"consumer warp0 waited on weights: {np.mean(cw) / clk * 1e3:.3f} ms/step; producer waited on free slots: {np.mean(pw) / clk * 1e3:.3f} ms/step")
    print("end of layer 1 across CTAs: max-min %.2f us, max-median %.2f us" % tuple(sk))
    lt = np.array(late) / 1e3  # [steps, 148] us relative to the median CTA
    m, sdv = lt.mean(0), lt.std(0)
    order = np.argsort(-m)
    print("slowest CTAs (mean lateness us +- std over steps):", ", ".join(f"{i}:{m[i]:.2f}+-{sdv[i]:.2f}" for i in order[:12]))
    print("systematic part: std of per-CTA means %.2f us; mean of per-CTA stds %.2f us" % (m.std(), sdv.mean()))
