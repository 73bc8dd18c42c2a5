"""Phase-level timing of the batched decode kernel (CTA 0's view, %globaltimer of the instrumented twin): per phase kind the time spent
in the gather (+ RMSNorm) that precedes it and in the phase body, averaged over layers and launches; and the product kernel's ms/launch for
several batch sizes.  Usage (GPU box): python tools/batch_probe.py [--steps 16] [--batches 1,4,8] [--advance 0]"""
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
ap.add_argument("--warm", type=int, default=4)
ap.add_argument("--batches", default="1,4,8")
ap.add_argument("--advance", type=int, default=0, help="decode this many extra tokens first (longer context)")
ap.add_argument("--percta", action="store_true", help="print which CTAs arrive last at the gathers of layer 1")
ap.add_argument("--lookahead", default="6", help="comma list of l2_lookahead_stages values to try")
args = ap.parse_args()

cfg = emma_x_config()
sd = make_state_dict(cfg, seed=0, device="cuda")
model = OpenVLAForActionPrediction(cfg, sd, max_batch=8, max_context=1024).to("cuda")
eng = model.engine
L = cfg.text_config.num_hidden_layers
lib = _lib.load()
KINDS = ["q", "k", "v", "att", "o", "gateup", "down"]
n_steps = 7 * L + 1

GRID = lib.emx_decode_grid()
for B, LA in [(int(b), int(la)) for b in args.batches.split(",") for la in args.lookahead.split(",")]:
    ids = torch.tensor([[1] + np.random.default_rng(1234).integers(3, 31744, 39).tolist()] * B, device="cuda")
    pv = torch.randn(B, 6, 224, 224, device="cuda").to(torch.bfloat16)
    eng.generate_batch(ids, pv, 2 + args.advance, eos_token_id=None)
    st = eng.b_state
    st[32 : 32 + B].fill_(10_000)  # limit: keep every sequence active
    p = eng._decode_batch_params(B)
    p.l2_lookahead_stages = LA
    for _ in range(args.warm):
        _lib.check(lib.emx_decode_batch_step(C.byref(p), _lib.stream()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        _lib.check(lib.emx_decode_batch_step(C.byref(p), _lib.stream()))
    e1.record()
    for _ in range(60):  # keep the GPU busy while the clocks are sampled
        _lib.check(lib.emx_decode_batch_step(C.byref(p), _lib.stream()))
    import subprocess
    clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader"],
                         capture_output=True, text=True).stdout.strip()
    torch.cuda.synchronize()
    print(f"   [clocks under load: {clk}]")
    ctx = int(st[8].item())
    print(f"== batch {B}, L2 look-ahead {LA} stages, context {ctx}: PRODUCT kernel {e0.elapsed_time(e1) / args.steps:.3f} ms/launch", flush=True)
    dbg = torch.zeros(2 * n_steps + 8 + 16 * GRID + 8, dtype=torch.int64, device="cuda")
    p.dbg = dbg.data_ptr()
    gath, body, tot, skew, entries = np.zeros(8), np.zeros(8), [], [], []
    for _ in range(args.steps):
        _lib.check(lib.emx_decode_batch_step(C.byref(p), _lib.stream()))
        torch.cuda.synchronize()
        t = dbg.cpu().numpy().astype(np.float64)
        tot.append((t[2 * n_steps] - t[0]) / 1e3)
        g = t[2 * n_steps + 8 : 2 * n_steps + 8 + 8 * GRID].reshape(GRID, 4, 2)
        entries.append(g[:, :, 0] - g[:, :, 0].min(axis=0, keepdims=True))
        skew.append([(g[:, k, 0].max() - g[:, k, 0].min(), g[:, k, 1].max() - g[:, k, 1].min(), np.median(g[:, k, 1] - g[:, k, 0]), (g[:, k, 1] - g[:, k, 0]).min()) for k in range(4)])
        for s in range(n_steps):
            k = 7 if s == n_steps - 1 else s % 7
            gath[k] += t[2 * s + 1] - t[2 * s]
            body[k] += t[2 * s + 2] - t[2 * s + 1]
    n = args.steps
    print(f"   instrumented twin: {np.mean(tot):.1f} us per launch (CTA 0, first gather -> end of lm_head)")
    print("   per layer (us):  " + "  ".join(f"{KINDS[k]}: gather {gath[k] / n / L / 1e3:.2f} + body {body[k] / n / L / 1e3:.2f}" for k in range(7)))
    print(f"   per launch (us): gathers {gath[:7].sum() / n / 1e3:.0f}, bodies {body[:7].sum() / n / 1e3:.0f}, lm_head gather {gath[7] / n / 1e3:.1f} + body {body[7] / n / 1e3:.1f}")
    gp = dbg.cpu().numpy()[2 * n_steps + 8 + 16 * GRID :].astype(np.float64) / n  # accumulated over the launches
    g2 = dbg.cpu().numpy()[2 * n_steps + 8 + 8 * GRID : 2 * n_steps + 8 + 16 * GRID].astype(np.float64).reshape(GRID, 4, 2)  # last launch
    g1 = dbg.cpu().numpy()[2 * n_steps + 8 : 2 * n_steps + 8 + 8 * GRID].astype(np.float64).reshape(GRID, 4, 2)
    for k, nm in enumerate(["q", "o", "gateup", "down"]):
        t0 = g1[:, k, 0].min()
        print(f"   {nm} (last launch, us after the first CTA entered): last entry {(g1[:, k, 0].max() - t0) / 1e3:.1f}, last CTA past its entry barrier {(g2[:, k, 0].max() - t0) / 1e3:.1f} "
              f"(median {(np.median(g2[:, k, 0]) - t0) / 1e3:.1f}), counter seen complete: first {(g2[:, k, 1].min() - t0) / 1e3:.1f} / last {(g2[:, k, 1].max() - t0) / 1e3:.1f}, "
              f"last exit {(g1[:, k, 1].max() - t0) / 1e3:.1f}")
    print(f"   gathers of CTA 0, thread 0, per launch (us): cbar {gp[0] / 1e3:.0f}, arrival counter {gp[1] / 1e3:.0f}, free slots {gp[2] / 1e3:.0f}, copy landed {gp[3] / 1e3:.0f}, "
          f"read+park {gp[4] / 1e3:.0f}, norm tail {gp[5] / 1e3:.0f}; units repaired by thread 0: {gp[6] * 1.0:.0f} in {4 * L} gathers")
    if args.percta:
        ent = np.mean(np.array(entries), axis=0) / 1e3  # [GRID, 4] entry time relative to the earliest CTA, us
        for k, nm in enumerate(["q", "o", "gateup", "down"]):
            order = np.argsort(-ent[:, k])
            print(f"   {nm}: latest CTAs " + " ".join(f"{c}:{ent[c, k]:.1f}" for c in order[:16]) + "  | earliest " + " ".join(f"{c}:{ent[c, k]:.1f}" for c in order[-6:]))
        print("   mean lateness per CTA over the 4 gathers, top 20: " + " ".join(f"{c}:{v:.1f}" for c, v in sorted(enumerate(ent.mean(1)), key=lambda x: -x[1])[:20]))
    sk = np.mean(np.array(skew), axis=0) / 1e3
    print("   layer 1, all CTAs (us): " + "  ".join(f"{nm}: entry spread {sk[k, 0]:.1f}, exit spread {sk[k, 1]:.1f}, gather median {sk[k, 2]:.1f} / min {sk[k, 3]:.1f}"
                                                    for k, nm in enumerate(["q", "o", "gateup", "down"])), flush=True)
