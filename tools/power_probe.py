"""Power / clock behaviour of the decode kernel: runs the decode step back to back for a few seconds and samples nvidia-smi.
Usage (GPU box): python tools/power_probe.py [--seconds 4] [--lookahead 256]"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emmax_b200 import OpenVLAForActionPrediction, _lib, emma_x_config
from emmax_b200.synthetic import make_state_dict

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=4.0)
ap.add_argument("--lookahead", type=int, default=256)
args = ap.parse_args()
cfg = emma_x_config()
sd = make_state_dict(cfg, seed=0, device="cuda")
model = OpenVLAForActionPrediction(cfg, sd, max_context=2048).to("cuda")
eng = model.engine
ids = torch.tensor([[1] + np.random.default_rng(1234).integers(3, 31744, 39).tolist()], device="cuda")
pv = torch.randn(1, 6, 224, 224, device="cuda").to(torch.bfloat16)
eng.generate(ids, pv, 2, eos_token_id=None)
lib = _lib.load()
p = eng._decode_params(0)
p.l2_lookahead_kb = args.lookahead
samples, stop = [], threading.Event()
Q = "power.draw,power.limit,clocks.sm,clocks.mem,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"


def sampler():
    while not stop.is_set():
        out = subprocess.run(["nvidia-smi", "--id=0", f"--query-gpu={Q}", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
        samples.append(out)
        stop.wait(0.25)


print("idle:", subprocess.run(["nvidia-smi", "--id=0", f"--query-gpu={Q}", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip())
th = threading.Thread(target=sampler, daemon=True)
th.start()
t0 = time.time()
n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < args.seconds:
    for _ in range(100):
        _lib.check(lib.emx_decode_step(C.byref(p), _lib.stream()))
    n += 100
    torch.cuda.synchronize()
    if int(eng.d_state[1].item()) > 1800:  # stay inside the context capacity
        eng.generate(ids, pv, 2, eos_token_id=None)
e1.record()
torch.cuda.synchronize()
stop.set()
th.join()
print(f"{n} decode steps, {e0.elapsed_time(e1) / n:.3f} ms/step incl. host gaps")
print("columns:", Q)
for s in samples:
    print(" ", s)
