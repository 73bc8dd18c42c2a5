"""One eager pass of the prefill body (vision towers, projector, Llama prefill) at batch B, to be run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel time shares of BASELINE.json configs[2] (see tools/ncu_summarize.py launches).
Usage (GPU box): ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv python tools/c3_pass.py [B]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from emmax_b200 import OpenVLAForActionPrediction, emma_x_config
from emmax_b200.synthetic import make_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = emma_x_config()
sd = make_state_dict(cfg, seed=0, device="cuda")
model = OpenVLAForActionPrediction(cfg, sd, max_context=320, max_batch=B).to("cuda")
eng = model.engine
n_ids = 40
ids = torch.tensor([[1] + np.random.default_rng(1234).integers(3, 31744, n_ids - 1).tolist()] * B, device="cuda")
pv = torch.randn(B, 6, 224, 224, device="cuda").to(torch.bfloat16)
ws = eng._workspace(B, n_ids)
ws["ids"].copy_(ids)
ws["pixels"].copy_(pv)
eng._vision(ws, B)
eng._projector(ws)
eng._llm_prefill(ws, B, n_ids)
torch.cuda.synchronize()
