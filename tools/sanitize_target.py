"""Small end-to-end target for `compute-sanitizer --tool memcheck python tools/sanitize_target.py` (GPU box): tiny configuration, prefill +
6 decode steps at a short (296) and a long (856, 4 TMEM passes per split) context. Last run: 0 errors (profiles/r01_sanitizer.txt)."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from emmax_b200 import OpenVLAForActionPrediction, tiny_config
from emmax_b200.synthetic import make_state_dict
cfg = tiny_config(); sd = make_state_dict(cfg, seed=0, device="cpu")
model = OpenVLAForActionPrediction(cfg, sd).to("cuda"); eng = model.engine
for n_ids in (40, 600):
    ids = torch.tensor([[1] + np.random.default_rng(1).integers(3, 300, n_ids - 1).tolist()], device="cuda")
    pv = torch.randn(1, 6, 224, 224, device="cuda").to(torch.bfloat16)
    out, _ = eng.generate(ids, pv, 6, eos_token_id=None, use_graph=False)
    torch.cuda.synchronize()
    print("ok", n_ids, out.tolist())
