"""Target for `compute-sanitizer --tool memcheck python tools/sanitize_prefill.py` (GPU box): the round-2 prefill kernels on small problems —
CTA-pair GEMM (forced for every M > 128) with each epilogue kind and M / N / K tails, single-CTA GEMM, tcgen05 attention for the three head
dims (causal and not, T with tails), vectorised RoPE + KV store — and one tiny continuous-batching stream."""
import os
import sys

sys.path.insert(0, os.getcwd())
import numpy as np
import torch

from emmax_b200 import OpenVLAForActionPrediction, tiny_config
from emmax_b200._lib import EPI_GELU, EPI_SWIGLU, call, ptr, stream
from emmax_b200.engine import Engine
from emmax_b200.synthetic import make_state_dict

BF = torch.bfloat16
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g, device="cuda").to(BF)  # noqa: E731
for mode in ("2", "0"):
    os.environ["EMX_GEMM_PAIR"] = mode
    for M, N, K in ((300, 520, 136), (777, 1000, 72), (129, 256, 64)):
        a, w = rnd(M, K), rnd(N, K)
        out = torch.zeros(M, N, dtype=BF, device="cuda")
        bias, ls, res = rnd(N), rnd(N), rnd(M, N)
        Engine.gemm(a, w, out)
        Engine.gemm(a, w, out, bias=bias, flags=EPI_GELU)
        buf = res.clone()
        Engine.gemm(a, w, buf, bias=bias, ls=ls, resid=buf)
        out2 = torch.zeros(M, N // 2, dtype=BF, device="cuda")
        Engine.gemm(a, w, out2, flags=EPI_SWIGLU)
        torch.cuda.synchronize()
        print("gemm ok", mode, M, N, K)
os.environ.pop("EMX_GEMM_PAIR")
for B, T, heads, hd, causal in ((1, 261, 2, 64, 0), (2, 256, 2, 72, 0), (1, 296, 2, 128, 1), (2, 37, 2, 128, 1), (1, 384, 1, 128, 1), (1, 5, 2, 72, 0)):
    qkv = rnd(B * T, 3 * heads * hd)
    out = torch.empty(B * T, heads * hd, dtype=BF, device="cuda")
    call("emx_attn_fwd", ptr(qkv), ptr(out), B, T, heads, hd, causal, hd ** -0.5, stream())
    torch.cuda.synchronize()
    print("attn ok", B, T, heads, hd, causal)
cfg = tiny_config()
sd = make_state_dict(cfg, seed=0, device="cpu")
model = OpenVLAForActionPrediction(cfg, sd, max_batch=8).to("cuda")
eng = model.engine
lens, limits = [12, 12, 20, 12, 31, 12, 12, 20, 20, 12], [3, 6, 2, 5, 6, 1, 4, 6, 3, 2]
rng = np.random.default_rng(1)
ids = [torch.tensor([[1] + rng.integers(3, 300, n - 1).tolist()], device="cuda") for n in lens]
pv = torch.randn(len(lens), 6, 224, 224, device="cuda").to(BF)
out = eng.serve([(ids[i], pv[i : i + 1], limits[i]) for i in range(len(lens))], eos_token_id=None, use_graph=False)
torch.cuda.synchronize()
print("serve ok", [o.numel() for o in out])
