"""Per-kernel parity (through the C ABI) against torch restatements of the reference ops, on seeded inputs.
bf16 outputs are compared with a 1-ulp-of-bf16 style tolerance (the kernels round at the same points as the eager ops
but sum in a different order); integer / fp64 de-tokeniser output must be bit-exact."""

import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


@pytest.fixture(scope="module")
def lib():
    from emmax_b200 import _lib

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(BF)


def assert_close_bf16(got, want, rel=2 ** -6, abs_=1e-3, frac=1.0, name="", scale=None):
    """|got - want| <= abs_ + rel * scale, scale = |want| unless given (sums with cancellation: pass operand magnitudes)."""
    got, want = got.float(), want.float()
    err = (got - want).abs()
    tol = abs_ + rel * (want.abs() if scale is None else scale.float())
    bad = (err > tol).float().mean().item()
    assert bad <= 1.0 - frac + 1e-12, f"{name}: {bad:.3%} of elements outside tolerance; max err {err.max().item():.4g}"


GEMM_SHAPES = [
    (261, 3072, 1024), (256, 1152, 4304), (256, 4304, 1152), (296, 12288, 4096), (256, 1024, 592), (1, 64, 64),
    (130, 136, 72), (300, 8704, 2176), (261, 1024, 4096), (5, 32064, 256), (512, 688 * 2, 256),
]  # fmt: skip


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_plain(lib, M, N, K):
    from emmax_b200.engine import Engine

    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    out = torch.zeros(M, N, dtype=BF, device="cuda")
    Engine.gemm(a, w, out)
    torch.cuda.synchronize()
    want = (a.float() @ w.float().T).to(BF)
    assert_close_bf16(out, want, name=f"gemm {M}x{N}x{K}")


PAIR_SHAPES = [(777, 1000, 136), (1024, 768, 64), (300, 520, 4096), (2368, 4096, 1024), (8352, 1024, 1024), (129, 256, 72)]


@pytest.mark.parametrize("M,N,K", PAIR_SHAPES)
def test_gemm_cta_pair(lib, M, N, K, monkeypatch):
    """The cta_group::2 kernel (256 x 256 tiles per CTA pair), forced for every M > 128: M / N / K tails, odd tile counts, and
    every fused epilogue, against the same torch restatement as the single-CTA kernel."""
    from emmax_b200._lib import EPI_GELU, EPI_SWIGLU
    from emmax_b200.engine import Engine

    monkeypatch.setenv("EMX_GEMM_PAIR", "2")
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    acc = a.float() @ w.float().T
    out = torch.zeros(M, N, dtype=BF, device="cuda")
    Engine.gemm(a, w, out)
    assert_close_bf16(out, acc.to(BF), name=f"pair gemm {M}x{N}x{K}")
    bias, ls, res = rnd(N, seed=5), rnd(N, seed=6).abs(), rnd(M, N, seed=7)
    Engine.gemm(a, w, out, bias=bias, flags=EPI_GELU)
    assert_close_bf16(out, torch.nn.functional.gelu((acc + bias.float()).to(BF)), name="pair bias+gelu")
    buf = res.clone()
    Engine.gemm(a, w, buf, bias=bias, ls=ls, resid=buf)
    want = ((acc + bias.float()).to(BF) * ls).to(BF) + res
    assert_close_bf16(buf, want, frac=0.99999, name="pair bias+ls+resid", scale=want.abs() + acc.abs() + res.abs().float() + 1)
    out2 = torch.zeros(M, N // 2, dtype=BF, device="cuda")
    Engine.gemm(a, w, out2, flags=EPI_SWIGLU)
    gu = acc.to(BF)
    # (a handful of the 10^7 products sit on a bf16 rounding boundary of gate or up: 1 ulp of the input moves the product by > tolerance)
    assert_close_bf16(out2, torch.nn.functional.silu(gu[:, 0::2]) * gu[:, 1::2], frac=0.99999, name="pair swiglu")
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(296, 4096, 4096), (261, 1024, 4096), (256, 1152, 1152), (296, 4096, 11008), (130, 136, 1032), (5, 520, 2048)])
def test_gemm_split_k(lib, M, N, K, monkeypatch):
    """emx_gemm_bf16_ws: small-M problems whose 128 x 128 tiles do not fill the machine are split along K (work item = tile x k-range, fp32
    partials in the caller's scratch, the last CTA of a tile adds them in split order and runs the fused epilogue). Same tolerances as the
    un-split kernel, every epilogue kind, and bit-identical results from run to run (the reduction order is fixed)."""
    from emmax_b200._lib import EPI_GELU, EPI_SWIGLU
    from emmax_b200.engine import Engine

    monkeypatch.setenv("EMX_GEMM_SPLITK", "1")  # off by default: measured slower than the un-split kernel (profiles/r02_splitk_negative.txt)
    scratch = torch.zeros(32 << 20, dtype=torch.uint8, device="cuda")
    a, w = rnd(M, K, seed=1), rnd(N, K, scale=K ** -0.5, seed=2)
    acc = a.float() @ w.float().T
    out = torch.zeros(M, N, dtype=BF, device="cuda")
    Engine.gemm(a, w, out, scratch=scratch)
    assert_close_bf16(out, acc.to(BF), name=f"split-K gemm {M}x{N}x{K}")
    again = torch.zeros_like(out)
    Engine.gemm(a, w, again, scratch=scratch)
    assert torch.equal(out, again), "split-K result differs from run to run"
    plain = torch.zeros_like(out)
    Engine.gemm(a, w, plain)  # un-split kernel: same product, different summation order
    assert_close_bf16(out, plain, name="split-K vs un-split")
    bias, ls, res = rnd(N, seed=5), rnd(N, seed=6).abs(), rnd(M, N, seed=7)
    Engine.gemm(a, w, out, bias=bias, flags=EPI_GELU, scratch=scratch)
    assert_close_bf16(out, torch.nn.functional.gelu((acc + bias.float()).to(BF)), name="split-K bias+gelu")
    buf = res.clone()
    Engine.gemm(a, w, buf, bias=bias, ls=ls, resid=buf, scratch=scratch)
    want = ((acc + bias.float()).to(BF) * ls).to(BF) + res
    assert_close_bf16(buf, want, frac=0.99999, name="split-K bias+ls+resid", scale=want.abs() + acc.abs() + res.abs().float() + 1)
    out2 = torch.zeros(M, N // 2, dtype=BF, device="cuda")
    Engine.gemm(a, w, out2, flags=EPI_SWIGLU, scratch=scratch)
    gu = acc.to(BF)
    assert_close_bf16(out2, torch.nn.functional.silu(gu[:, 0::2]) * gu[:, 1::2], frac=0.99999, name="split-K swiglu")
    assert int(scratch[:4096].view(torch.int32).abs().sum()) == 0, "tile counters must be back at zero after every launch"
    torch.cuda.synchronize()


def test_gemm_epilogues(lib):
    from emmax_b200._lib import EPI_GELU, EPI_SWIGLU
    from emmax_b200.engine import Engine

    M, N, K = 261, 1024, 512
    a, w = rnd(M, K, seed=3), rnd(N, K, scale=K ** -0.5, seed=4)
    bias, ls, res = rnd(N, seed=5), rnd(N, seed=6).abs(), rnd(M, N, seed=7)
    acc = a.float() @ w.float().T
    # bias + GELU
    out = torch.zeros(M, N, dtype=BF, device="cuda")
    Engine.gemm(a, w, out, bias=bias, flags=EPI_GELU)
    want = torch.nn.functional.gelu((acc + bias.float()).to(BF))
    assert_close_bf16(out, want, name="bias+gelu")
    # bias + LayerScale + residual, in place on the residual buffer (how the ViT block uses it)
    buf = res.clone()
    Engine.gemm(a, w, buf, bias=bias, ls=ls, resid=buf)
    want = ((acc + bias.float()).to(BF) * ls).to(BF) + res
    assert_close_bf16(buf, want, name="bias+ls+resid", scale=acc.abs() + res.abs().float() + 1)
    # residual only (Llama o_proj / down_proj)
    buf = res.clone()
    Engine.gemm(a, w, buf, resid=buf)
    assert_close_bf16(buf, acc.to(BF) + res, name="resid", scale=acc.abs() + res.abs().float())
    # broadcast residual rows (position embedding)
    pos = rnd(29, N, seed=8)
    out = torch.zeros(M, N, dtype=BF, device="cuda")
    Engine.gemm(a, w, out, bias=bias, resid=pos, resid_mod=29)
    want = (acc + bias.float()).to(BF) + pos[torch.arange(M, device="cuda") % 29]
    assert_close_bf16(out, want, name="resid_mod", scale=acc.abs() + 2)
    # SwiGLU over interleaved (gate, up) columns
    out = torch.zeros(M, N // 2, dtype=BF, device="cuda")
    Engine.gemm(a, w, out, flags=EPI_SWIGLU)
    gu = acc.to(BF)
    want = torch.nn.functional.silu(gu[:, 0::2]) * gu[:, 1::2]
    assert_close_bf16(out, want, name="swiglu")
    torch.cuda.synchronize()


def test_norms(lib):
    from emmax_b200._lib import call, ptr, stream

    for rows, dim in [(261, 1024), (256, 1152), (296, 4096), (3, 256), (7, 144)]:
        x, w, b = rnd(rows, dim, seed=1), (1 + 0.1 * rnd(dim, seed=2).float()).to(BF), rnd(dim, scale=0.1, seed=3)
        y = torch.empty_like(x)
        call("emx_layernorm", ptr(x), ptr(w), ptr(b), ptr(y), rows, dim, 1e-6, stream())
        want = torch.nn.functional.layer_norm(x, (dim,), w, b, 1e-6)
        assert_close_bf16(y, want, name=f"layernorm {rows}x{dim}")
        call("emx_rmsnorm", ptr(x), ptr(w), ptr(y), rows, dim, 1e-5, stream())
        xf = x.float()
        n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(BF)
        assert_close_bf16(y, w * n, name=f"rmsnorm {rows}x{dim}")
    torch.cuda.synchronize()


ATTN_CASES = [(1, 261, 16, 64, 0), (2, 256, 16, 72, 0), (1, 296, 32, 128, 1), (3, 37, 2, 128, 1), (1, 5, 2, 72, 0), (2, 384, 4, 128, 1),
              (2, 129, 3, 64, 1), (3, 300, 2, 72, 1), (1, 400, 2, 128, 1), (2, 385, 2, 64, 0), (5, 296, 8, 128, 0)]  # fmt: skip


@pytest.mark.parametrize("tc", ["1", "0"])
@pytest.mark.parametrize("B,T,heads,hd,causal", ATTN_CASES)
def test_attention(lib, B, T, heads, hd, causal, tc, monkeypatch):
    """tc=1: the tcgen05 kernel (scores resident in tensor memory) for T <= 384, the mma.sync kernel beyond; tc=0: mma.sync everywhere."""
    from emmax_b200._lib import call, ptr, stream

    monkeypatch.setenv("EMX_ATTN_TC", tc)
    qkv = rnd(B * T, 3 * heads * hd, seed=11)
    out = torch.empty(B * T, heads * hd, dtype=BF, device="cuda")
    call("emx_attn_fwd", ptr(qkv), ptr(out), B, T, heads, hd, causal, hd ** -0.5, stream())
    q, k, v = qkv.view(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4).float()
    want = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=bool(causal))
    want = want.transpose(1, 2).reshape(B * T, heads * hd)
    assert_close_bf16(out, want, rel=2 ** -6, abs_=4e-3, name="attention")


def test_rope_kvstore(lib):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb

    from emmax_b200._lib import call, ptr, stream

    B, T, heads, hd, page, max_pages = 2, 70, 4, 128, 64, 3
    qkv = rnd(B * T, 3 * heads * hd, seed=21)
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.int64, device="cuda").float() / hd))
    pos = torch.arange(256, device="cuda").float()
    fr = pos[:, None] * inv[None]
    cos_t, sin_t = fr.cos().to(BF).contiguous(), fr.sin().to(BF).contiguous()
    kc = torch.zeros(B * max_pages, heads, page, hd, dtype=BF, device="cuda")
    vc = torch.zeros_like(kc)
    # deliberately scrambled page ids
    tbl = torch.tensor([[4, 0, 2], [1, 5, 3]], dtype=torch.int32, device="cuda")
    ref = qkv.clone().view(B, T, 3, heads, hd)
    call("emx_rope_kvstore", ptr(qkv), B, T, heads, hd, ptr(cos_t), ptr(sin_t), 0, ptr(kc), ptr(vc), ptr(tbl), max_pages, page, stream())
    torch.cuda.synchronize()
    q, k, v = ref[:, :, 0].transpose(1, 2), ref[:, :, 1].transpose(1, 2), ref[:, :, 2].transpose(1, 2)  # [B, heads, T, hd]
    cos = torch.cat([cos_t[:T], cos_t[:T]], -1)[None].expand(B, -1, -1)
    sin = torch.cat([sin_t[:T], sin_t[:T]], -1)[None].expand(B, -1, -1)
    qe, ke = apply_rotary_pos_emb(q, k, cos, sin)
    got = qkv.view(B, T, 3, heads, hd)
    assert torch.equal(got[:, :, 0].transpose(1, 2), qe), "RoPE(q) must be bit-exact with transformers' bf16 formula"
    assert torch.equal(got[:, :, 1].transpose(1, 2), ke)
    for b in range(B):
        for t in (0, 1, 63, 64, 69):
            pg = int(tbl[b, t // page])
            assert torch.equal(kc[pg, :, t % page], ke[b, :, t])
            assert torch.equal(vc[pg, :, t % page], v[b, :, t])


@pytest.mark.parametrize("B,T", [(2, 296), (1, 296), (3, 200), (1, 40), (5, 131)])
def test_gemm_qkv_rope_fused_equals_two_kernels(lib, B, T, monkeypatch):
    """emx_gemm_qkv_rope: RoPE + paged-KV append fused into the q|k|v GEMM epilogue (CTA-pair kernel for M >= 512, single-CTA 128 x 256 tiles for
    a bs=1 prompt) must be BIT-identical to emx_gemm_bf16 + emx_rope_kvstore — packed qkv rows, K pages, V pages — with sequences that
    straddle tiles, a shuffled block table and a position offset; small problems take the two-kernel path inside the same call."""
    from emmax_b200._lib import call, ptr, stream

    heads, hd, K, page, max_pages, pos0 = 32, 128, 4096, 64, 8, 37
    H = heads * hd
    a, w = rnd(B * T, K, seed=51), rnd(3 * H, K, scale=K ** -0.5, seed=52)
    pos = torch.arange(pos0 + T + 8, device="cuda").float()
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device="cuda").float() / hd))
    cos, sin = torch.cos(pos[:, None] * inv).to(BF).contiguous(), torch.sin(pos[:, None] * inv).to(BF).contiguous()
    n_pages = B * max_pages
    table = torch.randperm(n_pages, generator=torch.Generator().manual_seed(3)).to(torch.int32).cuda().view(B, max_pages).contiguous()
    outs = []
    for fused in ("0", "1"):
        monkeypatch.setenv("EMX_QKV_ROPE_FUSED", fused)
        qkv = torch.zeros(B * T, 3 * H, dtype=BF, device="cuda")
        kc, vc = torch.zeros(n_pages, heads, page, hd, dtype=BF, device="cuda"), torch.zeros(n_pages, heads, page, hd, dtype=BF, device="cuda")
        call("emx_gemm_qkv_rope", ptr(a), K, ptr(w), K, ptr(qkv), B, T, heads, hd, K, ptr(cos), ptr(sin), pos0, ptr(kc), ptr(vc), ptr(table),
             max_pages, page, stream())
        torch.cuda.synchronize()
        outs.append((qkv, kc, vc))
    for name, x, y in zip(("qkv", "k pages", "v pages"), outs[0], outs[1]):
        assert torch.equal(x, y), f"{name}: fused epilogue differs from gemm + rope_kvstore ({(x != y).sum().item()} elements)"
    # and the two-kernel path is what the torch restatement says: v untouched, q / k rotated (rotate_half form, bf16 rounding points)
    qkv = outs[1][0].view(B, T, 3, heads, hd)
    lin = (a.float() @ w.float().T).to(BF).view(B, T, 3, heads, hd)
    assert_close_bf16(qkv[:, :, 2], lin[:, :, 2], name="v")
    c, s_ = cos[pos0 : pos0 + T][None, :, None, :].float(), sin[pos0 : pos0 + T][None, :, None, :].float()
    x1, x2 = lin[:, :, 0, :, : hd // 2].float(), lin[:, :, 0, :, hd // 2 :].float()
    want_q = torch.cat(((x1 * c).to(BF) + (-x2 * s_).to(BF), (x2 * c).to(BF) + (x1 * s_).to(BF)), dim=-1)
    assert_close_bf16(qkv[:, :, 0], want_q, frac=0.9999, name="q rope", scale=lin[:, :, 0].abs().float() + 1)
    # K rows of sequence b, token t live at page table[b, (pos0 + t) // 64], slot (pos0 + t) % 64
    b, t = B - 1, T - 1
    pg, slot = int(table[b, (pos0 + t) // page]), (pos0 + t) % page
    assert torch.equal(outs[1][1][pg, :, slot], qkv[b, t, 1]) and torch.equal(outs[1][2][pg, :, slot], qkv[b, t, 2])


def test_gemv_and_argmax(lib):
    from emmax_b200._lib import call, ptr, stream

    N, K = 32064, 4096
    w, x, r = rnd(N, K, scale=K ** -0.5, seed=31), rnd(K, seed=32), rnd(N, seed=33)
    y = torch.empty(N, dtype=BF, device="cuda")
    call("emx_gemv_bf16", ptr(w), K, ptr(x), ptr(y), ptr(r), N, K, stream())
    want = (w.float() @ x.float()).to(BF) + r
    assert_close_bf16(y, want, name="gemv+resid", scale=want.abs() + r.abs().float())
    logits = torch.empty(N, dtype=torch.float32, device="cuda")
    tok = torch.zeros(1, dtype=torch.int32, device="cuda")
    call("emx_lmhead_argmax", ptr(w), K, ptr(x), N, K, ptr(logits), ptr(tok), None, stream())
    torch.cuda.synchronize()
    assert int(tok) == int(torch.argmax(logits))
    assert_close_bf16(logits, (w.float() @ x.float()).to(BF), name="lm_head logits")
    # tie break: lowest index
    w2 = w.clone()
    w2[100] = w2[7]
    call("emx_lmhead_argmax", ptr(w2), K, ptr(w2[7].contiguous()), N, K, ptr(logits), ptr(tok), None, stream())
    torch.cuda.synchronize()
    assert int(tok) == int(torch.argmax(logits)) == 7


def test_batched_lm_head_rows(lib):
    """First tokens of a batched prefill: ONE tcgen05 GEMM of the B last hidden rows against lm_head + emx_argmax_rows_bf16, against the
    per-row GEMV path (emx_lmhead_argmax) and torch: bf16 logits of both within bf16 rounding, same argmax, lowest index on ties, the fp32
    copy exactly the bf16 values."""
    from emmax_b200._lib import call, ptr, stream
    from emmax_b200.engine import Engine

    B, N, K = 5, 32064, 4096
    w, x = rnd(N, K, scale=K ** -0.5, seed=35), rnd(B, K, seed=36)
    w[4242] = w[77]  # a tie: row 4242 repeats row 77 ...
    x[2] = w[77] * 8  # ... and sequence 2 points straight at it
    lb = torch.empty(B, N, dtype=BF, device="cuda")
    Engine.gemm(x, w, lb)
    l32 = torch.empty(B, N, dtype=torch.float32, device="cuda")
    tok = torch.full((B,), -1, dtype=torch.int32, device="cuda")
    call("emx_argmax_rows_bf16", ptr(lb), N, B, N, ptr(l32), ptr(tok), stream())
    torch.cuda.synchronize()
    assert torch.equal(l32, lb.float())
    want = (x.float() @ w.float().T).to(BF)
    assert_close_bf16(lb, want, name="batched lm_head logits")
    for b in range(B):
        assert int(tok[b]) == int(torch.argmax(l32[b])), b  # torch.argmax returns the first maximum too
        one = torch.empty(N, dtype=torch.float32, device="cuda")
        t1 = torch.zeros(1, dtype=torch.int32, device="cuda")
        call("emx_lmhead_argmax", ptr(w), K, ptr(x[b]), N, K, ptr(one), ptr(t1), None, stream())
        torch.cuda.synchronize()
        assert_close_bf16(l32[b], one, name=f"row {b}: GEMM vs GEMV logits")
    assert int(tok[2]) == 77 and float(l32[2, 77]) == float(l32[2, 4242])


def test_vit_frontend_kernels(lib):
    from emmax_b200._lib import call, ptr, stream

    B, D, P, prefix = 2, 128, 14, 5
    pix = rnd(B, 6, 224, 224, seed=41)
    kpad = 592
    out = torch.empty(B * 256, kpad, dtype=BF, device="cuda")
    call("emx_patch_im2col", ptr(pix), B, 6, 3, 224, 224, P, ptr(out), kpad, stream())
    want = torch.nn.functional.unfold(pix[:, 3:6].float(), kernel_size=P, stride=P).transpose(1, 2).reshape(B * 256, 588).to(BF)
    assert torch.equal(out[:, :588], want) and torch.all(out[:, 588:] == 0)
    pe, pos, pre = rnd(B * 256, D, seed=42), rnd(256, D, seed=43), rnd(prefix, D, seed=44)
    tok = torch.empty(B * 261, D, dtype=BF, device="cuda")
    call("emx_vit_assemble", ptr(pe), ptr(pos), ptr(pre), ptr(tok), B, 256, prefix, D, stream())
    want = torch.cat([pre[None].expand(B, -1, -1), pe.view(B, 256, D) + pos[None]], 1)
    assert torch.equal(tok.view(B, 261, D), want)
    feats = torch.zeros(B * 256, 272, dtype=BF, device="cuda")
    call("emx_vit_gather_features", ptr(tok), ptr(feats), B, 256, prefix, D, 272, 144, stream())
    torch.cuda.synchronize()
    assert torch.equal(feats[:, 144:], tok.view(B, 261, D)[:, prefix:].reshape(B * 256, D)) and torch.all(feats[:, :144] == 0)


def test_detokenize_bit_exact(lib, golden_dir):
    from emmax_b200._lib import call, ptr, stream

    with open(os.path.join(golden_dir, "detok_golden.json")) as f:
        g = json.load(f)
    ids = torch.tensor(g["ids"], dtype=torch.int32, device="cuda")
    norm = torch.empty(ids.numel(), dtype=torch.float64, device="cuda")
    call("emx_detokenize_actions", ptr(ids), ids.numel(), g["vocab_size"], 256, None, None, None, 0, ptr(norm), None, stream())
    want = np.array([float.fromhex(x) for x in g["decoded"]])
    assert np.array_equal(norm.cpu().numpy(), want), "device de-tokeniser must be bit-exact with the reference's numpy"
    u = g["unnorm"]
    ids = torch.tensor(u["ids"], dtype=torch.int32, device="cuda")
    q01 = torch.tensor(u["stats"]["q01"], dtype=torch.float64, device="cuda")
    q99 = torch.tensor(u["stats"]["q99"], dtype=torch.float64, device="cuda")
    mask = torch.tensor(u["stats"]["mask"], dtype=torch.uint8, device="cuda")
    norm, act = torch.empty(7, dtype=torch.float64, device="cuda"), torch.empty(7, dtype=torch.float64, device="cuda")
    call("emx_detokenize_actions", ptr(ids), 7, g["vocab_size"], 256, ptr(q01), ptr(q99), ptr(mask), 7, ptr(norm), ptr(act), stream())
    assert np.array_equal(act.cpu().numpy(), np.array([float.fromhex(x) for x in u["actions_hex"]]))


def test_preprocess_u8_bit_exact(lib):
    """GPU image transform == host PrismaticImageProcessor (torchvision to_tensor + normalize per backbone) + bf16 cast, bit for bit."""
    from PIL import Image

    from emmax_b200 import PrismaticImageProcessor

    proc = PrismaticImageProcessor()
    rng = np.random.default_rng(7)
    frames = rng.integers(0, 256, (3, 224, 224, 3), dtype=np.uint8)
    frames[0, :4] = 0
    frames[0, 4:8] = 255  # extremes
    want = torch.stack([proc.apply_transform(Image.fromarray(f)) for f in frames]).to(BF)
    got = proc.preprocess_device(torch.from_numpy(frames).cuda())
    assert got.shape == (3, 6, 224, 224) and got.dtype == BF
    assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16)), "device transform must be bit-exact with the host transform"
    with pytest.raises(ValueError):
        proc.preprocess_device(torch.zeros((1, 224, 224, 3), dtype=torch.float32, device="cuda"))


@pytest.mark.parametrize("h,w", [(256, 256), (480, 640), (200, 300)])
def test_resize_preprocess_u8_bit_exact(lib, h, w):
    """GPU resize (Pillow's antialiased bicubic in integer arithmetic) + normalise == the host processor, bit for bit, for frames that
    are not at the model's input size (256x256 sim frames, run_bridgev2_eval.py:161; camera frames)."""
    from PIL import Image

    from emmax_b200 import PrismaticImageProcessor

    proc = PrismaticImageProcessor()
    frames = np.random.default_rng(h * w).integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    want = torch.stack([proc.apply_transform(Image.fromarray(f)) for f in frames]).to(BF)
    got = proc.preprocess_device(torch.from_numpy(frames).cuda())
    assert got.shape == (2, 6, 224, 224)
    assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16)), "device resize + transform must be bit-exact with the host processor"


@pytest.mark.parametrize("h,w", [(480, 640), (300, 200), (224, 224), (257, 256)])
def test_letterbox_preprocess_bit_exact(lib, h, w):
    """`letterbox` strategy on the GPU (pad to square with int(255 * mean) of the last backbone, processing_prismatic.py:23-29, :130-131,
    then the antialiased bicubic resample) == the host processor, bit for bit; odd differences leave a non-square padded frame, as the
    reference's int((max - side) / 2) does."""
    from PIL import Image

    from emmax_b200 import PrismaticImageProcessor

    proc = PrismaticImageProcessor(image_resize_strategy="letterbox")
    frames = np.random.default_rng(h + w).integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    want = torch.stack([proc.apply_transform(Image.fromarray(f)) for f in frames]).to(BF)
    got = proc.preprocess_device(torch.from_numpy(frames).cuda())
    assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("h,w", [(256, 256), (480, 640), (224, 224)])
def test_center_crop_and_lanczos_twins_bit_exact(lib, h, w):
    """GPU twins of the robot loop's TensorFlow image steps == their numpy float32 host twins, bit for bit: the 0.9-area centre crop +
    bilinear resize (openvla_utils.py:81-124, :136-156) and resize_image's lanczos3 antialias resize (bridgev2_utils.py:152-166)."""
    from emmax_b200 import robot_utils as R

    rng = np.random.default_rng(h * 3 + w)
    frames = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    frames[0, : h // 8] = 255
    frames[0, h // 8 : h // 4] = 0
    got = R.center_crop_frame_device(torch.from_numpy(frames).cuda()).cpu().numpy()
    for b in range(2):
        assert np.array_equal(got[b], R.center_crop_frame(frames[b])), f"centre crop differs (frame {b})"
    assert np.array_equal(R.center_crop_frame_device(torch.from_numpy(frames[1]).cuda(), 0.8).cpu().numpy(), R.center_crop_frame(frames[1], 0.8))
    for size in ((224, 224), (128, 160)):
        got = R.lanczos3_resize_device(torch.from_numpy(frames[1]).cuda(), size).cpu().numpy()
        assert np.array_equal(got, R.lanczos3_resize(frames[1], size)), f"lanczos3 resize to {size} differs"
