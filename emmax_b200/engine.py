"""
GPU engine: packs a reference-named state dict into the layouts the kernels want and drives the hot path
(vision towers -> projector -> Llama prefill -> persistent decode steps) through the C ABI of libemmax.so.

PyTorch is used for device memory, streams and CUDA-graph capture only; every FLOP runs in the hand-written kernels.
Reference call stack being replaced: SURVEY.md §3.1 (HF path), i.e.
/root/reference/prismatic/extern/hf/modeling_prismatic.py:362-415 (multimodal forward), :325-341 (cached step).
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import ATT_MAX_SEGMENTS, EPI_GELU, EPI_SWIGLU, MAX_DECODE_BATCH, DecodeBatchParams, DecodeBatchState, DecodeParams, DecodeState, call, ptr, stream
from .configuration import OpenVLAConfig, ViTDims

BF16 = torch.bfloat16


def _on_engine_device(fn):
    """Run a method with the engine's GPU as the current CUDA device: the kernels are launched on `torch.cuda.current_stream()`, which
    belongs to the CURRENT device, while every pointer they receive lives on `self.device` (a process may hold engines on several GPUs)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)

    return wrapper


def admission_order(limits, order: str = "fifo") -> List[int]:
    """Order in which `Engine.serve` admits a stream of requests into the sequence slots. "fifo": arrival order. "longest_first": largest
    token limit first, ties in arrival order (longest-processing-time-first list scheduling: with S slots the stream then ends at most one
    SHORTEST request after the ideal sum(limits) / S launches, instead of one longest request after it)."""
    idx = list(range(len(limits)))
    if order == "fifo":
        return idx
    if order == "longest_first":
        return sorted(idx, key=lambda r: -int(limits[r]))  # sorted() is stable: ties keep arrival order
    raise ValueError(f"order must be 'fifo' or 'longest_first', got {order!r}")


def _ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass
class _ViTWeights:
    dims: ViTDims
    kpad: int
    w_pe: torch.Tensor  # [D, kpad]
    b_pe: torch.Tensor
    pos: torch.Tensor  # [n_patches, D]
    prefix: Optional[torch.Tensor]  # [n_prefix, D]
    blocks: List[Dict[str, torch.Tensor]]


class Engine:
    PAGE = 64

    def __init__(self, config: OpenVLAConfig, state_dict: Dict[str, torch.Tensor], device: torch.device,
                 max_batch: int = 1, max_context: int = 1024, kv_splits: int = 4) -> None:  # fmt: skip
        kv_splits = int(os.environ.get("EMX_KV_SPLITS", kv_splits))
        _lib.load()  # raises if libemmax.so is missing: there is no fallback path
        if device.type != "cuda":
            raise _lib.EmxError("emmax_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.config, self.device = config, device
        self.t = config.text_config
        self.max_batch, self.kv_splits = max_batch, kv_splits
        self.l2_lookahead_kb = int(os.environ.get("EMX_L2_LOOKAHEAD_KB", "256"))  # idle-triggered L2 prefetch window per CTA
        self.max_context = _ceil_to(max_context, self.PAGE)
        self._graphs: Dict[Tuple[int, int, int], Tuple[torch.cuda.CUDAGraph, int]] = {}
        self.last_decode = None
        self._side: Optional[torch.cuda.Stream] = None  # second vision tower at small batch
        with torch.cuda.device(device):
            self._pack(state_dict)
            self._alloc()

    # ------------------------------------------------------------------------------------------------------------
    def _pack(self, sd: Dict[str, torch.Tensor]) -> None:
        dev = self.device

        def take(name: str) -> torch.Tensor:
            return sd.pop(name).to(device=dev, dtype=BF16).contiguous()

        self.vits: List[_ViTWeights] = []
        for prefix, v in zip(("vision_backbone.featurizer.", "vision_backbone.fused_featurizer."), self.config.vision_dims):
            D, kk = v.embed_dim, 3 * v.patch_size * v.patch_size
            kpad = _ceil_to(kk, 8)
            w = take(prefix + "patch_embed.proj.weight").reshape(D, kk)
            w_pe = torch.zeros((D, kpad), dtype=BF16, device=dev)
            w_pe[:, :kk] = w
            pre = None
            if v.num_prefix_tokens > 0:
                pre = torch.cat([take(prefix + "cls_token")[0], take(prefix + "reg_token")[0]], dim=0).contiguous()
            blocks = []
            for i in range(v.depth):
                b = f"{prefix}blocks.{i}."
                names = ["norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
                         "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias"]  # fmt: skip
                if v.layerscale:
                    names += ["ls1.scale_factor", "ls2.scale_factor"]
                if i < v.used_depth:
                    blocks.append({n: take(b + n) for n in names})
                else:  # the last block cannot influence `get_intermediate_layers(n={depth-2})`: never uploaded
                    for n in names:
                        sd.pop(b + n, None)
            self.vits.append(_ViTWeights(v, kpad, w_pe, take(prefix + "patch_embed.proj.bias"), take(prefix + "pos_embed")[0], pre, blocks))
            for k in [k for k in sd if k.startswith(prefix)]:  # unused: final norm, SigLIP attn_pool
                sd.pop(k)

        self.proj = {n: take(f"projector.{n}") for n in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias")}

        t, p = self.t, "language_model.model."
        L, H, I = t.num_hidden_layers, t.hidden_size, t.intermediate_size
        self.embed = take(p + "embed_tokens.weight")
        self.w_qkv = torch.empty((L, 3 * H, H), dtype=BF16, device=dev)
        self.w_o = torch.empty((L, H, H), dtype=BF16, device=dev)
        self.w_gateup = torch.empty((L, 2 * I, H), dtype=BF16, device=dev)  # row 2i = gate_i, row 2i+1 = up_i
        self.w_down = torch.empty((L, H, I), dtype=BF16, device=dev)
        self.ln1 = torch.empty((L, H), dtype=BF16, device=dev)
        self.ln2 = torch.empty((L, H), dtype=BF16, device=dev)
        for i in range(L):
            b = f"{p}layers.{i}."
            for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
                self.w_qkv[i, j * H : (j + 1) * H] = sd.pop(b + f"self_attn.{n}.weight").to(dev)
            self.w_o[i] = sd.pop(b + "self_attn.o_proj.weight").to(dev)
            self.w_gateup[i, 0::2] = sd.pop(b + "mlp.gate_proj.weight").to(dev)
            self.w_gateup[i, 1::2] = sd.pop(b + "mlp.up_proj.weight").to(dev)
            self.w_down[i] = sd.pop(b + "mlp.down_proj.weight").to(dev)
            self.ln1[i] = sd.pop(b + "input_layernorm.weight").to(dev)
            self.ln2[i] = sd.pop(b + "post_attention_layernorm.weight").to(dev)
        self.final_norm = take(p + "norm.weight")
        self.lm_head = take("language_model.lm_head.weight")

        # RoPE tables exactly as transformers builds them (fp32 outer product -> cos/sin -> cast to the model dtype)
        hd = t.head_dim
        inv_freq = 1.0 / (t.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64, device=dev).float() / hd))
        pos = torch.arange(self.max_context, device=dev).float()
        freqs = pos[:, None] * inv_freq[None, :]
        self.cos_tab, self.sin_tab = freqs.cos().to(BF16).contiguous(), freqs.sin().to(BF16).contiguous()

    def _alloc(self) -> None:
        dev, t, B = self.device, self.t, self.max_batch
        H, I, L = t.hidden_size, t.intermediate_size, t.num_hidden_layers
        self.pages_per_seq = self.max_context // self.PAGE
        self.n_pages = self.pages_per_seq * B
        shape = (L, self.n_pages, t.num_attention_heads, self.PAGE, t.head_dim)
        self.k_cache = torch.zeros(shape, dtype=BF16, device=dev)
        self.v_cache = torch.zeros(shape, dtype=BF16, device=dev)
        self.block_table = torch.arange(self.n_pages, dtype=torch.int32, device=dev).reshape(B, self.pages_per_seq).contiguous()
        self.layer_cache_bytes = self.k_cache[0].numel() * 2
        # decode scratch
        grid = _lib.load().emx_decode_grid()
        # LL exchange buffers of the decode kernel: 8-byte units {payload | tag}, zeroed once (tag 0 is never valid)
        ll = lambda n: torch.zeros(n, dtype=torch.int64, device=dev)  # noqa: E731
        self.d_x, self.d_xo, self.d_attn = ll(H // 2), ll(H // 2), ll(H // 2)
        self.d_qkv, self.d_h = ll(3 * H // 2), ll(I // 2)
        self.d_part = ll(t.num_attention_heads * self.kv_splits * (t.head_dim + 2))
        self.d_argmax = ll(2 * grid)
        self.d_state = torch.zeros(C.sizeof(DecodeState) // 4, dtype=torch.int32, device=dev)
        self.max_new = self.max_context
        self.d_out_tokens = torch.zeros(self.max_new, dtype=torch.int32, device=dev)
        self.h_flag = torch.zeros(4, dtype=torch.int32).pin_memory()
        self._ws: Dict[Tuple[int, int], dict] = {}
        # split-K scratch (emx_gemm_bf16_ws): one per branch that may run concurrently inside the prefill graph (LLM + projector, each tower)
        self._scratch = [torch.zeros(32 << 20, dtype=torch.uint8, device=self.device) for _ in range(1 + len(self.vits))]
        self._slot_tables: Dict[Tuple[int, ...], torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------------------------
    # kernel wrappers
    # ------------------------------------------------------------------------------------------------------------
    @staticmethod
    def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, bias=None, ls=None, resid=None, resid_mod: int = 0, flags: int = 0,
             scratch: Optional[torch.Tensor] = None) -> None:
        """scratch: zero-initialised device buffer of this stream / graph branch: lets the library split K for small-M problems that
        do not fill the machine (emx_gemm_bf16_ws)."""
        M, K = a.shape
        N = w.shape[0]
        call("emx_gemm_bf16_ws", ptr(a), a.stride(0), ptr(w), w.stride(0), ptr(out), out.stride(0), M, N, K, ptr(bias), ptr(ls),
             ptr(resid), resid.stride(0) if resid is not None else 0, resid_mod, flags, ptr(scratch), scratch.numel() if scratch is not None else 0,
             stream())  # fmt: skip

    MAX_CACHED_SHAPES = 8  # distinct (batch, prompt length) workspaces + prefill graphs kept; a new shape costs one eager prefill + a capture

    def _workspace(self, B: int, n_ids: int) -> dict:
        key = (B, n_ids)
        if key in self._ws:
            self._ws[key] = self._ws.pop(key)  # most recently used last
            return self._ws[key]
        while len(self._ws) >= self.MAX_CACHED_SHAPES:
            # least recently used prompt shape: drop its workspace (~45 MB at S = 300 per batch item) and the graphs captured over it
            old = next(iter(self._ws))
            del self._ws[old]
            for gk in [gk for gk in self._graphs if gk[:2] == old]:
                del self._graphs[gk]
        dev, t, cfg = self.device, self.t, self.config
        P = cfg.num_patches
        S = n_ids + P
        ws: dict = {"S": S}
        e = lambda *s: torch.empty(s, dtype=BF16, device=dev)  # noqa: E731
        for i, vw in enumerate(self.vits):
            v = vw.dims
            T = v.num_tokens
            ws[f"v{i}"] = dict(im2col=e(B * P, vw.kpad), pe=e(B * P, v.embed_dim), tok=e(B * T, v.embed_dim), n=e(B * T, v.embed_dim),
                               qkv=e(B * T, 3 * v.embed_dim), att=e(B * T, v.embed_dim), hid=e(B * T, v.mlp_dim))  # fmt: skip
        vd, H, I = cfg.vision_embed_dim, t.hidden_size, t.intermediate_size
        ws.update(feats=e(B * P, vd), p1=e(B * P, 4 * vd), p2=e(B * P, H), patches=e(B * P, H),
                  x=e(B * S, H), n=e(B * S, H), qkv=e(B * S, 3 * H), att=e(B * S, H), h=e(B * S, I),
                  last_n=e(B, H), logits_bf16=e(B, t.vocab_size) if B > 1 else None,
                  pixels=torch.empty((B, 6, cfg.image_sizes[0], cfg.image_sizes[0]), dtype=BF16, device=dev),
                  ids=torch.empty((B, n_ids), dtype=torch.int64, device=dev),
                  logits=torch.empty((B, t.vocab_size), dtype=torch.float32, device=dev),
                  first=torch.empty(B, dtype=torch.int32, device=dev))  # fmt: skip
        self._ws[key] = ws
        return ws

    # ------------------------------------------------------------------------------------------------------------
    def _vision(self, ws: dict, B: int) -> None:
        """pixels [B,6,h,w] -> feats [B*P, vision_dim]  (modeling_prismatic.py:114-123).
        The two towers are independent until their features are concatenated: at small batch neither fills the GPU
        (24-96 CTAs per GEMM), so the second tower runs on a side stream (fork / join with stream waits, which a CUDA-graph
        capture records as parallel branches)."""
        main = torch.cuda.current_stream()
        d0 = self.vits[0].dims.embed_dim
        if len(self.vits) == 2 and B * self.vits[0].dims.num_tokens < 4096:
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(main)  # fork (after the pixel upload on `main`)
            with torch.cuda.stream(self._side):
                self._vision_tower(ws, B, 1, d0)
            self._vision_tower(ws, B, 0, 0)
            main.wait_stream(self._side)  # join
        else:
            col0 = 0
            for i, vw in enumerate(self.vits):
                self._vision_tower(ws, B, i, col0)
                col0 += vw.dims.embed_dim

    def _vision_tower(self, ws: dict, B: int, i: int, col0: int) -> None:
        cfg = self.config
        P, side = cfg.num_patches, cfg.image_sizes[0]
        vw = self.vits[i]
        v, w = vw.dims, ws[f"v{i}"]
        D, T, hd = v.embed_dim, v.num_tokens, v.head_dim
        sc = self._scratch[1 + i]  # this tower's split-K scratch (the towers run concurrently)
        call("emx_patch_im2col", ptr(ws["pixels"]), B, 6, 3 * i, side, side, v.patch_size, ptr(w["im2col"]), vw.kpad, stream())
        self.gemm(w["im2col"], vw.w_pe, w["pe"], bias=vw.b_pe, scratch=sc)
        call("emx_vit_assemble", ptr(w["pe"]), ptr(vw.pos), ptr(vw.prefix), ptr(w["tok"]), B, P, v.num_prefix_tokens, D, stream())
        for blk in vw.blocks:
            call("emx_layernorm", ptr(w["tok"]), ptr(blk["norm1.weight"]), ptr(blk["norm1.bias"]), ptr(w["n"]), B * T, D, v.ln_eps, stream())
            self.gemm(w["n"], blk["attn.qkv.weight"], w["qkv"], bias=blk["attn.qkv.bias"], scratch=sc)
            call("emx_attn_fwd", ptr(w["qkv"]), ptr(w["att"]), B, T, v.num_heads, hd, 0, hd**-0.5, stream())
            self.gemm(w["att"], blk["attn.proj.weight"], w["tok"], bias=blk["attn.proj.bias"], ls=blk.get("ls1.scale_factor"), resid=w["tok"], scratch=sc)
            call("emx_layernorm", ptr(w["tok"]), ptr(blk["norm2.weight"]), ptr(blk["norm2.bias"]), ptr(w["n"]), B * T, D, v.ln_eps, stream())
            self.gemm(w["n"], blk["mlp.fc1.weight"], w["hid"], bias=blk["mlp.fc1.bias"], flags=EPI_GELU, scratch=sc)
            self.gemm(w["hid"], blk["mlp.fc2.weight"], w["tok"], bias=blk["mlp.fc2.bias"], ls=blk.get("ls2.scale_factor"), resid=w["tok"], scratch=sc)
        call("emx_vit_gather_features", ptr(w["tok"]), ptr(ws["feats"]), B, P, v.num_prefix_tokens, D, cfg.vision_embed_dim, col0, stream())

    def _projector(self, ws: dict) -> None:
        """fc1 -> GELU -> fc2 -> GELU -> fc3  (modeling_prismatic.py:152-156)"""
        pj = self.proj
        self.gemm(ws["feats"], pj["fc1.weight"], ws["p1"], bias=pj["fc1.bias"], flags=EPI_GELU, scratch=self._scratch[0])
        self.gemm(ws["p1"], pj["fc2.weight"], ws["p2"], bias=pj["fc2.bias"], flags=EPI_GELU, scratch=self._scratch[0])
        self.gemm(ws["p2"], pj["fc3.weight"], ws["patches"], bias=pj["fc3.bias"], scratch=self._scratch[0])

    def _layer_cache(self, cache: torch.Tensor, layer: int) -> int:
        return cache.data_ptr() + layer * self.layer_cache_bytes

    def _slot_table(self, slot) -> int:
        """Device pointer of the block-table rows a prefill writes through: `slot` = first of B consecutive sequence slots, or an explicit
        tuple of slots (continuous batching refills whichever slots are free): their rows gathered into a small table of their own."""
        if isinstance(slot, int):
            return self.block_table[slot].data_ptr()
        key = tuple(int(b) for b in slot)
        if key not in self._slot_tables:
            if len(self._slot_tables) >= 64:
                self._slot_tables.pop(next(iter(self._slot_tables)))
            self._slot_tables[key] = self.block_table[list(key)].contiguous()
        return self._slot_tables[key].data_ptr()

    def _llm_prefill(self, ws: dict, B: int, n_ids: int, slot=0) -> None:
        """multimodal assembly + full-sequence Llama forward, KV written to the paged cache of sequences slot .. slot + B - 1
        (modeling_prismatic.py:380-415)"""
        t, cfg = self.t, self.config
        H, I, L, S = t.hidden_size, t.intermediate_size, t.num_hidden_layers, ws["S"]
        heads, hd = t.num_attention_heads, t.head_dim
        call("emx_embed_assemble", ptr(ws["ids"]), n_ids, ptr(self.embed), ptr(ws["patches"]), cfg.num_patches, ptr(ws["x"]), B, H, stream())
        tbl = self._slot_table(slot)
        for l in range(L):
            call("emx_rmsnorm", ptr(ws["x"]), ptr(self.ln1[l]), ptr(ws["n"]), B * S, H, t.rms_norm_eps, stream())
            # q|k|v projection with RoPE + paged-KV append fused into the GEMM epilogue (two kernels for shapes the fused epilogue does not cover)
            call("emx_gemm_qkv_rope", ptr(ws["n"]), ws["n"].stride(0), ptr(self.w_qkv[l]), self.w_qkv[l].stride(0), ptr(ws["qkv"]), B, S, heads, hd, H,
                 ptr(self.cos_tab), ptr(self.sin_tab), 0, self._layer_cache(self.k_cache, l), self._layer_cache(self.v_cache, l), tbl,
                 self.pages_per_seq, self.PAGE, stream())  # fmt: skip
            call("emx_attn_fwd", ptr(ws["qkv"]), ptr(ws["att"]), B, S, heads, hd, 1, hd**-0.5, stream())
            self.gemm(ws["att"], self.w_o[l], ws["x"], resid=ws["x"], scratch=self._scratch[0])
            call("emx_rmsnorm", ptr(ws["x"]), ptr(self.ln2[l]), ptr(ws["n"]), B * S, H, t.rms_norm_eps, stream())
            self.gemm(ws["n"], self.w_gateup[l], ws["h"], flags=EPI_SWIGLU, scratch=self._scratch[0])
            self.gemm(ws["h"], self.w_down[l], ws["x"], resid=ws["x"], scratch=self._scratch[0])
        # lm_head on the LAST position only (the reference computes all S rows and discards S-1 of them)
        x3 = ws["x"].view(B, S, H)
        for b in range(B):
            call("emx_rmsnorm", ptr(x3[b, S - 1]), ptr(self.final_norm), ptr(ws["last_n"][b]), 1, H, t.rms_norm_eps, stream())
            if B == 1:
                call("emx_lmhead_argmax", ptr(self.lm_head), H, ptr(ws["last_n"][b]), t.vocab_size, H, ptr(ws["logits"][b]),
                     ptr(ws["first"][b : b + 1]), None, stream())  # fmt: skip
        if B > 1:  # ONE pass over lm_head (262 MB) for the B last rows instead of B GEMVs: [B, H] x lm_head^T on the tensor cores, row-wise argmax
            self.gemm(ws["last_n"], self.lm_head, ws["logits_bf16"], scratch=self._scratch[0])
            call("emx_argmax_rows_bf16", ptr(ws["logits_bf16"]), t.vocab_size, B, t.vocab_size, ptr(ws["logits"]), ptr(ws["first"]), stream())

    def _prefill_body(self, ws: dict, B: int, n_ids: int, slot=0) -> None:
        self._vision(ws, B)
        self._projector(ws)
        self._llm_prefill(ws, B, n_ids, slot)

    # ------------------------------------------------------------------------------------------------------------
    @_on_engine_device
    @torch.no_grad()
    def prefill(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, use_graph: bool = True, slot=0) -> dict:
        """Vision + projector + LLM prefill for B same-length prompts, into the KV pages of sequences slot .. slot + B - 1 (`slot` an int)
        or of the B sequence slots listed in `slot` (a tuple). Returns the workspace (first tokens, logits, features)."""
        B, n_ids = input_ids.shape
        if isinstance(slot, int):
            if slot < 0 or slot + B > self.max_batch:
                raise ValueError(f"batch {B} at slot {slot} exceeds engine capacity {self.max_batch}")
        else:
            slot = tuple(int(b) for b in slot)
            if len(slot) != B or len(set(slot)) != B or min(slot) < 0 or max(slot) >= self.max_batch:
                raise ValueError(f"slots {slot}: need {B} distinct slots below the engine capacity {self.max_batch}")
        S = n_ids + self.config.num_patches
        if S + 1 > self.max_context:
            raise ValueError(f"prompt of {S} positions does not fit max_context={self.max_context}")
        ws = self._workspace(B, n_ids)
        ws["ids"].copy_(input_ids, non_blocking=True)
        ws["pixels"].copy_(pixel_values, non_blocking=True)
        if not use_graph:
            self._prefill_body(ws, B, n_ids, slot)
            return ws
        key = (B, n_ids, slot)
        if key not in self._graphs:
            # warm-up once outside capture (sets func attributes / driver entry points), then capture
            self._prefill_body(ws, B, n_ids, slot)
            torch.cuda.current_stream().synchronize()
            before = _lib.launch_count
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._prefill_body(ws, B, n_ids, slot)
            while len(self._graphs) >= 4 * self.MAX_CACHED_SHAPES:  # (several slot placements per prompt shape under continuous batching)
                del self._graphs[next(iter(self._graphs))]
            self._graphs[key] = (g, _lib.launch_count - before)
            _lib.count_launches(before - _lib.launch_count)  # capture enqueued nothing
        g, n_kernels = self._graphs[key]
        g.replay()
        _lib.count_launches(n_kernels)
        return ws

    def _decode_params(self, b: int = 0) -> DecodeParams:
        t = self.t
        p = DecodeParams()
        p.hidden, p.inter, p.heads, p.head_dim = t.hidden_size, t.intermediate_size, t.num_attention_heads, t.head_dim
        p.layers, p.vocab, p.rms_eps = t.num_hidden_layers, t.vocab_size, t.rms_norm_eps
        p.embed, p.w_qkv, p.w_o, p.w_gateup, p.w_down = ptr(self.embed), ptr(self.w_qkv), ptr(self.w_o), ptr(self.w_gateup), ptr(self.w_down)
        p.ln1, p.ln2, p.final_norm, p.lm_head = ptr(self.ln1), ptr(self.ln2), ptr(self.final_norm), ptr(self.lm_head)
        p.cos_tab, p.sin_tab = ptr(self.cos_tab), ptr(self.sin_tab)
        p.k_cache, p.v_cache = ptr(self.k_cache), ptr(self.v_cache)
        p.block_table = self.block_table[b].data_ptr()
        p.page_size, p.n_pages, p.max_pages = self.PAGE, self.n_pages, self.pages_per_seq
        p.x, p.xo, p.qkv, p.attn, p.h = ptr(self.d_x), ptr(self.d_xo), ptr(self.d_qkv), ptr(self.d_attn), ptr(self.d_h)
        p.part, p.argmax_part = ptr(self.d_part), ptr(self.d_argmax)
        p.out_tokens, p.logits_out = ptr(self.d_out_tokens), None
        p.eos_token, p.kv_splits, p.state = -1, self.kv_splits, ptr(self.d_state)
        p.dbg, p.l2_lookahead_kb, p.debug_flags = None, self.l2_lookahead_kb, 0  # profiling twins: tools/decode_probe.py sets these itself
        return p

    @_on_engine_device
    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, pixel_values: torch.Tensor, max_new_tokens: int, eos_token_id: Optional[int] = 2,
                 return_logits: bool = False, forced: Optional[List[int]] = None, use_graph: bool = True,
                 poll_every: int = 32) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:  # fmt: skip
        """Greedy decode for ONE sequence (the reference asserts bs == 1 for cached generation,
        modeling_prismatic.py:326, :460-463). Returns (new token ids [n] int32 on device, optional per-step logits)."""
        if input_ids.shape[0] != 1:
            raise ValueError("Generation with batch size > 1 is not currently supported!")
        n_ids = input_ids.shape[1]
        S = n_ids + self.config.num_patches
        if S + max_new_tokens > self.max_context:
            raise ValueError(f"{S} prompt positions + {max_new_tokens} new tokens exceed max_context={self.max_context}")
        ws = self.prefill(input_ids, pixel_values, use_graph=use_graph)
        V = self.t.vocab_size
        logits = torch.empty((max_new_tokens, V), dtype=torch.float32, device=self.device) if return_logits else None
        if return_logits:
            logits[0].copy_(ws["logits"][0])
        # hand-off: first token, position, counters (kernel-private words 4.. of the state are left alone)
        first = ws["first"][0:1]
        self.d_out_tokens[0:1].copy_(first)
        st = self.d_state
        st[0:1].copy_(first if forced is None else torch.tensor([forced[0]], dtype=torch.int32, device=self.device))
        st[1].fill_(S)
        st[2].fill_(1)
        st[3].zero_()
        if eos_token_id is not None and forced is None:
            st[3:4].copy_((first == eos_token_id).to(torch.int32))
        p = self._decode_params(0)
        p.eos_token = -1 if (eos_token_id is None or forced is not None) else int(eos_token_id)
        lib = _lib.load()
        s = stream()
        pending = None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        n_launched = 0
        for step in range(1, max_new_tokens):
            if return_logits:
                p.logits_out = logits[step].data_ptr()
            _lib.check(lib.emx_decode_step(C.byref(p), s))
            n_launched += 1
            if forced is not None:
                st[0].fill_(int(forced[step]))  # teacher forcing: overwrite the token the next step will embed
            if p.eos_token >= 0 and step % poll_every == 0:
                # lagging, non-blocking EOS poll: look at the flag copied one chunk ago, then queue a fresh copy
                if pending is not None and pending.query() and int(self.h_flag[3]) != 0:
                    break
                self.h_flag.copy_(st[0:4], non_blocking=True)
                pending = torch.cuda.Event()
                pending.record()
        ev1.record()
        _lib.count_launches(n_launched)
        self.last_decode = (ev0, ev1, n_launched, S)  # bench.py: average decode-step duration on this stream
        n = int(st[2].item())  # device -> host sync point (the reference syncs every token)
        return self.d_out_tokens[:n].clone(), (logits[:n] if return_logits else None)

    # ------------------------------------------------------------------------------------------------------------
    # batched decode: up to 8 sequences share one pass over the weights per token (emx_decode_batch_step)
    # ------------------------------------------------------------------------------------------------------------
    def _alloc_batch(self) -> None:
        if getattr(self, "b_state", None) is not None:
            return
        dev, t = self.device, self.t
        H, I = t.hidden_size, t.intermediate_size
        R = lambda n: _ceil_to(n, 16)  # noqa: E731
        ll = lambda n: torch.zeros(n, dtype=torch.int64, device=dev)  # noqa: E731
        MB = MAX_DECODE_BATCH
        grid = _lib.load().emx_decode_grid()
        self.b_x, self.b_xo, self.b_attn = ll(MB * R(H // 2)), ll(MB * R(H // 2)), ll(MB * R(H // 2))
        self.b_qkv, self.b_h = ll(MB * 3 * H // 2), ll(MB * R(I // 2))
        self.b_part = ll(MB * t.num_attention_heads * ATT_MAX_SEGMENTS * (t.head_dim + 2))
        self.b_argmax = ll(grid * MB * 2)
        self.b_sync = ll(2)
        self.b_state = torch.zeros(C.sizeof(DecodeBatchState) // 4, dtype=torch.int32, device=dev)
        self.b_out = torch.zeros((MB, self.max_new), dtype=torch.int32, device=dev)
        self.h_bflag = torch.zeros(C.sizeof(DecodeBatchState) // 4, dtype=torch.int32).pin_memory()

    @_on_engine_device
    @torch.no_grad()
    def serve(self, requests, eos_token_id: Optional[int] = 2, use_graph: bool = True, poll_every: int = 32,
              order: str = "fifo") -> List[torch.Tensor]:
        """Continuous batching over the sequence slots of the batched decode kernel: a stream of independent requests
        `(input_ids [1, n], pixel_values [1, 6, h, w], max_new_tokens)` is decoded min(8, max_batch) at a time, and a slot whose sequence
        has reached its token limit (or EOS) is refilled with the next request at once — prefilled straight into that slot's KV pages
        (same-length prompts admitted together share one batched prefill) — instead of idling until the longest sequence of its batch is
        done. With BASELINE.json configs[4]'s mix of 128- and 512-token requests a static batch of 8 runs half empty for 3/4 of its
        launches; here every launch carries 8 live sequences. Per-sequence results are those of `generate` (same kernels, same state
        machine: position, limit and EOS are per slot). Returns the generated ids per request, in request order.
        `order`: "fifo" admits requests in arrival order; "longest_first" (an offline batch whose token limits are known up front) admits
        the requests with the largest limits first, so that the tail of the stream is made of short requests and the slots drain
        together instead of the last long request decoding alone (see `admission_order`)."""
        MB = MAX_DECODE_BATCH
        S = min(MB, self.max_batch)
        P = self.config.num_patches
        reqs = []
        for i, (ids, pv, lim) in enumerate(requests):
            lim = int(lim)
            if ids.dim() != 2 or ids.shape[0] != 1 or pv.shape[0] != 1:
                raise ValueError(f"request {i}: input_ids [1, n] and pixel_values [1, ...] expected")
            if lim < 1 or lim > self.max_new or ids.shape[1] + P + lim > self.max_context:
                raise ValueError(f"request {i}: {ids.shape[1] + P} prompt positions + {lim} new tokens exceed max_context={self.max_context}")
            reqs.append((ids, pv, lim))
        queue = admission_order([r[2] for r in reqs], order)
        if self.pages_per_seq > 2 * ATT_MAX_SEGMENTS:
            raise ValueError(f"batched decode supports contexts up to {2 * ATT_MAX_SEGMENTS * self.PAGE} (max_context={self.max_context})")
        self._alloc_batch()
        dev = self.device
        st = self.b_state.view(-1)
        st[: 5 * MB].zero_()
        st[3 * MB : 4 * MB].fill_(1)  # every slot starts free = finished
        use_eos = eos_token_id is not None
        p = self._decode_batch_params(S)
        p.eos_token = int(eos_token_id) if use_eos else -1
        lib, s = _lib.load(), stream()
        slot_req, slot_left, slot_pos = [-1] * S, [0] * S, [0] * S
        results: List[Optional[torch.Tensor]] = [None] * len(reqs)
        nxt = done = n_launched = 0
        kv_reads = kv_writes = 0  # cached positions read / appended over all launches (host mirror; exact when no sequence stops at EOS)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        dec_ms_events = []

        def retire(b: int, n_tokens: int) -> None:
            nonlocal done
            results[slot_req[b]] = self.b_out[b, :n_tokens].clone()
            slot_req[b], slot_left[b] = -1, 0
            done += 1

        while done < len(reqs):
            # ---- admit: next requests into the free slots, same-length prompts in one batched prefill
            free = [b for b in range(S) if slot_req[b] < 0]
            while free and nxt < len(reqs):
                n_ids = reqs[queue[nxt]][0].shape[1]
                group = [queue[nxt]]
                while len(group) < len(free) and nxt + len(group) < len(reqs) and reqs[queue[nxt + len(group)]][0].shape[1] == n_ids:
                    group.append(queue[nxt + len(group)])
                slots = tuple(free[: len(group)])
                free = free[len(group) :]
                ids = torch.cat([reqs[r][0] for r in group]).to(dev)
                pv = torch.cat([reqs[r][1] for r in group]).to(dev, BF16)
                ws = self.prefill(ids, pv, use_graph=use_graph, slot=slots if len(slots) > 1 else slots[0])
                first = ws["first"][: len(group)]
                idx = torch.tensor(slots, dtype=torch.int64, device=dev)
                lims = torch.tensor([reqs[r][2] for r in group], dtype=torch.int32, device=dev)
                st.index_copy_(0, idx, first)
                st.index_copy_(0, idx + MB, torch.full((len(group),), n_ids + P, dtype=torch.int32, device=dev))
                st.index_copy_(0, idx + 2 * MB, torch.ones(len(group), dtype=torch.int32, device=dev))
                st.index_copy_(0, idx + 3 * MB, (first == eos_token_id).to(torch.int32) if use_eos else torch.zeros(len(group), dtype=torch.int32, device=dev))
                st.index_copy_(0, idx + 4 * MB, lims)
                self.b_out[:, 0].index_copy_(0, idx, first)
                for b, r in zip(slots, group):
                    slot_req[b], slot_left[b], slot_pos[b] = r, reqs[r][2] - 1, n_ids + P
                nxt += len(group)
            for b in range(S):  # a limit of 1 is complete after its prefill
                if slot_req[b] >= 0 and slot_left[b] == 0 and not use_eos:
                    retire(b, reqs[slot_req[b]][2])
            live = [b for b in range(S) if slot_req[b] >= 0]
            if not live:
                continue
            # ---- decode until the next slot frees up (or the next EOS poll)
            steps = min(slot_left[b] for b in live if slot_left[b] > 0) if any(slot_left[b] > 0 for b in live) else 0
            if use_eos:
                steps = min(steps, poll_every) if steps else 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                _lib.check(lib.emx_decode_batch_step(C.byref(p), s))
                for b in live:
                    if slot_left[b] > 0:
                        kv_reads += slot_pos[b]
                        kv_writes += 1
                        slot_pos[b] += 1
                        slot_left[b] -= 1
            e1.record()
            dec_ms_events.append((e0, e1))
            n_launched += steps
            if use_eos:
                hf = st[: 5 * MB].tolist()  # blocking read of the slot state once per poll interval
                for b in live:
                    if hf[3 * MB + b] != 0 or hf[2 * MB + b] >= reqs[slot_req[b]][2]:
                        retire(b, hf[2 * MB + b])
            else:
                for b in live:
                    if slot_left[b] == 0:
                        retire(b, reqs[slot_req[b]][2])
        ev1.record()
        _lib.count_launches(n_launched)
        self.last_serve = dict(events=(ev0, ev1), decode_events=dec_ms_events, launches=n_launched, kv_reads=kv_reads, kv_writes=kv_writes)
        return results  # type: ignore[return-value]

    def _decode_batch_params(self, B: int) -> DecodeBatchParams:
        t = self.t
        p = DecodeBatchParams()
        p.hidden, p.inter, p.heads, p.head_dim = t.hidden_size, t.intermediate_size, t.num_attention_heads, t.head_dim
        p.layers, p.vocab, p.rms_eps, p.batch = t.num_hidden_layers, t.vocab_size, t.rms_norm_eps, B
        p.embed, p.w_qkv, p.w_o, p.w_gateup, p.w_down = ptr(self.embed), ptr(self.w_qkv), ptr(self.w_o), ptr(self.w_gateup), ptr(self.w_down)
        p.ln1, p.ln2, p.final_norm, p.lm_head = ptr(self.ln1), ptr(self.ln2), ptr(self.final_norm), ptr(self.lm_head)
        p.cos_tab, p.sin_tab = ptr(self.cos_tab), ptr(self.sin_tab)
        p.k_cache, p.v_cache, p.block_table = ptr(self.k_cache), ptr(self.v_cache), ptr(self.block_table)
        p.page_size, p.n_pages, p.max_pages, p.out_stride = self.PAGE, self.n_pages, self.pages_per_seq, self.b_out.shape[1]
        p.x, p.xo, p.attn, p.qkv, p.h = ptr(self.b_x), ptr(self.b_xo), ptr(self.b_attn), ptr(self.b_qkv), ptr(self.b_h)
        p.part, p.argmax_part, p.sync = ptr(self.b_part), ptr(self.b_argmax), ptr(self.b_sync)
        p.out_tokens, p.logits_out, p.state, p.dbg = ptr(self.b_out), None, ptr(self.b_state), None
        p.eos_token = -1
        p.l2_lookahead_stages = int(os.environ.get("EMX_BATCH_L2_LOOKAHEAD_STAGES", "6"))
        return p

    @_on_engine_device
    @torch.no_grad()
    def generate_batch(self, input_ids, pixel_values: torch.Tensor, max_new_tokens, eos_token_id: Optional[int] = 2,
                       return_logits: bool = False, forced: Optional[List[List[int]]] = None, use_graph: bool = True,
                       poll_every: int = 32) -> Tuple[List[torch.Tensor], Optional[torch.Tensor]]:  # fmt: skip
        """Greedy decode of B <= 8 sequences at once — what the reference refuses to do (bs == 1 asserts, modeling_prismatic.py:326,
        :460-463). `input_ids`: [B, n_ids] tensor (one batched prefill) or a list of B [1, n_i] tensors of different lengths (prefilled one
        by one into their KV slots); `max_new_tokens`: int or one int per sequence (BASELINE.json configs[4]: mixed 128 / 512).
        Returns (list of B int32 id tensors, optional logits [steps, 8, vocab])."""
        same_len = isinstance(input_ids, torch.Tensor)
        B = input_ids.shape[0] if same_len else len(input_ids)
        if not 1 <= B <= min(MAX_DECODE_BATCH, self.max_batch):
            raise ValueError(f"batch {B} not in 1..{min(MAX_DECODE_BATCH, self.max_batch)} (engine max_batch={self.max_batch})")
        if self.pages_per_seq > 2 * ATT_MAX_SEGMENTS:
            raise ValueError(f"batched decode supports contexts up to {2 * ATT_MAX_SEGMENTS * self.PAGE} (max_context={self.max_context})")
        limits = [int(max_new_tokens)] * B if isinstance(max_new_tokens, int) else [int(x) for x in max_new_tokens]
        if len(limits) != B or min(limits) < 1:
            raise ValueError("max_new_tokens: one positive int, or one per sequence")
        P = self.config.num_patches
        lens = [input_ids.shape[1]] * B if same_len else [int(x.shape[1]) for x in input_ids]
        for b in range(B):
            if lens[b] + P + limits[b] > self.max_context:
                raise ValueError(f"sequence {b}: {lens[b] + P} prompt positions + {limits[b]} new tokens exceed max_context={self.max_context}")
        self._alloc_batch()
        V, dev = self.t.vocab_size, self.device
        T = max(limits)
        MB = MAX_DECODE_BATCH
        logits = torch.zeros((T, MB, V), dtype=torch.float32, device=dev) if return_logits else None
        first = torch.empty(B, dtype=torch.int32, device=dev)
        if same_len:
            ws = self.prefill(input_ids, pixel_values, use_graph=use_graph)
            first.copy_(ws["first"][:B])
            if return_logits:
                logits[0, :B].copy_(ws["logits"][:B])
        else:
            for b in range(B):
                ws = self.prefill(input_ids[b], pixel_values[b : b + 1], use_graph=use_graph, slot=b)
                first[b : b + 1].copy_(ws["first"][0:1])
                if return_logits:
                    logits[0, b].copy_(ws["logits"][0])
        st = self.b_state.view(-1)
        st[: 5 * MB].zero_()  # (the kernel-private epoch word stays)
        cur = first if forced is None else torch.tensor([f[0] for f in forced], dtype=torch.int32, device=dev)
        st[0:B].copy_(cur)
        st[MB : MB + B].copy_(torch.tensor([n + P for n in lens], dtype=torch.int32, device=dev))
        st[2 * MB : 2 * MB + B].fill_(1)
        use_eos = eos_token_id is not None and forced is None
        if use_eos:
            st[3 * MB : 3 * MB + B].copy_((first == eos_token_id).to(torch.int32))
        st[4 * MB : 4 * MB + B].copy_(torch.tensor(limits, dtype=torch.int32, device=dev))
        self.b_out[:B, 0].copy_(first)
        p = self._decode_batch_params(B)
        p.eos_token = int(eos_token_id) if use_eos else -1
        lib = _lib.load()
        s = stream()
        pending = None
        forced_t = None
        if forced is not None:  # teacher forcing: row `step` of [T, B] overwrites the tokens the next launch embeds
            forced_t = torch.zeros((T, B), dtype=torch.int32, device=dev)
            for b, f in enumerate(forced):
                forced_t[: len(f), b] = torch.tensor(f, dtype=torch.int32, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        n_launched = 0
        for step in range(1, T):
            if return_logits:
                p.logits_out = logits[step].data_ptr()
            _lib.check(lib.emx_decode_batch_step(C.byref(p), s))
            n_launched += 1
            if forced_t is not None:
                st[0:B].copy_(forced_t[step])
            if use_eos and step % poll_every == 0:
                # lagging, non-blocking poll: stop once every sequence is finished or at its limit (state copied one chunk ago)
                if pending is not None and pending.query():
                    hf = self.h_bflag
                    if all(int(hf[3 * MB + b]) != 0 or int(hf[2 * MB + b]) >= limits[b] for b in range(B)):
                        break
                self.h_bflag.copy_(st, non_blocking=True)
                pending = torch.cuda.Event()
                pending.record()
        ev1.record()
        _lib.count_launches(n_launched)
        self.last_decode_batch = (ev0, ev1, n_launched, [n + P for n in lens], limits)
        n_gen = st[2 * MB : 2 * MB + B].tolist()  # device -> host sync point
        return [self.b_out[b, : n_gen[b]].clone() for b in range(B)], logits
