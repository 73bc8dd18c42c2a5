"""
ctypes binding of libemmax.so (the C ABI in include/emmax.h).

The product path has NO fallback: if the shared library is missing or was not built for sm_100a, importing the engine
raises. PyTorch tensors stop here — only `data_ptr()`s, sizes and the current stream handle cross this line.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Any, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EMX_LIB", os.path.join(_HERE, "libemmax.so"))  # (EMX_LIB: A/B builds of the kernels in tools/ probes)

EPI_GELU = 1
EPI_SWIGLU = 2


class EmxError(RuntimeError):
    pass


class DecodeState(C.Structure):
    _fields_ = [
        ("cur_token", C.c_int32), ("pos", C.c_int32), ("n_generated", C.c_int32), ("finished", C.c_int32),
        ("barrier", C.c_uint32), ("epoch", C.c_uint32), ("pad0_", C.c_uint32 * 2), ("head_ticket", C.c_uint32 * 64),
    ]  # fmt: skip


class DecodeParams(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("inter", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("layers", C.c_int32), ("vocab", C.c_int32), ("rms_eps", C.c_float),
        ("embed", C.c_void_p), ("w_qkv", C.c_void_p), ("w_o", C.c_void_p), ("w_gateup", C.c_void_p),
        ("w_down", C.c_void_p), ("ln1", C.c_void_p), ("ln2", C.c_void_p), ("final_norm", C.c_void_p),
        ("lm_head", C.c_void_p), ("cos_tab", C.c_void_p), ("sin_tab", C.c_void_p),
        ("k_cache", C.c_void_p), ("v_cache", C.c_void_p), ("block_table", C.c_void_p),
        ("page_size", C.c_int32), ("n_pages", C.c_int32), ("max_pages", C.c_int32),
        ("x", C.c_void_p), ("xo", C.c_void_p), ("qkv", C.c_void_p), ("attn", C.c_void_p), ("h", C.c_void_p),
        ("part", C.c_void_p), ("argmax_part", C.c_void_p),
        ("out_tokens", C.c_void_p), ("logits_out", C.c_void_p),
        ("eos_token", C.c_int32), ("kv_splits", C.c_int32), ("state", C.c_void_p), ("dbg", C.c_void_p),
        ("l2_lookahead_kb", C.c_int32), ("debug_flags", C.c_int32),
    ]  # fmt: skip


MAX_DECODE_BATCH = 8
ATT_MAX_SEGMENTS = 16


class DecodeBatchState(C.Structure):
    _fields_ = [
        ("cur_token", C.c_int32 * 8), ("pos", C.c_int32 * 8), ("n_generated", C.c_int32 * 8), ("finished", C.c_int32 * 8),
        ("limit", C.c_int32 * 8), ("epoch", C.c_uint32), ("pad_", C.c_uint32 * 7),
    ]  # fmt: skip


class DecodeBatchParams(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("inter", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("layers", C.c_int32), ("vocab", C.c_int32), ("rms_eps", C.c_float), ("batch", C.c_int32),
        ("embed", C.c_void_p), ("w_qkv", C.c_void_p), ("w_o", C.c_void_p), ("w_gateup", C.c_void_p),
        ("w_down", C.c_void_p), ("ln1", C.c_void_p), ("ln2", C.c_void_p), ("final_norm", C.c_void_p),
        ("lm_head", C.c_void_p), ("cos_tab", C.c_void_p), ("sin_tab", C.c_void_p),
        ("k_cache", C.c_void_p), ("v_cache", C.c_void_p), ("block_table", C.c_void_p),
        ("page_size", C.c_int32), ("n_pages", C.c_int32), ("max_pages", C.c_int32), ("out_stride", C.c_int32),
        ("x", C.c_void_p), ("xo", C.c_void_p), ("attn", C.c_void_p), ("qkv", C.c_void_p), ("h", C.c_void_p),
        ("part", C.c_void_p), ("argmax_part", C.c_void_p), ("sync", C.c_void_p),
        ("out_tokens", C.c_void_p), ("logits_out", C.c_void_p), ("state", C.c_void_p), ("dbg", C.c_void_p),
        ("eos_token", C.c_int32), ("l2_lookahead_stages", C.c_int32),
    ]  # fmt: skip


_P, _I, _F, _L = C.c_void_p, C.c_int, C.c_float, C.c_long
_SIGNATURES = {
    "emx_last_error": (C.c_char_p, []),
    "emx_abi_version": (_I, []),
    "emx_arch": (C.c_char_p, []),
    "emx_gemm_bf16": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _P]),
    "emx_gemm_bf16_ws": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _P, _P, _P, _I, _I, _I, _P, _L, _P]),
    "emx_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "emx_rmsnorm": (_I, [_P, _P, _P, _I, _I, _F, _P]),
    "emx_preprocess_u8": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "emx_resize_preprocess_u8": (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "emx_crop_resize_u8": (_I, [_P, _I, _I, _I, _F, _F, _F, _F, _P, _I, _I, _P]),
    "emx_lanczos3_resize_u8": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P]),
    "emx_patch_im2col": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P]),
    "emx_vit_assemble": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "emx_vit_gather_features": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "emx_attn_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _F, _P]),
    "emx_rope_kvstore": (_I, [_P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _I, _I, _P]),
    "emx_gemm_qkv_rope": (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _I, _I, _P]),
    "emx_embed_assemble": (_I, [_P, _I, _P, _P, _I, _P, _I, _I, _P]),
    "emx_swiglu": (_I, [_P, _P, _I, _I, _P]),
    "emx_gemv_bf16": (_I, [_P, _I, _P, _P, _P, _I, _I, _P]),
    "emx_lmhead_argmax": (_I, [_P, _I, _P, _I, _I, _P, _P, _P, _P]),
    "emx_argmax_rows_bf16": (_I, [_P, _I, _I, _I, _P, _P, _P]),
    "emx_decode_step": (_I, [C.POINTER(DecodeParams), _P]),
    "emx_decode_batch_step": (_I, [C.POINTER(DecodeBatchParams), _P]),
    "emx_decode_batch_smem": (_I, []),
    "emx_decode_grid": (_I, []),
    "emx_decode_phase_rows": (_I, [_I, _I, _I, _I, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "emx_detokenize_actions": (_I, [_P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P]),
    "emx_debug_stream": (_I, [_P, C.c_long, _I, _I, C.c_long, _I, _I, _I, _I, _P]),
    "emx_debug_skeleton": (_I, [_P, C.c_long, _I, _P, _P, _I, _I, _I, C.c_long, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """dlopen libemmax.so and set prototypes. Raises EmxError when it is missing — there is no CPU/torch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EmxError(
            f"{LIB_PATH} not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C emmax_b200/csrc`); emmax_b200 has no fallback path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.emx_arch() != b"sm_100a":
        raise EmxError("libemmax.so was not built for sm_100a")
    _lib = lib
    return lib


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def check(rc: int) -> None:
    if rc != 0:
        raise EmxError(load().emx_last_error().decode() or f"libemmax error {rc}")


# kernels launched per C-ABI call (bookkeeping for bench.py's `gpu_launches`; graph replays add their captured count)
_KERNELS_PER_CALL = {"emx_lmhead_argmax": 2, "emx_lanczos3_resize_u8": 2, "emx_resize_preprocess_u8": 2}
launch_count = 0


def count_launches(n: int) -> None:
    global launch_count
    launch_count += n


def call(name: str, *args: Any) -> None:
    check(getattr(load(), name)(*args))
    count_launches(_KERNELS_PER_CALL.get(name, 1))
