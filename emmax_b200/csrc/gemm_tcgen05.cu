// emx_gemm_bf16 — C[M,N] = epilogue(A[M,K] · W[N,K]^T): the ViT / projector / Llama-prefill GEMM.
//
// Replaces (reference call sites; the arithmetic itself runs in cuBLASLt there):
//   timm ViT qkv / proj / fc1 / fc2 Linears     /root/reference/prismatic/extern/hf/modeling_prismatic.py:121
//   PrismaticProjector fc1/fc2/fc3              /root/reference/prismatic/extern/hf/modeling_prismatic.py:152-156
//   Llama q/k/v/o/gate/up/down at prefill       /root/reference/prismatic/extern/hf/modeling_prismatic.py:404-415
//
// B200 design: both operands are K-major bf16, so A and W tiles go HBM -> smem by TMA (128-byte swizzle, 64-element
// K slabs), are consumed straight from smem by tcgen05.mma (one elected thread, M=128 x N=BN x K=16 atoms) with the
// fp32 accumulator in TMEM, and the epilogue warps pull the tile out of TMEM with tcgen05.ld and apply the fused
// bias / GELU / LayerScale / residual / SwiGLU chain with the SAME bf16 rounding points as the unfused torch-eager
// reference (each reference op output is bf16).
//
// Persistent: one CTA per SM walks the output tiles (grouped order, see tile_coords: neighbouring CTAs share W and A blocks in L2); the fp32
// accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns), so the epilogue of tile i (TMEM read-out, fused math, global stores)
// overlaps the TMA / MMA main loop of tile i + 1 — the short-K ViT GEMMs (16-18 k-blocks per tile) were epilogue-bound without it.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..11 = epilogue (warp w reads
// TMEM lane quarter w mod 4; warps 4..7 take the left half of the tile's columns, 8..11 the right half: the GELU (erff) epilogue of the
// short-K ViT / projector tiles is longer than their main loop with four warps).
#include <stdlib.h>

#include "common.cuh"
#include "emmax.h"

namespace emx {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 8;                     // two per TMEM lane quarter, each covering half of the tile's columns
constexpr int kGemmThreads = (4 + kEpiWarps) * 32;  // warps 0..3: TMA producer, MMA issuer, TMEM allocator, idle

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 64) ? 8 : 6;  // ~192 KB of operands in flight either way
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Tile order: groups of kGroupM consecutive M tiles; inside a group the M index runs fastest, then N. The CTAs that run concurrently
// (one wave = 148 tiles or 74 pair tiles) then share ~8 A row blocks and ~9-18 W row blocks (a few tens of MB: L2-resident) instead of
// ALL of A (77 MB at M = 9472, K = 4096) against two W blocks, which re-streamed A from HBM once per pair of N tiles.
#ifndef EMX_GROUP_M
#define EMX_GROUP_M 8
#endif
constexpr int kGroupM = EMX_GROUP_M;
__device__ __forceinline__ void tile_coords(int tile, int mt, int nt, int& mi, int& ni) {
  const int per_group = kGroupM * nt;
  const int g = tile / per_group, r = tile - g * per_group;
  const int m_first = g * kGroupM;
  const int gm = min(kGroupM, mt - m_first);
  ni = r / gm;
  mi = m_first + (r - ni * gm);
}

struct EpiParams {
  const __nv_bfloat16* bias;   // [N] or null
  const __nv_bfloat16* ls;     // LayerScale [N] or null
  const __nv_bfloat16* resid;  // [M, ldr] or null
  int ldr;
  int resid_mod;  // >0: residual row = m % resid_mod (broadcast over batch: position embeddings)
  int flags;
};

struct EpiCtx {
  bool has_bias, has_ls, has_res, do_gelu, do_swiglu;
  __device__ explicit EpiCtx(const EpiParams& ep)
      : has_bias(ep.bias != nullptr), has_ls(ep.ls != nullptr), has_res(ep.resid != nullptr), do_gelu(ep.flags & EMX_EPI_GELU),
        do_swiglu(ep.flags & EMX_EPI_SWIGLU) {}
};

// 32 consecutive bf16 (a slice of bias / LayerScale, or of one residual row) as 4 x 16-byte loads when the slice is whole and
// 16-byte aligned, element by element at the N tail or for odd leading dimensions. Packed: w[i] = {elem 2i, elem 2i+1}.
__device__ __forceinline__ void ld32_bf16(const __nv_bfloat16* p, int valid, uint32_t (&w)[16]) {
  if (valid >= 32 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 u = reinterpret_cast<const uint4*>(p)[j];
      w[4 * j + 0] = u.x, w[4 * j + 1] = u.y, w[4 * j + 2] = u.z, w[4 * j + 3] = u.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t lo = 2 * j < valid ? static_cast<uint32_t>(__bfloat16_as_ushort(p[2 * j])) : 0u;
      const uint32_t hi = 2 * j + 1 < valid ? static_cast<uint32_t>(__bfloat16_as_ushort(p[2 * j + 1])) : 0u;
      w[j] = lo | (hi << 16);
    }
  }
}
__device__ __forceinline__ float bf16_of(const uint32_t (&w)[16], int j) { return (j & 1) ? bf16_hi(w[j >> 1]) : bf16_lo(w[j >> 1]); }

// tcgen05.wait::ld that also names the destination registers, so that no use of them can be scheduled in front of the wait
__device__ __forceinline__ void tmem_ld_wait_on(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])::"memory");
}

// One epilogue warp's share of an accumulator tile: lane i owns output row m (TMEM lane base + i), NC fp32 columns starting at TMEM
// address `taddr`, output columns n0 .. n0 + NC. Fused chain with the eager rounding points: bf16(acc + bias) -> bf16(GELU) ->
// bf16(* LayerScale) -> bf16(+ residual) | SwiGLU over interleaved (gate, up) column pairs.
// The residual slice of the row does not depend on the accumulator: EpiResid::load fetches ALL of it (16-byte loads) before the warp
// waits for the tile's MMAs, so the L2 round trips of the residual hide behind the main loop instead of being paid once per 32-column
// chunk (the ViT proj / fc2 and Llama o / down tiles were bound by exactly that chain). bias / LayerScale slices (same for every row,
// L1 hits) are fetched per chunk, in front of the TMEM wait.
// Epilogue specialisations (picked on the host from the flags; each kernel instantiation carries only its own chain):
//   EPI_ANY   : every combination, 32-column chunks in a rolled loop (bias / GELU / LayerScale / residual / SwiGLU tested at run time)
//   EPI_RESID : bias -> LayerScale -> residual (no GELU, no SwiGLU), chunk loop unrolled over the prefetched residual
//   EPI_SWIGLU: SwiGLU only
//   EPI_ROPE  : the Llama q|k|v projection of a prefill: RoPE on the q and k heads, K / V rows appended to the paged cache (see epilogue_rope_head)
enum EpiKind { EPI_ANY = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_ROPE = 3 };

// What the fused q|k|v epilogue needs beyond the GEMM itself (emx_gemm_qkv_rope; the un-fused twin is rope_kvstore_kernel in ops.cu).
struct RopeParams {
  const __nv_bfloat16* cos_tab;  // [max_pos][64] bf16
  const __nv_bfloat16* sin_tab;
  __nv_bfloat16* k_cache;        // this layer's [page][head][page_size][128]
  __nv_bfloat16* v_cache;
  const int32_t* block_table;    // [B][max_pages]
  int T, pos0, heads, max_pages, page_size;
};

template <int NC, int EPI>
struct EpiResid {
  uint32_t w[EPI == EPI_RESID ? NC / 32 : 1][16];
  __device__ __forceinline__ void load(const EpiParams& ep, const EpiCtx& ec, int m, int n0, int M, int N) {
    if constexpr (EPI == EPI_RESID) {
      if (m >= M) return;
      const __nv_bfloat16* row = ep.resid + static_cast<long>(ep.resid_mod > 0 ? m % ep.resid_mod : m) * ep.ldr;
#pragma unroll
      for (int c = 0; c < NC / 32; ++c) {
        const int nb = n0 + c * 32;
        if (nb < N) ld32_bf16(row + nb, N - nb, w[c]);
      }
    }
  }
};

// FROM_WS (split-K, last CTA of a tile): the accumulator chunk is the sum of the `splits` fp32 partials in the workspace, added in split
// order (deterministic), instead of a TMEM read.
template <int EPI, bool FROM_WS = false>
__device__ __forceinline__ void epilogue_chunk(uint32_t taddr, int m, int nb, bool row_ok, int N, __nv_bfloat16* C, int ldc, const EpiParams& ep,
                                               const EpiCtx& ec, const uint32_t (&wres)[16], long rrow, const float* wsp = nullptr, int splits = 0,
                                               long split_stride = 0) {
  uint32_t r[32];
  if constexpr (!FROM_WS) tmem_ld_32x32(taddr, r);
  const int valid = N - nb;  // >= 32: whole chunk
  uint32_t wb[16], wl[16], wr[16];
  if (EPI != EPI_SWIGLU && ec.has_bias) ld32_bf16(ep.bias + nb, valid, wb);
  if (EPI != EPI_SWIGLU && ec.has_ls) ld32_bf16(ep.ls + nb, valid, wl);
  if (EPI == EPI_ANY && ec.has_res && row_ok) ld32_bf16(ep.resid + rrow + nb, valid, wr);
  float v[32];
  if constexpr (FROM_WS) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    for (int sp = 0; sp < splits; ++sp) {
      const float4* src = reinterpret_cast<const float4*>(wsp + sp * split_stride);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 f = __ldcg(src + j);
        v[4 * j] += f.x, v[4 * j + 1] += f.y, v[4 * j + 2] += f.z, v[4 * j + 3] += f.w;
      }
    }
    if (!row_ok) return;
  } else {
    tmem_ld_wait_on(r);
    if (!row_ok) return;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  }
  if (EPI != EPI_SWIGLU && ec.has_bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += bf16_of(wb, j);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);  // Linear output (bias fused before rounding, as cuBLASLt)
  if (EPI == EPI_ANY && ec.do_gelu) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = bf16_round(gelu_erf(v[j]));
  }
  if (EPI != EPI_SWIGLU && ec.has_ls) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j] * bf16_of(wl, j));
  }
  if (EPI == EPI_RESID) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j] + bf16_of(wres, j));
  } else if (EPI == EPI_ANY && ec.has_res) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j] + bf16_of(wr, j));
  }
  if (EPI == EPI_SWIGLU || (EPI == EPI_ANY && ec.do_swiglu)) {
    // interleaved (gate, up) columns -> one output column per pair: bf16(bf16(silu(g)) * u)
    __nv_bfloat16* crow = C + static_cast<long>(m) * ldc + nb / 2;
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = pack_bf16(bf16_round(silu(v[4 * j])) * v[4 * j + 1], bf16_round(silu(v[4 * j + 2])) * v[4 * j + 3]);
    if (valid >= 32 && (reinterpret_cast<uintptr_t>(crow) & 15) == 0) {
      reinterpret_cast<uint4*>(crow)[0] = make_uint4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<uint4*>(crow)[1] = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (2 * j + 1 < valid) crow[j] = __ushort_as_bfloat16(static_cast<unsigned short>((j & 1) ? (o[j >> 1] >> 16) : (o[j >> 1] & 0xffffu)));
    }
  } else {
    __nv_bfloat16* crow = C + static_cast<long>(m) * ldc + nb;
    if (valid >= 32 && (reinterpret_cast<uintptr_t>(crow) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        reinterpret_cast<uint4*>(crow)[j] = make_uint4(pack_bf16(v[8 * j + 0], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                                                       pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) crow[j] = __float2bfloat16_rn(v[j]);
    }
  }
}

// EPI_ROPE: one epilogue thread = one row m (token t of sequence b) x one HEAD (the 128 accumulator columns at `taddr`, output columns
// n0 .. n0 + 127 of the packed q|k|v row; head_dim = 128, tiles are 256 columns wide, so a column half is exactly one head of q, k or v).
//   q, k : x_embed = bf16(bf16(x cos) + bf16(rotate_half(x) sin)), bf16 tables — the arithmetic of rope_kvstore_kernel (ops.cu), which is
//          transformers' apply_rotary_pos_emb in bf16 (modeling_prismatic.py:404-415 calls it through LlamaFlashAttention2); element d pairs
//          with d + 64, so the thread works on the 32-column chunks (c, c + 2) together;
//   k, v : also stored into the paged KV cache [page][head][slot][128] (what DynamicCache.update does in the reference);
//   all three land in the packed qkv buffer the prefill attention reads.
// Saves the separate RoPE / KV-store pass over the 12288-wide rows (read + write of q|k|v once more, 96 us per layer at bs = 32).
__device__ __forceinline__ void epilogue_rope_head(uint32_t taddr, int m, int n0, int M, __nv_bfloat16* C, int ldc, const RopeParams& rp) {
  const bool row_ok = m < M;
  const int Hd = rp.heads * 128;
  const int which = n0 / Hd, head = (n0 - which * Hd) >> 7;  // 0 = q, 1 = k, 2 = v
  const int b = row_ok ? m / rp.T : 0, t = row_ok ? m - b * rp.T : 0, pos = rp.pos0 + t;
  long cdst = 0;
  if (which > 0 && row_ok) {
    const int page = __ldg(rp.block_table + b * rp.max_pages + pos / rp.page_size), slot = pos % rp.page_size;
    cdst = ((static_cast<long>(page) * rp.heads + head) * rp.page_size + slot) * 128;
  }
  __nv_bfloat16* crow = C + static_cast<long>(m) * ldc + n0;
  __nv_bfloat16* cache = (which == 1) ? rp.k_cache : rp.v_cache;
#pragma unroll 1
  for (int c = 0; c < 2; ++c) {
    uint32_t lo[32], hi[32];
    tmem_ld_32x32(taddr + 32 * c, lo);
    tmem_ld_32x32(taddr + 64 + 32 * c, hi);
    uint32_t cw[16], sw[16];
    if (which < 2 && row_ok) {
      ld32_bf16(rp.cos_tab + static_cast<long>(pos) * 64 + 32 * c, 32, cw);
      ld32_bf16(rp.sin_tab + static_cast<long>(pos) * 64 + 32 * c, 32, sw);
    }
    tmem_ld_wait_on(lo);
    tmem_ld_wait_on(hi);
    if (!row_ok) continue;
    uint32_t olo[16], ohi[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float x10 = bf16_round(__uint_as_float(lo[2 * j])), x11 = bf16_round(__uint_as_float(lo[2 * j + 1]));  // the Linear's bf16 output
      const float x20 = bf16_round(__uint_as_float(hi[2 * j])), x21 = bf16_round(__uint_as_float(hi[2 * j + 1]));
      if (which < 2) {
        const float c0 = bf16_lo(cw[j]), c1 = bf16_hi(cw[j]), s0 = bf16_lo(sw[j]), s1 = bf16_hi(sw[j]);
        olo[j] = pack_bf16(bf16_round(x10 * c0) + bf16_round(-x20 * s0), bf16_round(x11 * c1) + bf16_round(-x21 * s1));
        ohi[j] = pack_bf16(bf16_round(x20 * c0) + bf16_round(x10 * s0), bf16_round(x21 * c1) + bf16_round(x11 * s1));
      } else {
        olo[j] = pack_bf16(x10, x11), ohi[j] = pack_bf16(x20, x21);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 vl = make_uint4(olo[4 * j], olo[4 * j + 1], olo[4 * j + 2], olo[4 * j + 3]);
      const uint4 vh = make_uint4(ohi[4 * j], ohi[4 * j + 1], ohi[4 * j + 2], ohi[4 * j + 3]);
      reinterpret_cast<uint4*>(crow + 32 * c)[j] = vl;
      reinterpret_cast<uint4*>(crow + 64 + 32 * c)[j] = vh;
      if (which > 0) {
        reinterpret_cast<uint4*>(cache + cdst + 32 * c)[j] = vl;
        reinterpret_cast<uint4*>(cache + cdst + 64 + 32 * c)[j] = vh;
      }
    }
  }
}

template <int NC, int EPI>
__device__ __forceinline__ void epilogue_rows(uint32_t taddr, int m, int n0, int M, int N, __nv_bfloat16* C, int ldc, const EpiParams& ep,
                                              const EpiCtx& ec, const EpiResid<NC, EPI>& res) {
  const bool row_ok = m < M;
  if constexpr (EPI == EPI_RESID) {
#pragma unroll
    for (int c = 0; c < NC / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= N) break;  // warp-uniform
      epilogue_chunk<EPI>(taddr + c * 32, m, nb, row_ok, N, C, ldc, ep, ec, res.w[c], 0);
    }
  } else {
    const long rrow = (EPI == EPI_ANY && ec.has_res) ? static_cast<long>(ep.resid_mod > 0 ? m % ep.resid_mod : m) * ep.ldr : 0;
#pragma unroll 1
    for (int c = 0; c < NC / 32; ++c) {
      const int nb = n0 + c * 32;
      if (nb >= N) break;  // warp-uniform
      epilogue_chunk<EPI>(taddr + c * 32, m, nb, row_ok, N, C, ldc, ep, ec, res.w[0], rrow);
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, __nv_bfloat16* C,
               int ldc, int M, int N, int K, EpiParams ep, RopeParams rp) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tmem_full = empty + Cfg::kStages;  // [2] MMA -> epilogue: accumulator buffer complete
  uint64_t* tmem_empty = tmem_full + 2;        // [2] epilogue -> MMA: accumulator buffer read out
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + BK - 1) / BK;
  const int mt = (M + BM - 1) / BM, nt = (N + BN - 1) / BN, n_tiles = mt * nt;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], kEpiWarps);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // ring iteration, continues across tiles: the producer runs ahead into the next tile
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int mi, ni;
        tile_coords(tile, mt, nt, mi, ni);
        const int m0 = mi * BM, n0 = ni * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, kb * BK, m0, &full[s]);
          tma_load_2d(sa + Cfg::kABytes, &tmB, kb * BK, n0, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      uint32_t it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const uint32_t buf = tl & 1, bph = (tl >> 1) & 1;
        mbar_wait(&tmem_empty[buf], bph ^ 1);  // the epilogue has read this accumulator buffer out (first two uses: free)
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_k128(sa), bdesc = umma_desc_k128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 elements (32 B) along K inside the 128-B swizzle span: +2 in the (addr >> 4) field
            umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty[s]);  // smem slot reusable once these MMAs retire
        }
        umma_commit(&tmem_full[buf]);  // accumulator complete
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3, ch = (warp - 4) >> 2;  // TMEM lane quarter this warp may read (warp id mod 4); column half it covers
    const EpiCtx ec(ep);
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      int mi, ni;
      tile_coords(tile, mt, nt, mi, ni);
      const int m0 = mi * BM, n0 = ni * BN + ch * (BN / 2);
      const uint32_t buf = tl & 1, bph = (tl >> 1) & 1;
      EpiResid<BN / 2, EPI> res;
      res.load(ep, ec, m0 + q * 32 + lane, n0, M, N);
      mbar_wait(&tmem_full[buf], bph);
      tc_fence_after();
      if constexpr (EPI == EPI_ROPE) {
        static_assert(EPI != EPI_ROPE || BN == 256, "EPI_ROPE: a column half must be one 128-wide head");
        epilogue_rope_head(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + ch * (BN / 2), m0 + q * 32 + lane, n0, M, C, ldc, rp);
      } else {
        epilogue_rows<BN / 2, EPI>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + ch * (BN / 2), m0 + q * 32 + lane, n0, M, N, C, ldc,
                              ep, ec, res);
      }
      // this warp's quarter of the accumulator buffer is read out: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Split-K variant for the small-M problems of a bs = 1 request whose 128 x 128 tiles do not fill the machine (Llama o_proj / down_proj at
// M = 296: 96 tiles; ViT proj / fc2 at M ~ 260: 18-24 tiles): work item = (tile, k-range). Every CTA dumps its fp32 partial tile into the
// caller's workspace; the CTA that arrives LAST at the tile's counter adds the partials in split order (deterministic, whoever is last) and
// runs the fused epilogue chain on the sum. Same pipeline as gemm_tn_kernel (BN = 128).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSplitCounterBytes = 4096;  // head of the workspace: one self-resetting arrival counter per tile (<= 1024 tiles)

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, __nv_bfloat16* C, int ldc, int M, int N,
                      int K, EpiParams ep, int splits, uint8_t* workspace) {
  constexpr int BN = 128;
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tmem_full = empty + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  volatile uint32_t* last_flag = tmem_holder + 1;
  unsigned int* counters = reinterpret_cast<unsigned int*>(workspace);
  float* wsf = reinterpret_cast<float*>(workspace + kSplitCounterBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (K + BK - 1) / BK;
  const int mt = (M + BM - 1) / BM, nt = (N + BN - 1) / BN, n_tiles = mt * nt, n_items = n_tiles * splits;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_holder, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tile = item % n_tiles, split = item / n_tiles;
        int mi, ni;
        tile_coords(tile, mt, nt, mi, ni);
        const int m0 = mi * BM, n0 = ni * BN;
        const int kb0 = split * nkb / splits, kb1 = (split + 1) * nkb / splits;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
          tma_load_2d(sa, &tmA, kb * BK, m0, &full[s]);
          tma_load_2d(sa + Cfg::kABytes, &tmB, kb * BK, n0, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      uint32_t it = 0, tl = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tl) {
        const int split = item / n_tiles;
        const int kb0 = split * nkb / splits, kb1 = (split + 1) * nkb / splits;
        const uint32_t buf = tl & 1, bph = (tl >> 1) & 1;
        mbar_wait(&tmem_empty[buf], bph ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_k128(sa), bdesc = umma_desc_k128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, ((kb - kb0) | k) != 0);
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3, ch = (warp - 4) >> 2;
    const EpiCtx ec(ep);
    const int row = q * 32 + lane;
    const long split_stride = static_cast<long>(BM) * BN;  // floats between the partials of one tile
    uint32_t tl = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tl) {
      const int tile = item % n_tiles, split = item / n_tiles;
      int mi, ni;
      tile_coords(tile, mt, nt, mi, ni);
      const int m0 = mi * BM, n0 = ni * BN + ch * (BN / 2);
      const uint32_t buf = tl & 1, bph = (tl >> 1) & 1;
      float* wtile = wsf + static_cast<long>(tile) * splits * split_stride + static_cast<long>(row) * BN + ch * (BN / 2);
      mbar_wait(&tmem_full[buf], bph);
      tc_fence_after();
      // ---- dump this k-range's partial (fp32, this thread's row, its half of the columns)
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + ch * (BN / 2);
#pragma unroll 1
      for (int c = 0; c < BN / 2 / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait_on(r);
        float4* dst = reinterpret_cast<float4*>(wtile + split * split_stride + c * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          __stcg(dst + j, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);  // accumulator buffer handed back: the MMA warp runs on
      // ---- arrive at the tile's counter; the last of the `splits` CTAs reduces and finishes the tile
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      if (threadIdx.x == 128) {
        const unsigned int old = atomicAdd(&counters[tile], 1u);
        const bool last = old == static_cast<unsigned int>(splits - 1);
        if (last) counters[tile] = 0;  // everybody has arrived: ready for the next launch
        *last_flag = last ? 1u : 0u;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      if (*last_flag) {
        __threadfence();
        const int m = m0 + row;
        EpiResid<BN / 2, EPI> res;
        res.load(ep, ec, m, n0, M, N);
        const bool row_ok = m < M;
        const long rrow = (EPI == EPI_ANY && ec.has_res) ? static_cast<long>(ep.resid_mod > 0 ? m % ep.resid_mod : m) * ep.ldr : 0;
#pragma unroll
        for (int c = 0; c < BN / 2 / 32; ++c) {
          const int nb = n0 + c * 32;
          if (nb < N)
            epilogue_chunk<EPI, true>(0, m, nb, row_ok, N, C, ldc, ep, ec, res.w[EPI == EPI_RESID ? c : 0], rrow, wtile + c * 32, splits, split_stride);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");  // last_flag is rewritten for the next item
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant for the large-M problems (bs = 32 prefill probe, batched ViT): cluster of 2 CTAs = one TPC, tcgen05 cta_group::2,
// output tile 256 x 256 per pair. Why: with cta_group::1 every SM pulls a 128-row A slab AND a 256-row W slab per k-block (48 KB per
// 128x256x64 MMA block = 96 B/clk/SM at full tensor rate), 2.3 x what the L2 -> SM fabric delivers chip-wide (~6300 B/clk / 148 SMs,
// B300_MICROARCH.md "LTS throughput cap"): the single-CTA kernel is L2-bound at ~45 % of the tensor peak. In a pair each SM stages
// only its own half of both operands (16 KB + 16 KB per k-block) and the tensor cores read the other half from the peer's shared
// memory: 64 B/clk/SM, and 6 ring stages instead of 4 in the same shared memory.
//   rank r of the pair: A rows m0 + 128 r .. +128, W rows n0 + 128 r .. +128, accumulator lanes = its 128 output rows x 256 columns,
//   double-buffered (2 x 256 TMEM columns). Barriers: full[s] lives in the LEADER (both CTAs' TMA bytes are credited to it), empty[s]
//   and tmem_full[b] are multicast commits to both CTAs, tmem_empty[b] is the leader's and counts the 8 epilogue warps of the pair.
// ---------------------------------------------------------------------------------------------------------------
struct PairCfg {
  static constexpr int BN = 256;
#ifndef EMX_PAIR_STAGES
#define EMX_PAIR_STAGES 6
#endif
  static constexpr int kStages = EMX_PAIR_STAGES;
  static constexpr int kABytes = BM * BK * 2;        // this CTA's 128 A rows
  static constexpr int kBBytes = (BN / 2) * BK * 2;  // this CTA's 128 W rows
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, __nv_bfloat16* C, int ldc, int M,
                    int N, int K, EpiParams ep, RopeParams rp) {
  using Cfg = PairCfg;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tmem_full = empty + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int nkb = (K + BK - 1) / BK;
  const int mt = (M + 2 * BM - 1) / (2 * BM), nt = (N + BN - 1) / BN, n_tiles = mt * nt;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);   // used in the leader only: its producer's arrive.expect_tx for the bytes of BOTH CTAs
      mbar_init(&empty[s], 1);  // multicast commit of the leader's MMA warp
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);   // multicast commit
      mbar_init(&tmem_empty[b], 2 * kEpiWarps);  // leader only: the epilogue warps of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_holder, 2 * BN);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / remote TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        int mi, ni;
        tile_coords(tile, mt, nt, mi, ni);
        const int m0 = mi * (2 * BM) + rank * BM, n0 = ni * BN + rank * (BN / 2);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&empty[s], ph ^ 1);  // own slot free (the commit is multicast)
          uint8_t* sa = smem + s * Cfg::kStageBytes;
          const uint32_t leader_full = mapa_u32(smem_u32(&full[s]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::kStageBytes);
          tma_load_2d_pair(sa, &tmA, kb * BK, m0, leader_full);
          tma_load_2d_pair(sa + Cfg::kABytes, &tmB, kb * BK, n0, leader_full);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
      uint32_t it = 0, tl = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs, ++tl) {
        const uint32_t buf = tl & 1, bph = (tl >> 1) & 1;
        mbar_wait(&tmem_empty[buf], bph ^ 1);  // both CTAs have read this accumulator buffer out
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % Cfg::kStages;
          const uint32_t ph = (it / Cfg::kStages) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_k128(sa), bdesc = umma_desc_k128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) umma_bf16_pair(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty[s]);
        }
        umma_commit_pair(&tmem_full[buf]);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3, ch = (warp - 4) >> 2;
    const EpiCtx ec(ep);
    uint32_t tl = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++tl) {
      int mi, ni;
      tile_coords(tile, mt, nt, mi, ni);
      const int m0 = mi * (2 * BM) + rank * BM, n0 = ni * BN + ch * (BN / 2);
      const uint32_t buf = tl & 1, bph = (tl >> 1) & 1;
      EpiResid<BN / 2, EPI> res;
      res.load(ep, ec, m0 + q * 32 + lane, n0, M, N);
      mbar_wait(&tmem_full[buf], bph);
      tc_fence_after();
      if constexpr (EPI == EPI_ROPE) {
        epilogue_rope_head(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + ch * (BN / 2), m0 + q * 32 + lane, n0, M, C, ldc, rp);
      } else {
        epilogue_rows<BN / 2, EPI>(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + ch * (BN / 2), m0 + q * 32 + lane, n0, M, N, C, ldc,
                              ep, ec, res);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[buf]), 0));
    }
  }
  tc_fence_before();
  cluster_sync_all();  // neither CTA may leave (shared memory, barriers, TMEM) while the other still depends on it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 2 * BN);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor [rows, cols] (cols contiguous, leading dim `ld` elements), box = [box_rows, 64 cols], 128-B swizzle,
// out-of-bounds elements read as zero (handles M / N / K tails).
static int make_tmap(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  EMX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EMX_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rows=%d cols=%d ld=%d base=%p", (int)r, rows, cols, ld, base);
  return 0;
}

static int epi_kind(const EpiParams& ep) {
  const bool gelu = ep.flags & EMX_EPI_GELU, swiglu = ep.flags & EMX_EPI_SWIGLU;
  if (swiglu && !gelu && !ep.bias && !ep.ls && !ep.resid) return EPI_SWIGLU;
  if (ep.resid && !gelu && !swiglu) return EPI_RESID;
  return EPI_ANY;
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, __nv_bfloat16* C, int ldc, int M, int N, int K, const EpiParams& ep,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  bool* attr_set = device_attr_flag(BN == 128 ? ATTR_GEMM128 : BN == 256 ? ATTR_GEMM256 : ATTR_GEMM64);  // per device: the opt-in is device state
  const int sms = device_sms();
  if (!attr_set || sms < 0) return -2;
  if (!*attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, EPI_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    *attr_set = true;
  }
  const int n_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int grid = n_tiles < sms ? n_tiles : sms;
  switch (epi_kind(ep)) {
    case EPI_RESID: gemm_tn_kernel<BN, EPI_RESID><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, RopeParams{}); break;
    case EPI_SWIGLU: gemm_tn_kernel<BN, EPI_SWIGLU><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, RopeParams{}); break;
    default: gemm_tn_kernel<BN, EPI_ANY><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, RopeParams{}); break;
  }
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int launch_gemm_splitk(const CUtensorMap& ta, const CUtensorMap& tb, __nv_bfloat16* C, int ldc, int M, int N, int K, const EpiParams& ep,
                              int splits, void* workspace, cudaStream_t stream) {
  using Cfg = GemmCfg<128>;
  bool* attr_set = device_attr_flag(ATTR_GEMM_SPLITK);
  const int sms = device_sms();
  if (!attr_set || sms < 0) return -2;
  if (!*attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_splitk_kernel<EPI_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_splitk_kernel<EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_splitk_kernel<EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    *attr_set = true;
  }
  const int n_items = ((M + BM - 1) / BM) * ((N + 127) / 128) * splits;
  const int grid = n_items < sms ? n_items : sms;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  switch (epi_kind(ep)) {
    case EPI_RESID: gemm_tn_splitk_kernel<EPI_RESID><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, splits, ws); break;
    case EPI_SWIGLU: gemm_tn_splitk_kernel<EPI_SWIGLU><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, splits, ws); break;
    default: gemm_tn_splitk_kernel<EPI_ANY><<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, splits, ws); break;
  }
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// k-ranges per 128 x 128 tile for an under-filled problem: minimise  rounds(tiles * s) * (k-blocks / s * t_kb + t_fixed) + t_reduce  with the
// measured per-CTA rates of this kernel (~0.43 us per 32 KB k-block: the L2 -> SM ingest of one SM; ~3 us of pipeline fill + epilogue per
// work item; ~1.5 us for the dump + reduction). 1 = do not split.
static int pick_splits(long tiles, int nkb, int sms, long workspace_bytes) {
  int best = 1;
  double best_t = 1e30;
  for (int s = 1; s <= 8 && s * 4 <= nkb; ++s) {
    if (s > 1 && (tiles > 1024 || static_cast<long>(kSplitCounterBytes) + tiles * s * BM * 128 * 4 > workspace_bytes)) break;
    const double rounds = static_cast<double>((tiles * s + sms - 1) / sms);
    const double t = rounds * (static_cast<double>(nkb) / s * 0.43 + 3.0) + (s > 1 ? 1.5 : 0.0);
    if (t < best_t * 0.93) best_t = t, best = s;  // a split has to buy at least 7 %
  }
  return best;
}

static int launch_gemm_pair(const CUtensorMap& ta, const CUtensorMap& tb, __nv_bfloat16* C, int ldc, int M, int N, int K, const EpiParams& ep,
                            cudaStream_t stream) {
  bool* attr_set = device_attr_flag(ATTR_GEMM_PAIR);
  const int sms = device_sms();
  if (!attr_set || sms < 0) return -2;
  if (!*attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_pair_kernel<EPI_ANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::kSmemBytes));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_pair_kernel<EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::kSmemBytes));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_pair_kernel<EPI_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::kSmemBytes));
    *attr_set = true;
  }
  const int n_tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + PairCfg::BN - 1) / PairCfg::BN);
  const int pairs = n_tiles < sms / 2 ? n_tiles : sms / 2;
  switch (epi_kind(ep)) {
    case EPI_RESID: gemm_tn_pair_kernel<EPI_RESID><<<2 * pairs, kGemmThreads, PairCfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, RopeParams{}); break;
    case EPI_SWIGLU: gemm_tn_pair_kernel<EPI_SWIGLU><<<2 * pairs, kGemmThreads, PairCfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, RopeParams{}); break;
    default: gemm_tn_pair_kernel<EPI_ANY><<<2 * pairs, kGemmThreads, PairCfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, RopeParams{}); break;
  }
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// fused q|k|v projection: CTA pairs for large M, the single-CTA 128 x 256 kernel otherwise (both have 128-column epilogue halves = one head)
static int launch_gemm_rope(bool pair, const CUtensorMap& ta, const CUtensorMap& tb, __nv_bfloat16* C, int ldc, int M, int N, int K,
                            const RopeParams& rp, cudaStream_t stream) {
  bool* attr_set = device_attr_flag(ATTR_GEMM_ROPE);
  const int sms = device_sms();
  if (!attr_set || sms < 0) return -2;
  if (!*attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_pair_kernel<EPI_ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg::kSmemBytes));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<256, EPI_ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256>::kSmemBytes));
    *attr_set = true;
  }
  const EpiParams ep{nullptr, nullptr, nullptr, 0, 0, 0};
  if (pair) {
    const int n_tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + PairCfg::BN - 1) / PairCfg::BN);
    const int pairs = n_tiles < sms / 2 ? n_tiles : sms / 2;
    gemm_tn_pair_kernel<EPI_ROPE><<<2 * pairs, kGemmThreads, PairCfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, rp);
  } else {
    const int n_tiles = ((M + BM - 1) / BM) * ((N + 255) / 256);
    const int grid = n_tiles < sms ? n_tiles : sms;
    gemm_tn_kernel<256, EPI_ROPE><<<grid, kGemmThreads, GemmCfg<256>::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep, rp);
  }
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace emx

extern "C" int emx_rope_kvstore(void* qkv, int B, int T, int heads, int hd, const void* cos_tab, const void* sin_tab, int pos0, void* k_cache,
                                void* v_cache, const int32_t* block_table, int max_pages, int page_size, cudaStream_t s);

extern "C" int emx_gemm_qkv_rope(const void* A, int lda, const void* W, int ldw, void* qkv, int B, int T, int heads, int head_dim, int K,
                                 const void* cos_tab, const void* sin_tab, int pos0, void* k_cache, void* v_cache, const int32_t* block_table,
                                 int max_pages, int page_size, cudaStream_t stream) {
  using namespace emx;
  EMX_REQUIRE(B > 0 && T > 0 && heads > 0 && head_dim > 0 && K > 0, "emx_gemm_qkv_rope: bad shape");
  EMX_REQUIRE(cos_tab && sin_tab && k_cache && v_cache && block_table && page_size > 0, "emx_gemm_qkv_rope: null pointer");
  const int M = B * T, N = 3 * heads * head_dim;
  const int sms = device_sms();
  if (sms < 0) return -2;
  const char* fe = getenv("EMX_QKV_ROPE_FUSED");  // A/B switch: 0 = always the two-kernel path
  const long pair_tiles = static_cast<long>((M + 2 * BM - 1) / (2 * BM)) * ((N + 255) / 256);
  const bool pair = pair_tiles >= sms / 2 && M >= 4 * BM;
  const bool wide = static_cast<long>((M + BM - 1) / BM) * ((N + 255) / 256) >= static_cast<long>(sms) * 90 / 100;
  const bool fused = head_dim == 128 && (pair || wide) && !(fe && fe[0] == '0') && lda % 8 == 0 && ldw % 8 == 0 &&
                     (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0;
  if (!fused) {  // shapes the fused epilogue does not cover (other head sizes, problems too small for 256-column tiles): GEMM, then the RoPE / KV pass
    if (int r = emx_gemm_bf16_ws(A, lda, W, ldw, qkv, N, M, N, K, nullptr, nullptr, nullptr, 0, 0, 0, nullptr, 0, stream)) return r;
    return emx_rope_kvstore(qkv, B, T, heads, head_dim, cos_tab, sin_tab, pos0, k_cache, v_cache, block_table, max_pages, page_size, stream);
  }
  CUtensorMap ta, tb;
  if (int r = make_tmap(&ta, A, M, K, lda, BM)) return r;
  if (int r = make_tmap(&tb, W, N, K, ldw, pair ? PairCfg::BN / 2 : 256)) return r;
  const RopeParams rp{static_cast<const __nv_bfloat16*>(cos_tab), static_cast<const __nv_bfloat16*>(sin_tab), static_cast<__nv_bfloat16*>(k_cache),
                      static_cast<__nv_bfloat16*>(v_cache), block_table, T, pos0, heads, max_pages, page_size};
  return launch_gemm_rope(pair, ta, tb, static_cast<__nv_bfloat16*>(qkv), N, M, N, K, rp, stream);
}

extern "C" int emx_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, const void* bias,
                             const void* layerscale, const void* resid, int ldr, int resid_mod, int flags, cudaStream_t stream) {
  return emx_gemm_bf16_ws(A, lda, W, ldw, C, ldc, M, N, K, bias, layerscale, resid, ldr, resid_mod, flags, nullptr, 0, stream);
}

extern "C" int emx_gemm_bf16_ws(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, const void* bias,
                                const void* layerscale, const void* resid, int ldr, int resid_mod, int flags, void* workspace,
                                long workspace_bytes, cudaStream_t stream) {
  using namespace emx;
  EMX_REQUIRE(M > 0 && N > 0 && K > 0, "emx_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K);
  EMX_REQUIRE(lda % 8 == 0 && ldw % 8 == 0, "emx_gemm_bf16: lda/ldw must be multiples of 8 elements (TMA 16-B strides): %d %d", lda, ldw);
  EMX_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0, "emx_gemm_bf16: A/W must be 16-B aligned");
  EMX_REQUIRE(!(flags & EMX_EPI_SWIGLU) || (N % 2 == 0), "emx_gemm_bf16: SwiGLU epilogue needs even N");
  EpiParams ep{static_cast<const __nv_bfloat16*>(bias), static_cast<const __nv_bfloat16*>(layerscale),
               static_cast<const __nv_bfloat16*>(resid), ldr, resid_mod, flags};
  CUtensorMap ta, tb;
  // BN = 128 keeps the grid at >= ~1 wave for the M <= 300 problems of a bs=1 request; BN = 256 for large M.
  const int sms = device_sms();
  if (sms < 0) return -2;
  // 128 x 256 tiles once they fill ~0.9 of a wave (each CTA then pulls 48 KB per 128x256x64 MMA block instead of 2 x 32 KB: the GEMMs of
  // this path are bound by the bytes an SM can pull from L2); EMX_GEMM_WIDE_MIN_PCT overrides the threshold (A/B runs)
  const char* we = getenv("EMX_GEMM_WIDE_MIN_PCT");
  const long wide_min = static_cast<long>(sms) * (we ? atoi(we) : 90) / 100;
  const bool wide = (static_cast<long>((M + BM - 1) / BM) * ((N + 255) / 256) >= wide_min);
  // CTA pairs (256 x 256 tiles) once M is large and there is at least one full round of pair tiles
  const char* pe = getenv("EMX_GEMM_PAIR");  // A/B switch for tools/gemm_probe.py: 0 = never, 2 = whenever M > 128
  const int pair_mode = pe ? pe[0] - '0' : 1;
  const long pair_tiles = static_cast<long>((M + 2 * BM - 1) / (2 * BM)) * ((N + 255) / 256);
  if ((pair_mode == 1 && pair_tiles >= sms / 2 && M >= 4 * BM) || (pair_mode == 2 && M > BM)) {
    if (int r = make_tmap(&ta, A, M, K, lda, BM)) return r;
    if (int r = make_tmap(&tb, W, N, K, ldw, PairCfg::BN / 2)) return r;
    return launch_gemm_pair(ta, tb, static_cast<__nv_bfloat16*>(C), ldc, M, N, K, ep, stream);
  }
  if (!wide && workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0) {
    // OFF unless EMX_GEMM_SPLITK=1: measured on B200 (profiles/r02_splitk_negative.txt) every bs = 1 problem got 7-11 us SLOWER with it - the
    // dump -> fence -> counter -> serial reduction tail costs ~10 us, and k-ranges on more SMs do not stream faster in proportion (the
    // weights come from HBM). Kept as a tested building block (tests/test_gpu_kernels.py::test_gemm_split_k) with that result recorded.
    const char* se = getenv("EMX_GEMM_SPLITK");
    const int splits = !(se && se[0] == '1') ? 1 : pick_splits(static_cast<long>((M + BM - 1) / BM) * ((N + 127) / 128), (K + BK - 1) / BK, sms, workspace_bytes);
    if (splits > 1) {
      if (int r = make_tmap(&ta, A, M, K, lda, BM)) return r;
      if (int r = make_tmap(&tb, W, N, K, ldw, 128)) return r;
      return launch_gemm_splitk(ta, tb, static_cast<__nv_bfloat16*>(C), ldc, M, N, K, ep, splits, workspace, stream);
    }
  }
  // 128 x 64 tiles when 128 x 128 tiles leave more than half of the SMs idle (ViT proj / fc2 and the projector's fc3 of a bs=1 request: 18-64 tiles):
  // twice the CTAs, each pulling 24 KB instead of 32 KB per k-block; EMX_GEMM_NARROW=0 switches it off (A/B runs)
  const char* ne = getenv("EMX_GEMM_NARROW");
  const bool narrow = !wide && !(ne && ne[0] == '0') && static_cast<long>((M + BM - 1) / BM) * ((N + 127) / 128) * 2 <= sms && N > 64;
  if (int r = make_tmap(&ta, A, M, K, lda, BM)) return r;
  if (int r = make_tmap(&tb, W, N, K, ldw, wide ? 256 : narrow ? 64 : 128)) return r;
  if (narrow) return launch_gemm<64>(ta, tb, static_cast<__nv_bfloat16*>(C), ldc, M, N, K, ep, stream);
  return wide ? launch_gemm<256>(ta, tb, static_cast<__nv_bfloat16*>(C), ldc, M, N, K, ep, stream)
              : launch_gemm<128>(ta, tb, static_cast<__nv_bfloat16*>(C), ldc, M, N, K, ep, stream);
}
