// emx_attn_fwd — FlashAttention-style (tiled keys, online softmax, no S x S matrix in memory) full-sequence attention
// for the ViT towers (non-causal, head_dim 64 / 72) and the Llama prefill (causal, head_dim 128).
//
// Replaces F.scaled_dot_product_attention inside timm `Attention.forward` (driven from
// /root/reference/prismatic/extern/hf/modeling_prismatic.py:121) and flash_attn_varlen_func inside transformers
// `LlamaFlashAttention2` (selected by attn_implementation="flash_attention_2", experiments/robot/openvla_utils.py:45).
//
// Numerics follow flash-attn: fp32 scores / running max / running sum, probabilities rounded to bf16 before the PV
// product, fp32 output accumulator normalised once at the end, bf16 output.
//
// v2 mapping (tensor cores): one CTA per (64- or 32-query block, head, batch), one warp per 16 queries. Q fragments stay in
// registers; K/V tiles of 64 keys are staged in shared memory (16-byte padded rows: conflict-free ldmatrix); S = Q K^T and
// O += P V run on mma.sync.m16n8k16 (bf16 -> fp32) with the FlashAttention-2 register hand-over of P (the accumulator layout of
// S is the A-operand layout of the next MMA); V fragments come from ldmatrix.trans. head_dim 72 is zero-padded to 80 in shared
// memory. Sequences here are <= ~300 tokens (5 key tiles), so the tile loop is not software-pipelined.
#include <cstdlib>

#include "common.cuh"
#include "emmax.h"

namespace emx {

constexpr int ATT_QB = 16;   // queries per CTA
constexpr int ATT_KT = 64;   // keys per tile
constexpr int ATT_WARPS = 4;

template <int HD>
__global__ void __launch_bounds__(ATT_WARPS * 32) attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                                  int T, int heads, int causal, float scale) {
  constexpr int KS = HD + 2;          // padded K row stride (elements): (HD/2 + 1) words is odd for 64/72/128
  constexpr int NP = HD / 2;          // bf16 pairs per row
  constexpr int PPL = (NP + 31) / 32; // pairs per lane in the PV phase
  __shared__ __align__(16) __nv_bfloat16 sK[ATT_KT * KS];
  __shared__ __align__(16) __nv_bfloat16 sV[ATT_KT * HD];
  __shared__ float sQ[ATT_WARPS][HD];
  __shared__ float sP[ATT_WARPS][ATT_KT];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_QB, head = blockIdx.y, b = blockIdx.z;
  const int Hd = heads * HD;
  const long rs = 3L * Hd;  // row stride of the packed qkv buffer
  const __nv_bfloat16* base = qkv + static_cast<long>(b) * T * rs;
  const int q_hi = min(q0 + ATT_QB, T);                    // exclusive
  const int k_end = causal ? q_hi : T;                     // keys this CTA needs
  constexpr int QPW = ATT_QB / ATT_WARPS;

  float m_run[QPW], l_run[QPW], acc[QPW][PPL][2];
#pragma unroll
  for (int i = 0; i < QPW; ++i) {
    m_run[i] = -INFINITY, l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < PPL; ++j) acc[i][j][0] = acc[i][j][1] = 0.f;
  }

  for (int kt = 0; kt < k_end; kt += ATT_KT) {
    const int nk = min(ATT_KT, k_end - kt);
    __syncthreads();  // previous tile fully consumed
    for (int i = threadIdx.x; i < nk * NP; i += blockDim.x) {
      const int r = i / NP, c = i % NP;
      const __nv_bfloat16* krow = base + static_cast<long>(kt + r) * rs + Hd + head * HD;
      reinterpret_cast<uint32_t*>(sK + r * KS)[c] = reinterpret_cast<const uint32_t*>(krow)[c];
      reinterpret_cast<uint32_t*>(sV + r * HD)[c] = reinterpret_cast<const uint32_t*>(krow + Hd)[c];
    }
    __syncthreads();

#pragma unroll
    for (int qi = 0; qi < QPW; ++qi) {
      const int q = q0 + warp * QPW + qi;
      if (q >= T) break;                       // warp-uniform
      if (causal && kt > q) continue;          // whole tile is in the future of this query
      {
        // (re)load q into this warp's smem slot as fp32 — cheap relative to the tile work
        const __nv_bfloat16* qrow = base + static_cast<long>(q) * rs + head * HD;
        for (int d = lane; d < HD; d += 32) sQ[warp][d] = ld_bf16(qrow + d);
        __syncwarp();
      }
      // scores: lane owns keys lane and lane+32
      float s[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int kk = lane + 32 * h;
        float a = 0.f;
        if (kk < nk) {
          const uint32_t* kr = reinterpret_cast<const uint32_t*>(sK + kk * KS);
#pragma unroll 8
          for (int c = 0; c < NP; ++c) {
            const uint32_t w = kr[c];
            a = fmaf(sQ[warp][2 * c], bf16_lo(w), a);
            a = fmaf(sQ[warp][2 * c + 1], bf16_hi(w), a);
          }
          a *= scale;
          if (causal && kt + kk > q) a = -INFINITY;
        } else {
          a = -INFINITY;
        }
        s[h] = a;
      }
      const float tmax = warp_max(fmaxf(s[0], s[1]));
      const float m_new = fmaxf(m_run[qi], tmax);
      const float corr = (m_run[qi] == -INFINITY) ? 0.f : __expf(m_run[qi] - m_new);
      const float p0 = (s[0] == -INFINITY) ? 0.f : __expf(s[0] - m_new);
      const float p1 = (s[1] == -INFINITY) ? 0.f : __expf(s[1] - m_new);
      l_run[qi] = l_run[qi] * corr + warp_sum(p0 + p1);
      m_run[qi] = m_new;
      sP[warp][lane] = bf16_round(p0);
      sP[warp][lane + 32] = bf16_round(p1);
      __syncwarp();
      // PV: lane owns bf16 pairs lane, lane+32, ...
#pragma unroll
      for (int j = 0; j < PPL; ++j) acc[qi][j][0] *= corr, acc[qi][j][1] *= corr;
      for (int kk = 0; kk < nk; ++kk) {
        const float p = sP[warp][kk];
        const uint32_t* vr = reinterpret_cast<const uint32_t*>(sV + kk * HD);
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
          const int c = lane + 32 * j;
          if (c < NP) {
            const uint32_t w = vr[c];
            acc[qi][j][0] = fmaf(p, bf16_lo(w), acc[qi][j][0]);
            acc[qi][j][1] = fmaf(p, bf16_hi(w), acc[qi][j][1]);
          }
        }
      }
      __syncwarp();
    }
  }

#pragma unroll
  for (int qi = 0; qi < QPW; ++qi) {
    const int q = q0 + warp * QPW + qi;
    if (q >= T) break;
    const float inv = 1.0f / l_run[qi];
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + (static_cast<long>(b) * T + q) * Hd + head * HD);
#pragma unroll
    for (int j = 0; j < PPL; ++j) {
      const int c = lane + 32 * j;
      if (c < NP) orow[c] = pack_bf16(acc[qi][j][0] * inv, acc[qi][j][1] * inv);
    }
  }
}


// ---- v2: tensor cores -----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int ATT2_KT = 64;  // keys per tile

template <int HD, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) attn_fwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T,
                                                                    int heads, int causal, float scale) {
  constexpr int HDP = (HD + 15) / 16 * 16;   // head_dim padded to the MMA k-step (72 -> 80)
  constexpr int RS = HDP * 2 + 16;           // shared-memory row stride in bytes (16-B skew: conflict-free ldmatrix)
  constexpr int KSTEPS = HDP / 16;           // k-steps of Q K^T  == pairs of 8-wide output tiles of P V
  constexpr int BM = NWARPS * 16;            // queries per CTA
  constexpr int CH = HD / 8;                 // 16-byte chunks per row in global memory
  constexpr int CHP = HDP / 8;
  __shared__ __align__(16) uint8_t sK[ATT2_KT * RS];
  __shared__ __align__(16) uint8_t sV[ATT2_KT * RS];
  static_assert(BM <= ATT2_KT, "the Q tile is staged through the K buffer");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * BM, head = blockIdx.y, b = blockIdx.z;
  const int Hd = heads * HD;
  const long rs = 3L * Hd;  // row stride (elements) of the packed qkv buffer: [q | k | v] x heads x HD
  const __nv_bfloat16* base = qkv + static_cast<long>(b) * T * rs + head * HD;

  // rows [r0, r0 + nrows) of q (which = 0), k (1) or v (2) -> smem tile, zero-filled beyond T and beyond HD
  auto load_tile = [&](uint8_t* dst, int which, int r0, int nrows) {
    for (int i = threadIdx.x; i < nrows * CHP; i += NWARPS * 32) {
      const int r = i / CHP, c = i % CHP;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r0 + r < T && c < CH) v = *reinterpret_cast<const uint4*>(base + static_cast<long>(r0 + r) * rs + which * Hd + c * 8);
      *reinterpret_cast<uint4*>(dst + r * RS + c * 16) = v;
    }
  };

  // ---- Q fragments (A operand), via the K buffer
  load_tile(sK, 0, q0, BM);
  __syncthreads();
  uint32_t qa[KSTEPS][4];
  {
    const uint32_t a_addr = smem_u32(sK) + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * RS + (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) ldsm_x4(a_addr + ks * 32, qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
  }

  float o[2 * KSTEPS][4];
#pragma unroll
  for (int n = 0; n < 2 * KSTEPS; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};  // rows g and g + 8 of this warp's 16 queries
  const int qrow0 = q0 + warp * 16 + g, qrow1 = qrow0 + 8;
  const int k_end = causal ? min(q0 + BM, T) : T;  // keys this CTA needs

  // ldmatrix lane offsets: K (B operand of Q K^T, non-transposed) and V (B operand of P V, transposed)
  const uint32_t k_lane = ((lane & 7) + (lane >> 4) * 8) * RS + ((lane >> 3) & 1) * 16;
  const uint32_t v_lane = ((lane & 7) + ((lane >> 3) & 1) * 8) * RS + (lane >> 4) * 16;

  for (int kt = 0; kt < k_end; kt += ATT2_KT) {
    __syncthreads();  // previous tile (or the Q staging) fully consumed
    load_tile(sK, 1, kt, ATT2_KT);
    load_tile(sV, 2, kt, ATT2_KT);
    __syncthreads();
    if (causal && kt > q0 + warp * 16 + 15) continue;  // the whole tile is in the future of this warp's queries (warp-uniform)

    // ---- S = Q K^T : 16 queries x 64 keys per warp
    float sc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(sK) + np * 16 * RS + k_lane + ks * 32, b0, b1, b2, b3);
        mma16816(sc[2 * np], qa[ks], b0, b1);
        mma16816(sc[2 * np + 1], qa[ks], b2, b3);
      }
    }
    // ---- scale, mask, online softmax (thread: rows g / g+8, keys 8n + 2t, 8n + 2t + 1)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int key = kt + 8 * n + 2 * t + e;
        const bool dead = key >= T;
        float s0 = sc[n][e] * scale, s1 = sc[n][2 + e] * scale;
        if (dead || (causal && key > qrow0)) s0 = -INFINITY;
        if (dead || (causal && key > qrow1)) s1 = -INFINITY;
        sc[n][e] = s0, sc[n][2 + e] = s1;
        mx0 = fmaxf(mx0, s0), mx1 = fmaxf(mx1, s1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)), mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)), mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
    const float c0 = (m_run[0] == -INFINITY) ? 0.f : __expf(m_run[0] - mn0), c1 = (m_run[1] == -INFINITY) ? 0.f : __expf(m_run[1] - mn1);
    m_run[0] = mn0, m_run[1] = mn1;
    float ls0 = 0.f, ls1 = 0.f;
    uint32_t pa[4][4];  // P as the A operand of the next MMA: k-step j covers keys 16j .. 16j + 15
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const float p00 = (sc[n][0] == -INFINITY) ? 0.f : __expf(sc[n][0] - mn0), p01 = (sc[n][1] == -INFINITY) ? 0.f : __expf(sc[n][1] - mn0);
      const float p10 = (sc[n][2] == -INFINITY) ? 0.f : __expf(sc[n][2] - mn1), p11 = (sc[n][3] == -INFINITY) ? 0.f : __expf(sc[n][3] - mn1);
      ls0 += p00 + p01, ls1 += p10 + p11;
      pa[n >> 1][(n & 1) * 2] = pack_bf16(p00, p01);      // a0 / a2: row g
      pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16(p10, p11);  // a1 / a3: row g + 8
    }
    l_run[0] = l_run[0] * c0 + ls0, l_run[1] = l_run[1] * c1 + ls1;  // per-thread partial row sums; reduced over the quad at the end
#pragma unroll
    for (int n = 0; n < 2 * KSTEPS; ++n) o[n][0] *= c0, o[n][1] *= c0, o[n][2] *= c1, o[n][3] *= c1;
    // ---- O += P V
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
      for (int np = 0; np < KSTEPS; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_trans(smem_u32(sV) + j * 16 * RS + v_lane + np * 32, b0, b1, b2, b3);
        mma16816(o[2 * np], pa[j], b0, b1);
        mma16816(o[2 * np + 1], pa[j], b2, b3);
      }
    }
  }

  // ---- normalise and store
  float l0 = l_run[0], l1 = l_run[1];
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1), l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1), l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  __nv_bfloat16* obase = out + static_cast<long>(b) * T * Hd + head * HD;
#pragma unroll
  for (int n = 0; n < 2 * KSTEPS; ++n) {
    const int d = 8 * n + 2 * t;
    if (d < HD) {
      if (qrow0 < T) *reinterpret_cast<uint32_t*>(obase + static_cast<long>(qrow0) * Hd + d) = pack_bf16(o[n][0] * i0, o[n][1] * i0);
      if (qrow1 < T) *reinterpret_cast<uint32_t*>(obase + static_cast<long>(qrow1) * Hd + d) = pack_bf16(o[n][2] * i1, o[n][3] * i1);
    }
  }
}

// ---- v3: tcgen05 tensor cores, scores resident in tensor memory -----------------------------------------------------------------
// One CTA per (128-query tile, head, batch item). The sequences of this path are short (ViT 256 / 261 tokens, prefill 256 + prompt),
// so a query tile's WHOLE score row fits in tensor memory next to the output accumulator (<= 384 fp32 score columns + <= 128 output
// columns of the 512): no online softmax, no accumulator rescaling.
//   control warp (warp 8): TMA loads (4-D tensor map over the packed qkv buffer {head_dim, 3*heads, T, B}: rows beyond T and columns
//     beyond head_dim are OUT OF BOUNDS and arrive as zeros, which is what pads head_dim 72 to 80 and the key count to a multiple of
//     64), then S = Q K^T as tcgen05.mma M=128 x N<=256 x K=16 steps (Q and K both K-major, 128-byte swizzle), later O = P V with P as
//     the K-major A operand and V — stored [key][head_dim], i.e. N-contiguous — as an MN-major B operand.
//   warps 0..7: thread (w mod 4, lane) owns query row 32 (w mod 4) + lane (= its TMEM lane) and the column half w / 4 of it (the two
//     halves exchange their row maximum and row sum through shared memory): pass 1 reads the score row (tcgen05.ld) for the row maximum, pass 2 forms
//     p = exp2((s - max) * scale * log2 e) in fp32, sums it in fp32, rounds p to bf16 (as flash-attn) and writes it into shared memory
//     in the swizzled K-major layout the tensor core reads (the K tile is dead by then: P overwrites it); after the PV commit the
//     same thread scales its output row by 1 / sum and stores it.
// V arrives while the scores are computed and soft-maxed. Keys > query (causal) or >= T get p = 0 exactly.
constexpr int ATC_QT = 128;        // queries per CTA = TMEM lanes
constexpr int ATC_MAX_KEYS = 384;  // score columns that fit next to the output accumulator
constexpr int ATC_O_COL = 384;     // TMEM column of the output accumulator
constexpr int ATC_SWARPS = 8;      // softmax / epilogue warps: two per TMEM lane quarter, each taking half of the row's columns
constexpr int ATC_CTRL = ATC_SWARPS;              // warp id of the control warp (TMA + MMA issue + TMEM allocation)
constexpr int ATC_THREADS = (ATC_SWARPS + 1) * 32;

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// MN-major operand (the N index is contiguous in shared memory), 128-byte swizzle: 64-element N blocks `lbo` bytes apart,
// groups of 8 K rows (128 B each) 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_mn128(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3fffu) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int HD>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* __restrict__ out, int T, int heads, int causal, float scale) {
  constexpr int HDP = (HD + 15) / 16 * 16;  // contraction length of Q K^T and N of P V (72 -> 80)
  constexpr int SLABS = (HD + 63) / 64;     // 64-column (128-byte) slabs of a q / k / v row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int keys_pad = (T + 63) / 64 * 64;                        // rows of the K / V tiles in shared memory
  const int kp_bytes = max(SLABS, 2) * keys_pad * 128;            // K tile, later P (keys_pad / 64 slabs of 128 rows x 128 B)
  uint8_t* sQ = smem;                                             // [SLABS][128 rows][128 B]
  uint8_t* sK = sQ + SLABS * ATC_QT * 128;                        // [SLABS][keys_pad rows][128 B]
  uint8_t* sV = sK + kp_bytes;                                    // [SLABS][keys_pad rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + SLABS * keys_pad * 128);
  uint64_t *bar_qk = bars, *bar_v = bars + 1, *bar_s = bars + 2, *bar_p = bars + 3, *bar_o = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);
  float* s_mx = reinterpret_cast<float*>(sQ);  // [2][128] partial row maxima of the two column halves: over the Q tile, dead once S is complete
  float* s_l = s_mx + 2 * ATC_QT;              // [2][128] partial row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATC_QT, head = blockIdx.y, b = blockIdx.z;
  const int Hd = heads * HD;
  const int keys_need = causal ? min(q0 + ATC_QT, T) : T;   // keys any query of this tile attends to
  const int keys_used = (keys_need + 15) / 16 * 16;         // score columns computed (multiple of the MMA N / K granule)
  const int key_boxes = (keys_need + 63) / 64;              // 64-row TMA boxes of K and of V

  if (warp == ATC_CTRL) {
    if (lane == 0) {
      prefetch_tmap(&tm);
      mbar_init(bar_qk, 1), mbar_init(bar_v, 1), mbar_init(bar_s, 1), mbar_init(bar_p, ATC_SWARPS * 32), mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == ATC_CTRL) {
    if (lane == 0) {
      // ---- loads: Q and K first (one barrier), V behind them
      mbar_arrive_expect_tx(bar_qk, static_cast<uint32_t>(SLABS * (2 + key_boxes) * 64 * 128));
      for (int s = 0; s < SLABS; ++s) {
        for (int r = 0; r < 2; ++r) tma_load_4d(sQ + (s * ATC_QT + r * 64) * 128, &tm, s * 64, head, q0 + r * 64, b, bar_qk);
        for (int r = 0; r < key_boxes; ++r) tma_load_4d(sK + (s * keys_pad + r * 64) * 128, &tm, s * 64, heads + head, r * 64, b, bar_qk);
      }
      mbar_arrive_expect_tx(bar_v, static_cast<uint32_t>(SLABS * key_boxes * 64 * 128));
      for (int s = 0; s < SLABS; ++s)
        for (int r = 0; r < key_boxes; ++r) tma_load_4d(sV + (s * keys_pad + r * 64) * 128, &tm, s * 64, 2 * heads + head, r * 64, b, bar_v);

      // ---- S = Q K^T
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      for (int n0 = 0; n0 < keys_used; n0 += 256) {
        const int nn = min(256, keys_used - n0);
        const uint32_t idesc = umma_idesc_bf16(ATC_QT, nn);
#pragma unroll
        for (int k = 0; k < HDP / 16; ++k) {
          const uint32_t qa = smem_u32(sQ) + (k >> 2) * (ATC_QT * 128) + (k & 3) * 32;
          const uint32_t ka = smem_u32(sK) + (k >> 2) * (keys_pad * 128) + n0 * 128 + (k & 3) * 32;
          umma_bf16(tmem_base + n0, umma_desc_k128(qa), umma_desc_k128(ka), idesc, k != 0);
        }
      }
      umma_commit(bar_s);

      // ---- O = P V once the softmax warps have written P (over the K tile) and V has landed
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_pv = umma_idesc_bf16(ATC_QT, HDP) | (1u << 16);  // B operand MN-major
      for (int k = 0; k < keys_used / 16; ++k) {
        const uint32_t pa = smem_u32(sK) + (k >> 2) * (ATC_QT * 128) + (k & 3) * 32;
        const uint32_t va = smem_u32(sV) + k * 2048;
        umma_bf16(tmem_base + ATC_O_COL, umma_desc_k128(pa), umma_desc_mn128(va, keys_pad * 128), idesc_pv, k != 0);
      }
      umma_commit(bar_o);
    }
    __syncwarp();
  } else {
    const int qt = warp & 3, half = warp >> 2;  // TMEM lane quarter (warp id mod 4) and which half of the columns this warp covers
    const int row = qt * 32 + lane, q = q0 + row;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(qt * 32) << 16);
    const int n_valid = causal ? min(q + 1, T) : T;  // keys [0, n_valid) count for this query
    const float sl2 = scale * 1.4426950408889634f;
    const int nch = (keys_used + 31) / 32, ch_split = (nch + 1) / 2;  // 32-column chunks of the score row: [0, ch_split) | [ch_split, nch)
    const int cb = half ? ch_split * 32 : 0, ce = half ? keys_used : min(keys_used, ch_split * 32);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // pass 1: row maximum of the raw scores (own half, then both halves through shared memory)
    float mx = -INFINITY;
    for (int c0 = cb; c0 < ce; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < n_valid) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
    s_mx[half * ATC_QT + row] = mx;
    asm volatile("bar.sync 1, %0;" ::"n"(ATC_SWARPS * 32) : "memory");
    mx = fmaxf(s_mx[row], s_mx[ATC_QT + row]);
    const float m2 = mx * sl2;
    // pass 2: p = exp2(s * sl2 - m2), fp32 row sum, bf16 P into the swizzled K-major A tile (slab = 64 keys, row pitch 128 B)
    float l = 0.f;
    for (int c0 = cb; c0 < ce; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(trow + c0, r);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float p0 = (c0 + 2 * j < n_valid) ? fast_exp2(fmaf(__uint_as_float(r[2 * j]), sl2, -m2)) : 0.f;
        const float p1 = (c0 + 2 * j + 1 < n_valid) ? fast_exp2(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -m2)) : 0.f;
        l += p0 + p1;
        pk[j] = pack_bf16(p0, p1);
      }
      uint8_t* prow = sK + (c0 >> 6) * (ATC_QT * 128) + row * 128;
      const int ch0 = (c0 & 63) >> 3;  // first 16-byte chunk of this 32-key group inside the 64-key slab row: 0 or 4
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(prow + (((ch0 + ch) ^ (row & 7)) << 4)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
    }
    s_l[half * ATC_QT + row] = l;
    fence_proxy_async();  // generic-proxy stores of P -> visible to the tensor core's async-proxy reads
    tc_fence_before();
    mbar_arrive(bar_p);
    asm volatile("bar.sync 1, %0;" ::"n"(ATC_SWARPS * 32) : "memory");  // both partial row sums are visible
    // epilogue: O row / l, own half of the output columns (32-column chunks)
    const float inv = 1.0f / (s_l[row] + s_l[ATC_QT + row]);
    constexpr int NCO = (HDP + 31) / 32, CO_SPLIT = (NCO + 1) / 2;
    mbar_wait(bar_o, 0);
    tc_fence_after();
    __nv_bfloat16* orow = out + (static_cast<long>(b) * T + q) * Hd + head * HD;
#pragma unroll
    for (int cc = 0; cc < CO_SPLIT; ++cc) {
      const int c0 = (half ? CO_SPLIT + cc : cc) * 32;
      if (c0 >= HDP || (half && CO_SPLIT + cc >= NCO)) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32(trow + ATC_O_COL + c0, r);
      tmem_ld_wait();
      if (q < T) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + 8 * j < HD)
            *reinterpret_cast<uint4*>(orow + c0 + 8 * j) =
                make_uint4(pack_bf16(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv),
                           pack_bf16(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv),
                           pack_bf16(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv),
                           pack_bf16(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == ATC_CTRL) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*PFN_encodeTiledA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int HD>
static int launch_attn_tc(const __nv_bfloat16* in, __nv_bfloat16* o, int B, int T, int heads, int causal, float scale, cudaStream_t s) {
  static PFN_encodeTiledA enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_encodeTiledA>(p);
  }
  EMX_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t Hd = static_cast<cuuint64_t>(heads) * HD;
  CUtensorMap tm;
  cuuint64_t dims[4] = {HD, static_cast<cuuint64_t>(3 * heads), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {HD * 2ull, 3 * Hd * 2, static_cast<cuuint64_t>(T) * 3 * Hd * 2};
  cuuint32_t box[4] = {64, 1, 64, 1}, estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(in), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EMX_REQUIRE(r == CUDA_SUCCESS, "emx_attn_fwd: cuTensorMapEncodeTiled failed (%d): B=%d T=%d heads=%d hd=%d", (int)r, B, T, heads, HD);
  constexpr int SLABS = (HD + 63) / 64;
  const int keys_pad = (T + 63) / 64 * 64;
  const int smem = SLABS * ATC_QT * 128 + (SLABS > 2 ? SLABS : 2) * keys_pad * 128 + SLABS * keys_pad * 128 + 64 + 1024;
  bool* attr_set = device_attr_flag(HD == 64 ? ATTR_ATTN_TC64 : HD == 72 ? ATTR_ATTN_TC72 : ATTR_ATTN_TC128);  // per device and instantiation
  if (!attr_set) return -2;
  if (!*attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    *attr_set = true;
  }
  attn_fwd_tc_kernel<HD><<<dim3((T + ATC_QT - 1) / ATC_QT, heads, B), ATC_THREADS, smem, s>>>(tm, o, T, heads, causal, scale);
  return 0;
}

template <int HD>
static int launch_attn_mma(const __nv_bfloat16* in, __nv_bfloat16* o, int B, int T, int heads, int causal, float scale, cudaStream_t s) {
  // 64-query CTAs when that still gives every SM two CTAs, else 32-query CTAs (bs = 1: 16-32 heads x 5 query blocks)
  const long ctas64 = static_cast<long>((T + 63) / 64) * heads * B;
  const int sms = device_sms();
  if (sms < 0) return -2;
  if (ctas64 >= 2L * sms) {
    attn_fwd_mma_kernel<HD, 4><<<dim3((T + 63) / 64, heads, B), 128, 0, s>>>(in, o, T, heads, causal, scale);
  } else {
    attn_fwd_mma_kernel<HD, 2><<<dim3((T + 31) / 32, heads, B), 64, 0, s>>>(in, o, T, heads, causal, scale);
  }
  return 0;
}

}  // namespace emx

extern "C" int emx_attn_fwd(const void* qkv, void* out, int B, int T, int heads, int head_dim, int causal, float scale, cudaStream_t s) {
  using namespace emx;
  EMX_REQUIRE(B > 0 && T > 0 && heads > 0, "emx_attn_fwd: empty problem");
  dim3 grid((T + ATT_QB - 1) / ATT_QB, heads, B);
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  static const bool simt = getenv("EMX_ATTN_SIMT") != nullptr;  // v1 (fp32 SIMT) kept as an A/B reference for the probes
  const char* tce = getenv("EMX_ATTN_TC");                      // A/B switch: 0 = mma.sync v2 for every shape
  const bool tc = !(tce && tce[0] == '0') && T <= ATC_MAX_KEYS && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (!simt && tc) {
    int r = 0;
    switch (head_dim) {
      case 64: r = launch_attn_tc<64>(in, o, B, T, heads, causal, scale, s); break;
      case 72: r = launch_attn_tc<72>(in, o, B, T, heads, causal, scale, s); break;
      case 128: r = launch_attn_tc<128>(in, o, B, T, heads, causal, scale, s); break;
      default: EMX_REQUIRE(false, "emx_attn_fwd: head_dim %d not supported (64, 72, 128)", head_dim);
    }
    if (r) return r;
    EMX_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (!simt) {
    switch (head_dim) {
      case 64: launch_attn_mma<64>(in, o, B, T, heads, causal, scale, s); break;
      case 72: launch_attn_mma<72>(in, o, B, T, heads, causal, scale, s); break;
      case 128: launch_attn_mma<128>(in, o, B, T, heads, causal, scale, s); break;
      default: EMX_REQUIRE(false, "emx_attn_fwd: head_dim %d not supported (64, 72, 128)", head_dim);
    }
    EMX_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  switch (head_dim) {
    case 64: attn_fwd_kernel<64><<<grid, ATT_WARPS * 32, 0, s>>>(in, o, T, heads, causal, scale); break;
    case 72: attn_fwd_kernel<72><<<grid, ATT_WARPS * 32, 0, s>>>(in, o, T, heads, causal, scale); break;
    case 128: attn_fwd_kernel<128><<<grid, ATT_WARPS * 32, 0, s>>>(in, o, T, heads, causal, scale); break;
    default: EMX_REQUIRE(false, "emx_attn_fwd: head_dim %d not supported (64, 72, 128)", head_dim);
  }
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
