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
