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
// v1 mapping (sequences here are <= ~300 tokens, so attention is < 2 % of the request's FLOPs): one CTA per
// (16-query block, head, batch); K/V tiles of 64 keys staged in padded shared memory (odd word stride ->
// conflict-free column reads); in the score phase a lane owns a key, in the PV phase a lane owns output dims.
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

}  // namespace emx

extern "C" int emx_attn_fwd(const void* qkv, void* out, int B, int T, int heads, int head_dim, int causal, float scale, cudaStream_t s) {
  using namespace emx;
  EMX_REQUIRE(B > 0 && T > 0 && heads > 0, "emx_attn_fwd: empty problem");
  dim3 grid((T + ATT_QB - 1) / ATT_QB, heads, B);
  const __nv_bfloat16* in = static_cast<const __nv_bfloat16*>(qkv);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  switch (head_dim) {
    case 64: attn_fwd_kernel<64><<<grid, ATT_WARPS * 32, 0, s>>>(in, o, T, heads, causal, scale); break;
    case 72: attn_fwd_kernel<72><<<grid, ATT_WARPS * 32, 0, s>>>(in, o, T, heads, causal, scale); break;
    case 128: attn_fwd_kernel<128><<<grid, ATT_WARPS * 32, 0, s>>>(in, o, T, heads, causal, scale); break;
    default: EMX_REQUIRE(false, "emx_attn_fwd: head_dim %d not supported (64, 72, 128)", head_dim);
  }
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
