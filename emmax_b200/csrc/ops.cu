// Row-wise / element-wise kernels of the ViT + prefill pipeline and the single-token building blocks.
// All are HBM/L2-bound: 16-byte vectorised coalesced loads, warp-shuffle reductions, fp32 math, and the same bf16
// rounding points as the torch-eager reference ops they replace (citations at each entry point in include/emmax.h).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"
#include "emmax.h"

namespace emx {

static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int g_sms[kMaxDevices];
static bool g_attr[kMaxDevices][ATTR_SLOTS];

int device_sms(int* device) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    set_last_error("device_sms: no current CUDA device (or index >= %d)", kMaxDevices);
    return -1;
  }
  if (g_sms[dev] == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
      set_last_error("device_sms: cudaDeviceGetAttribute failed for device %d", dev);
      return -1;
    }
    g_sms[dev] = sms;
  }
  if (device) *device = dev;
  return g_sms[dev];
}

bool* device_attr_flag(int slot) {
  int dev = 0;
  if (device_sms(&dev) < 0) return nullptr;
  return &g_attr[dev][slot];
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

// ---- LayerNorm: y = bf16(((x - mean) * rstd) * w + b), statistics in fp32 over the bf16 inputs ---------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                        const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, int dim,
                                                        float eps) {
  __shared__ float red[8];
  const long row = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * dim);
  const int nv = dim >> 3;
  float s = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    uint4 v = xr[i];
    s += bf16_lo(v.x) + bf16_hi(v.x) + bf16_lo(v.y) + bf16_hi(v.y) + bf16_lo(v.z) + bf16_hi(v.z) + bf16_lo(v.w) + bf16_hi(v.w);
  }
  const float mean = block_sum(s, red) / dim;
  float ss = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    uint4 v = xr[i];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = bf16_lo(u[j]) - mean, c = bf16_hi(u[j]) - mean;
      ss += a * a + c * c;
    }
  }
  const float rstd = 1.0f / sqrtf(block_sum(ss, red) / dim + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const uint4* br = reinterpret_cast<const uint4*>(b);
  uint4* yr = reinterpret_cast<uint4*>(y + row * dim);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    uint4 v = xr[i], ww = wr[i], bb = br[i], o;
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, uw[4] = {ww.x, ww.y, ww.z, ww.w}, ub[4] = {bb.x, bb.y, bb.z, bb.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float lo = (bf16_lo(u[j]) - mean) * rstd * bf16_lo(uw[j]) + bf16_lo(ub[j]);
      float hi = (bf16_hi(u[j]) - mean) * rstd * bf16_hi(uw[j]) + bf16_hi(ub[j]);
      r[j] = pack_bf16(lo, hi);
    }
    o.x = r[0], o.y = r[1], o.z = r[2], o.w = r[3];
    yr[i] = o;
  }
}

// One WARP per row for the ViT widths (dim <= 2048): the row lives in registers (one read of x), the two statistics are warp-shuffle
// reductions, no block barrier; 8 rows per 256-thread block. Same formula as layernorm_kernel (two-pass variance around the mean).
constexpr int LN_MAXV = 8;  // 16-byte vectors per lane: dim <= 32 * 8 * 8 = 2048
__global__ void __launch_bounds__(256) layernorm_warp_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                             const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ y, int rows,
                                                             int dim, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * 8 + warp;
  if (row >= rows) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * dim);
  const int nv = dim >> 3;
  uint4 v[LN_MAXV];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    const int i = lane + 32 * k;
    v[k] = i < nv ? xr[i] : make_uint4(0, 0, 0, 0);
    s += bf16_lo(v[k].x) + bf16_hi(v[k].x) + bf16_lo(v[k].y) + bf16_hi(v[k].y) + bf16_lo(v[k].z) + bf16_hi(v[k].z) + bf16_lo(v[k].w) + bf16_hi(v[k].w);
  }
  const float mean = warp_sum(s) / dim;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    if (lane + 32 * k < nv) {
      const uint32_t u[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(u[j]) - mean, c = bf16_hi(u[j]) - mean;
        ss += a * a + c * c;
      }
    }
  }
  const float rstd = 1.0f / sqrtf(warp_sum(ss) / dim + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  const uint4* br = reinterpret_cast<const uint4*>(b);
  uint4* yr = reinterpret_cast<uint4*>(y + row * dim);
#pragma unroll
  for (int k = 0; k < LN_MAXV; ++k) {
    const int i = lane + 32 * k;
    if (i < nv) {
      const uint4 ww = wr[i], bb = br[i];
      const uint32_t u[4] = {v[k].x, v[k].y, v[k].z, v[k].w}, uw[4] = {ww.x, ww.y, ww.z, ww.w}, ub[4] = {bb.x, bb.y, bb.z, bb.w};
      uint32_t r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = (bf16_lo(u[j]) - mean) * rstd * bf16_lo(uw[j]) + bf16_lo(ub[j]);
        const float hi = (bf16_hi(u[j]) - mean) * rstd * bf16_hi(uw[j]) + bf16_hi(ub[j]);
        r[j] = pack_bf16(lo, hi);
      }
      yr[i] = make_uint4(r[0], r[1], r[2], r[3]);
    }
  }
}

// ---- RMSNorm (LlamaRMSNorm): n = bf16(x * rsqrt(mean(x^2) + eps)); y = bf16(w * n) ----------------------------------
__global__ void __launch_bounds__(256) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                                      __nv_bfloat16* __restrict__ y, int dim, float eps) {
  __shared__ float red[8];
  const long row = blockIdx.x;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * dim);
  const int nv = dim >> 3;
  float ss = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    uint4 v = xr[i];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = bf16_lo(u[j]), c = bf16_hi(u[j]);
      ss += a * a + c * c;
    }
  }
  const float rs = 1.0f / sqrtf(block_sum(ss, red) / dim + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* yr = reinterpret_cast<uint4*>(y + row * dim);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    uint4 v = xr[i], ww = wr[i], o;
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, uw[4] = {ww.x, ww.y, ww.z, ww.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      r[j] = pack_bf16(bf16_lo(uw[j]) * bf16_round(bf16_lo(u[j]) * rs), bf16_hi(uw[j]) * bf16_round(bf16_hi(u[j]) * rs));
    o.x = r[0], o.y = r[1], o.z = r[2], o.w = r[3];
    yr[i] = o;
  }
}

// ---- patch im2col: out[(b*gh+py)*gw+px, c*P*P + ky*P + kx] = pixels[b, chan0+c, py*P+ky, px*P+kx]; zero pad to kpad ----
__global__ void patch_im2col_kernel(const __nv_bfloat16* __restrict__ pix, int c_total, int chan0, int H, int W, int P,
                                    __nv_bfloat16* __restrict__ out, int kpad) {
  const int gw = W / P, gh = H / P;
  const int row = blockIdx.x;  // b*gh*gw + py*gw + px
  const int b = row / (gh * gw), pp = row % (gh * gw), py = pp / gw, px = pp % gw;
  const int kk = 3 * P * P;
  for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
    __nv_bfloat16 v = __float2bfloat16_rn(0.f);
    if (k < kk) {
      const int c = k / (P * P), r = k % (P * P), ky = r / P, kx = r % P;
      v = pix[((static_cast<long>(b) * c_total + chan0 + c) * H + py * P + ky) * W + px * P + kx];
    }
    out[static_cast<long>(row) * kpad + k] = v;
  }
}

__global__ void vit_assemble_kernel(const __nv_bfloat16* __restrict__ patch_out, const __nv_bfloat16* __restrict__ pos,
                                    const __nv_bfloat16* __restrict__ prefix_tokens, __nv_bfloat16* __restrict__ tokens, int n_patches,
                                    int prefix, int D) {
  const int T = n_patches + prefix;
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  __nv_bfloat16* dst = tokens + static_cast<long>(blockIdx.x) * D;
  if (t < prefix) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] = prefix_tokens[t * D + d];
  } else {
    const int i = t - prefix;
    const __nv_bfloat16* src = patch_out + (static_cast<long>(b) * n_patches + i) * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) dst[d] = __float2bfloat16_rn(ld_bf16(src + d) + ld_bf16(pos + i * D + d));
  }
}

__global__ void vit_gather_kernel(const __nv_bfloat16* __restrict__ tokens, __nv_bfloat16* __restrict__ feat, int n_patches, int prefix,
                                  int D, int ldf, int col0) {
  const int b = blockIdx.x / n_patches, i = blockIdx.x % n_patches;
  const uint4* src = reinterpret_cast<const uint4*>(tokens + (static_cast<long>(b) * (n_patches + prefix) + prefix + i) * D);
  uint4* dst = reinterpret_cast<uint4*>(feat + static_cast<long>(blockIdx.x) * ldf + col0);
  for (int d = threadIdx.x; d < (D >> 3); d += blockDim.x) dst[d] = src[d];
}

// ---- RoPE + KV store -------------------------------------------------------------------------------------------------
// q_embed = bf16(bf16(q*cos) + bf16(rotate_half(q)*sin)) with bf16 cos/sin (transformers apply_rotary_pos_emb in bf16)
__global__ void rope_kvstore_kernel(__nv_bfloat16* __restrict__ qkv, int T, int heads, int hd, const __nv_bfloat16* __restrict__ cos_tab,
                                    const __nv_bfloat16* __restrict__ sin_tab, int pos0, __nv_bfloat16* __restrict__ k_cache,
                                    __nv_bfloat16* __restrict__ v_cache, const int32_t* __restrict__ block_table, int max_pages,
                                    int page_size) {
  const int row = blockIdx.x;  // b*T + t
  const int b = row / T, t = row % T, pos = pos0 + t;
  const int Hd = heads * hd, half = hd >> 1;
  __nv_bfloat16* q = qkv + static_cast<long>(row) * 3 * Hd;
  __nv_bfloat16* k = q + Hd;
  const __nv_bfloat16* v = q + 2 * Hd;
  const int page = block_table[b * max_pages + pos / page_size], slot = pos % page_size;
  if ((half & 7) == 0) {
    // 8 rotation pairs per thread: 16-byte loads / stores of both halves of q and k, of the bf16 cos / sin rows and of v
    const int hv = half >> 3;
    for (int i = threadIdx.x; i < heads * hv; i += blockDim.x) {
      const int h = i / hv, j = (i % hv) << 3;
      const uint4 cv = *reinterpret_cast<const uint4*>(cos_tab + static_cast<long>(pos) * half + j);
      const uint4 sv = *reinterpret_cast<const uint4*>(sin_tab + static_cast<long>(pos) * half + j);
      const int a = h * hd + j, bidx = a + half;
      const long dst = ((static_cast<long>(page) * heads + h) * page_size + slot) * hd + j;
      const uint32_t cw[4] = {cv.x, cv.y, cv.z, cv.w}, sw[4] = {sv.x, sv.y, sv.z, sv.w};
      auto rotate = [&](const uint4 lo, const uint4 hi, uint4& olo, uint4& ohi) {
        const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w}, u[4] = {hi.x, hi.y, hi.z, hi.w};
        uint32_t ol[4], ou[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float c0 = bf16_lo(cw[e]), c1 = bf16_hi(cw[e]), s0 = bf16_lo(sw[e]), s1 = bf16_hi(sw[e]);
          const float x10 = bf16_lo(l[e]), x11 = bf16_hi(l[e]), x20 = bf16_lo(u[e]), x21 = bf16_hi(u[e]);
          ol[e] = pack_bf16(bf16_round(x10 * c0) + bf16_round(-x20 * s0), bf16_round(x11 * c1) + bf16_round(-x21 * s1));
          ou[e] = pack_bf16(bf16_round(x20 * c0) + bf16_round(x10 * s0), bf16_round(x21 * c1) + bf16_round(x11 * s1));
        }
        olo = make_uint4(ol[0], ol[1], ol[2], ol[3]), ohi = make_uint4(ou[0], ou[1], ou[2], ou[3]);
      };
      uint4 r1, r2;
      rotate(*reinterpret_cast<const uint4*>(q + a), *reinterpret_cast<const uint4*>(q + bidx), r1, r2);
      *reinterpret_cast<uint4*>(q + a) = r1, *reinterpret_cast<uint4*>(q + bidx) = r2;
      rotate(*reinterpret_cast<const uint4*>(k + a), *reinterpret_cast<const uint4*>(k + bidx), r1, r2);
      *reinterpret_cast<uint4*>(k + a) = r1, *reinterpret_cast<uint4*>(k + bidx) = r2;
      *reinterpret_cast<uint4*>(k_cache + dst) = r1, *reinterpret_cast<uint4*>(k_cache + dst + half) = r2;
      *reinterpret_cast<uint4*>(v_cache + dst) = *reinterpret_cast<const uint4*>(v + a);
      *reinterpret_cast<uint4*>(v_cache + dst + half) = *reinterpret_cast<const uint4*>(v + bidx);
    }
    return;
  }
  for (int i = threadIdx.x; i < heads * half; i += blockDim.x) {
    const int h = i / half, j = i % half;
    const float c = ld_bf16(cos_tab + static_cast<long>(pos) * half + j), s = ld_bf16(sin_tab + static_cast<long>(pos) * half + j);
    const int a = h * hd + j, bidx = a + half;
    const long dst = ((static_cast<long>(page) * heads + h) * page_size + slot) * hd + j;
    {
      const float x1 = ld_bf16(q + a), x2 = ld_bf16(q + bidx);
      q[a] = __float2bfloat16_rn(bf16_round(x1 * c) + bf16_round(-x2 * s));
      q[bidx] = __float2bfloat16_rn(bf16_round(x2 * c) + bf16_round(x1 * s));
    }
    {
      const float x1 = ld_bf16(k + a), x2 = ld_bf16(k + bidx);
      const __nv_bfloat16 r1 = __float2bfloat16_rn(bf16_round(x1 * c) + bf16_round(-x2 * s));
      const __nv_bfloat16 r2 = __float2bfloat16_rn(bf16_round(x2 * c) + bf16_round(x1 * s));
      k[a] = r1, k[bidx] = r2;
      k_cache[dst] = r1, k_cache[dst + half] = r2;
    }
    v_cache[dst] = v[a], v_cache[dst + half] = v[bidx];
  }
}

__global__ void embed_assemble_kernel(const int64_t* __restrict__ ids, int n_ids, const __nv_bfloat16* __restrict__ embed,
                                      const __nv_bfloat16* __restrict__ patches, int n_patches, __nv_bfloat16* __restrict__ x, int H) {
  const int S = n_ids + n_patches;
  const int b = blockIdx.x / S, s = blockIdx.x % S;
  const uint4* src;
  if (s == 0)
    src = reinterpret_cast<const uint4*>(embed + ids[static_cast<long>(b) * n_ids] * H);
  else if (s <= n_patches)
    src = reinterpret_cast<const uint4*>(patches + (static_cast<long>(b) * n_patches + s - 1) * H);
  else
    src = reinterpret_cast<const uint4*>(embed + ids[static_cast<long>(b) * n_ids + s - n_patches] * H);
  uint4* dst = reinterpret_cast<uint4*>(x + static_cast<long>(blockIdx.x) * H);
  for (int d = threadIdx.x; d < (H >> 3); d += blockDim.x) dst[d] = src[d];
}

__global__ void swiglu_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ h, long total) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint32_t p = reinterpret_cast<const uint32_t*>(gu)[i];
  h[i] = __float2bfloat16_rn(bf16_round(silu(bf16_lo(p))) * bf16_hi(p));
}

// ---- GEMV building block: one warp per output row, 16-B streaming weight loads ------------------------------------
__global__ void __launch_bounds__(256) gemv_kernel(const __nv_bfloat16* __restrict__ W, int ldw, const __nv_bfloat16* __restrict__ x,
                                                   __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ resid,
                                                   float* __restrict__ y32, int N, int K) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  if (n >= N) return;
  const uint4* wr = reinterpret_cast<const uint4*>(W + static_cast<long>(n) * ldw);
  const uint4* xr = reinterpret_cast<const uint4*>(x);
  float acc = 0.f;
  const int nv = K >> 3;
#pragma unroll 4
  for (int i = lane; i < nv; i += 32) {
    const uint4 w = ldg_nc_v4(wr + i), xv = xr[i];
    acc = fmaf(bf16_lo(w.x), bf16_lo(xv.x), acc), acc = fmaf(bf16_hi(w.x), bf16_hi(xv.x), acc);
    acc = fmaf(bf16_lo(w.y), bf16_lo(xv.y), acc), acc = fmaf(bf16_hi(w.y), bf16_hi(xv.y), acc);
    acc = fmaf(bf16_lo(w.z), bf16_lo(xv.z), acc), acc = fmaf(bf16_hi(w.z), bf16_hi(xv.z), acc);
    acc = fmaf(bf16_lo(w.w), bf16_lo(xv.w), acc), acc = fmaf(bf16_hi(w.w), bf16_hi(xv.w), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float o = bf16_round(acc);
    if (y32) y32[n] = o;
    if (y) {
      if (resid) o = bf16_round(o + ld_bf16(resid + n));
      y[n] = __float2bfloat16_rn(o);
    }
  }
}

// single-CTA argmax, lowest index wins ties (torch.argmax / GenerationMixin greedy)
__global__ void __launch_bounds__(1024) argmax_kernel(const float* __restrict__ v, int n, int32_t* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int si[32];
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = v[i];
    if (x > best || (x == best && i < bi)) best = x, bi = i;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
  }
  if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = best, si[threadIdx.x >> 5] = bi;
  __syncthreads();
  if (threadIdx.x < 32) {
    best = (threadIdx.x < (blockDim.x >> 5)) ? sv[threadIdx.x] : -INFINITY;
    bi = (threadIdx.x < (blockDim.x >> 5)) ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
    }
    if (threadIdx.x == 0) *out = bi;
  }
}

// one CTA per row of a bf16 logits matrix: optional fp32 copy of the row + argmax, lowest index wins ties (first tokens of a batched prefill)
__global__ void __launch_bounds__(1024) argmax_rows_bf16_kernel(const __nv_bfloat16* __restrict__ v, int ld, int n, float* __restrict__ out32,
                                                                int32_t* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const __nv_bfloat16* row = v + static_cast<long>(blockIdx.x) * ld;
  float* o32 = out32 ? out32 + static_cast<long>(blockIdx.x) * n : nullptr;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = ld_bf16(row + i);
    if (o32) o32[i] = x;
    if (x > best) best = x, bi = i;  // indices ascend per thread: strict '>' keeps the lowest
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
  }
  if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = best, si[threadIdx.x >> 5] = bi;
  __syncthreads();
  if (threadIdx.x < 32) {
    best = sv[threadIdx.x], bi = si[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) best = ob, bi = oi;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = bi;
  }
}

// ---- action de-tokeniser: integer index math + fp64 table, bit-exact with numpy ----------------------------------
__global__ void detok_kernel(const int32_t* __restrict__ ids, int n, int vocab, int n_bins, const double* __restrict__ q01,
                             const double* __restrict__ q99, const uint8_t* __restrict__ mask, int adim, double* __restrict__ norm,
                             double* __restrict__ act) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int k = vocab - ids[i] - 1;
  k = min(max(k, 0), n_bins - 2);
  // np.linspace(-1, 1, n_bins)[k] = -1 + k*step (last point pinned to stop), centres = (e[k] + e[k+1]) / 2
  const double step = 2.0 / static_cast<double>(n_bins - 1);
  const double e0 = (k == n_bins - 1) ? 1.0 : __dadd_rn(-1.0, __dmul_rn(static_cast<double>(k), step));
  const double e1 = (k + 1 == n_bins - 1) ? 1.0 : __dadd_rn(-1.0, __dmul_rn(static_cast<double>(k + 1), step));
  const double c = __dmul_rn(__dadd_rn(e0, e1), 0.5);
  norm[i] = c;
  if (act) {
    const int d = i % adim;
    double a = c;
    if (mask == nullptr || mask[d]) {
      // 0.5 * (c + 1) * (q99 - q01) + q01, evaluated left to right as numpy does, no FMA contraction
      a = __dadd_rn(__dmul_rn(__dmul_rn(0.5, __dadd_rn(c, 1.0)), __dadd_rn(q99[d], -q01[d])), q01[d]);
    }
    act[i] = a;
  }
}


// uint8 HWC frame -> the [6, H, W] bf16 tensor the two ViT towers eat, for frames that already have the model's input size (the robot
// path pre-resizes, experiments/robot/bridge/bridgev2_utils.py:152-166; `resize-naive` on a 224x224 input is the identity).
// Per backbone b (DINOv2 first, then SigLIP), channel c: torchvision's to_tensor + normalize in fp32, op by op, then the bf16 cast of
// `.to(device, dtype=bfloat16)`: bf16(((float(u8) / 255) - mean[b][c]) / std[b][c])  — processing_prismatic.py:128-145.
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ hwc, __nv_bfloat16* __restrict__ out, int HW, int n_backbones,
                                     const float* __restrict__ mean, const float* __restrict__ stdv) {
  const int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const uint8_t* px = hwc + (static_cast<long>(b) * HW + i) * 3;
    for (int k = 0; k < n_backbones; ++k) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(px[c]), 255.0f), mean[3 * k + c]), stdv[3 * k + c]);
        out[(static_cast<long>(b) * 3 * n_backbones + 3 * k + c) * HW + i] = __float2bfloat16_rn(v);
      }
    }
  }
}

// PIL's antialiased resample (Pillow src/libImaging/Resample.c, ImagingResampleHorizontal/Vertical_8bpc) — what torchvision's
// `resize(PIL image, interpolation=BICUBIC, antialias=True)` runs at processing_prismatic.py:133 — restated in its own integer
// arithmetic so the GPU result is bit-identical: two separable passes with a uint8 intermediate, 22-bit fixed-point coefficients
// (computed on the host exactly as precompute_coeffs / normalize_coeffs_8bpc do), accumulator seeded with 1 << 21, arithmetic shift,
// clamp to [0, 255].
__global__ void resample_h_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int rows, int Win, int Wout,
                                     const int32_t* __restrict__ kk, const int32_t* __restrict__ bounds, int ksize) {
  // one thread per (row, output column); 3 interleaved channels
  const long n = static_cast<long>(rows) * Wout;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int xx = static_cast<int>(i % Wout);
    const long r = i / Wout;
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int32_t* k = kk + static_cast<long>(xx) * ksize;
    const uint8_t* px = in + (r * Win + xmin) * 3;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    for (int x = 0; x < xmax; ++x) {
      const int w = k[x];
      s0 += px[3 * x] * w, s1 += px[3 * x + 1] * w, s2 += px[3 * x + 2] * w;
    }
    uint8_t* o = out + i * 3;
    o[0] = static_cast<uint8_t>(min(max(s0 >> 22, 0), 255));
    o[1] = static_cast<uint8_t>(min(max(s1 >> 22, 0), 255));
    o[2] = static_cast<uint8_t>(min(max(s2 >> 22, 0), 255));
  }
}
// vertical pass fused with to_tensor + per-backbone normalize + bf16 cast (see preprocess_u8_kernel)
__global__ void resample_v_norm_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out, int Hin, int Hout, int W,
                                       const int32_t* __restrict__ kk, const int32_t* __restrict__ bounds, int ksize, int n_backbones,
                                       const float* __restrict__ mean, const float* __restrict__ stdv) {
  const int b = blockIdx.y;
  const int HW = Hout * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const int yy = i / W, x = i % W;
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const int32_t* k = kk + static_cast<long>(yy) * ksize;
    const uint8_t* px = in + ((static_cast<long>(b) * Hin + ymin) * W + x) * 3;
    int s[3] = {1 << 21, 1 << 21, 1 << 21};
    for (int y = 0; y < ymax; ++y) {
      const int w = k[y];
      const uint8_t* q = px + static_cast<long>(y) * W * 3;
      s[0] += q[0] * w, s[1] += q[1] * w, s[2] += q[2] * w;
    }
    for (int kb = 0; kb < n_backbones; ++kb) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float u8 = static_cast<float>(min(max(s[c] >> 22, 0), 255));
        const float v = __fdiv_rn(__fsub_rn(__fdiv_rn(u8, 255.0f), mean[3 * kb + c]), stdv[3 * kb + c]);
        out[(static_cast<long>(b) * 3 * n_backbones + 3 * kb + c) * HW + i] = __float2bfloat16_rn(v);
      }
    }
  }
}

}  // namespace emx

using namespace emx;

extern "C" const char* emx_last_error(void) { return g_err; }
extern "C" int emx_abi_version(void) { return 5; }
extern "C" const char* emx_arch(void) { return "sm_100a"; }

#define BF(p) static_cast<const __nv_bfloat16*>(p)
#define BFM(p) static_cast<__nv_bfloat16*>(p)

extern "C" int emx_layernorm(const void* x, const void* w, const void* b, void* y, int rows, int dim, float eps, cudaStream_t s) {
  EMX_REQUIRE(rows > 0 && dim % 8 == 0, "emx_layernorm: rows=%d dim=%d (dim must be a multiple of 8)", rows, dim);
  if (dim <= 32 * 8 * LN_MAXV)
    layernorm_warp_kernel<<<(rows + 7) / 8, 256, 0, s>>>(BF(x), BF(w), BF(b), BFM(y), rows, dim, eps);
  else
    layernorm_kernel<<<rows, 256, 0, s>>>(BF(x), BF(w), BF(b), BFM(y), dim, eps);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_rmsnorm(const void* x, const void* w, void* y, int rows, int dim, float eps, cudaStream_t s) {
  EMX_REQUIRE(rows > 0 && dim % 8 == 0, "emx_rmsnorm: rows=%d dim=%d (dim must be a multiple of 8)", rows, dim);
  rmsnorm_kernel<<<rows, 256, 0, s>>>(BF(x), BF(w), BFM(y), dim, eps);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_resize_preprocess_u8(const void* hwc, int B, int Hin, int Win, int Hout, int Wout, const int32_t* kk_h,
                                        const int32_t* bounds_h, int ksize_h, const int32_t* kk_v, const int32_t* bounds_v, int ksize_v,
                                        void* tmp, int n_backbones, const float* mean, const float* stdv, void* out, cudaStream_t s) {
  EMX_REQUIRE(B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && n_backbones >= 1 && n_backbones <= 2,
              "emx_resize_preprocess_u8: B=%d in=%dx%d out=%dx%d backbones=%d", B, Hin, Win, Hout, Wout, n_backbones);
  EMX_REQUIRE(hwc && kk_h && bounds_h && kk_v && bounds_v && tmp && mean && stdv && out, "emx_resize_preprocess_u8: null pointer");
  const long n = static_cast<long>(B) * Hin * Wout;
  resample_h_u8_kernel<<<static_cast<unsigned>(min((n + 255) / 256, 4096L)), 256, 0, s>>>(static_cast<const uint8_t*>(hwc), static_cast<uint8_t*>(tmp),
                                                                                          B * Hin, Win, Wout, kk_h, bounds_h, ksize_h);
  resample_v_norm_kernel<<<dim3((Hout * Wout + 255) / 256, B), 256, 0, s>>>(static_cast<const uint8_t*>(tmp), BFM(out), Hin, Hout, Wout, kk_v, bounds_v,
                                                                          ksize_v, n_backbones, mean, stdv);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_preprocess_u8(const void* hwc, int B, int H, int W, int n_backbones, const float* mean, const float* stdv, void* out,
                                 cudaStream_t s) {
  EMX_REQUIRE(B > 0 && H > 0 && W > 0 && n_backbones >= 1 && n_backbones <= 2, "emx_preprocess_u8: B=%d H=%d W=%d backbones=%d", B, H, W, n_backbones);
  EMX_REQUIRE(hwc && mean && stdv && out, "emx_preprocess_u8: null pointer");
  preprocess_u8_kernel<<<dim3((H * W + 255) / 256, B), 256, 0, s>>>(static_cast<const uint8_t*>(hwc), BFM(out), H * W, n_backbones, mean, stdv);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_patch_im2col(const void* pixels, int B, int c_total, int chan0, int H, int W, int patch, void* out, int kpad,
                                cudaStream_t s) {
  EMX_REQUIRE(B > 0 && H % patch == 0 && W % patch == 0 && kpad >= 3 * patch * patch, "emx_patch_im2col: bad geometry");
  patch_im2col_kernel<<<B * (H / patch) * (W / patch), 256, 0, s>>>(BF(pixels), c_total, chan0, H, W, patch, BFM(out), kpad);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_vit_assemble(const void* patch_out, const void* pos, const void* prefix_tokens, void* tokens, int B, int n_patches,
                                int prefix, int D, cudaStream_t s) {
  EMX_REQUIRE(B > 0 && (prefix == 0 || prefix_tokens), "emx_vit_assemble: prefix tokens missing");
  vit_assemble_kernel<<<B * (n_patches + prefix), 256, 0, s>>>(BF(patch_out), BF(pos), BF(prefix_tokens), BFM(tokens), n_patches, prefix, D);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_vit_gather_features(const void* tokens, void* features, int B, int n_patches, int prefix, int D, int ldf, int col0,
                                       cudaStream_t s) {
  EMX_REQUIRE(D % 8 == 0 && ldf % 8 == 0 && col0 % 8 == 0, "emx_vit_gather_features: D/ld/col0 must be multiples of 8");
  vit_gather_kernel<<<B * n_patches, 128, 0, s>>>(BF(tokens), BFM(features), n_patches, prefix, D, ldf, col0);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_rope_kvstore(void* qkv, int B, int T, int heads, int hd, const void* cos_tab, const void* sin_tab, int pos0,
                                void* k_cache, void* v_cache, const int32_t* block_table, int max_pages, int page_size, cudaStream_t s) {
  EMX_REQUIRE(B > 0 && T > 0 && hd % 2 == 0, "emx_rope_kvstore: bad shape");
  rope_kvstore_kernel<<<B * T, 256, 0, s>>>(BFM(qkv), T, heads, hd, BF(cos_tab), BF(sin_tab), pos0, BFM(k_cache), BFM(v_cache),
                                            block_table, max_pages, page_size);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_embed_assemble(const int64_t* ids, int n_ids, const void* embed, const void* patches, int n_patches, void* x, int B,
                                  int H, cudaStream_t s) {
  EMX_REQUIRE(B > 0 && n_ids >= 1 && H % 8 == 0, "emx_embed_assemble: bad shape");
  embed_assemble_kernel<<<B * (n_ids + n_patches), 256, 0, s>>>(ids, n_ids, BF(embed), BF(patches), n_patches, BFM(x), H);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_swiglu(const void* gate_up, void* h, int M, int I, cudaStream_t s) {
  const long total = static_cast<long>(M) * I;
  swiglu_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(BF(gate_up), BFM(h), total);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_gemv_bf16(const void* W, int ldw, const void* x, void* y, const void* resid, int N, int K, cudaStream_t s) {
  EMX_REQUIRE(K % 8 == 0 && ldw % 8 == 0, "emx_gemv_bf16: K and ldw must be multiples of 8");
  gemv_kernel<<<(N + 7) / 8, 256, 0, s>>>(BF(W), ldw, BF(x), BFM(y), BF(resid), nullptr, N, K);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_lmhead_argmax(const void* W, int ldw, const void* x, int N, int K, float* logits_out, int32_t* token_out, void* scratch,
                                 cudaStream_t s) {
  EMX_REQUIRE(K % 8 == 0 && ldw % 8 == 0, "emx_lmhead_argmax: K and ldw must be multiples of 8");
  float* logits = logits_out ? logits_out : static_cast<float*>(scratch);
  EMX_REQUIRE(logits != nullptr, "emx_lmhead_argmax: need logits_out or scratch (N floats)");
  gemv_kernel<<<(N + 7) / 8, 256, 0, s>>>(BF(W), ldw, BF(x), nullptr, nullptr, logits, N, K);
  argmax_kernel<<<1, 1024, 0, s>>>(logits, N, token_out);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_argmax_rows_bf16(const void* logits, int ld, int rows, int n, float* logits_out, int32_t* tokens_out, cudaStream_t s) {
  EMX_REQUIRE(logits && tokens_out && rows >= 1 && n >= 1 && ld >= n, "emx_argmax_rows_bf16: bad arguments");
  argmax_rows_bf16_kernel<<<rows, 1024, 0, s>>>(BF(logits), ld, n, logits_out, tokens_out);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
extern "C" int emx_detokenize_actions(const int32_t* ids, int n, int vocab_size, int n_bins, const double* q01, const double* q99,
                                      const uint8_t* mask, int action_dim, double* normalized, double* actions, cudaStream_t s) {
  EMX_REQUIRE(n > 0 && n_bins >= 2 && normalized, "emx_detokenize_actions: bad arguments");
  EMX_REQUIRE(actions == nullptr || (q01 && q99 && action_dim > 0), "emx_detokenize_actions: stats required for un-normalisation");
  detok_kernel<<<(n + 127) / 128, 128, 0, s>>>(ids, n, vocab_size, n_bins, q01, q99, mask, action_dim, normalized, actions);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
