// GPU twins of the robot loop's TensorFlow image steps, bit-exact with the numpy float32 restatements in
// emmax_b200/robot_utils.py (TensorFlow itself is absent from this image, so those restatements follow TensorFlow's published kernels:
// crop_and_resize_op.cc, scale_and_translate_op.cc, convert_image_dtype):
//   emx_crop_resize_u8      the `center_crop` branch of get_vla_action / get_seq_action
//                           (/root/reference/experiments/robot/openvla_utils.py:81-124, :136-156)
//   emx_lanczos3_resize_u8  `resize_image`'s tf.image.resize(method="lanczos3", antialias=True) + round + clip + uint8
//                           (/root/reference/experiments/robot/bridge/bridgev2_utils.py:152-166)
// Every float operation is an explicit round-to-nearest intrinsic in the order the host twin evaluates it (no FMA contraction).
#include "common.cuh"
#include "emmax.h"

namespace emx {

struct CropTap {
  int lo, hi;
  float lerp;
  bool inside;
};
// tf.image.crop_and_resize: pos = a * (n - 1) + i * scale, scale = (b - a) * (n - 1) / (m - 1)   (m > 1)
__device__ __forceinline__ CropTap crop_tap(int i, int n, int m, float a, float b) {
  const float nm1 = static_cast<float>(n - 1);
  float pos;
  if (m > 1) {
    const float scale = __fdiv_rn(__fmul_rn(__fsub_rn(b, a), nm1), static_cast<float>(m - 1));
    pos = __fadd_rn(__fmul_rn(a, nm1), __fmul_rn(static_cast<float>(i), scale));
  } else {
    pos = __fmul_rn(__fmul_rn(0.5f, __fadd_rn(a, b)), nm1);
  }
  CropTap t;
  t.inside = pos >= 0.f && pos <= nm1;
  const float lo = floorf(pos);
  t.lerp = __fsub_rn(pos, lo);
  t.lo = min(max(static_cast<int>(lo), 0), n - 1);
  t.hi = min(max(static_cast<int>(ceilf(pos)), 0), n - 1);
  return t;
}

__global__ void crop_resize_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W, int Ho, int Wo, float y1,
                                      float x1, float y2, float x2) {
  const int b = blockIdx.y;
  const float inv255 = __fdiv_rn(1.0f, 255.0f);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Ho * Wo; i += gridDim.x * blockDim.x) {
    const int yy = i / Wo, xx = i % Wo;
    const CropTap ty = crop_tap(yy, H, Ho, y1, y2), tx = crop_tap(xx, W, Wo, x1, x2);
    const uint8_t* img = in + static_cast<long>(b) * H * W * 3;
    uint8_t* o = out + (static_cast<long>(b) * Ho * Wo + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = 0.f;
      if (ty.inside && tx.inside) {
        const float tl = __fmul_rn(static_cast<float>(img[(ty.lo * W + tx.lo) * 3 + c]), inv255);
        const float tr = __fmul_rn(static_cast<float>(img[(ty.lo * W + tx.hi) * 3 + c]), inv255);
        const float bl = __fmul_rn(static_cast<float>(img[(ty.hi * W + tx.lo) * 3 + c]), inv255);
        const float br = __fmul_rn(static_cast<float>(img[(ty.hi * W + tx.hi) * 3 + c]), inv255);
        const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), tx.lerp));
        const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), tx.lerp));
        v = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), ty.lerp));
      }
      v = fminf(fmaxf(v, 0.f), 1.f);                                  // tf.clip_by_value(image, 0, 1)
      v = fminf(fmaxf(__fmul_rn(v, 255.5f), 0.f), 255.f);             // convert_image_dtype(float -> uint8, saturate=True): * (max + 0.5)
      o[c] = static_cast<uint8_t>(v);                                 // truncating cast
    }
  }
}

// horizontal pass: tmp[r][x][c] = sum_k in[r][min(start[x] + k, W - 1)][c] * w[x][k], accumulated in k order
__global__ void lanczos_h_kernel(const uint8_t* __restrict__ in, float* __restrict__ tmp, int H, int W, int Wo, const int32_t* __restrict__ start,
                                 const float* __restrict__ wt, int ks) {
  const long n = static_cast<long>(H) * Wo;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % Wo);
    const long r = i / Wo;
    const float* w = wt + static_cast<long>(x) * ks;
    const int s = start[x];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = 0; k < ks; ++k) {
      const uint8_t* px = in + (r * W + min(s + k, W - 1)) * 3;
      a0 = __fadd_rn(a0, __fmul_rn(static_cast<float>(px[0]), w[k]));
      a1 = __fadd_rn(a1, __fmul_rn(static_cast<float>(px[1]), w[k]));
      a2 = __fadd_rn(a2, __fmul_rn(static_cast<float>(px[2]), w[k]));
    }
    tmp[i * 3] = a0, tmp[i * 3 + 1] = a1, tmp[i * 3 + 2] = a2;
  }
}
// vertical pass + tf.round (half to even) + clip + uint8
__global__ void lanczos_v_kernel(const float* __restrict__ tmp, uint8_t* __restrict__ out, int H, int Ho, int Wo, const int32_t* __restrict__ start,
                                 const float* __restrict__ wt, int ks) {
  const long n = static_cast<long>(Ho) * Wo * 3;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long col = i % (static_cast<long>(Wo) * 3);
    const int y = static_cast<int>(i / (static_cast<long>(Wo) * 3));
    const float* w = wt + static_cast<long>(y) * ks;
    const int s = start[y];
    float a = 0.f;
    for (int k = 0; k < ks; ++k) a = __fadd_rn(a, __fmul_rn(tmp[static_cast<long>(min(s + k, H - 1)) * Wo * 3 + col], w[k]));
    out[i] = static_cast<uint8_t>(fminf(fmaxf(rintf(a), 0.f), 255.f));
  }
}

}  // namespace emx

using namespace emx;

extern "C" int emx_crop_resize_u8(const void* hwc, int B, int H, int W, float y1, float x1, float y2, float x2, void* out, int Ho, int Wo,
                                  cudaStream_t s) {
  EMX_REQUIRE(hwc && out && B > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "emx_crop_resize_u8: empty problem");
  dim3 grid((Ho * Wo + 255) / 256, B);
  crop_resize_u8_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(hwc), static_cast<uint8_t*>(out), H, W, Ho, Wo, y1, x1, y2, x2);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int emx_lanczos3_resize_u8(const void* hwc, int H, int W, int Ho, int Wo, const int32_t* start_h, const float* w_h, int ks_h,
                                      const int32_t* start_v, const float* w_v, int ks_v, void* tmp, void* out, cudaStream_t s) {
  EMX_REQUIRE(hwc && out && tmp && start_h && w_h && start_v && w_v && H > 0 && W > 0 && Ho > 0 && Wo > 0 && ks_h > 0 && ks_v > 0,
              "emx_lanczos3_resize_u8: empty problem");
  const long n1 = static_cast<long>(H) * Wo, n2 = static_cast<long>(Ho) * Wo * 3;
  lanczos_h_kernel<<<static_cast<unsigned>((n1 + 255) / 256), 256, 0, s>>>(static_cast<const uint8_t*>(hwc), static_cast<float*>(tmp), H, W, Wo, start_h, w_h, ks_h);
  EMX_CHECK_CUDA(cudaGetLastError());
  lanczos_v_kernel<<<static_cast<unsigned>((n2 + 255) / 256), 256, 0, s>>>(static_cast<const float*>(tmp), static_cast<uint8_t*>(out), H, Ho, Wo, start_v, w_v, ks_v);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}
