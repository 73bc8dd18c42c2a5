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
// Persistent: one CTA per SM walks the output tiles (M index fastest, so neighbouring CTAs share the W tile in L2); the fp32
// accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns), so the epilogue of tile i (TMEM read-out, fused math, global stores)
// overlaps the TMA / MMA main loop of tile i + 1 — the short-K ViT GEMMs (16-18 k-blocks per tile) were epilogue-bound without it.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..7 = epilogue.
#include "common.cuh"
#include "emmax.h"

namespace emx {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct EpiParams {
  const __nv_bfloat16* bias;   // [N] or null
  const __nv_bfloat16* ls;     // LayerScale [N] or null
  const __nv_bfloat16* resid;  // [M, ldr] or null
  int ldr;
  int resid_mod;  // >0: residual row = m % resid_mod (broadcast over batch: position embeddings)
  int flags;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, __nv_bfloat16* __restrict__ C,
               int ldc, int M, int N, int K, EpiParams ep) {
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
  const int mt = (M + BM - 1) / BM, n_tiles = mt * ((N + BN - 1) / BN);

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
      mbar_init(&tmem_empty[b], 4);  // one arrival per epilogue warp
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
        const int m0 = (tile % mt) * BM, n0 = (tile / mt) * BN;
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
    const int q = warp - 4;  // TMEM lane quarter this warp may read
    const bool has_bias = ep.bias != nullptr, has_ls = ep.ls != nullptr, has_res = ep.resid != nullptr;
    const bool do_gelu = ep.flags & EMX_EPI_GELU, do_swiglu = ep.flags & EMX_EPI_SWIGLU;
    uint32_t tl = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
    const int m0 = (tile % mt) * BM, n0 = (tile / mt) * BN;
    const uint32_t buf = tl & 1, bph = (tl >> 1) & 1;
    mbar_wait(&tmem_full[buf], bph);
    tc_fence_after();
    const int m = m0 + q * 32 + lane;
    const long rrow = has_res ? static_cast<long>(ep.resid_mod > 0 ? m % ep.resid_mod : m) * ep.ldr : 0;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + c * 32, r);
      tmem_ld_wait();
      const int nb = n0 + c * 32;
      if (m >= M || nb >= N) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      const bool fullchunk = nb + 32 <= N;
      if (has_bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (fullchunk || nb + j < N) v[j] += ld_bf16(ep.bias + nb + j);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]);  // Linear output (bias fused before rounding, as cuBLASLt)
      if (do_gelu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = bf16_round(gelu_erf(v[j]));
      }
      if (has_ls) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (fullchunk || nb + j < N) v[j] = bf16_round(v[j] * ld_bf16(ep.ls + nb + j));
      }
      if (has_res) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (fullchunk || nb + j < N) v[j] = bf16_round(v[j] + ld_bf16(ep.resid + rrow + nb + j));
      }
      if (do_swiglu) {
        // interleaved (gate, up) columns -> one output column per pair: bf16(bf16(silu(g)) * u)
        __nv_bfloat16* crow = C + static_cast<long>(m) * ldc + nb / 2;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (fullchunk || nb + 2 * j + 1 < N) crow[j] = __float2bfloat16_rn(bf16_round(silu(v[2 * j])) * v[2 * j + 1]);
      } else {
        __nv_bfloat16* crow = C + static_cast<long>(m) * ldc + nb;
        if (fullchunk && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
            o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
            reinterpret_cast<uint4*>(crow)[j] = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < N) crow[j] = __float2bfloat16_rn(v[j]);
        }
      }
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

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, __nv_bfloat16* C, int ldc, int M, int N, int K, const EpiParams& ep,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  bool* attr_set = device_attr_flag(BN == 128 ? ATTR_GEMM128 : ATTR_GEMM256);  // per device: the opt-in is device state
  const int sms = device_sms();
  if (!attr_set || sms < 0) return -2;
  if (!*attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    *attr_set = true;
  }
  const int n_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  gemm_tn_kernel<BN><<<n_tiles < sms ? n_tiles : sms, 256, Cfg::kSmemBytes, stream>>>(ta, tb, C, ldc, M, N, K, ep);
  EMX_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace emx

extern "C" int emx_gemm_bf16(const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K, const void* bias,
                             const void* layerscale, const void* resid, int ldr, int resid_mod, int flags, cudaStream_t stream) {
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
  const bool wide = (static_cast<long>((M + BM - 1) / BM) * ((N + 255) / 256) >= 2 * sms);
  if (int r = make_tmap(&ta, A, M, K, lda, BM)) return r;
  if (int r = make_tmap(&tb, W, N, K, ldw, wide ? 256 : 128)) return r;
  return wide ? launch_gemm<256>(ta, tb, static_cast<__nv_bfloat16*>(C), ldc, M, N, K, ep, stream)
              : launch_gemm<128>(ta, tb, static_cast<__nv_bfloat16*>(C), ldc, M, N, K, ep, stream);
}
