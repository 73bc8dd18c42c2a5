// emx_debug_stream — bandwidth probe for the weight-streaming mechanism of the decode kernel (not on the product path):
// every CTA pulls its contiguous share of `bytes` through a shared-memory ring with cp.async.bulk and drops it.
// Variants: rows x seg bytes per stage (seg contiguous bytes taken every `row_stride` bytes), number of stages.
#include "common.cuh"
#include "emmax.h"

namespace emx {

__global__ void __launch_bounds__(512, 1) stream_probe_kernel(const uint8_t* __restrict__ src, long bytes_per_pair, int rows, int seg,
                                                              long row_stride, int stages, int evict_first, int npairs) {
  extern __shared__ __align__(128) uint8_t smem_all[];
  const int stage_bytes = rows * seg;
  const int pair = threadIdx.x >> 6;
  const long ring_bytes = static_cast<long>(stages) * stage_bytes;
  uint8_t* smem = smem_all + pair * ring_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_all + npairs * ring_bytes) + pair * 2 * stages;
  uint64_t* empty = full + stages;
  const int warp = (threadIdx.x >> 5) & 1, lane = threadIdx.x & 31;
  if ((threadIdx.x & 63) == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // a "block" = rows x row_stride bytes of source, streamed as row_stride/seg stages of rows x seg
  const uint8_t* base = src + (static_cast<long>(blockIdx.x) * npairs + pair) * bytes_per_pair;
  const long block_bytes = static_cast<long>(rows) * row_stride;
  const long n_blocks = bytes_per_pair / block_bytes;
  const int segs_per_row = static_cast<int>(row_stride / seg);
  const long n_it = n_blocks * segs_per_row;
  if (warp == 0) {
    const uint64_t policy = evict_first ? l2_policy_evict_first() : l2_policy_evict_last();
    for (long it = 0; it < n_it; ++it) {
      const int slot = it % stages;
      const uint32_t ph = (it / stages) & 1;
      if (lane == 0) {
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&full[slot], stage_bytes);
      }
      __syncwarp();
      const long blk = it / segs_per_row, sg = it % segs_per_row;
      for (int r = lane; r < rows; r += 32)
        bulk_g2s(smem + static_cast<long>(slot) * stage_bytes + r * seg, base + blk * block_bytes + r * row_stride + sg * seg, seg, &full[slot], policy);
    }
  } else {
    for (long it = 0; it < n_it; ++it) {
      const int slot = it % stages;
      const uint32_t ph = (it / stages) & 1;
      mbar_wait(&full[slot], ph);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    }
  }
}

}  // namespace emx

extern "C" int emx_debug_stream(const void* src, long bytes, int rows, int seg, long row_stride, int stages, int evict_first, int grid,
                                int npairs, cudaStream_t stream) {
  using namespace emx;
  const int smem = npairs * (stages * rows * seg + 2 * stages * 8) + 128;
  EMX_REQUIRE(smem <= 227 * 1024 && seg % 16 == 0 && row_stride % seg == 0 && npairs >= 1 && npairs <= 8, "emx_debug_stream: bad geometry");
  EMX_CHECK_CUDA(cudaFuncSetAttribute(stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long block = static_cast<long>(rows) * row_stride;
  const long per_pair = bytes / grid / npairs / block * block;
  stream_probe_kernel<<<grid, 64 * npairs, smem, stream>>>(static_cast<const uint8_t*>(src), per_pair, rows, seg, row_stride, stages,
                                                           evict_first, npairs);
  EMX_CHECK_CUDA(cudaGetLastError());
  return static_cast<int>(per_pair / block);  // blocks per producer/consumer pair actually streamed (>= 0)
}
