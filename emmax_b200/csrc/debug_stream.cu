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
  const bool extra = pair >= npairs;  // padding warps (thread-count experiments): take part in the barrier, then leave
  const long ring_bytes = static_cast<long>(stages) * stage_bytes;
  uint8_t* smem = smem_all + pair * ring_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_all + npairs * ring_bytes) + pair * 2 * stages;
  uint64_t* empty = full + stages;
  const int warp = (threadIdx.x >> 5) & 1, lane = threadIdx.x & 31;
  if ((threadIdx.x & 63) == 0 && !extra) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (extra) return;
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
                                int npairs_and_pad, cudaStream_t stream) {
  using namespace emx;
  const int npairs = npairs_and_pad & 15, pad_warps = npairs_and_pad >> 4;
  const int smem = npairs * (stages * rows * seg + 2 * stages * 8) + 128;
  EMX_REQUIRE(smem <= 227 * 1024 && seg % 16 == 0 && row_stride % seg == 0 && npairs >= 1 && npairs <= 8, "emx_debug_stream: bad geometry");
  EMX_CHECK_CUDA(cudaFuncSetAttribute(stream_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long block = static_cast<long>(rows) * row_stride;
  const long per_pair = bytes / grid / npairs / block * block;
  stream_probe_kernel<<<grid, 64 * npairs + 32 * pad_warps, smem, stream>>>(static_cast<const uint8_t*>(src), per_pair, rows, seg, row_stride, stages,
                                                           evict_first, npairs);
  EMX_CHECK_CUDA(cudaGetLastError());
  return static_cast<int>(per_pair / block);  // blocks per producer/consumer pair actually streamed (>= 0)
}

// ---------------------------------------------------------------------------------------------------------------------
// emx_debug_skeleton — the DATAFLOW of the persistent decode kernel without its arithmetic (profiling aid): per CTA two
// producer warps stream a static byte schedule through the bulk-copy ring, eight consumer warps drain it (optionally
// burning `consume_cycles` per stage), and after every phase the consumers run a grid barrier of the chosen variant and
// then stall for `stall_ns[phase]` (the attention / vector-reload bubbles of the real kernel). Answers: what does the
// barrier cost under HBM load, how much of a stall does the ring hide, which barrier protocol is fastest.
// ---------------------------------------------------------------------------------------------------------------------
namespace emx {

struct SkelParams {
  const uint8_t* src;
  long region_bytes;      // per-CTA source region (offsets wrap inside it)
  int n_phases;
  int phase_stages[8];    // ring stages per phase per CTA
  int stall_ns[8];        // consumer stall after the barrier that ends the phase
  int reps;               // "layers"
  int rows, seg;          // stage = rows x seg bytes, rows contiguous segments `row_stride` apart
  long row_stride;
  int stages;
  int consume_cycles;
  int n_prod, n_cons;      // producer warps (1..2) and consumer warps (1..8) taking part
  int barrier_variant;    // 0 none, 1 fence+red.release / ld.acquire poll (v3 kernel), 2 red.release / relaxed poll + fence,
                          // 3 last arriver (atom) releases per-CTA flags, 4 like 2 with nanosleep back-off
  int pf_pace_ns;
  int prod_warp0;          // first producer warp (8 = like the decode kernel; 1 with n_cons == 1 = like the plain stream probe)
  int pf_stages, pf_mode;  // L2 prefetch warp: look-ahead in ring stages (0 = off); mode 0 bulk prefetch, 1 LSU prefetch per 128-B line,
                           // 2/3: the same two but issued only while the SM has no ring copy in flight, one stage per pf_pace_ns
  const float* weight;    // optional per-CTA work multiplier (mean 1): stages of a phase = round(phase_stages * weight), error carried
  int timers;             // 0: no %globaltimer reads inside the loop (pure latency measurements)
  uint32_t* sync;         // [0] ticket counter, [32 + 32*cta] per-CTA release flags (zeroed by the host)
  long long* out;         // [0..n_phases) CTA 0: summed phase time (ns), [8..16) summed barrier time, [16] total ns,
                          // [32 + cta] per-CTA total ns, [32 + 148 + cta] %smid
};

__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long gtime_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void cons_bar(int n_cons) { asm volatile("bar.sync 1, %0;" ::"r"(n_cons * 32) : "memory"); }

__device__ __forceinline__ void skel_barrier(const SkelParams& p, uint32_t& epoch) {
  ++epoch;
  const uint32_t target = epoch * gridDim.x;
  cons_bar(p.n_cons);
  if (threadIdx.x == 0) {
    uint32_t spins = 0;
    switch (p.barrier_variant) {
      case 1:
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.sync) : "memory");
        while (static_cast<int32_t>(ld_acquire_gpu_u32(p.sync) - target) < 0)
          if (++spins > EMX_SPIN_LIMIT) __trap();
        break;
      case 2:
      case 4:
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.sync) : "memory");
        while (static_cast<int32_t>(ld_relaxed_u32(p.sync) - target) < 0) {
          if (p.barrier_variant == 4) __nanosleep(64);
          if (++spins > EMX_SPIN_LIMIT) __trap();
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        break;
      case 3: {
        uint32_t old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.sync) : "memory");
        if (old + 1 == target) {
          for (uint32_t c = 0; c < gridDim.x; ++c) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.sync + 32 + 32 * c), "r"(epoch) : "memory");
        }
        const uint32_t* flag = p.sync + 32 + 32 * blockIdx.x;
        while (static_cast<int32_t>(ld_relaxed_u32(flag) - epoch) < 0)
          if (++spins > EMX_SPIN_LIMIT) __trap();
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        break;
      }
      case 5:
      case 6:
      case 7: {  // G counters on separate lines (arrivals serialise per address at L2); the poller sums all G
        const int G = p.barrier_variant == 5 ? 4 : p.barrier_variant == 6 ? 8 : 16;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.sync + 32 * (blockIdx.x % G)) : "memory");
        for (;;) {
          uint32_t sum = 0;
          for (int g = 0; g < G; ++g) sum += ld_relaxed_u32(p.sync + 32 * g);
          if (static_cast<int32_t>(sum - target) >= 0) break;
          if (++spins > EMX_SPIN_LIMIT) __trap();
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        break;
      }
      case 8: {  // two-level tree: 8 group counters, the last arriver of a group arrives at the root; everyone polls the root
        const int G = 8;
        const int g = blockIdx.x % G;
        const uint32_t gsize = (gridDim.x - g + G - 1) / G;
        uint32_t old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(p.sync + 32 * (1 + g)) : "memory");
        if (old + 1 == epoch * gsize) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.sync) : "memory");
        while (static_cast<int32_t>(ld_relaxed_u32(p.sync) - epoch * G) < 0)
          if (++spins > EMX_SPIN_LIMIT) __trap();
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        break;
      }
      case 9: {  // flags, no atomics: arrive = store to the CTA's own line; a master warp (CTA 0, warp 10) gathers and releases
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sync + 32 * (1 + blockIdx.x)), "r"(epoch) : "memory");
        const uint32_t* flag = p.sync + 32 * (1 + gridDim.x + blockIdx.x);
        while (static_cast<int32_t>(ld_relaxed_u32(flag) - epoch) < 0)
          if (++spins > EMX_SPIN_LIMIT) __trap();
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        break;
      }
      default: break;
    }
  }
  if (p.barrier_variant == 10) {  // all-to-all flags: every CTA stores its own flag and polls everybody's (one flag per thread)
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sync + 32 * (1 + blockIdx.x)), "r"(epoch) : "memory");
    uint32_t spins = 0;
    if (threadIdx.x < gridDim.x) {
      const uint32_t* flag = p.sync + 32 * (1 + threadIdx.x);
      while (static_cast<int32_t>(ld_relaxed_u32(flag) - epoch) < 0)
        if (++spins > EMX_SPIN_LIMIT) __trap();
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
  }
  cons_bar(p.n_cons);
}

// master warp of barrier variant 9: lane l watches the arrival flags of CTAs l, l+32, ...; when all reached `epoch`, the
// lanes store the release flags
__device__ void skel_master(const SkelParams& p, int lane, uint32_t n_barriers) {
  for (uint32_t epoch = 1; epoch <= n_barriers; ++epoch) {
    uint32_t spins = 0;
    for (;;) {
      bool ok = true;
      for (uint32_t c = lane; c < gridDim.x; c += 32) ok &= static_cast<int32_t>(ld_relaxed_u32(p.sync + 32 * (1 + c)) - epoch) >= 0;
      if (__all_sync(0xffffffffu, ok)) break;
      if (++spins > EMX_SPIN_LIMIT) __trap();
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    for (uint32_t c = lane; c < gridDim.x; c += 32)
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.sync + 32 * (1 + gridDim.x + c)), "r"(epoch) : "memory");
  }
}

__global__ void __launch_bounds__(352, 1) skeleton_kernel(const SkelParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int stage_bytes = p.rows * p.seg;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<long>(p.stages) * stage_bytes);
  uint64_t* empty = full + p.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], p.n_cons);
    *reinterpret_cast<volatile uint32_t*>(empty + p.stages) = 0;
    fence_mbar_init();
  }
  __syncthreads();
  const float wgt = p.weight ? p.weight[blockIdx.x] : 1.0f;
  // stages of (rep, phase) for this CTA: round(phase_stages * weight) with the rounding error carried forward
  auto stages_of = [&](int ph_i, float& debt) {
    const float want = p.phase_stages[ph_i] * wgt + debt;
    const int n = max(0, static_cast<int>(floorf(want + 0.5f)));
    debt = want - n;
    return n;
  };
  long n_it = 0;
  {
    float debt = 0.f;
    for (int rep = 0; rep < p.reps; ++rep)
      for (int i = 0; i < p.n_phases; ++i) n_it += stages_of(i, debt);
  }
  if (warp == 10) {  // service warp: barrier master (CTA 0, variant 9) or L2 prefetcher
    if (p.barrier_variant == 9 && blockIdx.x == 0) {
      skel_master(p, lane, static_cast<uint32_t>(p.reps) * p.n_phases);
      return;
    }
    if (p.pf_stages <= 0) return;
    const uint8_t* base = p.src + static_cast<long>(blockIdx.x) * p.region_bytes;
    const int segs_per_row = static_cast<int>(p.row_stride / p.seg);
    const long block_bytes = static_cast<long>(p.rows) * p.row_stride;
    const long n_blocks = p.region_bytes / block_bytes;
    volatile uint32_t* issued = reinterpret_cast<volatile uint32_t*>(empty + p.stages);  // stages issued by the producers
    auto prefetch_stage = [&](long it, int mode) {
      const long blk = (it / segs_per_row) % n_blocks, sg = it % segs_per_row;
      const uint8_t* src = base + blk * block_bytes + sg * p.seg;
      if (mode == 0) {
        for (int r = lane; r < p.rows; r += 32)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + r * p.row_stride), "r"(p.seg) : "memory");
      } else {
        for (int r = 0; r < p.rows; ++r)
          for (int off = lane * 128; off < p.seg; off += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + r * p.row_stride + off));
      }
    };
    if (p.pf_mode >= 2) {
      // idle-triggered: prefetch only while this SM has NO ring copy in flight (the newest issued stage has landed, i.e. the
      // consumers are stalled and the ring is full), paced at ~the SM's fair share of HBM, at most pf_stages ahead
      long next = 0;
      const long long pace_ns = p.pf_pace_ns;
      while (true) {
        const long iss = *issued;
        if (iss >= n_it) break;
        if (next < iss) next = iss;
        bool idle = false;
        if (iss > 0 && next < n_it && next - iss < p.pf_stages) {
          const long last = iss - 1;
          idle = mbar_try_wait(&full[last % p.stages], (last / p.stages) & 1);
        }
        if (!idle) {
          __nanosleep(200);
          continue;
        }
        const long long t0 = gtime_ns();
        prefetch_stage(next, p.pf_mode - 2);
        ++next;
        while (gtime_ns() - t0 < pace_ns) __nanosleep(100);
      }
      return;
    }
    for (long it = 0; it < n_it; ++it) {
      uint32_t spins = 0;
      while (it >= static_cast<long>(*issued) + p.pf_stages) {
        __nanosleep(100);
        if (++spins > EMX_SPIN_LIMIT) __trap();
      }
      if (it < static_cast<long>(*issued)) continue;  // the ring overtook us
      prefetch_stage(it, p.pf_mode);
    }
    return;
  }
  if (warp >= p.prod_warp0 && warp < p.prod_warp0 + 2) {  // producer warps, alternating stages
    const int pidx = warp - p.prod_warp0;
    if (pidx >= p.n_prod) return;
    volatile uint32_t* issued = reinterpret_cast<volatile uint32_t*>(empty + p.stages);
    const uint64_t policy = l2_policy_evict_first();
    const uint8_t* base = p.src + static_cast<long>(blockIdx.x) * p.region_bytes;
    const int segs_per_row = static_cast<int>(p.row_stride / p.seg);
    const long block_bytes = static_cast<long>(p.rows) * p.row_stride;
    const long n_blocks = p.region_bytes / block_bytes;
    for (long it = pidx; it < n_it; it += p.n_prod) {
      const int slot = it % p.stages;
      const uint32_t ph = (it / p.stages) & 1;
      if (lane == 0) {
        mbar_wait(&empty[slot], ph ^ 1);
        mbar_arrive_expect_tx(&full[slot], stage_bytes);
      }
      __syncwarp();
      const long blk = (it / segs_per_row) % n_blocks, sg = it % segs_per_row;
      for (int r = lane; r < p.rows; r += 32)
        bulk_g2s(smem + static_cast<long>(slot) * stage_bytes + r * p.seg, base + blk * block_bytes + r * p.row_stride + sg * p.seg, p.seg,
                 &full[slot], policy);
      if (lane == 0) *issued = static_cast<uint32_t>(it + 1);
    }
    return;
  }
  if (warp >= p.n_cons) return;  // (also the unused warps 8..9 when the producers sit at warp 1)
  uint32_t epoch = 0;
  long it = 0;
  long long t_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_bar[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_start = gtime_ns();
  float debt = 0.f;
  for (int rep = 0; rep < p.reps; ++rep) {
    for (int ph_i = 0; ph_i < p.n_phases; ++ph_i) {
      const long long t0 = p.timers ? gtime_ns() : 0;
      const int n_st = stages_of(ph_i, debt);
      for (int s = 0; s < n_st; ++s, ++it) {
        const int slot = it % p.stages;
        mbar_wait(&full[slot], (it / p.stages) & 1);
        if (p.consume_cycles > 0) {
          const long long c0 = clock64();
          while (clock64() - c0 < p.consume_cycles) {}
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
      const long long t1 = p.timers ? gtime_ns() : 0;
      if (p.barrier_variant) skel_barrier(p, epoch);
      const long long t2 = (p.timers || p.stall_ns[ph_i] > 0) ? gtime_ns() : 0;
      t_phase[ph_i] += t1 - t0, t_bar[ph_i] += t2 - t1;
      if (p.stall_ns[ph_i] > 0) {
        while (gtime_ns() - t2 < p.stall_ns[ph_i]) {}
      }
    }
  }
  if (threadIdx.x == 0) {
    const long long t_end = gtime_ns();
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.out[32 + blockIdx.x] = t_end - t_start, p.out[32 + gridDim.x + blockIdx.x] = smid;
    if (blockIdx.x == 0) {
      for (int i = 0; i < 8; ++i) p.out[i] = t_phase[i], p.out[8 + i] = t_bar[i];
      p.out[16] = t_end - t_start;
    }
  }
}

}  // namespace emx

extern "C" int emx_debug_skeleton(const void* src, long region_bytes, int n_phases, const int* phase_stages, const int* stall_ns, int reps,
                                  int rows, int seg, long row_stride, int stages, int consume_cycles, int barrier_variant, int n_prod,
                                  int n_cons, const float* weight, int timers, int pf_stages, int pf_mode, int pf_pace_ns, int launch_mode, void* sync, void* out, cudaStream_t stream) {
  using namespace emx;
  EMX_REQUIRE(n_phases >= 1 && n_phases <= 8 && seg % 16 == 0 && row_stride % seg == 0, "emx_debug_skeleton: bad arguments");
  SkelParams p;
  p.src = static_cast<const uint8_t*>(src), p.region_bytes = region_bytes, p.n_phases = n_phases;
  for (int i = 0; i < 8; ++i) p.phase_stages[i] = i < n_phases ? phase_stages[i] : 0, p.stall_ns[i] = i < n_phases ? stall_ns[i] : 0;
  p.reps = reps, p.rows = rows, p.seg = seg, p.row_stride = row_stride, p.stages = stages, p.consume_cycles = consume_cycles;
  p.prod_warp0 = (launch_mode & 2) ? 1 : 8;
  p.n_prod = n_prod, p.n_cons = n_cons, p.weight = weight, p.timers = timers, p.pf_stages = pf_stages, p.pf_mode = pf_mode, p.pf_pace_ns = pf_pace_ns;
  EMX_REQUIRE(n_prod >= 1 && n_prod <= 2 && n_cons >= 1 && n_cons <= 8, "emx_debug_skeleton: 1..2 producer and 1..8 consumer warps");
  p.barrier_variant = barrier_variant, p.sync = static_cast<uint32_t*>(sync), p.out = static_cast<long long*>(out);
  const int smem = stages * rows * seg + 2 * stages * 8 + 128;
  EMX_REQUIRE(smem <= 227 * 1024, "emx_debug_skeleton: ring of %d bytes does not fit", smem);
  EMX_CHECK_CUDA(cudaFuncSetAttribute(skeleton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  void* args[] = {&p};
  if (launch_mode & 1) {  // plain launch (only safe for barrier_variant 0 ... the grid is co-resident anyway at 1 CTA/SM)
    skeleton_kernel<<<kNumSMs, 352, smem, stream>>>(p);
    EMX_CHECK_CUDA(cudaGetLastError());
  } else {
    EMX_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(skeleton_kernel), dim3(kNumSMs), dim3(352), args, smem, stream));
  }
  return 0;
}
