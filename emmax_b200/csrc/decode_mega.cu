// emx_decode_step — one greedy decode step of Llama-2-7B for one sequence as ONE persistent kernel.
//
// Replaces the cached branch of PrismaticForConditionalGeneration.forward
// (/root/reference/prismatic/extern/hf/modeling_prismatic.py:325-341: language_model(input_ids[:, -1:], past_key_values))
// plus one iteration of GenerationMixin's greedy loop (called at modeling_prismatic.py:519), which in the reference is
// ~10^3 kernel launches per token, a full re-copy of the KV cache (DynamicCache torch.cat) and a host sync.
//
// B200 design (HBM-bound: 13.2 GB of bf16 weights per token):
//   * grid = one CTA per SM (148), 8 consumer warps + 1 producer warp, launched cooperatively (co-residency guaranteed);
//   * the producer warp walks the STATIC weight schedule of its CTA (layer -> qkv, o, gate/up, down -> row group ->
//     K chunk) and keeps a 6 x 32 KB shared-memory ring full with cp.async.bulk (TMA engine) copies, L2 evict-first;
//     it never waits for a grid barrier, so HBM keeps streaming while consumers synchronise or run attention;
//   * consumer warp w owns 2 rows of every 16-row group: 16-B conflict-free LDS of weights and of the bf16 activation
//     vector, fp32 FMA, one warp-shuffle reduction per row, fused epilogues (residual add, SwiGLU, argmax);
//   * RMSNorm is recomputed per CTA from the 8 KB residual vector (cheaper than a launch + barrier);
//   * attention: (head, kv-split) items across CTAs, RoPE + KV append fused in, last-arriving split combines;
//   * phases are separated by a ticket grid barrier (release/acquire at gpu scope); cross-CTA activations are read
//     with ld.global.cg (L1 bypass).
// Rounding points mirror the torch-eager reference (bf16 after every Linear / norm / residual add / activation).
#include "common.cuh"
#include "emmax.h"

namespace emx {

constexpr int DEC_CWARPS = 8;                    // consumer warps
constexpr int DEC_CTHREADS = DEC_CWARPS * 32;    // 256
constexpr int DEC_THREADS = DEC_CTHREADS + 32;   // + producer warp
constexpr int DEC_RPW = 2;                       // rows per consumer warp per group
constexpr int DEC_GROUP = DEC_CWARPS * DEC_RPW;  // 16 rows per ring stage
constexpr int DEC_KC = 1024;                     // K elements per ring stage (2 KB per row segment)
constexpr int DEC_STAGES = 6;
constexpr int DEC_STAGE_BYTES = DEC_GROUP * DEC_KC * 2;  // 32 KB
constexpr int DEC_XS_BYTES = 22528;                      // activation vector (bf16), up to 11264 elements
constexpr int DEC_MISC_BYTES = 2048;
constexpr int DEC_SMEM = DEC_STAGES * DEC_STAGE_BYTES + DEC_XS_BYTES + DEC_MISC_BYTES + 128;
constexpr int DEC_HD = 128;  // head_dim supported by the decode kernel (Llama-2)

enum PhaseKind { PH_QKV = 0, PH_O = 1, PH_GATEUP = 2, PH_DOWN = 3, PH_LMHEAD = 4 };

struct PhaseDesc {
  const __nv_bfloat16* W;
  int N, K;
};

__device__ __forceinline__ PhaseDesc phase_desc(const emx_decode_params& p, int layer, int kind) {
  const long H = p.hidden, I = p.inter;
  PhaseDesc d;
  switch (kind) {
    case PH_QKV: d.W = static_cast<const __nv_bfloat16*>(p.w_qkv) + layer * 3 * H * H, d.N = 3 * H, d.K = H; break;
    case PH_O: d.W = static_cast<const __nv_bfloat16*>(p.w_o) + layer * H * H, d.N = H, d.K = H; break;
    case PH_GATEUP: d.W = static_cast<const __nv_bfloat16*>(p.w_gateup) + layer * 2 * I * H, d.N = 2 * I, d.K = H; break;
    case PH_DOWN: d.W = static_cast<const __nv_bfloat16*>(p.w_down) + layer * H * I, d.N = H, d.K = I; break;
    default: d.W = static_cast<const __nv_bfloat16*>(p.lm_head), d.N = p.vocab, d.K = H; break;
  }
  return d;
}

// rows of a phase owned by this CTA, in units of DEC_RPW rows
__device__ __forceinline__ void cta_rows(int N, int& r_begin, int& r_end) {
  const long U = N / DEC_RPW;
  r_begin = static_cast<int>(U * blockIdx.x / gridDim.x) * DEC_RPW;
  r_end = static_cast<int>(U * (blockIdx.x + 1) / gridDim.x) * DEC_RPW;
}

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(DEC_CTHREADS) : "memory"); }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ticket barrier over all CTAs (consumer threads only; the producer warp never synchronises with the grid)
__device__ __forceinline__ void grid_sync(uint32_t* counter, uint32_t& target) {
  target += gridDim.x;
  cbar();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    uint32_t spins = 0;
    while (static_cast<int32_t>(ld_acquire_u32(counter) - target) < 0) {
      if (++spins > EMX_SPIN_LIMIT) __trap();
    }
    __threadfence();
  }
  cbar();
}

__device__ __forceinline__ float cblock_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  cbar();
  float t = (lane < DEC_CWARPS) ? red[lane] : 0.f;
  t = warp_sum(t);
  cbar();
  return t;
}
__device__ __forceinline__ float cblock_max(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  cbar();
  float t = (lane < DEC_CWARPS) ? red[lane] : -INFINITY;
  t = warp_max(t);
  cbar();
  return t;
}

// xs = bf16(w * bf16(x * rsqrt(mean(x^2) + eps)))  — LlamaRMSNorm, computed redundantly by every CTA
__device__ __forceinline__ void load_rmsnorm(const __nv_bfloat16* x, const __nv_bfloat16* w, __nv_bfloat16* xs, int H, float eps,
                                             float* red) {
  const int nv = H >> 3;
  float ss = 0.f;
  for (int i = threadIdx.x; i < nv; i += DEC_CTHREADS) {
    const uint4 v = ldg_cg_v4(reinterpret_cast<const uint4*>(x) + i);
    reinterpret_cast<uint4*>(xs)[i] = v;
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16_lo(u[j]), c = bf16_hi(u[j]);
      ss += a * a + c * c;
    }
  }
  const float rs = 1.0f / sqrtf(cblock_sum(ss, red) / H + eps);
  for (int i = threadIdx.x; i < nv; i += DEC_CTHREADS) {
    const uint4 v = reinterpret_cast<uint4*>(xs)[i];
    const uint4 ww = reinterpret_cast<const uint4*>(w)[i];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, uw[4] = {ww.x, ww.y, ww.z, ww.w};
    uint32_t r[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      r[j] = pack_bf16(bf16_lo(uw[j]) * bf16_round(bf16_lo(u[j]) * rs), bf16_hi(uw[j]) * bf16_round(bf16_hi(u[j]) * rs));
    reinterpret_cast<uint4*>(xs)[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
  cbar();
}

__device__ __forceinline__ void load_vec(const __nv_bfloat16* v, __nv_bfloat16* xs, int n) {
  for (int i = threadIdx.x; i < (n >> 3); i += DEC_CTHREADS) reinterpret_cast<uint4*>(xs)[i] = ldg_cg_v4(reinterpret_cast<const uint4*>(v) + i);
  cbar();
}

struct RingState {
  uint32_t it;  // stage counter, identical sequence in producer and consumers
  long long waited;  // profiling: cycles spent waiting on the ring (only meaningful when params.dbg != null)
};

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---- producer: stream one phase's rows of this CTA through the ring --------------------------------------------------
__device__ __forceinline__ void produce_phase(const PhaseDesc& d, uint8_t* ring, uint64_t* full, uint64_t* empty, RingState& rs,
                                              uint64_t policy, int lane) {
  int r_begin, r_end;
  cta_rows(d.N, r_begin, r_end);
  for (int r0 = r_begin; r0 < r_end; r0 += DEC_GROUP) {
    const int nrows = min(DEC_GROUP, r_end - r0);
    for (int k0 = 0; k0 < d.K; k0 += DEC_KC) {
      const int klen = min(DEC_KC, d.K - k0);
      const int slot = rs.it % DEC_STAGES;
      const uint32_t ph = (rs.it / DEC_STAGES) & 1;
      if (lane == 0) {
        const long long t0 = clock64();
        mbar_wait(&empty[slot], ph ^ 1);
        rs.waited += clock64() - t0;
        mbar_arrive_expect_tx(&full[slot], static_cast<uint32_t>(nrows) * klen * 2);
      }
      __syncwarp();
      if (lane < nrows)
        bulk_g2s(ring + slot * DEC_STAGE_BYTES + lane * (DEC_KC * 2), d.W + static_cast<long>(r0 + lane) * d.K + k0, klen * 2, &full[slot],
                 policy);
      ++rs.it;
    }
  }
}

// ---- consumer: dot products of this warp's 2 rows of every group against xs ------------------------------------------
template <typename Epi>
__device__ __forceinline__ void consume_phase(const PhaseDesc& d, const uint8_t* ring, uint64_t* full, uint64_t* empty, RingState& rs,
                                              const __nv_bfloat16* xs, int warp, int lane, Epi&& epi) {
  int r_begin, r_end;
  cta_rows(d.N, r_begin, r_end);
  for (int r0 = r_begin; r0 < r_end; r0 += DEC_GROUP) {
    const int nrows = min(DEC_GROUP, r_end - r0);
    const bool active = warp * DEC_RPW < nrows;
    float acc0 = 0.f, acc1 = 0.f;
    for (int k0 = 0; k0 < d.K; k0 += DEC_KC) {
      const int klen = min(DEC_KC, d.K - k0);
      const int slot = rs.it % DEC_STAGES;
      const uint32_t ph = (rs.it / DEC_STAGES) & 1;
      const long long t0 = clock64();
      mbar_wait(&full[slot], ph);
      rs.waited += clock64() - t0;
      if (active) {
        const uint4* w0 = reinterpret_cast<const uint4*>(ring + slot * DEC_STAGE_BYTES + (warp * DEC_RPW) * (DEC_KC * 2));
        const uint4* w1 = w0 + (DEC_KC * 2) / 16;
        const uint4* xv = reinterpret_cast<const uint4*>(xs + k0);
        const int nv = klen >> 3;
#pragma unroll 4
        for (int c = lane; c < nv; c += 32) {
          const uint4 x4 = xv[c], a = w0[c], b = w1[c];
          const uint32_t ux[4] = {x4.x, x4.y, x4.z, x4.w}, ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float xl = bf16_lo(ux[j]), xh = bf16_hi(ux[j]);
            acc0 = fmaf(bf16_lo(ua[j]), xl, acc0), acc0 = fmaf(bf16_hi(ua[j]), xh, acc0);
            acc1 = fmaf(bf16_lo(ub[j]), xl, acc1), acc1 = fmaf(bf16_hi(ub[j]), xh, acc1);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      ++rs.it;
    }
    if (active) {
      acc0 = warp_sum(acc0), acc1 = warp_sum(acc1);
      if (lane == 0) epi(r0 + warp * DEC_RPW, acc0, acc1);
    }
  }
}

// ---- attention for one (head, split) item ------------------------------------------------------------------------------
__device__ __forceinline__ long kv_row(const emx_decode_params& p, int layer, int head, int key) {
  const int page = p.block_table[key / p.page_size];
  const long layer_off = static_cast<long>(layer) * p.n_pages * p.heads * p.page_size * DEC_HD;
  return layer_off + ((static_cast<long>(page) * p.heads + head) * p.page_size + key % p.page_size) * DEC_HD;
}

__device__ void attention_item(const emx_decode_params& p, int layer, int head, int split, int pos, float* sm /*>= 5.5 KB*/, float* red) {
  constexpr int HALF = DEC_HD / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = pos + 1, S = p.kv_splits;
  const int k_begin = static_cast<int>(static_cast<long>(n) * split / S), k_end = static_cast<int>(static_cast<long>(n) * (split + 1) / S);
  const int nk = k_end - k_begin;
  float* sq = sm;                 // [128] rotated q
  float* sknew = sm + 128;        // [128] rotated new k
  float* svnew = sm + 256;        // [128] new v
  float* sacc = sm + 384;         // [4][128] PV partials
  float* sscore = sm + 896;       // [nk] scores / probabilities (host guarantees capacity)
  __nv_bfloat16* kc = static_cast<__nv_bfloat16*>(p.k_cache);
  __nv_bfloat16* vc = static_cast<__nv_bfloat16*>(p.v_cache);
  const __nv_bfloat16* qkv = static_cast<const __nv_bfloat16*>(p.qkv);
  const int H = p.hidden;
  const bool owns_new = (k_end == n);  // the split that contains the token being decoded

  if (tid < HALF) {
    const int j = tid;
    const float c = ld_bf16(static_cast<const __nv_bfloat16*>(p.cos_tab) + static_cast<long>(pos) * HALF + j);
    const float s = ld_bf16(static_cast<const __nv_bfloat16*>(p.sin_tab) + static_cast<long>(pos) * HALF + j);
    const float q1 = ldg_cg_bf16(qkv + head * DEC_HD + j), q2 = ldg_cg_bf16(qkv + head * DEC_HD + j + HALF);
    sq[j] = bf16_round(bf16_round(q1 * c) + bf16_round(-q2 * s));
    sq[j + HALF] = bf16_round(bf16_round(q2 * c) + bf16_round(q1 * s));
    if (owns_new) {
      const float k1 = ldg_cg_bf16(qkv + H + head * DEC_HD + j), k2 = ldg_cg_bf16(qkv + H + head * DEC_HD + j + HALF);
      const float r1 = bf16_round(bf16_round(k1 * c) + bf16_round(-k2 * s)), r2 = bf16_round(bf16_round(k2 * c) + bf16_round(k1 * s));
      const float v1 = ldg_cg_bf16(qkv + 2 * H + head * DEC_HD + j), v2 = ldg_cg_bf16(qkv + 2 * H + head * DEC_HD + j + HALF);
      sknew[j] = r1, sknew[j + HALF] = r2, svnew[j] = v1, svnew[j + HALF] = v2;
      const long dst = kv_row(p, layer, head, pos);
      kc[dst + j] = __float2bfloat16_rn(r1), kc[dst + j + HALF] = __float2bfloat16_rn(r2);
      vc[dst + j] = __float2bfloat16_rn(v1), vc[dst + j + HALF] = __float2bfloat16_rn(v2);
    }
  }
  cbar();

  // scores: one warp per key, lane owns 4 consecutive dims (8-byte coalesced loads of the 256-B K row)
  const float scale = rsqrtf(static_cast<float>(DEC_HD));
  const float q0 = sq[lane * 4], q1 = sq[lane * 4 + 1], q2 = sq[lane * 4 + 2], q3 = sq[lane * 4 + 3];
  float lmax = -INFINITY;
  for (int kk = warp; kk < nk; kk += DEC_CWARPS) {
    const int key = k_begin + kk;
    float d;
    if (key == pos) {
      d = q0 * sknew[lane * 4] + q1 * sknew[lane * 4 + 1] + q2 * sknew[lane * 4 + 2] + q3 * sknew[lane * 4 + 3];
    } else {
      const uint2 kv = ldg_cg_v2(kc + kv_row(p, layer, head, key) + lane * 4);
      d = q0 * bf16_lo(kv.x) + q1 * bf16_hi(kv.x) + q2 * bf16_lo(kv.y) + q3 * bf16_hi(kv.y);
    }
    d = warp_sum(d) * scale;
    if (lane == 0) sscore[kk] = d;
    lmax = fmaxf(lmax, d);
  }
  const float m = cblock_max(lmax, red);  // includes the barrier that publishes sscore
  float lsum = 0.f;
  for (int kk = tid; kk < nk; kk += DEC_CTHREADS) {
    const float pr = __expf(sscore[kk] - m);
    lsum += pr;
    sscore[kk] = bf16_round(pr);  // flash-attn: P is bf16 for the PV product, the row sum stays fp32
  }
  const float l = cblock_sum(lsum, red);

  // PV: thread = (key slice, dim pair); 64 threads read one 256-B V row coalesced
  const int pr_idx = tid & 63, slice = tid >> 6;
  float a0 = 0.f, a1 = 0.f;
  for (int kk = slice; kk < nk; kk += 4) {
    const int key = k_begin + kk;
    const float pw = sscore[kk];
    float v0, v1;
    if (key == pos) {
      v0 = svnew[2 * pr_idx], v1 = svnew[2 * pr_idx + 1];
    } else {
      const uint32_t w = ldg_cg_u32(vc + kv_row(p, layer, head, key) + 2 * pr_idx);
      v0 = bf16_lo(w), v1 = bf16_hi(w);
    }
    a0 = fmaf(pw, v0, a0), a1 = fmaf(pw, v1, a1);
  }
  sacc[slice * 128 + 2 * pr_idx] = a0, sacc[slice * 128 + 2 * pr_idx + 1] = a1;
  cbar();
  float* part = p.part + (static_cast<long>(head) * S + split) * (DEC_HD + 2);
  if (tid < DEC_HD) part[2 + tid] = sacc[tid] + sacc[128 + tid] + sacc[256 + tid] + sacc[384 + tid];
  if (tid == 0) part[0] = m, part[1] = l;

  // last-arriving split of this head combines the partials
  __shared__ uint32_t s_ticket;
  __threadfence();
  cbar();
  if (tid == 0) s_ticket = atomicAdd(&p.state->head_ticket[head], 1u);
  cbar();
  if ((s_ticket + 1) % S == 0) {
    __threadfence();
    if (tid < DEC_HD) {
      const float* ph = p.part + static_cast<long>(head) * S * (DEC_HD + 2);
      float M = -INFINITY;
      for (int s2 = 0; s2 < S; ++s2)
        if (ldg_cg_f32(ph + s2 * (DEC_HD + 2) + 1) > 0.f) M = fmaxf(M, ldg_cg_f32(ph + s2 * (DEC_HD + 2)));
      float num = 0.f, den = 0.f;
      for (int s2 = 0; s2 < S; ++s2) {
        const float ls = ldg_cg_f32(ph + s2 * (DEC_HD + 2) + 1);
        if (ls > 0.f) {
          const float w = __expf(ldg_cg_f32(ph + s2 * (DEC_HD + 2)) - M);
          num = fmaf(w, ldg_cg_f32(ph + s2 * (DEC_HD + 2) + 2 + tid), num);
          den = fmaf(w, ls, den);
        }
      }
      static_cast<__nv_bfloat16*>(p.attn)[head * DEC_HD + tid] = __float2bfloat16_rn(num / den);
    }
  }
  cbar();
}

// ---- the kernel ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(DEC_THREADS, 1) decode_step_kernel(const emx_decode_params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem + DEC_STAGES * DEC_STAGE_BYTES);
  float* misc = reinterpret_cast<float*>(smem + DEC_STAGES * DEC_STAGE_BYTES + DEC_XS_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + DEC_STAGES * DEC_STAGE_BYTES + DEC_XS_BYTES + DEC_MISC_BYTES);
  uint64_t* empty = full + DEC_STAGES;
  float* red = misc;                // [8]
  int* s_state = reinterpret_cast<int*>(misc + 16);  // [4]
  float* s_best = misc + 32;        // [8] + [8]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  emx_decode_state* st = p.state;
  if (tid == 0) {
    s_state[0] = static_cast<int>(ldg_cg_u32(&st->cur_token));
    s_state[1] = static_cast<int>(ldg_cg_u32(&st->pos));
    s_state[2] = static_cast<int>(ldg_cg_u32(&st->n_generated));
    s_state[3] = static_cast<int>(ldg_cg_u32(&st->finished));
    for (int s = 0; s < DEC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], DEC_CWARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int token = s_state[0], pos = s_state[1], n_gen = s_state[2];
  if (s_state[3]) return;  // sequence already hit EOS: nothing to do (uniform across the grid)

  const int L = p.layers, H = p.hidden;
  RingState rs{0, 0};
  long long* dbg = (blockIdx.x == 0) ? reinterpret_cast<long long*>(p.dbg) : nullptr;
  int dbg_i = 0;
  auto mark = [&]() {
    if (dbg && tid == 0) dbg[dbg_i] = global_ns();
    ++dbg_i;
  };

  if (warp == DEC_CWARPS) {
    // ===================== producer warp =====================
    const uint64_t policy = l2_policy_evict_first();
    for (int layer = 0; layer < L; ++layer)
      for (int kind = PH_QKV; kind <= PH_DOWN; ++kind) produce_phase(phase_desc(p, layer, kind), ring, full, empty, rs, policy, lane);
    produce_phase(phase_desc(p, 0, PH_LMHEAD), ring, full, empty, rs, policy, lane);
    if (dbg && lane == 0) dbg[15 * L + 9] = rs.waited;
    return;
  }

  // ===================== consumer warps =====================
  // barrier tickets: every non-finished launch performs exactly (5 L + 1) grid syncs
  const uint32_t n_sync = 5u * L + 1u;
  uint32_t target = ldg_cg_u32(&st->epoch) * n_sync * gridDim.x;

  __nv_bfloat16* x = static_cast<__nv_bfloat16*>(p.x);
  __nv_bfloat16* qkv = static_cast<__nv_bfloat16*>(p.qkv);
  __nv_bfloat16* hbuf = static_cast<__nv_bfloat16*>(p.h);
  const __nv_bfloat16* emb_row = static_cast<const __nv_bfloat16*>(p.embed) + static_cast<long>(token) * H;

  for (int layer = 0; layer < L; ++layer) {
    const __nv_bfloat16* resid_src = (layer == 0) ? emb_row : x;
    // ---- P1: RMSNorm + QKV ----
    mark();
    load_rmsnorm(resid_src, static_cast<const __nv_bfloat16*>(p.ln1) + static_cast<long>(layer) * H, xs, H, p.rms_eps, red);
    consume_phase(phase_desc(p, layer, PH_QKV), ring, full, empty, rs, xs, warp, lane, [&](int row, float a0, float a1) {
      *reinterpret_cast<uint32_t*>(qkv + row) = pack_bf16(a0, a1);
    });
    mark();
    grid_sync(&st->barrier, target);
    mark();
    // ---- P2: RoPE + KV append + split-KV attention ----
    for (int item = blockIdx.x; item < p.heads * p.kv_splits; item += gridDim.x)
      attention_item(p, layer, item / p.kv_splits, item % p.kv_splits, pos, reinterpret_cast<float*>(xs), red);
    mark();
    grid_sync(&st->barrier, target);
    mark();
    // ---- P3: o_proj + residual ----
    load_vec(static_cast<const __nv_bfloat16*>(p.attn), xs, H);
    mark();
    consume_phase(phase_desc(p, layer, PH_O), ring, full, empty, rs, xs, warp, lane, [&](int row, float a0, float a1) {
      const uint32_t r = ldg_cg_u32(resid_src + row);
      *reinterpret_cast<uint32_t*>(x + row) = pack_bf16(bf16_lo(r) + bf16_round(a0), bf16_hi(r) + bf16_round(a1));
    });
    mark();
    grid_sync(&st->barrier, target);
    mark();
    // ---- P4: RMSNorm + gate/up + SwiGLU ----
    load_rmsnorm(x, static_cast<const __nv_bfloat16*>(p.ln2) + static_cast<long>(layer) * H, xs, H, p.rms_eps, red);
    mark();
    consume_phase(phase_desc(p, layer, PH_GATEUP), ring, full, empty, rs, xs, warp, lane, [&](int row, float g, float u) {
      hbuf[row >> 1] = __float2bfloat16_rn(bf16_round(silu(bf16_round(g))) * bf16_round(u));
    });
    mark();
    grid_sync(&st->barrier, target);
    mark();
    // ---- P5: down_proj + residual ----
    load_vec(hbuf, xs, p.inter);
    mark();
    consume_phase(phase_desc(p, layer, PH_DOWN), ring, full, empty, rs, xs, warp, lane, [&](int row, float a0, float a1) {
      const uint32_t r = ldg_cg_u32(x + row);
      *reinterpret_cast<uint32_t*>(x + row) = pack_bf16(bf16_lo(r) + bf16_round(a0), bf16_hi(r) + bf16_round(a1));
    });
    mark();
    grid_sync(&st->barrier, target);
  }
  mark();

  // ---- final norm + lm_head + greedy argmax ----
  load_rmsnorm(x, static_cast<const __nv_bfloat16*>(p.final_norm), xs, H, p.rms_eps, red);
  mark();
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  consume_phase(phase_desc(p, 0, PH_LMHEAD), ring, full, empty, rs, xs, warp, lane, [&](int row, float a0, float a1) {
    const float v0 = bf16_round(a0), v1 = bf16_round(a1);
    if (p.logits_out) p.logits_out[row] = v0, p.logits_out[row + 1] = v1;
    if (v0 > best) best = v0, best_i = row;  // rows ascend within a warp: strict '>' keeps the lowest index
    if (v1 > best) best = v1, best_i = row + 1;
  });
  mark();
  if (dbg && tid == 0) dbg[15 * L + 8] = rs.waited;
  if (lane == 0) s_best[warp] = best, reinterpret_cast<int*>(s_best + 8)[warp] = best_i;
  cbar();
  if (tid == 0) {
    for (int w = 0; w < DEC_CWARPS; ++w) {
      const float v = s_best[w];
      const int i = reinterpret_cast<int*>(s_best + 8)[w];
      if (v > best || (v == best && i < best_i)) best = v, best_i = i;
    }
    p.argmax_part[2 * blockIdx.x] = best;
    reinterpret_cast<int*>(p.argmax_part)[2 * blockIdx.x + 1] = best_i;
  }
  grid_sync(&st->barrier, target);
  mark();
  if (blockIdx.x == 0 && warp == 0) {
    float b = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < static_cast<int>(gridDim.x); c += 32) {
      const float v = ldg_cg_f32(p.argmax_part + 2 * c);
      const int i = static_cast<int>(ldg_cg_u32(p.argmax_part + 2 * c + 1));
      if (v > b || (v == b && i < bi)) b = v, bi = i;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, b, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > b || (ob == b && oi < bi)) b = ob, bi = oi;
    }
    if (lane == 0) {
      p.out_tokens[n_gen] = bi;
      st->cur_token = bi;
      st->pos = pos + 1;
      st->n_generated = n_gen + 1;
      if (p.eos_token >= 0 && bi == p.eos_token) st->finished = 1;
      st->epoch = st->epoch + 1;
    }
  }
}

}  // namespace emx

extern "C" int emx_decode_grid(void) { return emx::kNumSMs; }

extern "C" int emx_decode_step(const emx_decode_params* params, cudaStream_t stream) {
  using namespace emx;
  const emx_decode_params& p = *params;
  EMX_REQUIRE(p.head_dim == DEC_HD, "emx_decode_step: head_dim %d not supported (128)", p.head_dim);
  EMX_REQUIRE(p.hidden % 8 == 0 && p.inter % 8 == 0 && p.vocab % 2 == 0, "emx_decode_step: hidden/inter must be multiples of 8, vocab even");
  EMX_REQUIRE(p.inter * 2 <= DEC_XS_BYTES && p.hidden * 2 <= DEC_XS_BYTES, "emx_decode_step: activation vector exceeds %d bytes", DEC_XS_BYTES);
  EMX_REQUIRE(p.kv_splits >= 1 && (p.kv_splits & (p.kv_splits - 1)) == 0, "emx_decode_step: kv_splits must be a power of two");
  EMX_REQUIRE(p.heads <= 64, "emx_decode_step: at most 64 heads");
  // score buffer lives in the activation area behind 896 floats of q/k/v/acc staging
  const int max_keys_per_split = (DEC_XS_BYTES / 4 - 896);
  EMX_REQUIRE((static_cast<long>(p.max_pages) * p.page_size + p.kv_splits - 1) / p.kv_splits + 1 <= max_keys_per_split,
              "emx_decode_step: context capacity %d x %d exceeds the per-split score buffer (%d keys)", p.max_pages, p.page_size,
              max_keys_per_split);
  static bool attr_set = false;
  static int grid = 0;
  if (!attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM));
    int dev = 0, sms = 0, per_sm = 0;
    EMX_CHECK_CUDA(cudaGetDevice(&dev));
    EMX_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    EMX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel, DEC_THREADS, DEC_SMEM));
    EMX_REQUIRE(per_sm >= 1, "emx_decode_step: kernel does not fit on an SM (smem %d)", DEC_SMEM);
    grid = sms;
    attr_set = true;
  }
  void* args[] = {const_cast<emx_decode_params*>(params)};
  EMX_CHECK_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(decode_step_kernel), dim3(grid), dim3(DEC_THREADS), args, DEC_SMEM, stream));
  return 0;
}
