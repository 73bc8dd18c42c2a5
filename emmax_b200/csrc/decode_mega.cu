// emx_decode_step — one greedy decode step of Llama-2-7B for one sequence as ONE persistent kernel.
//
// Replaces the cached branch of PrismaticForConditionalGeneration.forward
// (/root/reference/prismatic/extern/hf/modeling_prismatic.py:325-341: language_model(input_ids[:, -1:], past_key_values))
// plus one iteration of GenerationMixin's greedy loop (called at modeling_prismatic.py:519), which in the reference is
// ~10^3 kernel launches per token, a full re-copy of the KV cache (DynamicCache torch.cat) and a host sync.
//
// B200 design (HBM-bound: 13.2 GB of bf16 weights per token):
//   * grid = one CTA per SM (148) x 512 threads at 128 registers: 8 consumer warps + 4 attention warps + 2 producer warps +
//     1 L2-prefetch warp (+ 1 idle), launched cooperatively (co-residency is what makes the spin-waits below safe);
//   * the producer warps walk the STATIC weight schedule of their CTA (layer -> q, k, v, o, gate/up, down -> 16-row group ->
//     2048-column chunk) and keep a 3 x 64 KB shared-memory ring full with cp.async.bulk (TMA engine) copies, 16 x 4 KB per
//     instruction, L2 evict-first. Producers never wait for anything but a free ring slot;
//   * consumers: the 16 rows x 2048 columns of a stage are A operands of mma.sync.m16n8k16 (bf16, fp32 accumulate); each warp
//     owns a 256-column slice (ldmatrix from a 16-B padded, conflict-free row stride), the activation vector is the B operand
//     (one 16-B shared load per two k-steps thanks to a permuted layout), four independent accumulator chains; per-warp partial
//     row sums are handed to a ROTATING epilogue warp through named barriers (a fixed one would trail the others and hold
//     every ring slot longer);
//   * NO GRID BARRIERS. A 148-CTA barrier costs ~1.9 us on B200 even on an idle memory system (tools/skeleton_probe.py) and a
//     decode step needs 160 of them. Instead every cross-CTA vector (q|k|v, attention output, residual stream, SwiGLU output,
//     split-KV partials, argmax candidates) travels as 8-byte "LL" units {2 x bf16 | 32-bit tag} written with one 64-bit
//     store and read with polling 64-bit loads: data and flag arrive atomically, so no fence, no atomic and no barrier is
//     needed; a consumer waits exactly for the words it needs, ~one L2 round trip after they were produced. The tag encodes
//     (launch epoch, layer), so stale words of the previous layer / token never match. Reuse of a buffer is safe without
//     further synchronisation because every phase gathers a COMPLETE vector before it produces anything (see ll_gather);
//   * RMSNorm is recomputed per CTA from the 8 KB residual vector (cheaper than an extra exchange);
//   * attention runs on its OWN warps, concurrently with the consumers (see "attention warps" below): q rows are projected
//     first, the cached-key work happens in the shadow of the k/v weight stream out of K/V rows staged in TENSOR MEMORY a layer
//     ahead, and only the new token's k/v -> owner split -> head output -> o_proj chain is left on the critical path;
//   * an L2-prefetch warp keeps HBM busy while the consumers wait for an exchange and the ring is full.
// Rounding points mirror the torch-eager reference (bf16 after every Linear / norm / residual add / activation).
#include "decode_common.cuh"
#include "emmax.h"

namespace emx {

constexpr int DEC_PWARPS = 2;                    // producer warps (alternate ring stages)
constexpr int DEC_AWARPS = 4;                    // attention warps (their own dataflow, concurrent with the consumers)
constexpr int DEC_ATHREADS = DEC_AWARPS * 32;    // 128
constexpr int DEC_THREADS = 512;                 // 8 consumer + 4 attention + 2 producer + 1 L2-prefetch warp (+ 1 idle): 128 registers / thread
constexpr int DEC_MAX_RESID = 192;               // residual pairs of one CTA staged in shared memory
constexpr int DEC_XS_BYTES = 22528;                         // activation vector (bf16), up to 11264 elements
constexpr int DEC_MISC_BYTES = 4096;
constexpr int DEC_LN_BYTES = 8192;              // norm weights of the NEXT RMSNorm, fetched asynchronously a phase ahead (hidden <= 4096)
constexpr int DEC_SMEM = DEC_STAGES * DEC_STAGE_BYTES + DEC_XS_BYTES + DEC_MISC_BYTES + 128 + DEC_LN_BYTES;
constexpr int DEC_MAX_PAGES = 64;  // block-table entries staged in shared memory

// q, k and v rows are three phases of their own, q FIRST: every CTA finishes its share of the q rows a third of the way into the
// fused projection, so the attention warps get q ~7 us before k/v of the new token exist and do all the cached-key work in the shadow
// of the k/v weight stream.
constexpr int PH_STEPS = 6;  // weight phases per layer
enum PhaseKind { PH_Q = 0, PH_K = 1, PH_V = 2, PH_O = 3, PH_GATEUP = 4, PH_DOWN = 5, PH_LMHEAD = 6, PH_END = 7 };

struct PhaseDesc {
  const __nv_bfloat16* W;
  int N, K, kind, r_begin, r_end;  // [r_begin, r_end): rows of this phase owned by this CTA
};

// Per-CTA table of the five weight phases, built once per launch in shared memory (keeps 64-bit multiplies / divisions and a
// switch out of every phase transition: the kernel's instruction footprint matters, see the note at decode_step_kernel).
struct PhaseTab {
  const __nv_bfloat16* W[7];
  long layer_stride[7];  // elements between consecutive layers
  int N[7], K[7], r_begin[7], r_end[7];
};

// Rows of a phase owned by this CTA. Row pairs are never split (one LL unit = one row pair); gate/up rows come in groups of
// four (two SwiGLU outputs = one LL unit).
__device__ __forceinline__ void build_phase_tab(const emx_decode_params& p, PhaseTab& t, int kind) {
  const long H = p.hidden, I = p.inter;
  switch (kind) {
    case PH_Q:
    case PH_K:
    case PH_V: t.W[kind] = static_cast<const __nv_bfloat16*>(p.w_qkv) + kind * H * H, t.layer_stride[kind] = 3 * H * H, t.N[kind] = H, t.K[kind] = H; break;
    case PH_O: t.W[kind] = static_cast<const __nv_bfloat16*>(p.w_o), t.layer_stride[kind] = H * H, t.N[kind] = H, t.K[kind] = H; break;
    case PH_GATEUP: t.W[kind] = static_cast<const __nv_bfloat16*>(p.w_gateup), t.layer_stride[kind] = 2 * I * H, t.N[kind] = 2 * I, t.K[kind] = H; break;
    case PH_DOWN: t.W[kind] = static_cast<const __nv_bfloat16*>(p.w_down), t.layer_stride[kind] = H * I, t.N[kind] = H, t.K[kind] = I; break;
    default: t.W[kind] = static_cast<const __nv_bfloat16*>(p.lm_head), t.layer_stride[kind] = 0, t.N[kind] = p.vocab, t.K[kind] = H; break;
  }
  phase_rows(t.N[kind], (kind == PH_GATEUP) ? 4 : 2, blockIdx.x, gridDim.x, t.r_begin[kind], t.r_end[kind]);
}

__device__ __forceinline__ PhaseDesc phase_desc(const PhaseTab& t, int layer, int kind) {
  PhaseDesc d;
  d.W = t.W[kind] + layer * t.layer_stride[kind];
  d.N = t.N[kind], d.K = t.K[kind], d.kind = kind, d.r_begin = t.r_begin[kind], d.r_end = t.r_end[kind];
  return d;
}

// Gather a whole vector of `n_units` (even) LL units tagged `tag`: thread t takes the unit pairs t, t + 256, ...; all loads of a
// chunk of MAXP pairs are in flight before the first tag is checked, pairs that are not there yet are re-polled.
// sink(u, word) is called exactly once per unit (by the thread that fetched it).
// WHY BUFFER REUSE IS SAFE: a CTA produces outputs of phase n+1 only after gathering ALL of phase n's vector, so when any
// word of phase n+2 (or of the same phase one layer later) is overwritten, every CTA has long finished reading phase n.
template <int MAXP, typename Sink>
__device__ __forceinline__ void ll_gather(const uint64_t* buf, int n_units, uint32_t tag, bool check, Sink&& sink) {
  const int n_pairs = n_units >> 1;
  for (int base = 0; base < n_pairs; base += MAXP * DEC_CTHREADS) {
    uint64_t a[MAXP], b[MAXP];
    uint32_t pending = 0;
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int pr = base + i * DEC_CTHREADS + static_cast<int>(threadIdx.x);
      if (pr < n_pairs) {
        ll_load2(buf + 2 * pr, a[i], b[i]);
        pending |= 1u << i;
      }
    }
    uint32_t spins = 0;
    while (pending) {
#pragma unroll
      for (int i = 0; i < MAXP; ++i) {
        if (pending & (1u << i)) {
          const int pr = base + i * DEC_CTHREADS + static_cast<int>(threadIdx.x);
          if (!check || (static_cast<uint32_t>(a[i] >> 32) == tag && static_cast<uint32_t>(b[i] >> 32) == tag)) {
            pending &= ~(1u << i);
            sink(2 * pr, static_cast<uint32_t>(a[i]));
            sink(2 * pr + 1, static_cast<uint32_t>(b[i]));
          } else {
            ll_load2(buf + 2 * pr, a[i], b[i]);
          }
        }
      }
      if (++spins > EMX_SPIN_LIMIT) __trap();
    }
  }
}

// Block sum over the 8 consumer warps with ONE barrier: per-warp partials go to one of two 8-float buffers (alternating per call, so
// the next call's writes cannot overtake a slow reader of this call: there is at least one block barrier between two calls), every
// thread adds the 8 partials itself.
__device__ __forceinline__ float cblock_sum(float v, float* red, uint32_t parity) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* r = red + 8 * (parity & 1);
  if (lane == 0) r[warp] = v;
  cbar();
  const float4 a = *reinterpret_cast<const float4*>(r), b = *reinterpret_cast<const float4*>(r + 4);
  return ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
}

// The 8 KB norm-weight vectors are streamed once per token, so they are never L2-resident when they are needed, and an
// L2 prefetch hint is dropped while HBM is saturated. Each thread therefore copies exactly the words IT will need for the
// next RMSNorm into shared memory with cp.async a whole phase ahead (no cross-thread hand-off: the thread that copied a word
// is the thread that reads it after cp.async.wait_group).
__device__ __forceinline__ void ln_fetch_async(const __nv_bfloat16* w, uint2* ln_s, int H) {
  const int n_pairs = H >> 2;
  for (int pr = threadIdx.x; pr < n_pairs; pr += DEC_CTHREADS)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(ln_s + pr)), "l"(reinterpret_cast<const uint2*>(w) + pr) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Residual vector in (LL units, or a plain bf16 row for layer 0) -> LlamaRMSNorm -> xs, in ONE pass over registers:
//   xs = bf16(w * bf16(x * rsqrt(mean(x^2) + eps)))
// Thread t owns the unit pairs t, t + 256, ... (<= 4 pairs: hidden <= 4096, the host checks). The norm weights were copied to
// `ln_s` by ln_fetch_async; own(u, word) sees every raw word once.
template <typename Own>
__device__ __forceinline__ void gather_rmsnorm(const uint64_t* ll, const uint32_t* plain, int H, uint32_t tag, bool check,
                                               const uint2* ln_s, float eps, uint32_t* xs, float* red, uint32_t parity, Own&& own, long long* prof = nullptr) {
  const long long tp0 = prof ? clock64() : 0;
  constexpr int MAXP = 4;
  const int n_pairs = H >> 2;
  uint2 g[MAXP];
  uint64_t a[MAXP], b[MAXP];
  uint32_t pending = 0;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int pr = i * DEC_CTHREADS + static_cast<int>(threadIdx.x);
    g[i] = make_uint2(0, 0), a[i] = b[i] = 0;
    if (pr < n_pairs) {
      if (plain) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(plain) + pr);
        a[i] = v.x, b[i] = v.y;
      } else {
        ll_load2(ll + 2 * pr, a[i], b[i]);
        pending |= 1u << i;
      }
    }
  }
  uint32_t spins = 0;
  while (pending) {
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      if (pending & (1u << i)) {
        if (!check || (static_cast<uint32_t>(a[i] >> 32) == tag && static_cast<uint32_t>(b[i] >> 32) == tag))
          pending &= ~(1u << i);
        else
          ll_load2(ll + 2 * (i * DEC_CTHREADS + static_cast<int>(threadIdx.x)), a[i], b[i]);
      }
    }
    if (++spins > EMX_SPIN_LIMIT) __trap();
  }
  const long long tp1 = prof ? clock64() : 0;
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int pr = i * DEC_CTHREADS + static_cast<int>(threadIdx.x);
    if (pr < n_pairs) g[i] = ln_s[pr];
  }
  const long long tp2 = prof ? clock64() : 0;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) ss += sumsq2(static_cast<uint32_t>(a[i])) + sumsq2(static_cast<uint32_t>(b[i]));  // absent pairs are 0
  const float rs = 1.0f / sqrtf(cblock_sum(ss, red, parity) / H + eps);
  const long long tp3 = prof ? clock64() : 0;
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int pr = i * DEC_CTHREADS + static_cast<int>(threadIdx.x);
    if (pr < n_pairs) {
      const uint32_t v0 = static_cast<uint32_t>(a[i]), v1 = static_cast<uint32_t>(b[i]);
      own(2 * pr, v0), own(2 * pr + 1, v1);
      xs[xs_pos(2 * pr)] = pack_bf16(bf16_lo(g[i].x) * bf16_round(bf16_lo(v0) * rs), bf16_hi(g[i].x) * bf16_round(bf16_hi(v0) * rs));
      xs[xs_pos(2 * pr + 1)] = pack_bf16(bf16_lo(g[i].y) * bf16_round(bf16_lo(v1) * rs), bf16_hi(g[i].y) * bf16_round(bf16_hi(v1) * rs));
    }
  }
  cbar();
  if (prof) {
    const long long tp4 = clock64();
    prof[0] += tp1 - tp0, prof[1] += tp2 - tp1, prof[2] += tp3 - tp2, prof[3] += tp4 - tp3;
  }
}

// ---- static weight schedule of one CTA: (layer, kind) phases -> 16-row groups ------------------------------------------
struct SchedIter {
  int layer, kind, r, r_end, layers;
  PhaseDesc d;
  __device__ __forceinline__ void load_phase(const PhaseTab& t) {
    d = phase_desc(t, layer, kind);
    r = d.r_begin, r_end = d.r_end;
  }
  __device__ __forceinline__ bool done() const { return kind == PH_END; }
  __device__ __forceinline__ void next_phase(const PhaseTab& t) {
    if (kind == PH_LMHEAD) {
      kind = PH_END;
      return;
    }
    if (kind == PH_DOWN) {
      kind = PH_Q;
      if (++layer == layers) kind = PH_LMHEAD;
    } else {
      ++kind;
    }
    load_phase(t);
  }
  __device__ __forceinline__ void skip_empty(const PhaseTab& t) {
    while (!done() && r >= r_end) next_phase(t);
  }
  __device__ __forceinline__ void init(const PhaseTab& t, int n_layers) {
    layer = 0, kind = PH_Q, layers = n_layers;
    load_phase(t);
    skip_empty(t);
  }
  __device__ __forceinline__ void advance(const PhaseTab& t) {
    r += DEC_GROUP;
    skip_empty(t);
  }
  __device__ __forceinline__ int nrows() const { return min(DEC_GROUP, r_end - r); }
  __device__ __forceinline__ long group_bytes() const { return static_cast<long>(nrows()) * d.K * 2; }
};

// ---- producer warps -------------------------------------------------------------------------------------------------
// Both producer warps walk the same schedule; warp `pidx` issues the ring stages with it % DEC_PWARPS == pidx.
__device__ void producer_loop(const emx_decode_params& p, int dflags, const PhaseTab& tab, uint8_t* ring, uint64_t* full, uint64_t* empty,
                                              int lane, int pidx, volatile uint32_t* s_groups_issued, long long* dbg) {
  const uint64_t policy = (dflags & 4) ? l2_policy_evict_last() : l2_policy_evict_first();
  SchedIter cur;
  cur.init(tab, p.layers);
  uint32_t it = 0, groups = 0;
  long long waited = 0;
  while (!cur.done()) {
    const int nrows = cur.nrows();
    for (int k0 = 0; k0 < cur.d.K; k0 += DEC_KC) {
      const int klen = min(DEC_KC, cur.d.K - k0);
      const int slot = it % DEC_STAGES;
      const uint32_t ph = (it / DEC_STAGES) & 1;
      if ((it % DEC_PWARPS) == static_cast<uint32_t>(pidx)) {
        if (lane == 0) {
          const long long t0 = dbg ? clock64() : 0;
          mbar_wait(&empty[slot], ph ^ 1);
          if (dbg) waited += clock64() - t0;
          mbar_arrive_expect_tx(&full[slot], static_cast<uint32_t>(nrows) * klen * 2);
        }
        __syncwarp();
        if (lane < nrows)
          bulk_g2s(ring + slot * DEC_STAGE_BYTES + lane * DEC_ROWSTRIDE, cur.d.W + static_cast<long>(cur.r + lane) * cur.d.K + k0,
                   klen * 2, &full[slot], policy);
        if (lane == 0) s_groups_issued[pidx] = it + 1;  // ring stages issued by this producer (for the L2 prefetch warp)
      }
      ++it;
    }
    ++groups;
    cur.advance(tab);
  }
  if (dbg && lane == 0) dbg[15 * p.layers + 9 + pidx] = waited;
}

// ---- L2 prefetch warp ---------------------------------------------------------------------------------------------------
// The 192 KB ring only just covers the bandwidth-delay product of a saturated HBM (~45 KB/us per SM x ~4 us loaded latency),
// so whenever the consumers stall (grid barrier, attention) the ring fills, no new copy can be issued and HBM idles
// (tools/skeleton_probe.py: stalls are 100 % exposed without prefetch). This warp watches for exactly that condition — the
// newest ring copy of this CTA has LANDED, i.e. nothing of ours is in flight — and then pulls the next groups of the static
// schedule HBM -> L2 with cp.async.bulk.prefetch.L2, paced at about twice the SM's fair share, at most `l2_lookahead_kb`
// ahead of the ring. In the HBM-bound steady state it never triggers, so it costs nothing there.
__device__ void prefetch_loop(const emx_decode_params& p, int dflags, const PhaseTab& tab, int lane, volatile uint32_t* s_issued, uint64_t* full,
                                              long long* dbg) {
  const long lookahead = static_cast<long>(p.l2_lookahead_kb) * 1024;
  if (lookahead <= 0) return;
  const long long pace_ns = ((dflags >> 8) & 0xfff) ? ((dflags >> 8) & 0xfff) * 10 : 700;  // per 64 KB while the ring is idle
  const long long pace_catchup_ns = (dflags >> 20) ? (dflags >> 20) * 10 : 1300;            // per 64 KB while the ring drains L2
  const bool catchup = dflags & 8;  // measured neutral-to-slightly-negative on B200 (profiles/r01_decode_v5_prefetch_modes.txt): off by default
  SchedIter cur, pf;
  cur.init(tab, p.layers);
  pf.init(tab, p.layers);
  uint32_t cur_stage = 0;  // first ring-stage index of the group `cur` points at
  long ahead = 0;          // bytes between the start of `cur` and the start of `pf`
  long pf_off = 0;         // progress inside the group `pf` points at
  long pf_total = 0;
  bool live = false;       // data prefetched after the last idle period is still ahead of the ring
  while (!pf.done()) {
    // follow the producers: s_issued = number of ring stages issued so far
    const uint32_t issued = max(s_issued[0], s_issued[1]);
    while (!cur.done()) {
      const uint32_t st = (cur.d.K + DEC_KC - 1) / DEC_KC;
      if (cur_stage + st > issued) break;
      cur_stage += st;
      ahead -= cur.group_bytes();
      cur.advance(tab);
    }
    if (cur.done()) break;  // everything is in the ring already
    // lead of the prefetch point over the ring's issue point, in bytes
    const uint32_t st_cur = (cur.d.K + DEC_KC - 1) / DEC_KC;
    const long ring_off = static_cast<long>(issued - cur_stage) * (cur.group_bytes() / st_cur);
    long lead = ahead + pf_off - ring_off;
    if (ahead <= 0 || lead <= 0) {  // the ring has caught up with (or passed) the prefetch point: restart after the group it is loading
      pf = cur, pf_off = 0;
      ahead = pf.group_bytes();
      pf.advance(tab);
      if (pf.done()) break;
      live = false;
      lead = ahead - ring_off;
    }
    // Two triggers. IDLE: the newest ring copy of this CTA has landed, nothing of ours is in flight (consumers are stalled and the
    // ring is full). CATCH-UP: the consumers are running again and the ring is refilling out of the L2 lines prefetched during the
    // stall (lead > 0) — those copies cost no HBM time, so HBM would idle until the ring reaches the end of the prefetched
    // region; keep streaming ahead at about the SM's fair share of HBM until the ring overtakes the prefetch point, which is
    // the HBM-bound steady state where this warp stays out of the way.
    bool go = false, idle = false;
    if (issued > 0 && lead < lookahead) {
      if (catchup && live) {
        go = true;
      } else {
        const uint32_t last = issued - 1;
        go = idle = mbar_try_wait(&full[last % DEC_STAGES], (last / DEC_STAGES) & 1);
      }
    }
    if (!go) {
      __nanosleep(200);
      continue;
    }
    const long long t0 = global_ns();
    const long gb = pf.group_bytes();
    const long n = min(static_cast<long>(64 * 1024), gb - pf_off);
    {
      const char* src = reinterpret_cast<const char*>(pf.d.W + static_cast<long>(pf.r) * pf.d.K) + pf_off;
      const long per_lane = ((n + 31) / 32 + 15) & ~15L;
      const long off = per_lane * lane;
      if (off < n) prefetch_l2(src + off, static_cast<uint32_t>(min(per_lane, n - off)));
    }
    live = true;
    pf_total += n;
    pf_off += n;
    if (pf_off >= gb) {
      ahead += gb;
      pf_off = 0;
      pf.advance(tab);
    }
    const long long wait_ns = (idle ? pace_ns : pace_catchup_ns) * n / (64 * 1024);
    while (global_ns() - t0 < wait_ns) __nanosleep(100);
  }
  if (dbg && lane == 0) dbg[15 * p.layers + 13] = pf_total;
}

// ---- consumer: tensor-core dot products of one phase ---------------------------------------------------------------------
struct ConsumerState {
  uint32_t it;
  uint32_t group;  // selects the partial-sum buffer / named barrier
  long long waited;        // cycles in mbar_wait(full) (thread 0: warp 0)
  long long t_sync, t_epi;  // cycles warp 0 spent waiting for the other warps' partial sums / in the epilogue
};

// epi(row, v0, v1, valid) is called for row pairs (row even) by lanes 0..7 of ONE warp per row group (all eight lanes, converged), rows ascending
// per lane; valid == false marks lanes beyond the last row of a short group
template <bool PROF, typename Epi>
__device__ __forceinline__ void consume_phase(const PhaseDesc& d, const uint8_t* ring, uint64_t* full, uint64_t* empty, ConsumerState& cs,
                                              const __nv_bfloat16* xs, float* part, int warp, int lane, int debug_flags, Epi&& epi) {
  const int r_begin = d.r_begin, r_end = d.r_end;
  // ldmatrix.x4 row address of this lane: matrices (rows 0-7 | 8-15) x (cols 0-7 | 8-15) of a 16x16 A tile
  const uint32_t a_lane_off = ((lane & 7) + ((lane >> 3) & 1) * 8) * DEC_ROWSTRIDE + (lane >> 4) * 16;
  const int kbeg = warp * DEC_KW;
  for (int r0 = r_begin; r0 < r_end; r0 += DEC_GROUP) {
    const int nrows = min(DEC_GROUP, r_end - r0);
    // four independent accumulator chains (k-step % 4): the HMMA dependency latency is off the critical path
    float c[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.f;
    for (int k0 = 0; k0 < d.K; k0 += DEC_KC) {
      const int klen = min(DEC_KC, d.K - k0);
      const int slot = cs.it % DEC_STAGES;
      const uint32_t ph = (cs.it / DEC_STAGES) & 1;
      const int ksteps = min(DEC_KW / 16, (klen - kbeg) / 16);  // <= 0: this warp's slice is past the K tail
      const uint4* xw = reinterpret_cast<const uint4*>(xs + k0 + kbeg) + (lane & 3);
      const long long t0 = PROF ? clock64() : 0;
      mbar_wait(&full[slot], ph);
      if (PROF) cs.waited += clock64() - t0;
      if (ksteps > 0 && !(debug_flags & 64)) {
        const uint32_t a_base = smem_u32(ring + slot * DEC_STAGE_BYTES) + a_lane_off + kbeg * 2;
        if (ksteps == DEC_KW / 16) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two batches of 8 k-steps: 8 ldmatrix in flight, then 8 HMMA
            uint32_t a[8][4];
            uint4 b[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = xw[(h * 4 + j) * 4];
#pragma unroll
            for (int j = 0; j < 8; ++j) ldmatrix_x4(a_base + (h * 8 + j) * 32, a[j][0], a[j][1], a[j][2], a[j][3]);
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const uint4 bb = b[j / 2];
              mma_bf16_16816(c[j & 3], a[j][0], a[j][1], a[j][2], a[j][3], bb.x, bb.y);
              mma_bf16_16816(c[(j + 1) & 3], a[j + 1][0], a[j + 1][1], a[j + 1][2], a[j + 1][3], bb.z, bb.w);
            }
          }
        } else {
          for (int j = 0; j < ksteps; ++j) {
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4(a_base + j * 32, a0, a1, a2, a3);
            const uint4 bb = xw[(j >> 1) * 4];
            mma_bf16_16816(c[0], a0, a1, a2, a3, (j & 1) ? bb.z : bb.x, (j & 1) ? bb.w : bb.y);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      ++cs.it;
    }
    // column 0 of the accumulator tile lives in lanes with lane % 4 == 0: c[0] -> row lane/4, c[2] -> row lane/4 + 8
    const uint32_t buf = cs.group % DEC_PARTBUFS;
    float* pb = part + buf * (DEC_CWARPS * DEC_GROUP);
    if ((lane & 3) == 0) {
      pb[warp * DEC_GROUP + (lane >> 2)] = (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
      pb[warp * DEC_GROUP + (lane >> 2) + 8] = (c[0][2] + c[1][2]) + (c[2][2] + c[3][2]);
    }
    // the epilogue duty rotates over the warps (group g -> warp g % 8): the ring slot of a stage is released by the SLOWEST warp, so a
    // fixed epilogue warp would trail the others by one epilogue per row group and hold every slot that much longer
    if (warp == static_cast<int>(cs.group % DEC_CWARPS)) {
      const long long t1 = PROF ? clock64() : 0;
      part_sync(buf);
      const long long t2 = PROF ? clock64() : 0;
      if (lane < 8) {
        float v0 = 0.f, v1 = 0.f;
#pragma unroll
        for (int w = 0; w < DEC_CWARPS; ++w) v0 += pb[w * DEC_GROUP + 2 * lane], v1 += pb[w * DEC_GROUP + 2 * lane + 1];
        epi(r0 + 2 * lane, v0, v1, 2 * lane < nrows);
      }
      __syncwarp();
      if (PROF) cs.t_sync += t2 - t1, cs.t_epi += clock64() - t2;
    } else {
      part_arrive(buf);
    }
    ++cs.group;
  }
}

// ---- attention for one (head, split) item ------------------------------------------------------------------------------
// Paged KV cache addressing. page_size is a power of two (host-checked); KvAddr holds everything that is constant for one
// (layer, head), so that a row offset costs one table load, one IMAD.WIDE and a shift-add.
struct KvAddr {
  long base;         // element offset of (layer, page 0, head, slot 0)
  int page_stride;   // elements between consecutive pages: heads * page_size * head_dim
  int shift, mask;
  const int32_t* table;  // block table of the sequence, staged in SHARED memory at kernel start (no dependent global load per row)
  __device__ __forceinline__ long row(int key) const {
    return base + static_cast<long>(table[key >> shift]) * page_stride + ((key & mask) << 7);  // << 7: * DEC_HD
  }
};
__device__ __forceinline__ KvAddr kv_addr(const emx_decode_params& p, const int32_t* s_table, int layer, int head) {
  KvAddr a;
  a.shift = 31 - __clz(p.page_size), a.mask = p.page_size - 1, a.table = s_table;
  a.page_stride = (p.heads << a.shift) * DEC_HD;
  a.base = (static_cast<long>(layer) * p.n_pages * p.heads + head) * (static_cast<long>(DEC_HD) << a.shift);
  return a;
}

// ---- attention warps -------------------------------------------------------------------------------------------------------
// One (head, kv-split) item per CTA and layer, run by 4 dedicated warps CONCURRENTLY with the consumer warps. Per layer:
//   1. (long before q exists: the cached rows do not depend on the token being decoded) every cached K / V row of the item is
//      pulled HBM -> registers -> TENSOR MEMORY. The kernel's GEMVs run on mma.sync, so the SM's 256 KB of TMEM is free; each
//      attention thread uses its own TMEM lane as 1 KB of extra register file (tcgen05.st / tcgen05.ld, 32 lanes x 32 columns
//      per pass). All loaded-HBM latency (2-3 us per dependent access while the weight stream saturates the memory system) is paid
//      here, in the ~45 us these warps would otherwise idle;
//   2. q arrives as LL units a third of the way into the q|k|v projection -> RoPE -> scores of all cached keys, softmax
//      statistics, P·V out of TMEM — while the consumers of every CTA are still streaming the k and v rows;
//   3. the splits that do not contain the new token publish their partial (m, l, acc) right away; the LAST split (which does)
//      waits for k, v of the new token (RoPE, KV append, one online-softmax step), merges the other splits' partials — long
//      there by then — and publishes the head's output.
// The consumers only ever see the finished attention vector (ll_gather before o_proj).
constexpr int ATT_PASS = 64;                     // keys per pass: 2 threads per key (K), 4 key slices x 32 quads (V)
constexpr int ATT_MAX_PASSES = 8;                // <= 512 cached keys per split (host-checked): contexts up to 2048 = the reference's llm_max_length
constexpr int ATT_KPT = ATT_PASS * ATT_MAX_PASSES / DEC_ATHREADS;  // keys per thread in the softmax-statistics step
constexpr int ATT_SQ = 0, ATT_SQ2 = 68 /* second half of q, skewed by 4 banks */, ATT_SVNEW = 136, ATT_SACC = 264 /*[4][128]*/, ATT_SSCORE = 776;
constexpr int ATT_SM_FLOATS = ATT_SSCORE + ATT_PASS * ATT_MAX_PASSES;  // 1032 floats
constexpr int ATT_SM_OFFSET = 16384;             // inside the activation area: bytes [16 K, 22 K) are only used by the down_proj input
constexpr int ATT_TMEM_COLS = 2 * 32 * ATT_MAX_PASSES;  // K passes | V passes, 32 columns (128 B per thread) each

__device__ __forceinline__ void abar() { asm volatile("bar.sync 6, %0;" ::"n"(DEC_ATHREADS) : "memory"); }

__device__ __forceinline__ void att_load_k(uint32_t (&b)[32], const __nv_bfloat16* kc, const KvAddr& ka, int k_begin, int n_old, int pass, int atid) {
  const int idx = pass * ATT_PASS + (atid >> 1);
  if (idx < n_old) {
    const uint4* kr = reinterpret_cast<const uint4*>(kc + ka.row(k_begin + idx) + (atid & 1) * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 v = ldg_nc_v4(kr + j);
      b[4 * j] = v.x, b[4 * j + 1] = v.y, b[4 * j + 2] = v.z, b[4 * j + 3] = v.w;
    }
  }
}
__device__ __forceinline__ void att_load_v(uint32_t (&b)[32], const __nv_bfloat16* vc, const KvAddr& ka, int k_begin, int n_old, int pass, int atid) {
  const int slice = atid >> 5, quad = atid & 31;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int idx = pass * ATT_PASS + slice + 4 * u;
    uint2 v = make_uint2(0, 0);
    if (idx < n_old) v = ldg_nc_v2(vc + ka.row(k_begin + idx) + 4 * quad);
    b[2 * u] = v.x, b[2 * u + 1] = v.y;
  }
}
__device__ __forceinline__ void att_score(const uint32_t (&b)[32], float* sm, int n_old, int pass, int atid, float scale) {
  const int idx = pass * ATT_PASS + (atid >> 1);
  const float* q = sm + ((atid & 1) ? ATT_SQ2 : ATT_SQ);
  float d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 qa = *reinterpret_cast<const float4*>(q + 8 * j), qb = *reinterpret_cast<const float4*>(q + 8 * j + 4);
    d0 = fmaf(qa.x, bf16_lo(b[4 * j]), d0), d1 = fmaf(qa.y, bf16_hi(b[4 * j]), d1);
    d0 = fmaf(qa.z, bf16_lo(b[4 * j + 1]), d0), d1 = fmaf(qa.w, bf16_hi(b[4 * j + 1]), d1);
    d0 = fmaf(qb.x, bf16_lo(b[4 * j + 2]), d0), d1 = fmaf(qb.y, bf16_hi(b[4 * j + 2]), d1);
    d0 = fmaf(qb.z, bf16_lo(b[4 * j + 3]), d0), d1 = fmaf(qb.w, bf16_hi(b[4 * j + 3]), d1);
  }
  float d = d0 + d1;
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  if (idx < n_old && !(atid & 1)) sm[ATT_SSCORE + idx] = d * scale;
}
__device__ __forceinline__ void att_pv(const uint32_t (&b)[32], const float* sm, int n_old, int pass, int atid, float (&a)[4]) {
  const int slice = atid >> 5;
#pragma unroll
  for (int u = 0; u < 16; ++u) {
    const int idx = pass * ATT_PASS + slice + 4 * u;
    const float pw = (idx < n_old) ? sm[ATT_SSCORE + idx] : 0.f;
    a[0] = fmaf(pw, bf16_lo(b[2 * u]), a[0]), a[1] = fmaf(pw, bf16_hi(b[2 * u]), a[1]);
    a[2] = fmaf(pw, bf16_lo(b[2 * u + 1]), a[2]), a[3] = fmaf(pw, bf16_hi(b[2 * u + 1]), a[3]);
  }
}

template <bool PROF>
__device__ void attention_loop(const emx_decode_params& p, int dflags, const int32_t* s_table, const uint32_t* s_rope, int pos, uint32_t tag0, bool check,
                               float* sm, float* red, uint32_t* tmem_holder, volatile int* s_step, int atid) {
  constexpr int HALF = DEC_HD / 2;
  const int item = blockIdx.x, S = p.kv_splits;
  if (item >= p.heads * S || (dflags & 2)) return;
  const int head = item / S, split = item % S;
  const int awarp = atid >> 5, lane = atid & 31;
  const int n = pos + 1;
  const int k_begin = static_cast<int>(static_cast<long>(n) * split / S), k_end = static_cast<int>(static_cast<long>(n) * (split + 1) / S);
  const bool owns_new = (split == S - 1);         // the last split contains the token being decoded (its last key) and combines
  const int n_old = k_end - k_begin - (owns_new ? 1 : 0);  // cached keys of this split: k_begin .. k_begin + n_old - 1
  const int passes = (n_old + ATT_PASS - 1) / ATT_PASS;
  const __nv_bfloat16* kc = static_cast<const __nv_bfloat16*>(p.k_cache);
  const __nv_bfloat16* vc = static_cast<const __nv_bfloat16*>(p.v_cache);
  const uint64_t* qkv = static_cast<const uint64_t*>(p.qkv);
  const int H = p.hidden, L = p.layers;
  const float scale = rsqrtf(static_cast<float>(DEC_HD));
  long long* dbg = (PROF && blockIdx.x == 0 && atid == 0 && p.dbg) ? reinterpret_cast<long long*>(p.dbg) + 15 * L + 16 + 148 : nullptr;
  long long t_q = 0, t_old = 0, t_new = 0, t_pub = 0, t_sc = 0, t_sm = 0, t_pre = 0;

  // tensor memory: K pass i at columns [32 i, 32 i + 32), V pass i at [32 (MAX_PASSES + i), ...); a warp owns TMEM lanes 32 (warp % 4) ..
  if (awarp == 0) tmem_alloc(tmem_holder, ATT_TMEM_COLS);
  tc_fence_before();
  abar();
  tc_fence_after();
  const uint32_t tbase = *tmem_holder + (static_cast<uint32_t>(((DEC_CWARPS + awarp) & 3) * 32) << 16);

  uint32_t buf[32];
  auto stage = [&](int layer) {  // HBM -> registers -> TMEM for every cached row of the item (the long-latency part, off the critical path)
    const KvAddr a = kv_addr(p, s_table, layer, head);
#pragma unroll 1
    for (int ps = 0; ps < passes; ++ps) {
      att_load_k(buf, kc, a, k_begin, n_old, ps, atid);
      tmem_st_32x32(tbase + 32 * ps, buf);
      att_load_v(buf, vc, a, k_begin, n_old, ps, atid);
      tmem_st_32x32(tbase + 32 * (ATT_MAX_PASSES + ps), buf);
    }
    tmem_st_wait();
  };
  stage(0);

#pragma unroll 1
  for (int layer = 0; layer < L; ++layer) {
    const uint32_t tag = tag0 + layer;
    const long long ts0 = PROF ? global_ns() : 0;
    // ---- q: LL units (unit = 2 consecutive elements); lane t of warp 0 owns units t and t + 32, i.e. elements (2t, 2t+1) and their
    // rotate_half partners (2t+64, 2t+65)
    if (awarp == 0) {
      const uint64_t* src = qkv + head * HALF;
      uint32_t lo, hi;
      ll_wait2<true>(src + lane, src + lane + 32, tag, check, lo, hi);  // ~40 us of waiting per layer
      const float x1a = bf16_lo(lo), x1b = bf16_hi(lo), x2a = bf16_lo(hi), x2b = bf16_hi(hi);
      const uint32_t cw = s_rope[lane], sw = s_rope[32 + lane];  // bf16 cos / sin of this position
      const float ca = bf16_lo(cw), cb = bf16_hi(cw), sa = bf16_lo(sw), sb = bf16_hi(sw);
      // x_embed = bf16(bf16(x*cos) + bf16(rotate_half(x)*sin)), rotate_half(x) = [-x2, x1]
      sm[ATT_SQ + 2 * lane] = bf16_round(bf16_round(x1a * ca) + bf16_round(-x2a * sa));
      sm[ATT_SQ + 2 * lane + 1] = bf16_round(bf16_round(x1b * cb) + bf16_round(-x2b * sb));
      sm[ATT_SQ2 + 2 * lane] = bf16_round(bf16_round(x2a * ca) + bf16_round(x1a * sa));
      sm[ATT_SQ2 + 2 * lane + 1] = bf16_round(bf16_round(x2b * cb) + bf16_round(x1b * sb));
    }
    abar();
    const long long ts1 = PROF ? global_ns() : 0;

    // ---- cached keys, out of TMEM (loops are NOT unrolled: one instance of each helper keeps the instruction footprint of these
    // warps small next to the consumers' hot loop)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float m = -INFINITY, l = 0.f;
    if (n_old > 0) {
#pragma unroll 1
      for (int ps = 0; ps < passes; ++ps) {
        tmem_ld_32x32(tbase + 32 * ps, buf);
        tmem_ld_wait();
        att_score(buf, sm, n_old, ps, atid, scale);
      }
      abar();
      if (PROF) t_sc += global_ns() - ts1;
      // softmax statistics over the cached keys: thread t owns keys t, t + 128, ...
      float sv[ATT_KPT];
      float wm = -INFINITY;
#pragma unroll
      for (int i = 0; i < ATT_KPT; ++i) {
        const int k = atid + i * DEC_ATHREADS;
        sv[i] = (k < n_old) ? sm[ATT_SSCORE + k] : -INFINITY;
        wm = fmaxf(wm, sv[i]);
      }
      wm = warp_max(wm);
      if (lane == 0) red[awarp] = wm;
      abar();
      m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
      float wl = 0.f;
#pragma unroll
      for (int i = 0; i < ATT_KPT; ++i) {
        const int k = atid + i * DEC_ATHREADS;
        if (k < n_old) {
          const float pk = __expf(sv[i] - m);
          wl += pk;
          sm[ATT_SSCORE + k] = bf16_round(pk);  // flash-attn: P is bf16 for the PV product, the row sum stays fp32
        }
      }
      wl = warp_sum(wl);
      if (lane == 0) red[4 + awarp] = wl;
      abar();  // also publishes the probabilities
      l = (red[4] + red[5]) + (red[6] + red[7]);
      if (PROF) t_sm += global_ns() - ts1;
#pragma unroll 1
      for (int ps = 0; ps < passes; ++ps) {
        tmem_ld_32x32(tbase + 32 * (ATT_MAX_PASSES + ps), buf);
        tmem_ld_wait();
        att_pv(buf, sm, n_old, ps, atid, acc);
      }
    }
    *reinterpret_cast<float4*>(sm + ATT_SACC + (atid >> 5) * 128 + 4 * lane) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    abar();
    // threads 0..63 reduce the 4 key slices for output elements (2t, 2t + 1)
    float n0 = 0.f, n1 = 0.f;
    if (atid < HALF) {
#pragma unroll
      for (int s2 = 0; s2 < DEC_AWARPS; ++s2) {
        const float2 v = *reinterpret_cast<const float2*>(sm + ATT_SACC + s2 * 128 + 2 * atid);
        n0 += v.x, n1 += v.y;
      }
    }
    const long long ts2 = PROF ? global_ns() : 0;

    uint64_t* part = static_cast<uint64_t*>(p.part) + static_cast<long>(head) * S * (DEC_HD + 2);
    if (!owns_new) {
      // hand the partial (m, l, acc[128]) to the combining CTA of this head as LL units
      uint64_t* mine = part + split * (DEC_HD + 2);
      if (atid < HALF) ll_store(mine + 2 + 2 * atid, __float_as_uint(n0), tag), ll_store(mine + 3 + 2 * atid, __float_as_uint(n1), tag);
      if (atid == 0) ll_store(mine, __float_as_uint(m), tag), ll_store(mine + 1, __float_as_uint(l), tag);
    } else {
      // ---- the other splits' partials first: they are published while the k/v rows are still being projected, so merging them
      // now keeps their L2 round trip off the critical path (a split with l == 0 is empty)
      float M = (l > 0.f) ? m : -INFINITY, den = l;
      if (atid < HALF) {
#pragma unroll 1
        for (int s0 = 0; s0 < S - 1; s0 += 3) {  // up to three other splits at a time: all six unit pairs in flight before the first tag check
          const int cnt = min(3, S - 1 - s0);
          uint32_t w[12];
          ll_fetch_pairs<6>(
              [&](int i) -> const uint64_t* {
                const uint64_t* ph = part + (s0 + min(i >> 1, cnt - 1)) * (DEC_HD + 2);
                return (i & 1) ? ph + 2 + 2 * atid : ph;
              },
              w, tag, check);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float ms = __uint_as_float(w[4 * c]), ls = __uint_as_float(w[4 * c + 1]);
            if (c < cnt && ls > 0.f) {
              const float Mn = fmaxf(M, ms);
              const float wo = (M == -INFINITY) ? 0.f : __expf(M - Mn), wn = __expf(ms - Mn);
              n0 = n0 * wo + __uint_as_float(w[4 * c + 2]) * wn, n1 = n1 * wo + __uint_as_float(w[4 * c + 3]) * wn, den = den * wo + ls * wn;
              M = Mn;
            }
          }
        }
      }
      const long long ts3 = PROF ? global_ns() : 0;
      if (PROF) t_pub += ts3 - ts2;
      // ---- the token being decoded: k (warp 0) and v (warp 1) arrive at the end of the projection
      if (awarp < 2) {
        const uint64_t* src = qkv + (awarp + 1) * (H / 2) + head * HALF;
        uint32_t lo, hi;
        ll_wait2(src + lane, src + lane + 32, tag, check, lo, hi);
        const long dst = kv_addr(p, s_table, layer, head).row(pos);
        if (awarp == 1) {
          sm[ATT_SVNEW + 2 * lane] = bf16_lo(lo), sm[ATT_SVNEW + 2 * lane + 1] = bf16_hi(lo);
          sm[ATT_SVNEW + 2 * lane + HALF] = bf16_lo(hi), sm[ATT_SVNEW + 2 * lane + 1 + HALF] = bf16_hi(hi);
          __nv_bfloat16* vw = static_cast<__nv_bfloat16*>(p.v_cache);
          reinterpret_cast<uint32_t*>(vw + dst)[lane] = lo, reinterpret_cast<uint32_t*>(vw + dst + HALF)[lane] = hi;
        } else {
          const float x1a = bf16_lo(lo), x1b = bf16_hi(lo), x2a = bf16_lo(hi), x2b = bf16_hi(hi);
          const uint32_t cw = s_rope[lane], sw = s_rope[32 + lane];
          const float ca = bf16_lo(cw), cb = bf16_hi(cw), sa = bf16_lo(sw), sb = bf16_hi(sw);
          const float r1a = bf16_round(bf16_round(x1a * ca) + bf16_round(-x2a * sa)), r1b = bf16_round(bf16_round(x1b * cb) + bf16_round(-x2b * sb));
          const float r2a = bf16_round(bf16_round(x2a * ca) + bf16_round(x1a * sa)), r2b = bf16_round(bf16_round(x2b * cb) + bf16_round(x1b * sb));
          __nv_bfloat16* kw = static_cast<__nv_bfloat16*>(p.k_cache);
          reinterpret_cast<uint32_t*>(kw + dst)[lane] = pack_bf16(r1a, r1b), reinterpret_cast<uint32_t*>(kw + dst + HALF)[lane] = pack_bf16(r2a, r2b);
          float d = sm[ATT_SQ + 2 * lane] * r1a;
          d = fmaf(sm[ATT_SQ + 2 * lane + 1], r1b, d), d = fmaf(sm[ATT_SQ2 + 2 * lane], r2a, d), d = fmaf(sm[ATT_SQ2 + 2 * lane + 1], r2b, d);
          d = warp_sum(d);
          if (lane == 0) red[8] = d * scale;
        }
      }
      abar();
      if (PROF) t_new += global_ns() - ts3;
      if (atid < HALF) {
        // one online-softmax step with the new key
        const float s_new = red[8];
        const float Mn = fmaxf(M, s_new);
        const float wo = (M == -INFINITY) ? 0.f : __expf(M - Mn), pn = __expf(s_new - Mn), pb = bf16_round(pn);
        n0 = n0 * wo + pb * sm[ATT_SVNEW + 2 * atid], n1 = n1 * wo + pb * sm[ATT_SVNEW + 2 * atid + 1];
        den = den * wo + pn;
        ll_store(static_cast<uint64_t*>(p.attn) + head * HALF + atid, pack_bf16(n0 / den, n1 / den), tag);
      }
    }
    const long long ts4 = PROF ? global_ns() : 0;
    abar();  // sq / sacc / scores are rewritten by the next layer; everyone is done with the TMEM rows
    if (layer + 1 < L) {
      // The staging loads (up to 24 outstanding 16-B loads per thread) share the SM's load path with the consumers' LL polling:
      // issued right away they sit in front of the o_proj-output gather and delay it by ~2 us. Hold them until the consumers are
      // inside the gate/up weight phase (22 us without a single global load of theirs; staging takes 8-13 us).
      if (atid == 0 && !(dflags & 32)) {
        uint32_t spins = 0;
        while (*s_step < PH_STEPS * layer + PH_GATEUP) {
          __nanosleep(200);
          if (++spins > EMX_SPIN_LIMIT) __trap();
        }
      }
      abar();
      stage(layer + 1);
    }
    if (PROF) t_q += ts1 - ts0, t_old += ts2 - ts1, t_pre += global_ns() - ts4;
  }
  if (PROF && dbg) dbg[0] = t_q, dbg[1] = t_old, dbg[2] = t_new, dbg[3] = t_pub, dbg[4] = t_sc, dbg[5] = t_sm, dbg[6] = t_pre;
  tc_fence_before();
  abar();
  if (awarp == 0) tmem_dealloc(*tmem_holder, ATT_TMEM_COLS);
}

// ---- the kernel ----------------------------------------------------------------------------------------------------
// MODE 0: the PRODUCT kernel — `debug_flags` and `dbg` are not even read, every wait is checked, every store happens.
// MODE 1: profiling twin that honours `debug_flags` (timing experiments whose results may be garbage), no instrumentation.
// MODE 2: instrumented twin used when `dbg` is given (phase timestamps, wait counters; also honours `debug_flags`); it costs registers.
template <int MODE>
__global__ void __launch_bounds__(DEC_THREADS, 1) decode_step_kernel(const emx_decode_params p) {
  constexpr bool PROF = (MODE == 2);
  const int dflags = (MODE >= 1) ? p.debug_flags : 0;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  uint32_t* xs = reinterpret_cast<uint32_t*>(smem + DEC_STAGES * DEC_STAGE_BYTES);  // activation vector, bf16 pairs, xs_pos order
  float* misc = reinterpret_cast<float*>(smem + DEC_STAGES * DEC_STAGE_BYTES + DEC_XS_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + DEC_STAGES * DEC_STAGE_BYTES + DEC_XS_BYTES + DEC_MISC_BYTES);
  uint64_t* empty = full + DEC_STAGES;
  uint2* ln_s = reinterpret_cast<uint2*>(smem + DEC_STAGES * DEC_STAGE_BYTES + DEC_XS_BYTES + DEC_MISC_BYTES + 128);
  float* red = misc;                                 // [8]
  int* s_state = reinterpret_cast<int*>(misc + 16);  // [5]
  float* part = misc + 64;                           // [DEC_PARTBUFS][8 warps][16 rows] partial row sums
  uint32_t* s_resid = reinterpret_cast<uint32_t*>(misc + 64 + DEC_PARTBUFS * DEC_CWARPS * DEC_GROUP);  // [DEC_MAX_RESID] residual bf16 pairs of this CTA's rows
  volatile uint32_t* s_issued = reinterpret_cast<volatile uint32_t*>(misc + 24);  // [2] producers -> prefetch warp
  PhaseTab& tab = *reinterpret_cast<PhaseTab*>(misc + 64 + DEC_PARTBUFS * DEC_CWARPS * DEC_GROUP + DEC_MAX_RESID);  // 8-byte aligned
  static_assert(sizeof(PhaseTab) == 224, "PhaseTab layout");
  int32_t* s_table = reinterpret_cast<int32_t*>(misc + 64 + DEC_PARTBUFS * DEC_CWARPS * DEC_GROUP + DEC_MAX_RESID + 56);  // [DEC_MAX_PAGES] block table
  uint32_t* s_rope = reinterpret_cast<uint32_t*>(s_table + DEC_MAX_PAGES);  // [32] cos pairs | [32] sin pairs of this position (bf16)
  static_assert((64 + DEC_PARTBUFS * DEC_CWARPS * DEC_GROUP + DEC_MAX_RESID + 56 + DEC_MAX_PAGES + 64) * 4 <= DEC_MISC_BYTES, "misc area overflow");
  float* att_red = misc + 48;  // [9] attention warps: per-warp maxima / sums, score of the new key
  volatile int* s_step = reinterpret_cast<volatile int*>(misc + 58);  // consumers -> attention warps: the (layer, phase) step being consumed
  float* att_sm = reinterpret_cast<float*>(smem + DEC_STAGES * DEC_STAGE_BYTES + ATT_SM_OFFSET);
  static_assert(ATT_SM_OFFSET + ATT_SM_FLOATS * 4 <= DEC_XS_BYTES, "attention scratch exceeds the activation area");

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  emx_decode_state* st = p.state;
  if (tid >= 32 && tid < 39) build_phase_tab(p, tab, tid - 32);
  if (tid >= 64 && tid < 64 + p.max_pages) s_table[tid - 64] = __ldg(p.block_table + (tid - 64));
  if (tid >= 128 && tid < 192) {  // RoPE row of the position being decoded: the same for all layers
    const int i = tid - 128;
    const long ppos = static_cast<long>(ldg_cg_u32(&st->pos)) * (DEC_HD / 2);
    const __nv_bfloat16* tabp = static_cast<const __nv_bfloat16*>(i < 32 ? p.cos_tab : p.sin_tab);
    s_rope[i] = __ldg(reinterpret_cast<const uint32_t*>(tabp + ppos) + (i & 31));
  }
  if (tid == 0) {
    s_state[0] = static_cast<int>(ldg_cg_u32(&st->cur_token));
    s_state[1] = static_cast<int>(ldg_cg_u32(&st->pos));
    s_state[2] = static_cast<int>(ldg_cg_u32(&st->n_generated));
    s_state[3] = static_cast<int>(ldg_cg_u32(&st->finished));
    s_state[4] = static_cast<int>(ldg_cg_u32(&st->epoch));
    s_issued[0] = 0, s_issued[1] = 0;
    *reinterpret_cast<volatile int*>(misc + 58) = -1;
    for (int s = 0; s < DEC_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], DEC_CWARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int token = min(max(s_state[0], 0), p.vocab - 1), pos = s_state[1], n_gen = s_state[2];  // (clamped: a profiling mode that skips the LL waits may have written garbage)
  // sequence already hit EOS, or the KV cache of this sequence is full (a caller driving its own loop past the capacity must not
  // corrupt memory: the block table, the RoPE tables and out_tokens all end at max_pages * page_size): nothing to do, uniformly
  if (s_state[3] || pos >= p.max_pages * p.page_size) return;

  const int L = p.layers, H = p.hidden;
  long long* dbg = (PROF && blockIdx.x == 0) ? reinterpret_cast<long long*>(p.dbg) : nullptr;

  // LL tags of this launch: tag0 + l for everything exchanged inside layer l (and for the residual stream ENTERING layer l);
  // tag0 + L enters the final norm, tag0 + L + 1 carries the argmax candidates. Never 0, unique across launches.
  const uint32_t tag0 = static_cast<uint32_t>(s_state[4]) * static_cast<uint32_t>(L + 2) + 1u;
  const bool check = !(dflags & 1);
  if (warp >= DEC_CWARPS + DEC_AWARPS + DEC_PWARPS) {
    if (warp == DEC_CWARPS + DEC_AWARPS + DEC_PWARPS) prefetch_loop(p, dflags, tab, lane, s_issued, full, dbg);
    return;  // (the 16th warp only pads the block to 512 threads = 128 registers per thread)
  }
  if (warp >= DEC_CWARPS + DEC_AWARPS) {
    producer_loop(p, dflags, tab, ring, full, empty, lane, warp - DEC_CWARPS - DEC_AWARPS, s_issued, dbg);
    return;
  }
  if (warp >= DEC_CWARPS) {
    attention_loop<PROF>(p, dflags, s_table, s_rope, pos, tag0, check, att_sm, att_red, reinterpret_cast<uint32_t*>(misc + 60), s_step, tid - DEC_CTHREADS);
    return;
  }

  // ===================== consumer warps =====================
  // NOTE on code size: the L1.5 instruction cache is 32 KB (2048 instructions) and an instruction miss goes to an L2 that is
  // saturated with weight traffic (~0.5-1 us). Everything that runs once per phase is therefore instantiated ONCE: the layers
  // are a flat loop over (layer, phase) steps with one gather + RMSNorm, one plain gather, one attention and one consume body.
  int dbg_i = 0;
  auto mark = [&]() {
    if (PROF) {
      if (dbg && tid == 0) dbg[dbg_i] = global_ns();
      ++dbg_i;
    }
  };
  ConsumerState cs{0, 0, 0, 0, 0};
  long long gprof[4] = {0, 0, 0, 0};  // PROF: cycles of the gather + RMSNorm steps (loads back | ln wait | sum | normalise + barrier)
  const bool drop = dflags & 16;

  uint64_t* xd = static_cast<uint64_t*>(p.x);      // residual stream after down_proj   [H/2] units
  uint64_t* xo = static_cast<uint64_t*>(p.xo);     // residual stream after o_proj      [H/2]
  uint64_t* qkv = static_cast<uint64_t*>(p.qkv);   // [3H/2]: q | k | v
  uint64_t* attn = static_cast<uint64_t*>(p.attn); // [H/2]
  uint64_t* hbuf = static_cast<uint64_t*>(p.h);    // [inter/2]
  const uint32_t* emb_row = reinterpret_cast<const uint32_t*>(static_cast<const __nv_bfloat16*>(p.embed) + static_cast<long>(token) * H);
  const int rb = tab.r_begin[PH_O], rb2 = rb >> 1, re2 = tab.r_end[PH_O] >> 1;  // this CTA's rows of the residual-producing phases

  // this CTA's own rows of the residual stream are kept for the residual add of the next epilogue
  auto own = [&](int u, uint32_t w) {
    if (u >= rb2 && u < re2) s_resid[u - rb2] = w;
  };
  float best = -INFINITY;
  int best_i = 0x7fffffff;

  ln_fetch_async(static_cast<const __nv_bfloat16*>(p.ln1), ln_s, H);
  const int n_steps = PH_STEPS * L + 1;
  int layer = 0, kind = PH_Q;
#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    const uint32_t tag = tag0 + layer;  // (the lm_head step has layer == L)
    mark();
    if (kind == PH_Q || kind == PH_GATEUP || kind == PH_LMHEAD) {
      // ---- residual stream in + RMSNorm ----
      const uint64_t* src = (kind == PH_GATEUP) ? xo : xd;
      gather_rmsnorm(src, step == 0 ? emb_row : nullptr, H, tag, check, ln_s, p.rms_eps, xs, red, kind == PH_GATEUP ? 1u : 0u, own, (PROF && dbg && tid == 0) ? gprof : nullptr);
      if (kind != PH_LMHEAD) {  // norm weights of the next RMSNorm: ln2 of this layer, ln1 of the next one, the final norm
        const __nv_bfloat16* next_w = (kind == PH_Q)     ? static_cast<const __nv_bfloat16*>(p.ln2) + static_cast<long>(layer) * H
                                      : (layer + 1 < L) ? static_cast<const __nv_bfloat16*>(p.ln1) + static_cast<long>(layer + 1) * H
                                                        : static_cast<const __nv_bfloat16*>(p.final_norm);
        ln_fetch_async(next_w, ln_s, H);
      }
    } else if (kind == PH_O || kind == PH_DOWN) {
      // ---- a plain vector in: the attention output (published by the attention warps of the split-0 CTAs) for o_proj, the
      // SwiGLU output for down_proj ----
      ll_gather<11>(kind == PH_O ? attn : hbuf, (kind == PH_O ? H : p.inter) >> 1, tag, check, [&](int u, uint32_t w) { xs[xs_pos(u)] = w; });
      cbar();
    }  // PH_K, PH_V: same input vector as PH_Q
    mark();
    if (tid == 0) *s_step = step;
    consume_phase<PROF>(phase_desc(tab, layer, kind), ring, full, empty, cs, reinterpret_cast<const __nv_bfloat16*>(xs), part, warp, lane,
                        dflags, [&](int row, float a0, float a1, bool valid) {
                          // lanes 0..7 of warp 0, converged; `kind` is uniform
                          if (kind <= PH_V) {
                            if (valid) ll_store(qkv + kind * (H >> 1) + (row >> 1), pack_bf16(a0, a1), tag, drop);
                          } else if (kind == PH_GATEUP) {
                            // lane i holds (gate, up) of output row/2; two outputs make one LL unit
                            const float hv = bf16_round(bf16_round(silu(bf16_round(a0))) * bf16_round(a1));
                            const float hn = __shfl_down_sync(0xffu, hv, 1);
                            if (valid && !(lane & 1)) ll_store(hbuf + (row >> 2), pack_bf16(hv, hn), tag, drop);
                          } else if (kind == PH_LMHEAD) {
                            if (valid) {
                              const float v0 = bf16_round(a0), v1 = bf16_round(a1);
                              if (p.logits_out) p.logits_out[row] = v0, p.logits_out[row + 1] = v1;
                              if (v0 > best) best = v0, best_i = row;  // rows ascend per thread: strict '>' keeps the lowest index
                              if (v1 > best) best = v1, best_i = row + 1;
                            }
                          } else if (valid) {  // o_proj / down_proj: + residual; down_proj feeds the NEXT layer (tag + 1)
                            const uint32_t r = s_resid[(row - rb) >> 1];
                            ll_store((kind == PH_O ? xo : xd) + (row >> 1), pack_bf16(bf16_lo(r) + bf16_round(a0), bf16_hi(r) + bf16_round(a1)),
                                     kind == PH_O ? tag : tag + 1, drop);
                          }
                        });
    if (PROF && p.dbg && step == 2 * PH_STEPS - 1 && tid == 0) reinterpret_cast<long long*>(p.dbg)[15 * L + 16 + blockIdx.x] = global_ns();  // end of layer 1
    if (++kind == PH_LMHEAD) {
      kind = PH_Q;
      if (++layer == L) kind = PH_LMHEAD;
    }
  }
  mark();
  if (PROF && dbg && tid == 0) dbg[15 * L + 4] = gprof[0], dbg[15 * L + 5] = gprof[1], dbg[15 * L + 6] = gprof[2], dbg[15 * L + 7] = gprof[3];
  if (PROF && dbg && tid == 0) dbg[15 * L + 8] = cs.waited, dbg[15 * L + 11] = cs.t_sync, dbg[15 * L + 12] = cs.t_epi;
  cbar();  // every epilogue is done: the partial-sum buffers are free
  float* s_best = part;  // [64] values + [64] indices: lanes 0..7 of every warp hold candidates (rotating epilogue duty)
  if (lane < 8) s_best[warp * 8 + lane] = best, reinterpret_cast<int*>(s_best + 64)[warp * 8 + lane] = best_i;
  cbar();
  uint64_t* cand = static_cast<uint64_t*>(p.argmax_part);  // [grid][2] LL units: value bits, index
  if (tid == 0) {
    for (int w = 1; w < 64; ++w) {
      const float v = s_best[w];
      const int i = reinterpret_cast<int*>(s_best + 64)[w];
      if (v > best || (v == best && i < best_i)) best = v, best_i = i;
    }
    ll_store(cand + 2 * blockIdx.x, __float_as_uint(best), tag0 + L + 1);
    ll_store(cand + 2 * blockIdx.x + 1, static_cast<uint32_t>(best_i), tag0 + L + 1);
  }
  if (blockIdx.x == 0 && warp == 0) {
    float b = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < static_cast<int>(gridDim.x); c += 32) {
      const float v = __uint_as_float(ll_wait(cand + 2 * c, tag0 + L + 1, check));
      const int i = static_cast<int>(ll_wait(cand + 2 * c + 1, tag0 + L + 1, check));
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
      st->epoch = static_cast<uint32_t>(s_state[4]) + 1u;
    }
    mark();
  }
}

}  // namespace emx

// CTAs per launch = SMs of the current device (148 on B200; also the answer when no device is visible, e.g. the CPU test suite)
extern "C" int emx_decode_grid(void) {
  const int sms = emx::device_sms();
  return sms > 0 ? sms : emx::kNumSMs;
}

extern "C" int emx_decode_phase_rows(int n_rows, int granule, int cta, int grid, int* r_begin, int* r_end) {
  EMX_REQUIRE(n_rows > 0 && (granule == 2 || granule == 4) && grid > 0 && cta >= 0 && cta < grid && r_begin && r_end && n_rows % granule == 0,
              "emx_decode_phase_rows: n_rows=%d granule=%d cta=%d grid=%d", n_rows, granule, cta, grid);
  emx::phase_rows(n_rows, static_cast<uint32_t>(granule), static_cast<uint32_t>(cta), static_cast<uint32_t>(grid), *r_begin, *r_end);
  return 0;
}

extern "C" int emx_decode_step(const emx_decode_params* params, cudaStream_t stream) {
  using namespace emx;
  const emx_decode_params& p = *params;
  EMX_REQUIRE(p.head_dim == DEC_HD, "emx_decode_step: head_dim %d not supported (128)", p.head_dim);
  EMX_REQUIRE(p.hidden % 16 == 0 && p.inter % 16 == 0 && p.vocab % 2 == 0, "emx_decode_step: hidden/inter must be multiples of 16, vocab even");
  EMX_REQUIRE(p.x && p.xo && p.qkv && p.attn && p.h && p.part && p.argmax_part && p.state, "emx_decode_step: null scratch pointer");
  EMX_REQUIRE(p.inter * 2 <= DEC_XS_BYTES && p.hidden * 2 <= DEC_XS_BYTES, "emx_decode_step: activation vector exceeds %d bytes", DEC_XS_BYTES);
  EMX_REQUIRE(p.page_size > 0 && (p.page_size & (p.page_size - 1)) == 0, "emx_decode_step: page_size must be a power of two");
  EMX_REQUIRE(p.kv_splits >= 1 && p.kv_splits <= 8 && (p.kv_splits & (p.kv_splits - 1)) == 0, "emx_decode_step: kv_splits must be 1, 2, 4 or 8");
  EMX_REQUIRE(p.hidden <= 16 * DEC_CTHREADS && p.hidden * 2 <= DEC_LN_BYTES, "emx_decode_step: hidden > %d not supported by the fused gather + RMSNorm", DEC_LN_BYTES / 2);
  EMX_REQUIRE(p.max_pages <= DEC_MAX_PAGES, "emx_decode_step: block table of %d pages exceeds %d", p.max_pages, DEC_MAX_PAGES);
  const int max_keys_per_split = ATT_PASS * ATT_MAX_PASSES;
  EMX_REQUIRE((static_cast<long>(p.max_pages) * p.page_size + p.kv_splits - 1) / p.kv_splits <= max_keys_per_split,
              "emx_decode_step: context capacity %d x %d exceeds the per-split score buffer (%d keys)", p.max_pages, p.page_size,
              max_keys_per_split);
  int dev = 0;
  const int grid = device_sms(&dev);
  bool* attr_set = device_attr_flag(ATTR_DECODE);
  if (grid < 0 || !attr_set) return -2;
  if (!*attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(decode_step_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM));
    int per_sm = 0;
    EMX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_step_kernel<2>, DEC_THREADS, DEC_SMEM));
    EMX_REQUIRE(per_sm >= 1, "emx_decode_step: kernel does not fit on an SM (smem %d)", DEC_SMEM);
    *attr_set = true;
  }
  EMX_REQUIRE(p.heads * p.kv_splits <= grid, "emx_decode_step: heads x kv_splits must not exceed the grid of %d CTAs (one attention item per CTA)", grid);
  EMX_REQUIRE(p.hidden / 2 / grid + 2 <= DEC_MAX_RESID, "emx_decode_step: hidden too large for the residual staging buffer");
  void* args[] = {const_cast<emx_decode_params*>(params)};
  // the product kernel unless the caller asks for a profiling twin: `dbg` -> instrumented, `debug_flags` alone -> flag-honouring
  void* fn = p.dbg ? reinterpret_cast<void*>(decode_step_kernel<2>)
                   : (p.debug_flags ? reinterpret_cast<void*>(decode_step_kernel<1>) : reinterpret_cast<void*>(decode_step_kernel<0>));
  EMX_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(DEC_THREADS), args, DEC_SMEM, stream));
  return 0;
}
