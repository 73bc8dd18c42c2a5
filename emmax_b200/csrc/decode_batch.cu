// emx_decode_batch_step — one greedy decode step of Llama-2-7B for UP TO 8 SEQUENCES as ONE persistent kernel.
//
// The reference cannot do this at all: its cached branch asserts batch size 1
// (/root/reference/prismatic/extern/hf/modeling_prismatic.py:326, :460-463), so N robots / N simulator environments cost N full
// passes over the 13.2 GB of weights per token. Single-token decode is HBM-bound and the weight stream is 97 % of its bytes: here
// one pass over the weights serves 8 sequences (BASELINE.json configs[4]: bs=64 over 8 GPUs = 8 per GPU, mixed 128/512 new tokens).
//
// Same skeleton as decode_mega.cu (one CTA per SM, static per-CTA weight schedule streamed by two producer warps through a
// 3 x 64 KB cp.async.bulk ring, mma.sync.m16n8k16 consumers, no grid barriers: every cross-CTA vector travels as 8-byte LL units),
// re-thought for a batch:
//   * the 8 columns of the m16n8k16 B operand — 7 of which are dead weight at batch 1 — are the 8 sequences: thread (g, t) of a
//     consumer warp feeds column g = sequence g;
//   * 8 activation vectors (8 x 22 KB for down_proj) do not fit next to the ring in shared memory. They live in TENSOR MEMORY, already
//     in B-fragment order: a thread gathers exactly the LL units ITS fragments need (the exchange buffers are laid out so that these
//     are 16-byte runs), normalises them and parks them in its own TMEM lane (tcgen05.st, up to 192 of its 256 columns); the hot loop
//     pulls 32 columns per ring stage back with one tcgen05.ld. No shared memory, no bank conflicts, no block barrier for the hand-over;
//   * attention is a phase of the same pipeline: the cached K/V pages of all (sequence, head) rows are one linear list of page-pair
//     items, split evenly over the CTAs whatever the individual context lengths are (flash-decoding); the producers stream each
//     CTA's share through the weight ring (16 KB bulk copies straight out of the paged cache), every consumer warp keeps an online-softmax
//     state for its 16 keys of the stage, and a rotating warp merges the 8 warp states at the end of a row segment and either
//     publishes the partial (m, l, acc[128]) or — in the CTA that owns the row's FIRST item — folds in the other segments and the
//     token being decoded (RoPE, KV append, one more online-softmax step) and publishes the head output;
//   * per-sequence position, block table, RoPE row, token limit and EOS state; sequences that are finished (or slots beyond `batch`)
//     are inactive: nobody waits for their units, nothing of theirs is stored.
// Rounding points mirror the torch-eager reference exactly as decode_mega.cu does.
#include <cstdio>
#include "decode_common.cuh"
#include "emmax.h"

namespace emx {

constexpr int DB_MAXB = EMX_DECODE_MAX_BATCH;     // sequences per launch == N of the MMA atom
static_assert(DB_MAXB == 8, "the batch is the N dimension of mma.m16n8k16");
constexpr int DB_PWARPS = 2;                      // producer warps (alternate ring stages)
constexpr int DB_THREADS = (DEC_CWARPS + DB_PWARPS + 1) * 32;  // 352: 8 consumer + 2 producer + 1 L2-prefetch warp
constexpr int DB_MAX_PAGES = 64;                  // block-table entries per sequence staged in shared memory
constexpr int DB_MAX_RESID = 64;                  // residual units (row pairs) of this CTA's rows, per sequence
constexpr int DB_ATT_WSTRIDE = 132;               // floats per warp in an attention partial buffer: acc[128], m, l, pad
constexpr int DB_PART_FLOATS = 1088;              // one partial buffer: GEMV [8 warps][16 rows][8 seqs] = 1024 | attention 8 x 132 = 1056
constexpr int DB_MAXSEG = EMX_DECODE_ATT_MAX_SEGMENTS;  // row segments (CTAs) one (sequence, head) row can be split into
constexpr int DB_LN_BYTES = 8192;                 // norm weights of the next RMSNorm (hidden <= 4096)
constexpr int DB_PAGE = 64;                       // KV page = 64 tokens (one 16 KB bulk copy per head and page)
constexpr int DB_PAGE_BYTES = DB_PAGE * DEC_HD * 2;
constexpr int DB_TMEM_COLS = 512;
constexpr int DB_PARTU = DEC_HD + 2;              // LL units of one split-KV partial: m, l, acc[128] (fp32 payloads)

constexpr int BPH_STEPS = 7;  // phases per layer
enum BatchPhase { BPH_Q = 0, BPH_K, BPH_V, BPH_ATT, BPH_O, BPH_GATEUP, BPH_DOWN, BPH_LMHEAD, BPH_KINDS };

struct __align__(16) BatchShared {
  uint64_t full[DEC_STAGES], empty[DEC_STAGES];
  float red[2][DEC_CWARPS][DB_MAXB];  // RMSNorm: per-warp sums of squares
  int tok[DB_MAXB], pos[DB_MAXB], ngen[DB_MAXB];
  uint32_t active_mask, epoch, tmem_base, pad0;
  uint32_t issued[2], pad1[2];  // producers -> L2-prefetch warp: ring stages issued so far
  int att_pp[DB_MAXB];       // page pairs (items) per (sequence, head) row
  int att_off[DB_MAXB + 1];  // first item of sequence i in the linear item list
  int att_i0, att_i1, att_T;  // this CTA's items [i0, i1) of T
  const __nv_bfloat16* W[BPH_KINDS];
  long layer_stride[BPH_KINDS];
  int N[BPH_KINDS], K[BPH_KINDS], r_begin[BPH_KINDS], r_end[BPH_KINDS];
  int32_t table[DB_MAXB][DB_MAX_PAGES];
  uint32_t rope[DB_MAXB][64];  // [32] cos pairs | [32] sin pairs (bf16) of each sequence's position
  uint32_t resid[DB_MAXB][DB_MAX_RESID];
};
constexpr int DB_SMEM = DEC_STAGES * DEC_STAGE_BYTES + DEC_PARTBUFS * DB_PART_FLOATS * 4 + DB_LN_BYTES + static_cast<int>(sizeof(BatchShared));
static_assert(DB_SMEM <= 232448, "decode_batch_kernel: shared memory exceeds 227 KB");

// Exchange-buffer position of unit u (= bf16 pair u of a vector): inside every block of 16 units the 4 x 4 matrix is transposed, so that
// the units a thread needs for two consecutive k-steps of mma.m16n8k16 (u = 16 blk + 4 m + t, m = 0..3, t = lane % 4) are the 32-byte run
// [16 blk + 4 t, 16 blk + 4 t + 4).
__device__ __forceinline__ int ll_pos(int u) { return (u & ~15) | ((u & 3) << 2) | ((u >> 2) & 3); }

__device__ __forceinline__ bool tag_ok(uint64_t v, uint32_t tag) { return static_cast<uint32_t>(v >> 32) == tag; }

__device__ __forceinline__ void build_phase(const emx_decode_batch_params& p, BatchShared& sh, int kind) {
  const long H = p.hidden, I = p.inter;
  const __nv_bfloat16* W = nullptr;
  long ls = 0;
  int N = 0, K = 0;
  switch (kind) {
    case BPH_Q:
    case BPH_K:
    case BPH_V: W = static_cast<const __nv_bfloat16*>(p.w_qkv) + kind * H * H, ls = 3 * H * H, N = H, K = H; break;
    case BPH_O: W = static_cast<const __nv_bfloat16*>(p.w_o), ls = H * H, N = H, K = H; break;
    case BPH_GATEUP: W = static_cast<const __nv_bfloat16*>(p.w_gateup), ls = 2 * I * H, N = 2 * I, K = H; break;
    case BPH_DOWN: W = static_cast<const __nv_bfloat16*>(p.w_down), ls = H * I, N = H, K = I; break;
    case BPH_LMHEAD: W = static_cast<const __nv_bfloat16*>(p.lm_head), ls = 0, N = p.vocab, K = H; break;
    default: break;  // BPH_ATT: no weights
  }
  sh.W[kind] = W, sh.layer_stride[kind] = ls, sh.N[kind] = N, sh.K[kind] = K;
  int rb = 0, re = 0;
  if (N > 0) phase_rows(N, (kind == BPH_GATEUP) ? 4 : 2, blockIdx.x, gridDim.x, rb, re);
  sh.r_begin[kind] = rb, sh.r_end[kind] = re;
}

// ---- attention items ---------------------------------------------------------------------------------------------------
// Item g of the linear list -> (sequence, head, page pair j of that row). att_off is non-decreasing; inactive sequences own no items.
struct AttItem {
  int seq, head, j, pp;
};
__device__ __forceinline__ AttItem att_item(const BatchShared& sh, int g) {
  AttItem it;
  int n = 0;
#pragma unroll
  for (int i = 1; i < DB_MAXB; ++i) n += (g >= sh.att_off[i]) ? 1 : 0;
  it.seq = n, it.pp = sh.att_pp[n];
  const int rem = g - sh.att_off[n];
  it.head = rem / it.pp, it.j = rem - it.head * it.pp;
  return it;
}
// element offset of (layer, page, head, slot 0) in the paged cache [layer][page][head][64][128]
__device__ __forceinline__ long kv_page_off(const emx_decode_batch_params& p, int layer, int page, int head) {
  return ((static_cast<long>(layer) * p.n_pages + page) * p.heads + head) * (DB_PAGE * DEC_HD);
}

// ring stages the gather of a K-element vector occupies (see "gathers" below; their producer is consumer thread 0: the producer warps and the
// prefetch warp skip them)
__host__ __device__ __forceinline__ int gather_stages(int K) { return (K + DEC_KC - 1) / DEC_KC; }

// ---- producer warps, L2-prefetch warp ----------------------------------------------------------------------------------------
// All walk the same schedule (layer -> q, k, v rows -> this CTA's K/V page pairs -> o, gate/up, down rows; lm_head rows).
// PREFETCH == false: producer warp `pidx` issues the ring stages with it % 2 == pidx; it never waits for anything but a free ring slot.
// PREFETCH == true: the ring only covers ~4 us of a saturated HBM, so whenever the consumers stall (an exchange, a gather) the ring fills,
// no new copy can be issued and HBM idles. This warp watches for exactly that — the newest ring copy of this CTA has LANDED, i.e. nothing
// of ours is in flight — and then pulls the next stages of the schedule HBM -> L2 (cp.async.bulk.prefetch.L2), paced at about twice the
// SM's fair share of HBM, at most `l2_lookahead_stages` ahead of the ring. In the HBM-bound steady state it never triggers.
struct PfGate {
  volatile uint32_t* issued;  // [2] ring stages issued so far by the two producer warps
  uint64_t* full;
  uint32_t lookahead;
  // true: prefetch stage `it` now; false: the ring got there first, skip it
  __device__ __forceinline__ bool admit(uint32_t it) const {
    uint32_t spins = 0;
    for (;;) {
      const uint32_t iss = max(issued[0], issued[1]);
      if (it < iss) return false;
      if (iss > 0 && it < iss + lookahead) {
        const uint32_t last = iss - 1;
        if (mbar_try_wait(&full[last % DEC_STAGES], (last / DEC_STAGES) & 1)) return true;
      }
      __nanosleep(200);
      if (++spins > EMX_SPIN_LIMIT) __trap();
    }
  }
};
__device__ __forceinline__ void pf_pace(long long t0, int bytes) {
  const long long wait_ns = 700LL * bytes / (64 * 1024);
  while (global_ns() - t0 < wait_ns) __nanosleep(100);
}

template <bool PREFETCH>
__device__ __noinline__ void batch_producer(const emx_decode_batch_params& p, const BatchShared& sh, uint8_t* ring, uint64_t* full, uint64_t* empty, int lane,
                                            int pidx, volatile uint32_t* s_issued) {
  const uint64_t policy = l2_policy_evict_first();
  const __nv_bfloat16* kc = static_cast<const __nv_bfloat16*>(p.k_cache);
  const __nv_bfloat16* vc = static_cast<const __nv_bfloat16*>(p.v_cache);
  const PfGate gate{s_issued, full, static_cast<uint32_t>(max(p.l2_lookahead_stages, 0))};
  uint32_t it = 0;
  const int L = p.layers;
  for (int layer = 0; layer <= L; ++layer) {
    const int k_first = (layer == L) ? BPH_LMHEAD : BPH_Q, k_last = (layer == L) ? BPH_LMHEAD : BPH_DOWN;
    for (int kind = k_first; kind <= k_last; ++kind) {
      if (kind == BPH_ATT) {
        for (int g = sh.att_i0; g < sh.att_i1; ++g, ++it) {
          if (!PREFETCH && (it % DB_PWARPS) != static_cast<uint32_t>(pidx)) continue;
          if (PREFETCH && !gate.admit(it)) continue;
          const AttItem a = att_item(sh, g);
          const int pages = (sh.pos[a.seq] + DB_PAGE - 1) / DB_PAGE;
          const int np = min(2, pages - 2 * a.j);  // pages of this item: K page(s) at [0, 32 K), V page(s) at [32 K, 64 K) of the stage
          const int is_v = lane >= np ? 1 : 0, pg = lane - is_v * np;
          const long off = (lane < 2 * np) ? kv_page_off(p, layer, sh.table[a.seq][2 * a.j + pg], a.head) : 0;
          if (PREFETCH) {
            const long long t0 = global_ns();
            if (lane < 2 * np) prefetch_l2((is_v ? vc : kc) + off, DB_PAGE_BYTES);
            pf_pace(t0, 2 * np * DB_PAGE_BYTES);
            continue;
          }
          const int slot = it % DEC_STAGES;
          const uint32_t ph = (it / DEC_STAGES) & 1;
          if (lane == 0) {
            mbar_wait(&empty[slot], ph ^ 1);
            mbar_arrive_expect_tx(&full[slot], static_cast<uint32_t>(np) * 2 * DB_PAGE_BYTES);
          }
          __syncwarp();
          if (lane < 2 * np)
            bulk_g2s(ring + slot * DEC_STAGE_BYTES + is_v * 2 * DB_PAGE_BYTES + pg * DB_PAGE_BYTES, (is_v ? vc : kc) + off, DB_PAGE_BYTES, &full[slot], policy);
          if (lane == 0) s_issued[pidx] = it + 1;
        }
        continue;
      }
      const int K = sh.K[kind], r_end = sh.r_end[kind];
      // the gather in front of this phase occupies ring slots of its own
      // (k / v rows and the embedding rows of layer 0 have none). The producers issue nothing for them but still wait for each of
      // these slots in turn: an mbarrier wait carries ONE parity bit, so a warp must observe every phase of a slot's `empty` barrier
      // in order — skipping the wait would let the test for the slot's NEXT use pass two phases early.
      if (kind != BPH_K && kind != BPH_V && !(layer == 0 && kind == BPH_Q)) {
        for (int v = 0; v < gather_stages(K); ++v, ++it) {
          if (!PREFETCH && lane == 0) mbar_wait(&empty[it % DEC_STAGES], ((it / DEC_STAGES) & 1) ^ 1);
        }
        __syncwarp();
      }
      const __nv_bfloat16* W = sh.W[kind] + layer * sh.layer_stride[kind];
      for (int r = sh.r_begin[kind]; r < r_end; r += DEC_GROUP) {
        const int nrows = min(DEC_GROUP, r_end - r);
        for (int k0 = 0; k0 < K; k0 += DEC_KC, ++it) {
          if (!PREFETCH && (it % DB_PWARPS) != static_cast<uint32_t>(pidx)) continue;
          if (PREFETCH && !gate.admit(it)) continue;
          const int klen = min(DEC_KC, K - k0);
          const __nv_bfloat16* src = W + static_cast<long>(r + lane) * K + k0;
          if (PREFETCH) {
            const long long t0 = global_ns();
            if (lane < nrows) prefetch_l2(src, klen * 2);
            pf_pace(t0, nrows * klen * 2);
            continue;
          }
          const int slot = it % DEC_STAGES;
          const uint32_t ph = (it / DEC_STAGES) & 1;
          if (lane == 0) {
            mbar_wait(&empty[slot], ph ^ 1);
            mbar_arrive_expect_tx(&full[slot], static_cast<uint32_t>(nrows) * klen * 2);
          }
          __syncwarp();
          if (lane < nrows) bulk_g2s(ring + slot * DEC_STAGE_BYTES + lane * DEC_ROWSTRIDE, src, klen * 2, &full[slot], policy);
          if (lane == 0) s_issued[pidx] = it + 1;
        }
      }
    }
  }
}

// ---- gathers: LL exchange buffer -> ring slots (TMA) -> this thread's B fragments -> tensor memory ---------------------------------
// Thread (warp w, lane = 4 g + t) feeds sequence g. For ring-stage chunk c (2048 columns) warp w multiplies columns
// [2048 c + 256 w, + 256): 16 k-steps, i.e. 8 blocks of 32 columns = 16 units each, of which this thread needs units 4 m + t (m = 0..3):
// one 32-byte run per block in ll_pos order ("pairs" of units; pair i belongs to k-step i). The 32 payload words of a chunk are TMEM
// columns [32 c, 32 c + 32) of the thread's lane, in the order the hot loop wants them: word 2 i = b0, 2 i + 1 = b1 of k-step i.
//
// 8 sequences make a gather 128 KB (hidden) to 344 KB (intermediate) of LL units per CTA: polled from registers that is many dependent
// L2 round trips (~2 us each under the weight stream) and a large unrolled instruction footprint (an instruction-cache miss costs as
// much as a data miss here). Instead the vector travels like the weights: consumer thread 0 points the TMA engine at the LL buffers
// (one cp.async.bulk per sequence and chunk, 8 KB) and lands them in the NEXT RING STAGES — every byte in flight at once, no registers, a
// few instructions. The producer warps skip these stages of the schedule; for the consumers they are ordinary stages (wait `full`, read,
// release `empty` per warp), so the ring refills with the phase's first weight stages behind them. Threads read their own units from
// shared memory (conflict-free: sequences are skewed by 16 bytes) and check the tags; a unit that was not published yet when the copy
// read it (the arrival counter that times the copy is relaxed) is fetched by its reader straight from the L2 (ll_repair).
constexpr int DB_GSEQ_STRIDE = 1024 * 8 + 16;  // bytes between sequences inside a ring slot: one chunk = 1024 units, + 16 B bank skew
static_assert(DB_MAXB * DB_GSEQ_STRIDE <= DEC_STAGE_BYTES, "a gather chunk of 8 sequences must fit one ring stage");

__device__ __forceinline__ int chunk_ksteps(int K, int c, int warp) { return max(0, min(DEC_KW / 16, (K - c * DEC_KC - warp * DEC_KW) / 16)); }
__device__ __forceinline__ int half_ksteps(int K, int hc, int warp) { return max(0, min(8, chunk_ksteps(K, hc >> 1, warp) - 8 * (hc & 1))); }
// natural unit index of payload word q (0..15) of half chunk hc for (warp, t)
__device__ __forceinline__ int g_unit(int hc, int warp, int t, int q) { return 16 * (64 * (hc >> 1) + 8 * warp + 4 * (hc & 1) + (q >> 2)) + 4 * (q & 3) + t; }

// Slow path of a gather: the pair of units at `gp` carried a stale tag in the bulk copy (the arrival counter is relaxed, so a publisher's
// stores may trail its arrival). Poll the pair itself; the tags are the proof.
__device__ __noinline__ uint2 ll_repair(const uint64_t* gp, uint32_t tag) {
  uint64_t a, b;
  uint32_t spins = 0;
  for (;;) {
    ll_load2(gp, a, b);
    if (tag_ok(a, tag) && tag_ok(b, tag)) break;
    __nanosleep(100);
    if (++spins > EMX_SPIN_LIMIT) __trap();
  }
  return make_uint2(static_cast<uint32_t>(a), static_cast<uint32_t>(b));
}

__device__ __forceinline__ void ln_fetch_async_b(const __nv_bfloat16* w, uint32_t* ln_s, int H) {
  for (int i = threadIdx.x; i < (H >> 3); i += DEC_CTHREADS)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ln_s + 4 * i)), "l"(reinterpret_cast<const uint4*>(w) + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

struct BCons {
  uint32_t it;     // ring stage counter (weight stages, K/V items and the stages of the gathers)
  uint32_t group;  // row groups / attention flushes so far: selects the partial buffer, its named barrier and the rotating warp
  uint32_t gathers;  // exchanges (gathers out of an LL buffer) done so far in this launch
};

// One gather: K-element vectors of all sequences (LL units at buf + n * seq_stride, ll_pos order) -> TMEM. norm: LlamaRMSNorm on the way,
//   y = bf16(w * bf16(x * rsqrt(mean(x^2) + eps)))   (pass 1 parks the raw words in TMEM while summing squares, pass 2 normalises in place)
// and the CTA's own rows of the residual stream are kept in shared memory for the residual add of the next o_proj / down_proj epilogue.
// embed != nullptr (layer 0): the vectors are plain bf16 rows of the embedding table, no exchange, no ring slots.
__device__ __noinline__ void gather_b(const uint64_t* buf, long seq_stride, int K, bool norm, const __nv_bfloat16* embed, uint32_t tag, uint32_t tm, BatchShared& sh,
                                      uint8_t* ring, BCons& cs, const uint32_t* ln_s, float eps, uint32_t parity, int rb2, int re2, int warp, int lane,
                                      void* sync_cnt, uint32_t gathers_per_launch, long long* gprof, long long* gcta) {
  // gcta (instrumented twin, thread 0 of every CTA, layer 1 only): [0] after the entry barrier, [1] arrival counter complete
  // gprof (instrumented twin, thread 0 only): [0] cbar, [1] arrival counter, [2] first chunks issued (free slots), [3] first chunk landed, [4] read + park (incl. the later chunks' landing),
  // [5] norm tail, [6] units this thread had to repair (stale tag in the copy)
  const int n = lane >> 2, t = lane & 3;
  const bool act = (sh.active_mask >> n) & 1;
  const int nc = gather_stages(K);
  float ss = 0.f;
  if (embed) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(embed + static_cast<long>(sh.tok[n]) * K);
#pragma unroll 1
    for (int hc = 0; hc < 2 * nc; ++hc) {
      const int ks = half_ksteps(K, hc, warp);
      if (ks > 0) {
        uint32_t r[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          r[q] = (act && (q >> 1) < ks) ? __ldg(row + g_unit(hc, warp, t, q)) : 0u;
          ss += sumsq2(r[q]);
          const int u = g_unit(hc, warp, t, q);
          if (act && (q >> 1) < ks && u >= rb2 && u < re2) sh.resid[n][u - rb2] = r[q];
        }
        tmem_st_32x16(tm + 16 * hc, r);
      }
    }
  } else {
    const int units = static_cast<int>(seq_stride);  // per sequence, a multiple of 16
    const uint32_t ring_s = smem_u32(ring);
    // Arrival counter of the exchange: the vector was produced by the phase every CTA has just finished, so "all CTAs have arrived here"
    // means "every unit is published". It only decides WHEN the copies are worth issuing (a copy issued earlier would carry stale
    // units for certain); the tags stay the proof.
    long long tg = gprof ? global_ns() : 0;
    auto lap = [&](int k) {
      if (gprof) {
        const long long now = global_ns();
        gprof[k] += now - tg, tg = now;
      }
    };
    // chunk c of the vector -> ring stage cs.it + c: this thread is the producer of that stage (the producer warps skip it)
    auto issue_chunk = [&](int c) {
      const uint32_t it = cs.it + c, slot = it % DEC_STAGES;
      mbar_wait(&sh.empty[slot], ((it / DEC_STAGES) & 1) ^ 1);  // all 8 consumer warps are done with the slot's previous stage
      const uint32_t cbytes = static_cast<uint32_t>(min(1024, units - 1024 * c) * 8);
      mbar_arrive_expect_tx(&sh.full[slot], static_cast<uint32_t>(__popc(sh.active_mask)) * cbytes);
      const uint64_t policy = l2_policy_evict_last();
      for (int q = 0; q < DB_MAXB; ++q)
        if ((sh.active_mask >> q) & 1) bulk_g2s(ring + slot * DEC_STAGE_BYTES + q * DB_GSEQ_STRIDE, buf + q * seq_stride + 1024 * c, cbytes, &sh.full[slot], policy);
    };
    cbar();  // every warp of this CTA is done with the previous phase: its epilogues' / combines' stores are issued
    lap(0);
    if (gcta && threadIdx.x == 0) gcta[0] = global_ns();
    if (threadIdx.x == 0) {
      unsigned long long* cnt = static_cast<unsigned long long*>(sync_cnt);
      // relaxed on purpose: a release would put a MEMBAR.GPU (microseconds while the SM's bulk copies are in flight) on the critical path of
      // every exchange. This CTA's LL stores were issued before the barrier above and drain to the L2 ahead of this reduction in practice;
      // if one ever lags, its tag is stale in the copy and the reader fetches that unit itself (ll_repair).
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(cnt) : "memory");
      const unsigned long long want = (static_cast<unsigned long long>(sh.epoch) * gathers_per_launch + cs.gathers + 1ull) * gridDim.x;
      uint32_t spins = 0;
      for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(cnt) : "memory");
        if (v >= want) break;
        __nanosleep(64);
        if (++spins > EMX_SPIN_LIMIT) __trap();
      }
      asm volatile("fence.proxy.async;" ::: "memory");  // the bulk copies below read what the generic proxy has just observed
      lap(1);
      if (gcta) gcta[1] = global_ns();
      for (int c = 0; c < min(nc, DEC_STAGES); ++c) issue_chunk(c);
      lap(2);
    }
    ++cs.gathers;
    // The chunks are ORDINARY ring stages from here on: wait for `full`, read, release `empty` per warp. A slot is handed back as soon as
    // the 8 warps have read it, so the producer warps refill the ring with the first weight stages of the phase while the later chunks
    // are still being parked (the hot loop used to start on an empty ring, one loaded-L2 latency after the gather).
#pragma unroll 1
    for (int c = 0; c < nc; ++c) {
      const uint32_t it = cs.it + c, slot = it % DEC_STAGES;
      mbar_wait(&sh.full[slot], (it / DEC_STAGES) & 1);
      if (c == 0 && threadIdx.x == 0) lap(3);
#pragma unroll 1
      for (int hh = 0; hh < 2; ++hh) {
        const int hc = 2 * c + hh;
        const int ks = half_ksteps(K, hc, warp);
        if (ks > 0) {  // warp-uniform
          const uint32_t base = ring_s + slot * DEC_STAGE_BYTES + n * DB_GSEQ_STRIDE + 128 * (8 * warp + 4 * hh) + 32 * t;
          uint32_t r[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint4 v = make_uint4(0u, tag, 0u, tag);
            if (act && i < ks) v = lds128(base + 128 * (i >> 1) + 16 * (i & 1));  // {payload, tag, payload, tag}
            if ((v.y != tag) | (v.w != tag)) {  // (never in practice) not published when the copy read it: fetch the pair from the L2
              const uint2 fix = ll_repair(buf + n * seq_stride + 1024 * c + 16 * (8 * warp + 4 * hh + (i >> 1)) + 4 * t + 2 * (i & 1), tag);
              v.x = fix.x, v.z = fix.y;
              if (gprof) gprof[6] += 1;
            }
            r[2 * i] = v.x, r[2 * i + 1] = v.z;
          }
          if (norm) {
#pragma unroll
            for (int q = 0; q < 16; ++q) ss += sumsq2(r[q]);
            const int ub = g_unit(hc, warp, 0, 0);  // this trip covers units [ub, ub + 64): own rows of the residual stream among them? (warp-uniform)
            if (ub < re2 && ub + 64 > rb2) {
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const int u = g_unit(hc, warp, t, q);
                if (act && (q >> 1) < ks && u >= rb2 && u < re2) sh.resid[n][u - rb2] = r[q];
              }
            }
          }
          tmem_st_32x16(tm + 16 * hc, r);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.empty[slot]);
      if (threadIdx.x == 0 && c + DEC_STAGES < nc) issue_chunk(c + DEC_STAGES);  // (waits for the slowest warp's release of chunk c)
    }
    if (threadIdx.x == 0) lap(4);
    cs.it += nc;
  }
  if (!norm) {
    tmem_st_wait();
    return;
  }
  const long long tg_tail = gprof ? global_ns() : 0;
  // sum of squares of sequence n: the 4 lanes of a quad, then the 8 warps
  ss += __shfl_xor_sync(0xffffffffu, ss, 1);
  ss += __shfl_xor_sync(0xffffffffu, ss, 2);
  if (t == 0) sh.red[parity & 1][warp][n] = ss;
  asm volatile("cp.async.wait_group 0;" ::: "memory");  // this thread's share of the norm weights has landed
  tmem_st_wait();
  cbar();  // partial sums and everybody's norm weights are visible
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < DEC_CWARPS; ++w) tot += sh.red[parity & 1][w][n];
  const float rs = 1.0f / sqrtf(tot / K + eps);
#pragma unroll 1
  for (int hc = 0; hc < 2 * nc; ++hc) {
    const int ks = half_ksteps(K, hc, warp);
    if (ks > 0) {
      uint32_t r[16];
      tmem_ld_32x16(tm + 16 * hc, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        if ((q >> 1) < ks) {
          // bf16(w * bf16(x * rs)) on both halves of the word: fp32 multiply, ONE packed round (cvt.rn.bf16x2.f32), then a packed bf16
          // multiply — the product of two bf16 is exact in fp32, so HMUL2.BF16's single rounding is the reference's
          const uint32_t g = ln_s[g_unit(hc, warp, t, q)], v = r[q];
          const __nv_bfloat162 y = __floats2bfloat162_rn(bf16_lo(v) * rs, bf16_hi(v) * rs);
          const __nv_bfloat162 o = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&g), y);
          r[q] = *reinterpret_cast<const uint32_t*>(&o);
        }
      }
      tmem_st_32x16(tm + 16 * hc, r);
    }
  }
  tmem_st_wait();
  cbar();  // everybody is done with ln_s: the next norm's weights may be fetched into it
  if (gprof && threadIdx.x == 0 && !embed) gprof[5] += global_ns() - tg_tail;
}

// ---- consumer: tensor-core dot products of one weight phase, 8 sequences at once ---------------------------------------------
// epi(row, v[4], n) is called by all 32 lanes of ONE warp per row group: lane = 8 a + n handles rows row .. row + 3 (row = r0 + 4 a) of
// sequence n; rows >= r_end (short last group) must be ignored by the callee.
template <typename Epi>
__device__ __forceinline__ void consume_phase_b(int K, int r_begin, int r_end, const uint8_t* ring, uint64_t* full, uint64_t* empty, BCons& cs, uint32_t tm,
                                                float* part, int warp, int lane, Epi&& epi) {
  const uint32_t a_lane_off = ((lane & 7) + ((lane >> 3) & 1) * 8) * DEC_ROWSTRIDE + (lane >> 4) * 16;
  const int kbeg = warp * DEC_KW;
  for (int r0 = r_begin; r0 < r_end; r0 += DEC_GROUP) {
    float c[4][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.f;
    int chunk = 0;
    for (int k0 = 0; k0 < K; k0 += DEC_KC, ++chunk) {
      const int klen = min(DEC_KC, K - k0);
      const int slot = cs.it % DEC_STAGES;
      const uint32_t ph = (cs.it / DEC_STAGES) & 1;
      const int ksteps = min(DEC_KW / 16, (klen - kbeg) / 16);  // <= 0: this warp's slice is past the K tail
      uint32_t b[32];
      if (ksteps > 0) tmem_ld_32x32(tm + 32 * chunk, b);  // in flight while the stage lands
      mbar_wait(&full[slot], ph);
      if (ksteps > 0) {
        tmem_ld_wait();
        const uint32_t a_base = smem_u32(ring + slot * DEC_STAGE_BYTES) + a_lane_off + kbeg * 2;
        if (ksteps == DEC_KW / 16) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {  // two batches of 8 k-steps: 8 ldmatrix in flight, then 8 HMMA on four accumulator chains
            uint32_t a[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) ldmatrix_x4(a_base + (h * 8 + j) * 32, a[j][0], a[j][1], a[j][2], a[j][3]);
#pragma unroll
            for (int j = 0; j < 8; ++j) mma_bf16_16816(c[j & 3], a[j][0], a[j][1], a[j][2], a[j][3], b[2 * (h * 8 + j)], b[2 * (h * 8 + j) + 1]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < DEC_KW / 16; ++j) {
            if (j < ksteps) {
              uint32_t a0, a1, a2, a3;
              ldmatrix_x4(a_base + j * 32, a0, a1, a2, a3);
              mma_bf16_16816(c[j & 3], a0, a1, a2, a3, b[2 * j], b[2 * j + 1]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      ++cs.it;
    }
    // accumulator tile: c[.][0..1] -> row lane/4, sequences 2 t, 2 t + 1; c[.][2..3] -> row lane/4 + 8
    const uint32_t buf = cs.group % DEC_PARTBUFS;
    float* pb = part + buf * DB_PART_FLOATS;
    {
      const int g = lane >> 2, t = lane & 3;
      *reinterpret_cast<float2*>(pb + warp * 128 + g * 8 + 2 * t) = make_float2((c[0][0] + c[1][0]) + (c[2][0] + c[3][0]), (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]));
      *reinterpret_cast<float2*>(pb + warp * 128 + (g + 8) * 8 + 2 * t) =
          make_float2((c[0][2] + c[1][2]) + (c[2][2] + c[3][2]), (c[0][3] + c[1][3]) + (c[2][3] + c[3][3]));
    }
    if (warp == static_cast<int>(cs.group % DEC_CWARPS)) {  // rotating epilogue duty (see decode_mega.cu)
      part_sync(buf);
      const int n = lane & 7, a = lane >> 3;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int w = 0; w < DEC_CWARPS; ++w) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += pb[w * 128 + (4 * a + i) * 8 + n];
      }
      epi(r0 + 4 * a, v, n);
      __syncwarp();
    } else {
      part_arrive(buf);
    }
    ++cs.group;
  }
}

// ---- attention phase -----------------------------------------------------------------------------------------------------
struct AttState {
  float m, l, acc[4], q[4];
  int row;  // (sequence * heads + head) the state and q belong to, -1: none
};

// 4 consecutive elements (4 lane .. 4 lane + 3) of a 128-wide head vector out of two LL units (one 16-byte load per lane, issued early and
// polled when the values are needed), optionally through RoPE (rotate_half form: the partner element d +- 64 lives in lane ^ 16).
// x_embed = bf16(bf16(x cos) + bf16(rotate_half(x) sin)), tables are bf16.
struct Head4 {
  const uint64_t* src;
  uint64_t a, b;
  __device__ __forceinline__ void issue(const uint64_t* units, int lane) {
    src = units + 2 * lane;
    ll_load2(src, a, b);
  }
  __device__ __forceinline__ void finish(const uint32_t* rope, int lane, uint32_t tag, float (&out)[4]) {
    uint32_t spins = 0;
    while (!(tag_ok(a, tag) && tag_ok(b, tag))) {
      ll_load2(src, a, b);
      if (++spins > EMX_SPIN_LIMIT) __trap();
    }
    const uint32_t w0 = static_cast<uint32_t>(a), w1 = static_cast<uint32_t>(b);
    const float x[4] = {bf16_lo(w0), bf16_hi(w0), bf16_lo(w1), bf16_hi(w1)};
    if (!rope) {
#pragma unroll
      for (int e = 0; e < 4; ++e) out[e] = x[e];
      return;
    }
    const int ci = (2 * lane) & 31;
    const uint32_t cw[2] = {rope[ci], rope[ci + 1]}, sw[2] = {rope[32 + ci], rope[32 + ci + 1]};
    const float sgn = (lane < 16) ? -1.f : 1.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pr = __shfl_xor_sync(0xffffffffu, x[e], 16);
      const float cs = (e & 1) ? bf16_hi(cw[e >> 1]) : bf16_lo(cw[e >> 1]), sn = (e & 1) ? bf16_hi(sw[e >> 1]) : bf16_lo(sw[e >> 1]);
      out[e] = bf16_round(bf16_round(x[e] * cs) + bf16_round(sgn * pr * sn));
    }
  }
};

// One online-softmax step of this warp over its 16 keys of the stage (keys 16 warp .. 16 warp + 15 of the item's 128; K rows at
// [page][row][128] from byte 0, V rows from byte 32 K). Lane l owns head elements 4 l .. 4 l + 3 of q, K, V and the accumulator.
__device__ __forceinline__ void att_block(const uint8_t* stage, int warp, int lane, int nvalid, float scale, AttState& s) {
  const uint32_t kb = smem_u32(stage) + (warp >> 2) * DB_PAGE_BYTES + (warp & 3) * 16 * (DEC_HD * 2) + lane * 8;
  const uint32_t vb = kb + 2 * DB_PAGE_BYTES;
  float pr[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint2 kw = lds64(kb + i * (DEC_HD * 2));
    pr[i] = fmaf(s.q[3], bf16_hi(kw.y), fmaf(s.q[2], bf16_lo(kw.y), fmaf(s.q[1], bf16_hi(kw.x), s.q[0] * bf16_lo(kw.x))));
  }
  // transposing butterfly: 16 partial dot products per lane -> the complete score of key (lane >> 1) in every lane
  float v8[8], v4[4], v2[2], sc;
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = (up ? pr[i + 8] : pr[i]) + __shfl_xor_sync(0xffffffffu, up ? pr[i] : pr[i + 8], 16);
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) v4[i] = (up ? v8[i + 4] : v8[i]) + __shfl_xor_sync(0xffffffffu, up ? v8[i] : v8[i + 4], 8);
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) v2[i] = (up ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, up ? v4[i] : v4[i + 2], 4);
  }
  {
    const bool up = lane & 2;
    sc = (up ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, up ? v2[0] : v2[1], 2);
  }
  sc += __shfl_xor_sync(0xffffffffu, sc, 1);
  const bool valid = (lane >> 1) < nvalid;
  sc = valid ? sc * scale : -INFINITY;
  const float m_new = fmaxf(s.m, warp_max(sc));  // finite: nvalid >= 1
  const float pk = valid ? __expf(sc - m_new) : 0.f;
  const float lsum = warp_sum((lane & 1) ? 0.f : pk);  // every key sits in two lanes
  const float corr = __expf(s.m - m_new);              // exp(-inf) = 0 on the first block
  s.l = s.l * corr + lsum, s.m = m_new;
#pragma unroll
  for (int e = 0; e < 4; ++e) s.acc[e] *= corr;
  const float pb = bf16_round(pk);  // flash-attn: P is bf16 for the PV product, the row sum stays fp32
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float pi = __shfl_sync(0xffffffffu, pb, 2 * i);
    if (i < nvalid) {  // rows beyond the valid keys may hold anything
      const uint2 vw = lds64(vb + i * (DEC_HD * 2));
      s.acc[0] = fmaf(pi, bf16_lo(vw.x), s.acc[0]), s.acc[1] = fmaf(pi, bf16_hi(vw.x), s.acc[1]);
      s.acc[2] = fmaf(pi, bf16_lo(vw.y), s.acc[2]), s.acc[3] = fmaf(pi, bf16_hi(vw.y), s.acc[3]);
    }
  }
}

__device__ __forceinline__ void att_merge(float& M, float& den, float (&num)[4], float m2, float l2, const float (&a2)[4]) {
  if (l2 > 0.f) {
    const float Mn = fmaxf(M, m2);
    const float wo = (M == -INFINITY) ? 0.f : __expf(M - Mn), wn = __expf(m2 - Mn);
#pragma unroll
    for (int e = 0; e < 4; ++e) num[e] = num[e] * wo + a2[e] * wn;
    den = den * wo + l2 * wn, M = Mn;
  }
}

// End of a row segment, run by the rotating warp once all 8 warp states are in `pb`: merge them. The CTA that owns the row's FIRST item
// combines: it folds in the segments of the CTAs after it and the token being decoded (RoPE, KV append, one more online-softmax step)
// and publishes the head output in o_proj's exchange order; every other segment is published as a partial. Publishers never wait, and a
// combiner only waits for segments that were started at the same time as its own, so no CTA ever waits for another one's WHOLE share
// (the mirror-image choice — combining in the CTA that owns the row's last item — chains every CTA behind its predecessor).
// What the finishing warp of a row it COMBINES fetches when the row starts, so that the L2 round trips are over when the row ends (a flush
// that waits for them holds this warp back for ~2 us under the weight stream, and the ring with it: a slot is released by the slowest warp):
// k and v of the token being decoded, and the partial of the next CTA's segment of the row (published at the START of that CTA's share,
// i.e. usually long before; if it is not there yet, att_finish polls for it as before).
struct AttPre {
  Head4 hk, hv;
  uint64_t sa[3], sb[3];
  bool kv, seg;
};

__device__ void att_finish(const emx_decode_batch_params& p, const BatchShared& sh, const float* pb, int lane, const AttItem& it, int g, int layer,
                           uint32_t tag, const float (&q)[4], float scale, AttPre& pre) {
  const int n = it.seq, H = p.hidden;
  const int row_start = sh.att_off[n] + it.head * it.pp, row_last = row_start + it.pp - 1;
  const int seg_start = max(sh.att_i0, row_start);
  const bool combiner = (seg_start == row_start);
  const uint64_t* qkv = static_cast<const uint64_t*>(p.qkv) + static_cast<long>(n) * (3 * H / 2);
  Head4& hk = pre.hk;
  Head4& hv = pre.hv;
  if (combiner && !pre.kv) {  // k and v of the token being decoded (normally fetched when the row started)
    hk.issue(qkv + H / 2 + it.head * (DEC_HD / 2), lane);
    hv.issue(qkv + H + it.head * (DEC_HD / 2), lane);
  }
  float M = -INFINITY, den = 0.f, num[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int w = 0; w < DEC_CWARPS; ++w) {
    const float4 a = *reinterpret_cast<const float4*>(pb + w * DB_ATT_WSTRIDE + 4 * lane);
    const float a4[4] = {a.x, a.y, a.z, a.w};
    att_merge(M, den, num, pb[w * DB_ATT_WSTRIDE + DEC_HD], pb[w * DB_ATT_WSTRIDE + DEC_HD + 1], a4);
  }
  uint64_t* rowpart = static_cast<uint64_t*>(p.part) + (static_cast<long>(n) * p.heads + it.head) * (DB_MAXSEG * DB_PARTU);
  if (!combiner) {
    uint64_t* dst = rowpart + (seg_start - row_start) * DB_PARTU;
    if (lane == 0) ll_store(dst, __float_as_uint(M), tag), ll_store(dst + 1, __float_as_uint(den), tag);
#pragma unroll
    for (int e = 0; e < 4; ++e) ll_store(dst + 2 + 4 * lane + e, __float_as_uint(num[e]), tag);
    return;
  }
  // ---- segments of the CTAs after this one (owner(g) = the CTA c with T c / G <= g < T (c + 1) / G); g is the last item of ours
  const long T = sh.att_T, G = gridDim.x;
  for (int gc = g + 1; gc <= row_last;) {
    const long c = ((gc + 1) * G - 1) / T;
    const uint64_t* src = rowpart + (gc - row_start) * DB_PARTU;
    uint32_t w[6];
    bool have = pre.seg && gc == g + 1;  // the segment fetched ahead is the one that follows ours
    if (have) {
#pragma unroll
      for (int i = 0; i < 3; ++i) have &= tag_ok(pre.sa[i], tag) && tag_ok(pre.sb[i], tag);
    }
    if (have) {
#pragma unroll
      for (int i = 0; i < 3; ++i) w[2 * i] = static_cast<uint32_t>(pre.sa[i]), w[2 * i + 1] = static_cast<uint32_t>(pre.sb[i]);
    } else {
      ll_fetch_pairs<3>([&](int i) -> const uint64_t* { return i == 0 ? src : src + 2 + 4 * lane + 2 * (i - 1); }, w, tag, true);
    }
    const float a4[4] = {__uint_as_float(w[2]), __uint_as_float(w[3]), __uint_as_float(w[4]), __uint_as_float(w[5])};
    att_merge(M, den, num, __uint_as_float(w[0]), __uint_as_float(w[1]), a4);
    gc = static_cast<int>(T * (c + 1) / G);
  }
  // ---- the token being decoded: k (RoPE) and v arrive from the projection phases of this layer
  const int pos = sh.pos[n];
  float kr[4], vn[4];
  hk.finish(sh.rope[n], lane, tag, kr);
  hv.finish(nullptr, lane, tag, vn);
  const long dst = kv_page_off(p, layer, sh.table[n][pos / DB_PAGE], it.head) + (pos % DB_PAGE) * DEC_HD + 4 * lane;
  *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.k_cache) + dst) = make_uint2(pack_bf16(kr[0], kr[1]), pack_bf16(kr[2], kr[3]));
  *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p.v_cache) + dst) = make_uint2(pack_bf16(vn[0], vn[1]), pack_bf16(vn[2], vn[3]));
  const float s_new = warp_sum(fmaf(q[3], kr[3], fmaf(q[2], kr[2], fmaf(q[1], kr[1], q[0] * kr[0])))) * scale;
  const float Mn = fmaxf(M, s_new);
  const float wo = (M == -INFINITY) ? 0.f : __expf(M - Mn), pn = __expf(s_new - Mn), pnb = bf16_round(pn);
  den = den * wo + pn;
#pragma unroll
  for (int e = 0; e < 4; ++e) num[e] = (num[e] * wo + pnb * vn[e]) / den;
  uint64_t* attn = static_cast<uint64_t*>(p.attn) + static_cast<long>(n) * ((H / 2 + 15) & ~15);
  const int u = it.head * (DEC_HD / 2) + 2 * lane;
  ll_store(attn + ll_pos(u), pack_bf16(num[0], num[1]), tag);
  ll_store(attn + ll_pos(u + 1), pack_bf16(num[2], num[3]), tag);
}

__device__ __noinline__ void attention_phase_b(const emx_decode_batch_params& p, const BatchShared& sh, const uint8_t* ring, uint64_t* full, uint64_t* empty, BCons& cs,
                                  float* part, int layer, uint32_t tag, int warp, int lane) {
  const float scale = rsqrtf(static_cast<float>(DEC_HD));
  const int i0 = sh.att_i0, i1 = sh.att_i1;
  AttState s;
  s.row = -1, s.m = -INFINITY, s.l = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) s.acc[e] = 0.f, s.q[e] = 0.f;
  const uint64_t* qkv = static_cast<const uint64_t*>(p.qkv);
  const int H = p.hidden;
  Head4 hq;
  bool hq_issued = false;
  AttPre pre;
  pre.kv = pre.seg = false;
#pragma unroll 1
  for (int g = i0; g < i1; ++g) {
    const AttItem it = att_item(sh, g);
    const int row = it.seq * p.heads + it.head;
    const bool new_row = row != s.row;
    if (new_row) {
      // q of (sequence, head) — projected two phases ago, long there. For the first row of the phase its L2 round trip overlaps the wait for
      // the stage; for every later row it was issued while the last item of the row before was still streaming (see below)
      if (!hq_issued) hq.issue(qkv + static_cast<long>(it.seq) * (3 * H / 2) + it.head * (DEC_HD / 2), lane);
      // the warp that will finish this row (the flush below: cs.group does not change before it) fetches what a COMBINER needs now
      const int row_start = sh.att_off[it.seq] + it.head * it.pp;
      pre.kv = pre.seg = false;
      if (warp == static_cast<int>(cs.group % DEC_CWARPS) && g == row_start) {
        const uint64_t* qs = qkv + static_cast<long>(it.seq) * (3 * H / 2);
        pre.hk.issue(qs + H / 2 + it.head * (DEC_HD / 2), lane);
        pre.hv.issue(qs + H + it.head * (DEC_HD / 2), lane);
        pre.kv = true;
        if (row_start + it.pp > i1) {  // the row runs on into the next CTA's share: that segment starts at item i1
          const uint64_t* src = static_cast<const uint64_t*>(p.part) + (static_cast<long>(it.seq) * p.heads + it.head) * (DB_MAXSEG * DB_PARTU) +
                                static_cast<long>(i1 - row_start) * DB_PARTU;
          ll_load2(src, pre.sa[0], pre.sb[0]);
          ll_load2(src + 2 + 4 * lane, pre.sa[1], pre.sb[1]);
          ll_load2(src + 2 + 4 * lane + 2, pre.sa[2], pre.sb[2]);
          pre.seg = true;
        }
      }
    }
    const int slot = cs.it % DEC_STAGES;
    const uint32_t ph = (cs.it / DEC_STAGES) & 1;
    mbar_wait(&full[slot], ph);
    if (new_row) {
      s.row = row, s.m = -INFINITY, s.l = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) s.acc[e] = 0.f;
      hq.finish(sh.rope[it.seq], lane, tag, s.q);
      hq_issued = false;
    }
    const int nvalid = min(16, sh.pos[it.seq] - (128 * it.j + 16 * warp));  // cached keys are positions 0 .. pos - 1
    if (nvalid > 0) att_block(ring + slot * DEC_STAGE_BYTES, warp, lane, nvalid, scale, s);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[slot]);
    ++cs.it;
    if (it.j == it.pp - 1 || g == i1 - 1) {  // end of the row, or of this CTA's share of it
      if (g + 1 < i1) {  // the next item starts a new row: its q travels while this row is flushed and the next stage lands
        const AttItem nx = att_item(sh, g + 1);
        hq.issue(qkv + static_cast<long>(nx.seq) * (3 * H / 2) + nx.head * (DEC_HD / 2), lane);
        hq_issued = true;
      }
      const uint32_t buf = cs.group % DEC_PARTBUFS;
      float* pb = part + buf * DB_PART_FLOATS;
      *reinterpret_cast<float4*>(pb + warp * DB_ATT_WSTRIDE + 4 * lane) = make_float4(s.acc[0], s.acc[1], s.acc[2], s.acc[3]);
      if (lane == 0) pb[warp * DB_ATT_WSTRIDE + DEC_HD] = s.m, pb[warp * DB_ATT_WSTRIDE + DEC_HD + 1] = s.l;
      if (warp == static_cast<int>(cs.group % DEC_CWARPS)) {
        part_sync(buf);
        att_finish(p, sh, pb, lane, it, g, layer, tag, s.q, scale, pre);
        __syncwarp();
      } else {
        part_arrive(buf);
      }
      ++cs.group;
      s.row = -1;  // (a new segment of the same row cannot follow inside one CTA, but the state must restart)
    }
  }
}

// ---- the kernel ----------------------------------------------------------------------------------------------------------
template <bool PROF>
__global__ void __launch_bounds__(DB_THREADS, 1) decode_batch_kernel(const emx_decode_batch_params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* ring = smem;
  float* part = reinterpret_cast<float*>(smem + DEC_STAGES * DEC_STAGE_BYTES);
  uint32_t* ln_s = reinterpret_cast<uint32_t*>(part + DEC_PARTBUFS * DB_PART_FLOATS);
  BatchShared& sh = *reinterpret_cast<BatchShared*>(reinterpret_cast<uint8_t*>(ln_s) + DB_LN_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  emx_decode_batch_state* st = p.state;

  if (tid < DB_MAXB) {
    const int n = tid;
    const int tok = static_cast<int>(ldg_cg_u32(&st->cur_token[n])), pos = static_cast<int>(ldg_cg_u32(&st->pos[n]));
    const int ngen = static_cast<int>(ldg_cg_u32(&st->n_generated[n])), fin = static_cast<int>(ldg_cg_u32(&st->finished[n]));
    const int lim = static_cast<int>(ldg_cg_u32(&st->limit[n]));
    // active: a live sequence with room for one more token (KV cache, RoPE tables and out_tokens all end at these limits) and at
    // least one cached key (a prefill always leaves some)
    const bool act = n < p.batch && !fin && ngen < lim && ngen < p.out_stride && pos >= 1 && pos < p.max_pages * DB_PAGE;
    sh.tok[n] = min(max(tok, 0), p.vocab - 1), sh.pos[n] = act ? pos : 0, sh.ngen[n] = ngen;
    const uint32_t m = __ballot_sync(0xffu, act);
    if (tid == 0) sh.active_mask = m, sh.epoch = ldg_cg_u32(&st->epoch), sh.issued[0] = 0, sh.issued[1] = 0;
  }
  if (tid >= 32 && tid < 32 + BPH_KINDS) build_phase(p, sh, tid - 32);
  if (tid == 64) {
    for (int s = 0; s < DEC_STAGES; ++s) {
      mbar_init(&sh.full[s], 1);
      mbar_init(&sh.empty[s], DEC_CWARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const uint32_t active_mask = sh.active_mask;
  if (!active_mask) return;  // every sequence is finished: nothing to do, uniformly
  for (int i = tid; i < DB_MAXB * p.max_pages; i += DB_THREADS) {
    const int n = i / p.max_pages, pg = i - n * p.max_pages;
    sh.table[n][pg] = ((active_mask >> n) & 1) ? __ldg(p.block_table + i) : 0;
  }
  for (int i = tid; i < DB_MAXB * 64; i += DB_THREADS) {
    const int n = i >> 6, j = i & 63;
    const uint32_t* tabp = reinterpret_cast<const uint32_t*>(j < 32 ? p.cos_tab : p.sin_tab);
    sh.rope[n][j] = ((active_mask >> n) & 1) ? __ldg(tabp + static_cast<long>(sh.pos[n]) * (DEC_HD / 4) + (j & 31)) : 0u;
  }
  if (tid == 0) {
    int T = 0;
    for (int n = 0; n < DB_MAXB; ++n) {
      const int pages = (sh.pos[n] + DB_PAGE - 1) / DB_PAGE;  // 0 for inactive sequences
      sh.att_pp[n] = (pages + 1) / 2, sh.att_off[n] = T;
      T += sh.att_pp[n] * p.heads;
    }
    sh.att_off[DB_MAXB] = T, sh.att_T = T;
    sh.att_i0 = static_cast<int>(static_cast<long>(T) * blockIdx.x / gridDim.x);
    sh.att_i1 = static_cast<int>(static_cast<long>(T) * (blockIdx.x + 1) / gridDim.x);
  }
  if (warp == 0) tmem_alloc(&sh.tmem_base, DB_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= DEC_CWARPS + DB_PWARPS) {
    if (p.l2_lookahead_stages > 0) batch_producer<true>(p, sh, ring, sh.full, sh.empty, lane, 0, sh.issued);
    return;
  }
  if (warp >= DEC_CWARPS) {
    batch_producer<false>(p, sh, ring, sh.full, sh.empty, lane, warp - DEC_CWARPS, sh.issued);
    return;
  }

  // ===================== consumer warps =====================
  const int L = p.layers, H = p.hidden, I = p.inter;
  const long sH = (H / 2 + 15) & ~15, sI = (I / 2 + 15) & ~15;  // per-sequence strides (units) of the ll_pos-ordered buffers
  const uint32_t tag0 = sh.epoch * static_cast<uint32_t>(L + 2) + 1u;  // tag0 + l: layer l; + L: final norm; + L + 1: argmax candidates
  // this thread's TMEM lane and its half of the 512 columns (two consumer warps share a lane quarter)
  const uint32_t tm = sh.tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>((warp >> 2) * 256);
  uint64_t* xd = static_cast<uint64_t*>(p.x);
  uint64_t* xo = static_cast<uint64_t*>(p.xo);
  uint64_t* qkv = static_cast<uint64_t*>(p.qkv);
  uint64_t* attn = static_cast<uint64_t*>(p.attn);
  uint64_t* hbuf = static_cast<uint64_t*>(p.h);
  const int rb = sh.r_begin[BPH_O], rb2 = rb >> 1, re2 = sh.r_end[BPH_O] >> 1;
  long long* dbg = (PROF && blockIdx.x == 0 && tid == 0) ? reinterpret_cast<long long*>(p.dbg) : nullptr;

  BCons cs{0, 0, 0};
  long long* gprof = (PROF && dbg) ? dbg + 2 * (BPH_STEPS * L + 1) + 8 + 16 * gridDim.x : nullptr;  // (host zeroes it)
  float best = -INFINITY;  // lm_head: this lane's best logit of ITS sequence (lane & 7)
  int best_i = 0x7fffffff;

  ln_fetch_async_b(static_cast<const __nv_bfloat16*>(p.ln1), ln_s, H);
  const int n_steps = BPH_STEPS * L + 1;
  int layer = 0, kind = BPH_Q;
#pragma unroll 1
  for (int step = 0; step < n_steps; ++step) {
    const uint32_t tag = tag0 + layer;  // (the lm_head step has layer == L)
    if (PROF && dbg) dbg[2 * step] = global_ns();
    // every CTA: entry / exit times of the four gathers of layer 1 (skew of the exchanges)
    const int gslot = (kind == BPH_Q) ? 0 : (kind == BPH_O) ? 1 : (kind == BPH_GATEUP) ? 2 : (kind == BPH_DOWN) ? 3 : -1;
    long long* gdbg = (PROF && p.dbg && tid == 0 && layer == 1 && gslot >= 0) ? reinterpret_cast<long long*>(p.dbg) + 2 * n_steps + 8 + 8 * blockIdx.x + 2 * gslot : nullptr;
    if (PROF && gdbg) gdbg[0] = global_ns();
    if (kind == BPH_Q || kind == BPH_GATEUP || kind == BPH_LMHEAD) {  // residual stream in + RMSNorm
      gather_b(kind == BPH_GATEUP ? xo : xd, sH, H, true, step == 0 ? static_cast<const __nv_bfloat16*>(p.embed) : nullptr, tag, tm, sh, ring, cs, ln_s,
               p.rms_eps, kind == BPH_GATEUP ? 1u : 0u, rb2, re2, warp, lane, p.sync, 4u * L, gprof, gdbg ? gdbg + 8 * gridDim.x : nullptr);
      if (kind != BPH_LMHEAD) {
        const __nv_bfloat16* next_w = (kind == BPH_Q)    ? static_cast<const __nv_bfloat16*>(p.ln2) + static_cast<long>(layer) * H
                                      : (layer + 1 < L) ? static_cast<const __nv_bfloat16*>(p.ln1) + static_cast<long>(layer + 1) * H
                                                        : static_cast<const __nv_bfloat16*>(p.final_norm);
        ln_fetch_async_b(next_w, ln_s, H);
      }
    } else if (kind == BPH_O || kind == BPH_DOWN) {  // a plain vector in: the attention output for o_proj, the SwiGLU output for down_proj
      const bool o = (kind == BPH_O);
      gather_b(o ? attn : hbuf, o ? sH : sI, o ? H : I, false, nullptr, tag, tm, sh, ring, cs, ln_s, 0.f, 0u, rb2, re2, warp, lane, p.sync, 4u * L, gprof, gdbg ? gdbg + 8 * gridDim.x : nullptr);
    }
    if (PROF && dbg) dbg[2 * step + 1] = global_ns();
    if (PROF && gdbg) gdbg[1] = global_ns();
    if (kind == BPH_ATT) {
      attention_phase_b(p, sh, ring, sh.full, sh.empty, cs, part, layer, tag, warp, lane);
    } else {
      consume_phase_b(sh.K[kind], sh.r_begin[kind], sh.r_end[kind], ring, sh.full, sh.empty, cs, tm, part, warp, lane,
                      [&](int row, const float (&v)[4], int n) {
                        const int r_end = sh.r_end[kind];
                        if (!((active_mask >> n) & 1) || row >= r_end) return;
                        const bool two = row + 2 < r_end;  // rows row + 2, row + 3 exist (always, for gate/up quads)
                        if (kind <= BPH_V) {
                          uint64_t* dst = qkv + static_cast<long>(n) * (3 * H / 2) + kind * (H >> 1) + (row >> 1);
                          ll_store(dst, pack_bf16(v[0], v[1]), tag);
                          if (two) ll_store(dst + 1, pack_bf16(v[2], v[3]), tag);
                        } else if (kind == BPH_GATEUP) {
                          // rows (gate_i, up_i, gate_i+1, up_i+1): two SwiGLU outputs = one unit
                          const float h0 = bf16_round(bf16_round(silu(bf16_round(v[0]))) * bf16_round(v[1]));
                          const float h1 = bf16_round(bf16_round(silu(bf16_round(v[2]))) * bf16_round(v[3]));
                          ll_store(hbuf + n * sI + ll_pos(row >> 2), pack_bf16(h0, h1), tag);
                        } else if (kind == BPH_LMHEAD) {
#pragma unroll
                          for (int i = 0; i < 4; ++i) {
                            if (row + i < r_end) {
                              const float lv = bf16_round(v[i]);
                              if (p.logits_out) p.logits_out[static_cast<long>(n) * p.vocab + row + i] = lv;
                              if (lv > best) best = lv, best_i = row + i;  // rows ascend per thread: strict '>' keeps the lowest index
                            }
                          }
                        } else {  // o_proj / down_proj: + residual; down_proj feeds the NEXT layer (tag + 1)
                          uint64_t* dst = (kind == BPH_O ? xo : xd) + n * sH;
                          const uint32_t t2 = (kind == BPH_O) ? tag : tag + 1;
                          const int u = row >> 1;
                          const uint32_t r0 = sh.resid[n][u - rb2];
                          ll_store(dst + ll_pos(u), pack_bf16(bf16_lo(r0) + bf16_round(v[0]), bf16_hi(r0) + bf16_round(v[1])), t2);
                          if (two) {
                            const uint32_t r1 = sh.resid[n][u + 1 - rb2];
                            ll_store(dst + ll_pos(u + 1), pack_bf16(bf16_lo(r1) + bf16_round(v[2]), bf16_hi(r1) + bf16_round(v[3])), t2);
                          }
                        }
                      });
    }
    if (++kind == BPH_LMHEAD) {
      kind = BPH_Q;
      if (++layer == L) kind = BPH_LMHEAD;
    }
  }
  if (PROF && dbg) dbg[2 * n_steps] = global_ns();

  // ---- greedy argmax per sequence: lanes -> CTA -> grid (lowest index wins ties, as torch.argmax) ----
  cbar();  // every epilogue is done: the partial buffers are free
  float* s_bv = part;
  int* s_bi = reinterpret_cast<int*>(part + DEC_CTHREADS);
  s_bv[tid] = best, s_bi[tid] = best_i;
  cbar();
  const uint32_t tag_a = tag0 + L + 1;
  uint64_t* cand = static_cast<uint64_t*>(p.argmax_part);  // [grid][8][2] units: value bits, index
  if (tid < DB_MAXB) {
    float b = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = 0; k < DEC_CTHREADS / DB_MAXB; ++k) {  // lanes with lane % 8 == tid of every warp
      const float v = s_bv[8 * k + tid];
      const int i = s_bi[8 * k + tid];
      if (v > b || (v == b && i < bi)) b = v, bi = i;
    }
    ll_store(cand + (blockIdx.x * DB_MAXB + tid) * 2, __float_as_uint(b), tag_a);
    ll_store(cand + (blockIdx.x * DB_MAXB + tid) * 2 + 1, static_cast<uint32_t>(bi), tag_a);
  }
  if (blockIdx.x == 0) {
    const int n = warp;  // consumer warp n of CTA 0 finishes sequence n
    if ((active_mask >> n) & 1) {
      float b = -INFINITY;
      int bi = 0x7fffffff;
      for (int c = lane; c < static_cast<int>(gridDim.x); c += 32) {
        const float v = __uint_as_float(ll_wait(cand + (c * DB_MAXB + n) * 2, tag_a, true));
        const int i = static_cast<int>(ll_wait(cand + (c * DB_MAXB + n) * 2 + 1, tag_a, true));
        if (v > b || (v == b && i < bi)) b = v, bi = i;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, b, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > b || (ob == b && oi < bi)) b = ob, bi = oi;
      }
      if (lane == 0) {
        const int ng = sh.ngen[n];
        p.out_tokens[static_cast<long>(n) * p.out_stride + ng] = bi;
        st->cur_token[n] = bi;
        st->pos[n] = sh.pos[n] + 1;
        st->n_generated[n] = ng + 1;
        if (p.eos_token >= 0 && bi == p.eos_token) st->finished[n] = 1;
      }
    }
    if (tid == 0) st->epoch = sh.epoch + 1u;
  }
  tc_fence_before();
  cbar();
  if (warp == 0) tmem_dealloc(sh.tmem_base, DB_TMEM_COLS);
}

}  // namespace emx

extern "C" int emx_decode_batch_step(const emx_decode_batch_params* params, cudaStream_t stream) {
  using namespace emx;
  const emx_decode_batch_params& p = *params;
  EMX_REQUIRE(p.head_dim == DEC_HD, "emx_decode_batch_step: head_dim %d not supported (128)", p.head_dim);
  EMX_REQUIRE(p.batch >= 1 && p.batch <= DB_MAXB, "emx_decode_batch_step: batch %d not in 1..%d", p.batch, DB_MAXB);
  EMX_REQUIRE(p.hidden % 16 == 0 && p.inter % 16 == 0 && p.vocab % 2 == 0, "emx_decode_batch_step: hidden/inter must be multiples of 16, vocab even");
  EMX_REQUIRE(p.heads * DEC_HD == p.hidden, "emx_decode_batch_step: heads x head_dim must equal hidden");
  EMX_REQUIRE(p.x && p.xo && p.qkv && p.attn && p.h && p.part && p.argmax_part && p.state && p.out_tokens && p.block_table && p.sync,
              "emx_decode_batch_step: null pointer");
  EMX_REQUIRE(p.hidden * 2 <= DB_LN_BYTES, "emx_decode_batch_step: hidden > %d not supported by the fused gather + RMSNorm", DB_LN_BYTES / 2);
  EMX_REQUIRE(p.inter <= 8 * DEC_KC && p.hidden <= 8 * DEC_KC, "emx_decode_batch_step: activation vector exceeds the 256 TMEM columns of a thread");
  EMX_REQUIRE(p.page_size == DB_PAGE, "emx_decode_batch_step: page_size must be %d", DB_PAGE);
  EMX_REQUIRE(p.max_pages >= 1 && p.max_pages <= DB_MAX_PAGES && p.max_pages <= 2 * DB_MAXSEG,
              "emx_decode_batch_step: block table of %d pages exceeds %d", p.max_pages, 2 * DB_MAXSEG);
  EMX_REQUIRE(p.out_stride >= 1, "emx_decode_batch_step: out_stride");
  int dev = 0;
  const int grid = device_sms(&dev);
  bool* attr_set = device_attr_flag(ATTR_DECODE_BATCH);
  if (grid < 0 || !attr_set) return -2;
  if (!*attr_set) {
    EMX_CHECK_CUDA(cudaFuncSetAttribute(decode_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_SMEM));
    EMX_CHECK_CUDA(cudaFuncSetAttribute(decode_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, DB_SMEM));
    int per_sm = 0;
    EMX_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_batch_kernel<true>, DB_THREADS, DB_SMEM));
    EMX_REQUIRE(per_sm >= 1, "emx_decode_batch_step: kernel does not fit on an SM (smem %d)", DB_SMEM);
    *attr_set = true;
  }
  EMX_REQUIRE(p.hidden / 2 / grid + 2 <= DB_MAX_RESID, "emx_decode_batch_step: hidden too large for the residual staging buffer");
  void* args[] = {const_cast<emx_decode_batch_params*>(params)};
  void* fn = p.dbg ? reinterpret_cast<void*>(decode_batch_kernel<true>) : reinterpret_cast<void*>(decode_batch_kernel<false>);
  EMX_CHECK_CUDA(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(DB_THREADS), args, DB_SMEM, stream));
  return 0;
}

extern "C" int emx_decode_batch_smem(void) { return emx::DB_SMEM; }
