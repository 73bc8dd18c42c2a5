// Building blocks shared by the two persistent decode kernels (decode_mega.cu: one sequence per launch, attention on dedicated warps
// out of tensor memory; decode_batch.cu: up to 8 sequences per launch, K/V streamed through the weight ring): the static row partition,
// "LL" exchange units, the permuted activation layout, mma.sync / ldmatrix wrappers.
#pragma once
#include "common.cuh"

namespace emx {

constexpr int DEC_CWARPS = 8;                    // consumer warps
constexpr int DEC_CTHREADS = DEC_CWARPS * 32;    // 256
constexpr int DEC_GROUP = 16;                    // rows per ring stage == M of the MMA atom
constexpr int DEC_KC = 2048;                     // K elements per ring stage (4 KB per row segment)
constexpr int DEC_KW = DEC_KC / DEC_CWARPS;      // 256 columns per consumer warp per stage
constexpr int DEC_ROWSTRIDE = DEC_KC * 2 + 16;   // padded row stride (bytes): ldmatrix rows land in distinct bank groups
constexpr int DEC_STAGES = 3;
constexpr int DEC_STAGE_BYTES = DEC_GROUP * DEC_ROWSTRIDE;  // 65792
constexpr int DEC_PARTBUFS = 4;                  // partial-sum buffers (a warp is never more than 3 ring stages ahead of the epilogue warp)
constexpr int DEC_HD = 128;                      // head_dim supported by the decode kernels (Llama-2)

// [r_begin, r_end) of an n_rows-row phase owned by CTA `cta` of `grid`: contiguous, balanced to one granule (row pairs are one LL unit;
// gate/up rows come in quads = two SwiGLU outputs = one LL unit), together covering every row exactly once. Host-callable so that the
// CPU test suite checks the very formula the kernel uses (emx_decode_phase_rows).
__host__ __device__ __forceinline__ void phase_rows(int n_rows, uint32_t granule, uint32_t cta, uint32_t grid, int& r_begin, int& r_end) {
  const uint32_t U = static_cast<uint32_t>(n_rows) / granule;  // U * grid < 2^32
  r_begin = static_cast<int>(U * cta / grid * granule);
  r_end = static_cast<int>(U * (cta + 1) / grid * granule);
}

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(DEC_CTHREADS) : "memory"); }

// ---- LL units: {32-bit payload | 32-bit tag} in one naturally aligned 64-bit word ---------------------------------------
// A 64-bit scalar store / load is single-copy atomic, so a reader that sees the expected tag also sees the payload.
__device__ __forceinline__ void ll_store(uint64_t* unit, uint32_t data, uint32_t tag, bool drop = false) {
  if (drop) return;
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(unit), "l"((static_cast<uint64_t>(tag) << 32) | data) : "memory");
}
__device__ __forceinline__ uint64_t ll_load(const uint64_t* unit) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(unit) : "memory");
  return v;
}
__device__ __forceinline__ void ll_load2(const uint64_t* unit, uint64_t& a, uint64_t& b) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(unit) : "memory");
}
// spin until `unit` carries `tag` (check == false: profiling modes whose results are garbage anyway)
__device__ __forceinline__ uint32_t ll_wait(const uint64_t* unit, uint32_t tag, bool check) {
  uint64_t v = ll_load(unit);
  uint32_t spins = 0;
  while (check && static_cast<uint32_t>(v >> 32) != tag) {
    v = ll_load(unit);
    if (++spins > EMX_SPIN_LIMIT) __trap();
  }
  return static_cast<uint32_t>(v);
}

// two units at once: both loads are in flight before the first tag is checked (one L2 round trip instead of two when both are there)
template <bool BACKOFF = false>
__device__ __forceinline__ void ll_wait2(const uint64_t* ua, const uint64_t* ub, uint32_t tag, bool check, uint32_t& a, uint32_t& b) {
  uint64_t va = ll_load(ua), vb = ll_load(ub);
  uint32_t spins = 0;
  while (check && (static_cast<uint32_t>(va >> 32) != tag || static_cast<uint32_t>(vb >> 32) != tag)) {
    if (BACKOFF) __nanosleep(96);  // long waits (tens of us): do not burn issue slots and L2 requests spinning
    if (static_cast<uint32_t>(va >> 32) != tag) va = ll_load(ua);
    if (static_cast<uint32_t>(vb >> 32) != tag) vb = ll_load(ub);
    if (++spins > EMX_SPIN_LIMIT) __trap();
  }
  a = static_cast<uint32_t>(va), b = static_cast<uint32_t>(vb);
}

// NP unit pairs at computed addresses: all loads in flight before the first tag is checked, missing pairs are re-polled.
template <int NP, typename Addr>
__device__ __forceinline__ void ll_fetch_pairs(Addr&& addr, uint32_t (&out)[2 * NP], uint32_t tag, bool check) {
  uint64_t a[NP], b[NP];
  uint32_t pending = (1u << NP) - 1;
#pragma unroll
  for (int i = 0; i < NP; ++i) ll_load2(addr(i), a[i], b[i]);
  uint32_t spins = 0;
  while (pending) {
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      if (pending & (1u << i)) {
        if (!check || (static_cast<uint32_t>(a[i] >> 32) == tag && static_cast<uint32_t>(b[i] >> 32) == tag)) pending &= ~(1u << i);
        else ll_load2(addr(i), a[i], b[i]);
      }
    }
    if (++spins > EMX_SPIN_LIMIT) __trap();
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) out[2 * i] = static_cast<uint32_t>(a[i]), out[2 * i + 1] = static_cast<uint32_t>(b[i]);
}

// Activation vector layout in shared memory: inside every block of 16 words (32 bf16) the 4 x 4 word matrix is transposed.
// The two B fragments a lane needs for TWO consecutive k-steps of mma.m16n8k16 (words 8j+t, 8j+4+t, 8j+8+t, 8j+12+t,
// t = lane%4) are then one 16-byte shared load. xs_pos maps a word index of the vector to its position.
__device__ __forceinline__ int xs_pos(int w) { return (w & ~15) + 4 * (w & 3) + 2 * ((w >> 3) & 1) + ((w >> 2) & 1); }

__device__ __forceinline__ float sumsq2(uint32_t w) {
  const float a = bf16_lo(w), c = bf16_hi(w);
  return a * a + c * c;
}

__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void prefetch_l2(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// explicit shared-space loads from a 32-bit shared address (smem_u32 of the base, taken once): a dereference of a generic pointer into
// shared memory makes the compiler rebuild the shared window address (S2R SR_CgaCtaId + LEA) in front of every load
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}


// Hand-off of per-warp partial row sums to the epilogue warp of a row group: the others arrive and run on, the epilogue warp waits. One
// named barrier per partial buffer; a warp can be at most DEC_STAGES ring stages (< DEC_PARTBUFS row groups) ahead of the epilogue warp,
// so neither a buffer nor a barrier id is reused before the epilogue is done with it.
__device__ __forceinline__ void part_arrive(uint32_t buf) { asm volatile("bar.arrive %0, %1;" ::"r"(2u + buf), "n"(DEC_CTHREADS) : "memory"); }
__device__ __forceinline__ void part_sync(uint32_t buf) { asm volatile("bar.sync %0, %1;" ::"r"(2u + buf), "n"(DEC_CTHREADS) : "memory"); }

}  // namespace emx
