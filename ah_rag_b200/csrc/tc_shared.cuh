// tc_shared.cuh -- pieces shared by the single-CTA (topk_tc.cu) and CTA-pair (topk_tc2.cu) variants of K1.
#pragma once
#include "common.cuh"

namespace ahr {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 64;
constexpr int TC_THREADS = 192;
constexpr uint32_t TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
constexpr uint32_t TC_B_BYTES = TC_BN * TC_BK * 2;  // 32 KB
constexpr uint32_t TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;

struct TcParams {
  int64_t n_rows;
  int64_t n_ctiles;  // ceil(n_rows / 256)
  int32_t nq;
  int32_t dim;
  int32_t n_qtiles;  // ceil(nq / 128)
  int32_t n_slabs;   // S
  int32_t k;         // entries kept per (query, slab), <= KLIST
  float *part_scores;    // [nq_pad][S][k]
  int32_t *part_ids;     // row index inside this index (local), -1 = empty
  float *dbg_scores;     // optional [nq][n_rows] dump of the raw accumulators (tests)
  unsigned long long *stats;  // optional [grid][8] cycle counters: where each role waits (AHR_TC_STATS=1)
};

// mbarrier wait that optionally accounts the cycles spent waiting
__device__ __forceinline__ void mbar_wait_acct(uint64_t *bar, uint32_t parity, unsigned long long *acc) {
  if (mbar_try_wait(bar, parity)) return;
  if (acc) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
    }
    *acc += (unsigned long long)(clock64() - t0);
  } else {
    while (!mbar_try_wait(bar, parity)) {
    }
  }
}

__device__ __forceinline__ void slab_range(const TcParams &p, int slab, int64_t &t0, int64_t &t1) {
  t0 = (p.n_ctiles * slab) / p.n_slabs;
  t1 = (p.n_ctiles * (slab + 1)) / p.n_slabs;
}

// ---------------------------------------------------------------------------------------------
// Per-row candidate lists of the epilogue.
//
// One list per query row (= TMEM lane = epilogue thread), KLIST 64-bit keys, UNSORTED, row-major in shared
// memory (keys[row][KLIST]); the owning thread keeps the smallest kept score `thr` and its slot in registers.
//   key = orderable(score) << 32 | ~row_id      (orderable: monotone uint32 image of the float)
// so the minimum key is the lowest score and, among equal scores, the LARGEST row id -- the entry to evict
// (lower ids survive ties).  Empty slots hold key(-inf, id -1), the smallest key of all.
// The scan of the accumulator is per thread (one max per column, one compare per 32 columns); an actual
// insertion is WARP-COOPERATIVE: the 32 lanes load the row's keys (2 each), min-reduce with shuffles and hand
// the new minimum back to the owner.  The first version inserted per thread (sorted shift, later replace-min):
// with ~4 insertions per tile per warp, each running alone on 1 of 32 lanes through generic-space loads, the
// epilogue cost 18-22 k cycles per tile against 8 k for the MMA (profiles/r1_k1_epilogue_history.md).
// select.cu sorts the flushed lists, so order inside a list is irrelevant.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_orderable(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float orderable_float(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
__device__ __forceinline__ unsigned long long make_key(float s, int32_t id) {
  return ((unsigned long long)float_orderable(s) << 32) | (unsigned long long)(~(uint32_t)id);
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

// All 32 lanes call.  Replaces the entry at `slot` of the list at shared address `list_addr` by `newkey` and
// returns the slot of the new minimum among the first k entries.
template <int KLIST>
static __device__ __noinline__ int warp_list_replace_min(uint32_t list_addr, int k, unsigned long long newkey, int slot) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) sts_u64(list_addr + slot * 8, newkey);
  __syncwarp();
  unsigned long long key[KLIST / 32];
  unsigned long long m = ~0ull;
#pragma unroll
  for (int i = 0; i < KLIST / 32; ++i) {
    key[i] = (lane + 32 * i < k) ? lds_u64(list_addr + (lane + 32 * i) * 8) : ~0ull;
    m = key[i] < m ? key[i] : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, m, o);
    m = other < m ? other : m;
  }
  int found = -1;
#pragma unroll
  for (int i = KLIST / 32 - 1; i >= 0; --i) {
    const unsigned hit = __ballot_sync(0xffffffffu, key[i] == m);
    if (hit) found = 32 * i + (__ffs(hit) - 1);
  }
  return found;
}

// Slow path of the scan, ONE copy in the binary (the unrolled scan must stay small: the first fused epilogue was
// ~9000 straight-line instructions per tile and ran 80 % stalled on instruction fetch).  Called by all 32 lanes
// when some lane saw a column above its threshold.  v[32] = this lane's scores of the chunk (local memory: the
// column to insert is only known at run time), n_valid = columns of the chunk that exist in the corpus.
// Per row, candidates are taken in ascending column order, so equal scores keep the lower row id.
template <int KLIST>
static __device__ __noinline__ void epilogue_insert_chunk(const float *v, int n_valid, bool row_valid, int k, int32_t id0,
                                                          uint32_t lists_warp, float &thr, int &slot) {
  const int lane = threadIdx.x & 31;
  unsigned bits = 0;
  if (row_valid) {
#pragma unroll 8
    for (int j = 0; j < 32; ++j) bits |= (v[j] > thr ? 1u : 0u) << j;
    if (n_valid < 32) bits &= (1u << n_valid) - 1u;  // TMA zero-filled rows past the end are not candidates
  }
  for (;;) {
    const unsigned lanes = __ballot_sync(0xffffffffu, bits != 0);
    if (!lanes) break;
    const int owner = __ffs(lanes) - 1;
    const int j_own = __ffs(bits) - 1;  // lowest pending column of this lane (meaningful in the owner)
    const int j = __shfl_sync(0xffffffffu, j_own, owner);
    const float s = __shfl_sync(0xffffffffu, v[j & 31], owner);
    const float thr_o = __shfl_sync(0xffffffffu, thr, owner);
    if (lane == owner) bits &= bits - 1;
    if (!(s > thr_o)) continue;  // the owner's threshold rose since the bit was set
    const int slot_o = __shfl_sync(0xffffffffu, slot, owner);
    const uint32_t list = lists_warp + (uint32_t)owner * (KLIST * 8);
    const int new_slot = warp_list_replace_min<KLIST>(list, k, make_key(s, id0 + j), slot_o);
    const unsigned long long new_min = lds_u64(list + new_slot * 8);  // same address in every lane: broadcast
    if (lane == owner) {
      slot = new_slot;
      thr = orderable_float((uint32_t)(new_min >> 32));
    }
  }
}

// Scan one accumulator tile (BN fp32 columns of this thread's TMEM lane) against the row's threshold.
//   t_acc      TMEM address of (this warp's lane quadrant, first column of the accumulator stage)
//   lists_warp shared address of the 32 lists of this warp (row-major, KLIST keys each)
//   release()  called once every column of the tile is in registers: hands the TMEM stage back to the MMA
// Fast path per 32 columns: one TMEM load (double buffered), 16 three-input max, one compare, one vote.
template <int KLIST, int BN, bool DBG, typename Release>
__device__ __forceinline__ void epilogue_scan_tile(uint32_t t_acc, int ncols, bool row_valid, int k, int64_t col0,
                                                   int64_t q_global, const TcParams &p, uint32_t lists_warp, float &thr,
                                                   int &slot, unsigned long long *n_slow, Release release) {
  uint32_t va[32], vb[32];
  tmem_ld_32x32b_x32(t_acc, va);
  auto scan = [&](uint32_t(&v)[32], int c) {
    if (c >= ncols) return;  // warp-uniform: ragged last corpus tile
    float m = __uint_as_float(v[0]);
#pragma unroll
    for (int j = 1; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
    if (DBG || __any_sync(0xffffffffu, row_valid && m > thr)) {
      float vl[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) vl[j] = __uint_as_float(v[j]);
      const int n_valid = min(32, ncols - c);
      if (DBG && row_valid)
        for (int j = 0; j < n_valid; ++j) p.dbg_scores[q_global * p.n_rows + col0 + c + j] = vl[j];
      if (n_slow) ++*n_slow;
      epilogue_insert_chunk<KLIST>(vl, n_valid, row_valid, k, (int32_t)(col0 + c), lists_warp, thr, slot);
    }
  };
#pragma unroll 1
  for (int c = 0; c < BN; c += 64) {  // two chunks per iteration: the register buffers alternate statically
    tmem_ld_wait();
    tmem_ld_32x32b_x32(t_acc + (uint32_t)(c + 32), vb);
    scan(va, c);
    tmem_ld_wait();
    if (c + 64 < BN) {
      tmem_ld_32x32b_x32(t_acc + (uint32_t)(c + 64), va);
    } else {
      release();  // the whole accumulator is in registers
    }
    scan(vb, c + 32);
  }
}

// 32 lists of a warp <- empty
template <int KLIST>
__device__ __forceinline__ void lists_reset(uint32_t lists_warp) {
  const int lane = threadIdx.x & 31;
  const unsigned long long empty = make_key(-INFINITY, -1);
  for (int i = lane; i < 32 * KLIST; i += 32) sts_u64(lists_warp + i * 8, empty);
  __syncwarp();
}

// 32 lists of a warp -> part[q][slab][k] (coalesced: the warp walks one row's list at a time)
template <int KLIST>
__device__ __forceinline__ void lists_flush(uint32_t lists_warp, int k, int64_t q_first, const TcParams &p, int slab) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  for (int r = 0; r < 32; ++r) {
    const int64_t q = q_first + r;
    if (q >= p.nq) break;
    const int64_t base = (q * p.n_slabs + slab) * k;
    for (int j = lane; j < k; j += 32) {
      const unsigned long long key = lds_u64(lists_warp + (uint32_t)(r * KLIST + j) * 8);
      p.part_scores[base + j] = orderable_float((uint32_t)(key >> 32));
      p.part_ids[base + j] = (int32_t)(~(uint32_t)key);
    }
  }
}

}  // namespace ahr
