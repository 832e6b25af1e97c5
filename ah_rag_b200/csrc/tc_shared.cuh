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
  int32_t n_seg;     // 1: plain bf16 operands.  3: fp32 operands split as hi + lo bf16 -- the K loop runs over
                     // (q_hi, x_hi), (q_lo, x_hi), (q_hi, x_lo), i.e. q.x up to the 2^-18-relative lo*lo term
  float *part_scores;    // [nq_pad][S][k]
  int32_t *part_ids;     // row index inside this index (local), -1 = empty
  float *dbg_scores;     // optional [nq][n_rows] dump of the raw accumulators (tests)
  unsigned long long *stats;  // optional [grid][8] cycle counters: where each role waits (AHR_TC_STATS=1)
  int32_t tile_stride;   // corpus tile i of this launch is tile i * tile_stride of the index (sample pass: > 1)
  int32_t thr_stride;    // thr_src[q * thr_stride] = k-th best score of query q over a SAMPLE of the corpus rows
  const float *thr_src;  // optional: seeds every list's admission threshold (see seed_threshold)
  uint32_t *thr_shared;  // optional [n_slabs][q_rows]: what every slab has published for every query (see "shared bound")
  int32_t q_rows;        // row pitch of thr_shared (queries padded to whole tiles)
  int32_t bound_m;       // each slab publishes the bound_m-th best score of its list: n_slabs * bound_m >= k
  int32_t list_filter;   // 1 (default): a full list is first filtered against its risen threshold; AHR_TC_LIST_FILTER=0: A/B knob
  int32_t qtile_major;   // item order of the pair kernel: 0 = slab-major, 1 = query-tile-major (see tc_item_decode)
  int32_t partial_m;     // > 0: every sorted list also publishes its partial_m-th best at thr_shared + n_slabs * q_rows
                         // ("partial bound", see shared_bound_peek); 0 = off
};

// item -> (slab, query tile).  Slab-major (item = slab * n_qtiles + qtile) runs the query tiles of one slab side by side:
// the corpus tile is read once from HBM for all of them.  Query-tile-major (item = qtile * n_slabs + slab) runs the S slabs
// of one query tile side by side: their lists publish to the shared bound TOGETHER -- with slab-major order and more items
// than CTA pairs (c3: 144 items on 74 pairs) a query's slabs are spread over two rounds, its bound stays at "not all slabs
// have published" for the whole first round, and every list of that round learns its threshold alone (~500 appends and
// ~35 sorts per row).  Measured on c3 (profiles/r2_k1_q256_wait_counters.md): 2.6 % fewer cycles (MMA wait for TMEM 3.9 M ->
// 2.4 M, slow chunks 8324 -> 5010) but 3 % MORE time -- fewer query tiles share a slab position, the corpus is read twice,
// and under the power cap the extra HBM traffic costs more clock than the saved cycles are worth.  Kept as an experiment
// knob (AHR_TC_ITEM_ORDER=1), off by default.
__device__ __forceinline__ void tc_item_decode(const TcParams &p, int item, int &slab, int &qt) {
  if (p.qtile_major) {
    qt = item / p.n_slabs;
    slab = item - qt * p.n_slabs;
  } else {
    slab = item / p.n_qtiles;
    qt = item - slab * p.n_qtiles;
  }
}

inline int tc_list_filter_default() {
  static const int v = getenv("AHR_TC_LIST_FILTER") ? atoi(getenv("AHR_TC_LIST_FILTER")) : 1;
  return v;
}

// Optional knobs of a K1 launch (host side).
struct TcExtra {
  int32_t q_rows = 0;        // rows addressable behind dev_q (>= nq; padded so TMA never fills out-of-bounds rows)
  int32_t tile_stride = 1;
  const float *thr_src = nullptr;
  int32_t thr_stride = 0;
  uint32_t *thr_shared = nullptr;  // zero-initialised by the caller before the launch
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
// One list per query row (= TMEM lane = epilogue thread): KLIST 64-bit keys, row-major in shared memory
// (keys[row][KLIST]); the owning thread keeps in registers the admission threshold `thr` and the fill count `cnt`.
//   key = orderable(score) << 32 | ~row_id      (orderable: monotone uint32 image of the float)
// so a larger key is a better candidate and, among equal scores, the LOWER row id is the larger key.
//
//   append      per thread, O(1): a column with score > thr is stored at keys[row][cnt++].  Every lane appends to
//               its own row in the same instruction -- no serialisation, however many rows have candidates.
//   compaction  when a row reaches KLIST entries the warp sorts that row (bitonic network over 2 keys per lane,
//               shuffles only), keeps the best k, and the owner's threshold becomes the k-th score.  Amortised over
//               KLIST - k appends.  The threshold is therefore slightly stale between compactions (admits a few
//               extra candidates, never drops one).
//
// History (profiles/r1_k1_epilogue_history.md): sorted per-thread insertion -> per-thread replace-min -> warp-
// cooperative replace-min per candidate (each ~65 instructions, serialised over rows: the per-item fill phase cost
// 0.8x the MMA time on 1.25M-row shards) -> this scheme.  select.cu orders the flushed lists.
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

// Sample-seeded admission threshold.  thr_src holds, per query, the k-th best score over a subset of the corpus rows
// (a strided sample of tiles scored by this same kernel).  The global k-th best score is >= that value, so a row
// scoring below it can never reach the global top-k: every list starts at the largest float below it instead of -inf
// (the sample rows themselves are scanned again by the main pass and are admitted, scores being bit-identical).
// Without it every (slab, query) list re-learns its threshold from scratch: ~k ln(rows per slab / k) appends and a
// compaction every KLIST - k of them, front-loaded on every item.
__device__ __forceinline__ float seed_threshold(const TcParams &p, int64_t q, bool row_valid) {
  if (p.thr_src == nullptr || !row_valid) return -INFINITY;
  const float s = p.thr_src[q * p.thr_stride];
  if (!(s > -INFINITY)) return -INFINITY;       // fewer than k sample rows (or NaN): no bound
  if (s == 0.0f) return -1.401298464e-45f;      // +-0 compare equal: step below both
  return orderable_float(float_orderable(s) - 1u);
}

// Shared admission bound.  Every (slab, query) list learns its own threshold from scratch: ~k ln(rows per slab / k) appends
// per row, front-loaded, and a chunk is "slow" as soon as ONE of the warp's 32 rows has a candidate.  With many short slabs
// (a 256-query batch over a 12.5M-row shard runs 74 slabs of 169k rows) 57 % of all chunks were slow and the MMA issuer
// spent >= 24 % of the kernel waiting for the epilogue to hand TMEM back (profiles/r2_k1_q256_wait_counters_before.txt).
//
// Slabs are disjoint row ranges.  If every slab s holds at least m rows scoring >= b_s for a query, the corpus holds at
// least S * m >= k rows scoring >= B = min_s b_s, so the query's global k-th best is >= B and a row scoring strictly below
// B can be dropped by EVERY slab.  Each list publishes its m-th best score, m = ceil(k / S), whenever it is compacted
// (thr_shared[slab][q], orderable uint image of the float, 0 = nothing yet, single writer, monotone); the epilogue threads
// re-read the S values of their query every few tiles and raise their threshold to just below the minimum -- rows scoring
// EXACTLY B are kept (they may win the id tie-break), hence float_below().  Until all S slabs have published the minimum
// is 0 = no bound.  Relaxed accesses: a stale value is a valid, merely weaker, bound, so the result -- the global top-k
// under the total order -- does not depend on timing.  With S = 74, k = 50 this is the minimum of the slabs' running
// maxima, about the 74th best score seen so far over the WHOLE corpus, against the 50th best of one slab before.
__device__ __forceinline__ float float_below(float s) {
  if (s == 0.0f) return -1.401298464e-45f;  // +-0 compare equal: step below both
  return orderable_float(float_orderable(s) - 1u);
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// minimum over the slabs of what they published for query q (0 as soon as one slab has not published).
//
// Partial bound.  With more items than CTA pairs (c3: 9 slabs x 16 query tiles on 74 pairs) the slabs of a query run in
// several rounds, and until the last round starts "every slab has published" is false: the lists of the earlier rounds
// learn their thresholds alone (~500 appends, ~35 sorts per row).  The argument behind the bound does not need every slab:
// if P published slabs each hold >= m' rows scoring >= b'_s and P * m' >= k, the corpus holds >= k rows scoring >=
// min b'_s.  So a sorted list also publishes its m'-th best, m' = ceil(k / slabs that run entirely in the first round)
// (c3: 4 slabs, m' = 13), in a second array, and while the full bound is missing the minimum over the slabs that HAVE
// published counts as soon as there are enough of them.  Weaker than the full bound (13th instead of 6th best), which takes
// over once it exists.
__device__ __forceinline__ uint32_t shared_bound_peek(const TcParams &p, int64_t q, bool active) {
  if (!active) return 0;
  uint32_t b0 = 0xFFFFFFFFu, b1 = 0xFFFFFFFFu;
  const uint32_t *src = p.thr_shared + q;
  int s = 0;
  for (; s + 1 < p.n_slabs; s += 2) {  // two independent chains; the loads pipeline in L2
    b0 = min(b0, ld_relaxed_u32(src + (int64_t)s * p.q_rows));
    b1 = min(b1, ld_relaxed_u32(src + (int64_t)(s + 1) * p.q_rows));
  }
  if (s < p.n_slabs) b0 = min(b0, ld_relaxed_u32(src + (int64_t)s * p.q_rows));
  const uint32_t full = min(b0, b1);
  if (full != 0 || p.partial_m <= 0) return full;
  src += (int64_t)p.n_slabs * p.q_rows;
  uint32_t b = 0xFFFFFFFFu;
  int have = 0;
  for (s = 0; s < p.n_slabs; ++s) {
    const uint32_t v = ld_relaxed_u32(src + (int64_t)s * p.q_rows);
    if (v != 0) { b = min(b, v); ++have; }
  }
  return have * p.partial_m >= p.k ? b : 0;
}
__device__ __forceinline__ float shared_bound_apply(uint32_t bits, float thr) {
  // 0x007FFFFF is the image of -inf: anything at or below it carries no information
  return bits > 0x00800000u ? fmaxf(thr, float_below(orderable_float(bits))) : thr;
}

// All 32 lanes call.  Sorts the first `cnt` entries of the list at shared address `list_addr` (best first), pads the
// rest with the empty key, writes all KLIST slots back.  Returns the score of the k-th entry (-inf if cnt < k).
template <int KLIST>
static __device__ __noinline__ float warp_list_compact(uint32_t list_addr, int cnt, int k) {
  constexpr int NPL = KLIST / 32;  // keys per lane; element index of register r in lane l is r * 32 + l
  const int lane = threadIdx.x & 31;
  const unsigned long long empty = make_key(-INFINITY, -1);
  unsigned long long key[NPL];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NPL; ++r) key[r] = (r * 32 + lane < cnt) ? lds_u64(list_addr + (r * 32 + lane) * 8) : empty;
#pragma unroll
  for (int size = 2; size <= KLIST; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {  // partner lives in another register of the same lane
        const int rs = stride / 32;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          if ((r & rs) == 0) {
            const bool desc = (((r * 32 + lane) & size) == 0);  // this block sorts best-first
            const unsigned long long lo = key[r], hi = key[r ^ rs];
            const bool swap = desc ? (lo < hi) : (lo > hi);
            key[r] = swap ? hi : lo;
            key[r ^ rs] = swap ? lo : hi;
          }
        }
      } else {  // partner lives in lane ^ stride, same register
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[r], stride);
          const bool lower = (lane & stride) == 0;
          const bool desc = (((r * 32 + lane) & size) == 0);
          const bool keep_max = lower == desc;  // in a best-first block the lower index keeps the larger key
          const unsigned long long mx = key[r] > other ? key[r] : other;
          const unsigned long long mn = key[r] > other ? other : key[r];
          key[r] = keep_max ? mx : mn;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NPL; ++r) sts_u64(list_addr + (r * 32 + lane) * 8, key[r]);
  unsigned long long kth = empty;
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const unsigned long long c = __shfl_sync(0xffffffffu, key[r], (k - 1) & 31);
    if (((k - 1) >> 5) == r) kth = c;
  }
  __syncwarp();
  return cnt >= k ? orderable_float((uint32_t)(kth >> 32)) : -INFINITY;
}

// All 32 lanes call.  Drops from a FULL list every entry that no longer beats `thr` (the row's threshold has risen since
// the entry was appended: a compaction, or the shared bound that arrives while an item's first tile is being scanned)
// and closes the gaps, order preserved.  ~25 instructions against ~400 for the sort: when it frees enough slots the sort
// is skipped.  Returns the number of entries left.
template <int KLIST>
static __device__ __noinline__ int warp_list_filter(uint32_t list_addr, float thr) {
  constexpr int NPL = KLIST / 32;
  const int lane = threadIdx.x & 31;
  unsigned long long key[NPL];
  __syncwarp();
#pragma unroll
  for (int r = 0; r < NPL; ++r) key[r] = lds_u64(list_addr + (r * 32 + lane) * 8);
  __syncwarp();  // every slot is in registers before any is overwritten
  int base = 0;
#pragma unroll
  for (int r = 0; r < NPL; ++r) {
    const bool keep = orderable_float((uint32_t)(key[r] >> 32)) > thr;  // the admission rule of the append path
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) sts_u64(list_addr + (base + __popc(m & ((1u << lane) - 1u))) * 8, key[r]);
    base += __popc(m);
  }
  __syncwarp();
  return base;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Rare path, ONE copy in the binary: some lane's list is full (and may still have `pending` candidate columns of
// the current chunk).  The warp compacts the full rows, their owners re-filter the pending columns against the
// refreshed threshold and store what still qualifies; repeats until no list is full.
// scratch_lane: shared address of this lane's column of the warp's chunk scratch (score of column j at + j*128).
template <int KLIST>
static __device__ __noinline__ void epilogue_overflow(uint32_t scratch_lane, unsigned pending, int k, int32_t id0,
                                                      uint32_t lists_warp, float &thr_ref, int &cnt_ref,
                                                      uint32_t *bound_slot, int bound_m, bool filter,
                                                      int64_t partial_off, int partial_m) {
  const int lane = threadIdx.x & 31;
  const uint32_t my_list = lists_warp + (uint32_t)lane * (KLIST * 8);
  float thr = thr_ref;
  int cnt = cnt_ref;
  for (;;) {
    while (pending != 0 && cnt < KLIST) {  // per lane, ascending column order: store what fits
      const int j = __ffs(pending) - 1;
      pending &= pending - 1;
      const float s = lds_f32(scratch_lane + j * 128);
      if (s > thr) {  // re-check: a compaction may have raised this row's threshold
        sts_u64(my_list + cnt * 8, make_key(s, id0 + j));
        ++cnt;
      }
    }
    unsigned full = __ballot_sync(0xffffffffu, cnt == KLIST);
    if (!full) break;  // no full list left, hence nothing pending either
    while (full) {
      const int owner = __ffs(full) - 1;
      full &= full - 1;
      // entries appended under an older, weaker threshold first: often most of the list (first tile of an item, once the
      // other slabs have published), and then there is nothing to sort
      const int left = filter ? warp_list_filter<KLIST>(lists_warp + (uint32_t)owner * (KLIST * 8), __shfl_sync(0xffffffffu, thr, owner)) : KLIST;
      if (left <= KLIST - KLIST / 8) {
        if (lane == owner) cnt = left;
        continue;
      }
      const float kth = warp_list_compact<KLIST>(lists_warp + (uint32_t)owner * (KLIST * 8), left, k);
      if (lane == owner) {
        cnt = min(left, k);
        thr = fmaxf(thr, kth);  // never below a bound taken from thr_shared
        // the list is sorted now: its bound_m-th best score goes to the other slabs (single writer, only ever grows)
        if (bound_slot != nullptr) {
          st_relaxed_u32(bound_slot, (uint32_t)(lds_u64(my_list + (bound_m - 1) * 8) >> 32));
          // (partial bound: cnt >= partial_m here -- the sort ran on more than KLIST - KLIST / 8 >= k >= partial_m entries)
          if (partial_m > 0 && cnt >= partial_m)
            st_relaxed_u32(bound_slot + partial_off, (uint32_t)(lds_u64(my_list + (partial_m - 1) * 8) >> 32));
        }
      }
    }
  }
  thr_ref = thr;
  cnt_ref = cnt;
}

// Scan one accumulator tile (BN fp32 columns of this thread's TMEM lane) against the row's threshold.
//   t_acc        TMEM address of (this warp's lane quadrant, first column of the accumulator stage)
//   lists_warp   shared address of the 32 lists of this warp (row-major, KLIST keys each)
//   scratch_warp shared address of this warp's 4 KB chunk scratch ([32 columns][32 lanes] fp32)
//   release()    called once every column of the tile is in registers: hands the TMEM stage back to the MMA
// Per 32 columns: one TMEM load (double buffered), a 4-way max tree, one compare.  A thread whose maximum beats its
// threshold builds the bitmask of its candidate columns from registers, parks the chunk in the scratch (the column
// to store is only known at run time; local memory would go to L2 in this kernel, shared memory is 30 cycles) and
// appends its candidates itself -- no warp cooperation, no function call.  One vote per chunk detects full lists,
// which take the out-of-line path above.
template <int KLIST, int BN, bool DBG, typename Release>
__device__ __forceinline__ void epilogue_scan_tile(uint32_t t_acc, int ncols, bool row_valid, int k, int64_t col0,
                                                   int64_t q_global, const TcParams &p, uint32_t lists_warp,
                                                   uint32_t scratch_warp, float &thr, int &cnt,
                                                   unsigned long long *n_slow, uint32_t *bound_slot, int bound_m,
                                                   float &best, bool young, Release release) {
  const int lane = threadIdx.x & 31;
  const uint32_t my_list = lists_warp + (uint32_t)lane * (KLIST * 8);
  const uint32_t scratch_lane = scratch_warp + (uint32_t)lane * 4;
  uint32_t va[32], vb[32];
  tmem_ld_32x32b_x32(t_acc, va);
  auto scan = [&](uint32_t(&v)[32], int c) {
    if (c >= ncols) return;  // warp-uniform: ragged last corpus tile
    // four independent max chains (a single warp per scheduler is latency-bound on dependent chains)
    float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(v[2]), m3 = __uint_as_float(v[3]);
#pragma unroll
    for (int j = 4; j < 32; j += 4) {
      m0 = fmaxf(m0, __uint_as_float(v[j]));
      m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
      m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
      m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
    }
    const float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    unsigned pending = 0;
    const bool hit = row_valid && m > thr;
    if (DBG || __any_sync(0xffffffffu, hit)) {
      if (n_slow) ++n_slow[0];
      // park the chunk: conflict-free (lane = bank), static register indices
#pragma unroll
      for (int j = 0; j < 32; ++j) sts_f32(scratch_lane + j * 128, __uint_as_float(v[j]));
      const int n_valid = min(32, ncols - c);  // TMA zero-filled rows past the end are not candidates
      if (DBG && row_valid)
        for (int j = 0; j < n_valid; ++j) p.dbg_scores[q_global * p.n_rows + col0 + c + j] = lds_f32(scratch_lane + j * 128);
      if (hit) {  // divergent: only the threads with a candidate in this chunk
        unsigned bits = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) bits |= (__uint_as_float(v[j]) > thr ? 1u : 0u) << j;
        if (n_valid < 32) bits &= (1u << n_valid) - 1u;
        const int32_t id0 = (int32_t)(col0 + c);
        float cmax = -INFINITY;
        while (bits != 0 && cnt < KLIST) {  // ascending column order: equal scores keep the lower row id
          const int j = __ffs(bits) - 1;
          bits &= bits - 1;
          const float sc = lds_f32(scratch_lane + j * 128);
          AHR_DEV_ASSERT(cnt >= 0 && cnt < KLIST && j >= 0 && j < 32);
          sts_u64(my_list + cnt * 8, make_key(sc, id0 + j));
          cmax = fmaxf(cmax, sc);
          ++cnt;
        }
        pending = bits;
        // bound_m == 1: the list's best score IS what this slab publishes -- no need to wait for the next compaction
        // (lists that rarely fill up would otherwise sit on a stale value)
        if (bound_m == 1 && bound_slot != nullptr && cmax > best) {
          best = cmax;
          st_relaxed_u32(bound_slot, float_orderable(cmax));
        }
      }
      // a full list is compacted right away, even with nothing pending: its threshold is stale.  In the first tile of an
      // item the lists fill before the first look at the shared bound (64 columns, all admitted): look now, so that the
      // overflow path can drop what the bound excludes instead of sorting it
      if (__any_sync(0xffffffffu, cnt == KLIST)) {
        // (only where a list publishes its running maximum, bound_m == 1: with bound_m > 1 the value to publish is read
        // from the SORTED list, and skipping sorts starved the other slabs of bounds -- Q = 512 / 1024 ran 3 % slower)
        const bool filter = p.list_filter != 0 && (bound_m == 1 || p.list_filter == 2);  // 2: experiment knob, every list
        if (young && filter) thr = shared_bound_apply(shared_bound_peek(p, q_global, bound_slot != nullptr), thr);
        epilogue_overflow<KLIST>(scratch_lane, pending, k, (int32_t)(col0 + c), lists_warp, thr, cnt, bound_slot, bound_m, filter,
                                 (int64_t)p.n_slabs * p.q_rows, p.partial_m);
      }
    }
  };
#pragma unroll 1
  for (int c = 0; c < BN; c += 64) {  // two chunks per iteration: the register buffers alternate statically
    // first tile of an item: no bound exists when it starts (all slabs start together), and without one its 256 columns
    // cost ~130 appends and ~6 compactions per row -- most of what is left of the MMA's wait for TMEM.  The other slabs
    // publish within their first chunk, so look again every 64 columns.
    if (young && c > 0) thr = shared_bound_apply(shared_bound_peek(p, q_global, bound_slot != nullptr), thr);
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

// Final compaction of the warp's 32 lists and flush to part[q][slab][k] (sorted best first, padded with empties;
// coalesced: the warp walks one row's list at a time).
template <int KLIST>
__device__ __forceinline__ void lists_flush(uint32_t lists_warp, int k, int cnt, int64_t q_first, const TcParams &p,
                                            int slab) {
  const int lane = threadIdx.x & 31;
  for (int r = 0; r < 32; ++r) {
    const int64_t q = q_first + r;
    if (q >= p.nq) break;  // warp-uniform
    const int cnt_r = __shfl_sync(0xffffffffu, cnt, r);
    const uint32_t list = lists_warp + (uint32_t)r * (KLIST * 8);
    warp_list_compact<KLIST>(list, cnt_r, k);
    const int64_t base = (q * p.n_slabs + slab) * k;
    for (int j = lane; j < k; j += 32) {
      const unsigned long long key = lds_u64(list + j * 8);
      p.part_scores[base + j] = orderable_float((uint32_t)(key >> 32));
      p.part_ids[base + j] = (int32_t)(~(uint32_t)key);
    }
  }
  __syncwarp();
}

}  // namespace ahr
