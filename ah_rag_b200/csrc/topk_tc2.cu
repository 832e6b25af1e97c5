// topk_tc2.cu -- K1, CTA-pair variant: the fused similarity GEMM + top-k epilogue on tcgen05.mma.cta_group::2.
//
// Why pairs: with one CTA per tile every SM has to pull A (16 KB) + B (32 KB) per 64-wide k-block from L2 =
// 96 B/clk/SM at full MMA rate, and only 3 x 48 KB stages fit beside the 64 KB of per-row lists -- the ncu
// capture of the single-CTA kernel (profiles/r1_k1_v1_1cta_ncu.md) shows a latency-bound pipeline at 16-35 %
// tensor-pipe activity.  A CTA pair computes one 256 x 256 tile: each CTA stages its own 128 queries (A,
// 16 KB) and HALF of the 256 corpus rows (B, 16 KB); the tensor core reads both halves of B from the two
// SMs.  32 KB per stage -> 4 stages in flight (128 KB, leaving room for L1) and 64 B/clk/SM of L2 traffic.
//
//   pair (cluster of 2 CTAs) : item = (slab, 256-query tile); item -> cluster round-robin, slab-major
//   CTA r of the pair        : query rows  qt*256 + r*128 .. +128   (its TMEM lanes / its epilogue / its lists)
//                              corpus rows ct*256 + r*128 .. +128   (its half of every B tile)
//   leader CTA (rank 0)      : issues every tcgen05.mma; its full barrier collects the TMA bytes of BOTH CTAs;
//                              tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs;
//                              the epilogue warps of both CTAs release the accumulator on the leader's barrier.
//   operand rings            : A (queries, re-read from L2 for every corpus tile) and B (corpus, streamed from HBM)
//                              have SEPARATE rings: NA slots of A, NB > NA slots of B.  One producer thread issues
//                              B(t) and then A(t - (NB - NA)): both wait for the same event (the MMA finished k-block
//                              t - NB), so nothing blocks, but the HBM stream runs NB k-blocks ahead while the L2
//                              stream only needs NA.  Measured (profiles/r1_k1_epilogue_history.md, "rings"):
//                              NA = 3 starves the tensor-bound workload (L2 latency under load > 3 k-blocks) and a
//                              deeper corpus ring does not help the HBM-bound one, so the default is NA = NB = 4;
//                              what did help both was taking the integer divisions out of the producer's loop.
//                              L2 prefetch (cp.async.bulk.prefetch.tensor) was tried as well and is slower.
#include "common.cuh"
#include "kernels.h"
#include "tc_shared.cuh"

namespace ahr {

constexpr int T2_BM = 128;                      // query rows per CTA
constexpr int T2_BN = 256;                      // corpus rows per tile (pair)
constexpr int T2_BN_HALF = 128;                 // corpus rows staged by each CTA
constexpr uint32_t T2_A_BYTES = T2_BM * TC_BK * 2;        // 16 KB
constexpr uint32_t T2_B_BYTES = T2_BN_HALF * TC_BK * 2;   // 16 KB

template <int KLIST, int NA, int NB, bool DBG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1)
topk_tc2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                const __grid_constant__ CUtensorMap map_q_lo, const __grid_constant__ CUtensorMap map_x_lo, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  static_assert(NA >= 1 && NB >= 1, "ring depths");
  constexpr bool B_LEADS = NB >= NA;
  constexpr int ND = B_LEADS ? NB : NA;         // slots of the `done` ring = depth of the deeper operand ring
  constexpr int LAG = B_LEADS ? NB - NA : NA - NB;  // the producer issues block t - LAG of the shallower ring right after block t
  uint8_t *a_tiles = smem;                                // [NA][16 KB]
  uint8_t *b_tiles = smem + NA * T2_A_BYTES;              // [NB][16 KB]
  unsigned long long *lists = reinterpret_cast<unsigned long long *>(b_tiles + NB * T2_B_BYTES);  // [128][KLIST] keys
  float *scratch = reinterpret_cast<float *>(lists + KLIST * T2_BM);  // [4 warps][32 columns][32 lanes] chunk scratch
  uint64_t *bars = reinterpret_cast<uint64_t *>(scratch + 4 * 32 * 32);
  uint64_t *full_a = bars;               // [NA] leader's copy is the live one (bytes of both CTAs)
  uint64_t *full_b = full_a + NA;        // [NB]
  uint64_t *done = full_b + NB;          // [ND] per CTA: the MMA finished k-block t (slot t mod ND): frees B slot t mod NB
                                         //      and A slot t mod NA -- ONE commit per k-block (a second one per
                                         //      k-block cost 30 % on the tensor-bound workload)
  uint64_t *tfull = done + ND;           // [2] per CTA: accumulator stage complete
  uint64_t *tempty = tfull + 2;          // [2] leader's copy: both epilogues drained the stage
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int n_items = p.n_slabs * p.n_qtiles;  // n_qtiles counts 256-query tiles here
  const int kb_per_seg = (p.dim + TC_BK - 1) / TC_BK;
  const int n_kblocks = kb_per_seg * p.n_seg;  // k-blocks per tile over all operand segments

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_x);
    prefetch_tensormap(&map_q_lo);
    prefetch_tensormap(&map_x_lo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NA; ++s) mbar_init(&full_a[s], 1);  // the leader's arrive.expect_tx; TMA bytes of both CTAs
      for (int s = 0; s < NB; ++s) mbar_init(&full_b[s], 1);
      for (int s = 0; s < ND; ++s) mbar_init(&done[s], 1);  // one multicast commit
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);   // one multicast commit
        mbar_init(&tempty[a], 8);  // 4 epilogue warps x 2 CTAs
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc_2sm<512>(tmem_slot);  // same warp id, same smem offset in both CTAs
  }
  tc_fence_before();
  cluster_sync_all();  // barriers initialised and TMEM allocated in both CTAs before any remote access
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      unsigned long long w_empty = 0;
      unsigned long long *acct = p.stats ? &w_empty : nullptr;
      // a cursor walks this pair's (item, corpus tile, segment, k-block) sequence; the A cursor trails the B cursor by
      // LAG.  No division on the per-k-block path: this single thread has to issue a block every ~500 cycles and an
      // integer division is ~100 cycles of dependent latency (three of them cost 26 % on the tensor-bound workload).
      struct Cursor {
        int item, seg, kb;
        int32_t q_row0;
        int64_t ct, t1;
        bool done;
      };
      auto open_item = [&](Cursor &c) {
        c.done = c.item >= n_items;
        if (c.done) return;
        int slab, qt;
        tc_item_decode(p, c.item, slab, qt);
        int64_t t0;
        slab_range(p, slab, t0, c.t1);
        c.ct = t0;
        c.seg = 0;
        c.kb = 0;
        c.q_row0 = qt * (2 * T2_BM) + (int32_t)rank * T2_BM;
      };
      auto advance = [&](Cursor &c) {
        if (++c.kb < kb_per_seg) return;
        c.kb = 0;
        if (++c.seg < p.n_seg) return;
        c.seg = 0;
        if (++c.ct < c.t1) return;
        c.item += n_clusters;
        open_item(c);
      };
      Cursor ca{cluster_id, 0, 0, 0, 0, 0, false}, cb{cluster_id, 0, 0, 0, 0, 0, false};
      open_item(ca);
      open_item(cb);
      int sa = 0, sb = 0, sd = 0;
      uint32_t pd = 0;
      const uint32_t full_a0 = mapa_u32(smem_u32(&full_a[0]), 0), full_b0 = mapa_u32(smem_u32(&full_b[0]), 0);  // leader's
      auto issue_a = [&]() {
        const CUtensorMap *mq = ca.seg == 1 ? &map_q_lo : &map_q;  // (q_hi,x_hi) (q_lo,x_hi) (q_hi,x_lo)
        if (leader) mbar_arrive_expect_tx(&full_a[sa], 2 * T2_A_BYTES);
        // queries are re-read by every corpus tile: keep them in L2
        tma_load_2d_2sm(a_tiles + sa * T2_A_BYTES, mq, full_a0 + sa * 8, ca.kb * TC_BK, ca.q_row0, kEvictLast);
        if (++sa == NA) sa = 0;
        advance(ca);
      };
      auto issue_b = [&]() {
        const CUtensorMap *mx = cb.seg == 2 ? &map_x_lo : &map_x;
        const int32_t x_row0 = (int32_t)cb.ct * (p.tile_stride * T2_BN) + (int32_t)rank * T2_BN_HALF;
        if (leader) mbar_arrive_expect_tx(&full_b[sb], 2 * T2_B_BYTES);
        // normal L2 policy: the pairs sweeping the same slab for other query tiles hit these lines (evict-last / evict-first
        // measured: no change in time or DRAM traffic, profiles/r2_k1_traffic.md)
        tma_load_2d_2sm(b_tiles + sb * T2_B_BYTES, mx, full_b0 + sb * 8, cb.kb * TC_BK, x_row0, kEvictNormal);
        if (++sb == NB) sb = 0;
        advance(cb);
      };
      // The deeper ring leads.  Its block t waits for "k-block t - ND finished" (done[t mod ND]); the block t - LAG of
      // the other ring needs "k-block t - LAG - (ND - LAG) finished" -- the same event, so it is issued right behind
      // without a wait of its own.
      Cursor &cl = B_LEADS ? cb : ca, &ct_ = B_LEADS ? ca : cb;
      int lead = 0;
      while (!cl.done) {
        mbar_wait_acct(&done[sd], pd ^ 1, acct);
        if (++sd == ND) { sd = 0; pd ^= 1; }
        if (B_LEADS) issue_b(); else issue_a();
        if (lead < LAG) ++lead;
        else if (B_LEADS) issue_a();
        else issue_b();
      }
      while (!ct_.done) {  // the last LAG blocks of the trailing ring: keep walking the `done` ring
        mbar_wait_acct(&done[sd], pd ^ 1, acct);
        if (++sd == ND) { sd = 0; pd ^= 1; }
        if (B_LEADS) issue_a(); else issue_b();
      }
      if (p.stats) p.stats[blockIdx.x * 8 + 0] = w_empty;
    }
    __syncwarp();  // reconverge before the block/cluster-wide barrier below
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * T2_BM, T2_BN);  // M = 256 across the pair, N = 256
      int sa = 0, sb = 0, sd = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      unsigned long long w_full = 0, w_tempty = 0;
      const long long t_begin = clock64();
      for (int item = cluster_id; item < n_items; item += n_clusters) {
        int slab, qt_unused;
        tc_item_decode(p, item, slab, qt_unused);
        int64_t t0, t1;
        slab_range(p, slab, t0, t1);
        for (int64_t ct = t0; ct < t1; ++ct) {
          mbar_wait_acct(&tempty[acc], acc_phase ^ 1, p.stats ? &w_tempty : nullptr);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * T2_BN;
          for (int kb = 0; kb < n_kblocks; ++kb) {
            mbar_wait_acct(&full_b[sb], pb, p.stats ? &w_full : nullptr);
            mbar_wait_acct(&full_a[sa], pa, p.stats ? &w_full : nullptr);
            tc_fence_after();
            const uint64_t a_desc = make_smem_desc_sw128(smem_u32(a_tiles + sa * T2_A_BYTES));
            const uint64_t b_desc = make_smem_desc_sw128(smem_u32(b_tiles + sb * T2_B_BYTES));
#pragma unroll
            for (int k4 = 0; k4 < TC_BK / 16; ++k4)
              umma_bf16_ss_2sm(d_tmem, a_desc + (uint64_t)(k4 * 2), b_desc + (uint64_t)(k4 * 2), idesc,
                               (kb | k4) != 0 ? 1u : 0u);
            umma_commit_2sm(&done[sd], 0x3);  // frees A slot sa and B slot sb in both CTAs
            if (++sd == ND) sd = 0;
            if (++sa == NA) { sa = 0; pa ^= 1; }
            if (++sb == NB) { sb = 0; pb ^= 1; }
          }
          umma_commit_2sm(&tfull[acc], 0x3);  // accumulator ready in both CTAs
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      if (p.stats) {
        p.stats[blockIdx.x * 8 + 1] = w_full;
        p.stats[blockIdx.x * 8 + 2] = w_tempty;
        p.stats[blockIdx.x * 8 + 3] = (unsigned long long)(clock64() - t_begin);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (both CTAs): TMEM -> registers -> per-row top-k =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access; thread = query row = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t lists_warp = smem_u32(lists) + (uint32_t)(quad * 32) * (KLIST * 8);
    const uint32_t scratch_warp = smem_u32(scratch) + (uint32_t)quad * (32 * 32 * 4);
    const int k = p.k;
    int acc = 0;
    uint32_t acc_phase = 0;
    unsigned long long w_tfull = 0, n_slow[4] = {0, 0, 0, 0};
    unsigned long long *acct = p.stats ? &w_tfull : nullptr;  // warp-uniform (tcgen05.wait::ld is .aligned); thread 64 reports
    for (int item = cluster_id; item < n_items; item += n_clusters) {
      int slab, qt;
      tc_item_decode(p, item, slab, qt);
      int64_t t0, t1;
      slab_range(p, slab, t0, t1);
      const int64_t q_first = (int64_t)qt * (2 * T2_BM) + (int64_t)rank * T2_BM + quad * 32;
      const int64_t q_global = q_first + lane;
      const bool row_valid = q_global < p.nq;
      float thr = seed_threshold(p, q_global, row_valid);  // admission threshold of this thread's row
      int cnt = 0;            // entries in this thread's list
      float best = -INFINITY;  // best score appended so far (published when each slab publishes its top-1)
      const bool bound_on = p.thr_shared != nullptr && row_valid;
      uint32_t *bound_slot = bound_on ? p.thr_shared + (int64_t)slab * p.q_rows + q_global : nullptr;
      for (int64_t ct = t0; ct < t1; ++ct) {
        // every tile while the lists are young, every 8th afterwards; issued before the wait: the L2 latency hides behind it
        const int64_t age = ct - t0;
        const uint32_t bound_bits = shared_bound_peek(p, q_global, bound_on && (age < 8 || (age & 7) == 0));
        mbar_wait_acct(&tfull[acc], acc_phase, acct);
        tc_fence_after();
        thr = shared_bound_apply(bound_bits, thr);
        const int64_t col0 = ct * p.tile_stride * T2_BN;
        const int ncols = (int)min((int64_t)T2_BN, p.n_rows - col0);
        const uint32_t tempty_leader = mapa_u32(smem_u32(&tempty[acc]), 0);
        epilogue_scan_tile<KLIST, T2_BN, DBG>(t_lane + (uint32_t)(acc * T2_BN), ncols, row_valid, k, col0, q_global, p,
                                              lists_warp, scratch_warp, thr, cnt, p.stats ? n_slow : nullptr, bound_slot, p.bound_m, best, age == 0, [&]() {  // uniform: every lane counts
                                                tc_fence_before();
                                                __syncwarp();
                                                if (lane == 0) mbar_arrive_cluster(tempty_leader);  // leader's barrier
                                              });
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      lists_flush<KLIST>(lists_warp, k, cnt, q_first, p, slab);
    }
    if (acct && threadIdx.x == 64) {
      p.stats[blockIdx.x * 8 + 4] = w_tfull;
      p.stats[blockIdx.x * 8 + 5] = n_slow[0];
      p.stats[blockIdx.x * 8 + 6] = n_slow[2];  // cycles inside the slow path
      p.stats[blockIdx.x * 8 + 7] = n_slow[3];  // row compactions
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be reading our smem / signalling our barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int make_bf16_tensor_map(CUtensorMap *map, const void *dev_ptr, int64_t rows, int32_t dim, int32_t box_rows);

int tc2_make_corpus_map(CUtensorMap *map, const void *dev_rows, int64_t n_rows, int32_t dim) {
  return make_bf16_tensor_map(map, dev_rows, n_rows, dim, T2_BN_HALF);
}

template <int KLIST, int NA, int NB>
static constexpr size_t tc2_smem_bytes() {
  return 1024 + (size_t)NA * T2_A_BYTES + (size_t)NB * T2_B_BYTES + (size_t)KLIST * T2_BM * 8 + 4 * 32 * 32 * 4 + 256;
}

template <int KLIST, int NA, int NB, bool DBG>
static int tc2_launch_t(const CUtensorMap &map_q, const CUtensorMap &map_x, const CUtensorMap &map_q_lo,
                        const CUtensorMap &map_x_lo, const TcParams &p, int n_clusters,
                        cudaStream_t stream) {
  static std::atomic<uint64_t> configured{0};  // bit d: attribute set on device d (function attributes are per device)
  constexpr size_t smem = tc2_smem_bytes<KLIST, NA, NB>();
  static_assert(smem <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
  int dev = 0;
  AHR_CUDA_CHECK(cudaGetDevice(&dev));
  if (!((configured.load(std::memory_order_relaxed) >> (dev & 63)) & 1)) {
    AHR_CUDA_CHECK(cudaFuncSetAttribute(topk_tc2_kernel<KLIST, NA, NB, DBG>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
  }
  topk_tc2_kernel<KLIST, NA, NB, DBG><<<2 * n_clusters, TC_THREADS, smem, stream>>>(map_q, map_x, map_q_lo, map_x_lo, p);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// Launch the CTA-pair K1.  part_* must hold nq * n_slabs * k entries; n_pairs = SM count / 2.
int tc2_topk_partials(const CUtensorMap &map_x, const CUtensorMap *map_x_lo, const void *dev_q, const void *dev_q_lo,
                      int64_t n_rows, int32_t dim, int32_t nq, int32_t k, int32_t n_slabs, int n_pairs,
                      float *part_scores, int32_t *part_ids, float *dbg_scores,
                      unsigned long long *stats, const TcExtra &ex, cudaStream_t stream) {
  AHR_REQUIRE(k >= 1 && k <= 120, "tensor-core top-k supports 1 <= k <= 120 (got %d)", k);
  AHR_REQUIRE(dim % 8 == 0, "bf16 index needs dim %% 8 == 0 for TMA (got %d)", dim);
  AHR_REQUIRE(n_rows < (int64_t)INT32_MAX - T2_BN, "rows per index shard must fit int32 (got %lld)", (long long)n_rows);
  CUtensorMap map_q, map_q_lo;
  const int32_t q_rows = std::max(ex.q_rows, nq);
  int rc = make_bf16_tensor_map(&map_q, dev_q, q_rows, dim, T2_BM);
  if (rc != AHR_OK) return rc;
  const bool split = map_x_lo != nullptr && dev_q_lo != nullptr;  // fp32 operands as hi + lo bf16 halves
  if (split) {
    rc = make_bf16_tensor_map(&map_q_lo, dev_q_lo, q_rows, dim, T2_BM);
    if (rc != AHR_OK) return rc;
  } else {
    map_q_lo = map_q;
  }
  const CUtensorMap &mx_lo = split ? *map_x_lo : map_x;
  TcParams p;
  p.n_seg = split ? 3 : 1;
  p.n_rows = n_rows;
  p.tile_stride = std::max(1, ex.tile_stride);
  p.n_ctiles = ((n_rows + T2_BN - 1) / T2_BN + p.tile_stride - 1) / p.tile_stride;
  p.thr_src = ex.thr_src;
  p.thr_stride = ex.thr_stride;
  p.thr_shared = n_slabs > 1 ? ex.thr_shared : nullptr;  // one slab: its own threshold is the bound
  p.q_rows = std::max(ex.q_rows, nq);
  p.bound_m = (k + n_slabs - 1) / n_slabs;
  p.list_filter = tc_list_filter_default();
  {
    // slab-major unless AHR_TC_ITEM_ORDER=1 (experiment knob, see tc_item_decode)
    static const int order_env = getenv("AHR_TC_ITEM_ORDER") ? atoi(getenv("AHR_TC_ITEM_ORDER")) : -1;
    p.qtile_major = order_env >= 0 ? (order_env != 0) : 0;
  }
  p.nq = nq;
  p.dim = dim;
  p.n_qtiles = (nq + 2 * T2_BM - 1) / (2 * T2_BM);
  p.n_slabs = n_slabs;
  p.k = k;
  p.part_scores = part_scores;
  p.part_ids = part_ids;
  p.dbg_scores = dbg_scores;
  p.stats = stats;
  const int items = p.n_slabs * p.n_qtiles;
  int n_clusters = n_pairs < items ? n_pairs : items;
  if (n_clusters < 1) n_clusters = 1;
  {
    // partial bound (tc_shared.cuh): only when a query's slabs run in several rounds and lists publish from sorts
    static const int partial_env = getenv("AHR_TC_PARTIAL_BOUND") ? atoi(getenv("AHR_TC_PARTIAL_BOUND")) : 1;
    const int first_round_slabs = n_clusters / p.n_qtiles;  // slabs whose items all start in the first round (slab-major)
    p.partial_m = 0;
    if (partial_env && p.thr_shared != nullptr && items > n_clusters && !p.qtile_major && p.bound_m > 1 && first_round_slabs >= 1) {
      const int m2 = (k + first_round_slabs - 1) / first_round_slabs;
      if (m2 > p.bound_m && m2 <= k) p.partial_m = m2;
    }
  }
  if (dbg_scores) {
    AHR_REQUIRE(k <= 56, "debug score dump supports k <= 56");
    return tc2_launch_t<64, 4, 4, true>(map_q, map_x, map_q_lo, mx_lo, p, n_clusters, stream);
  }
  if (k <= 56) {  // the list needs room beyond k: KLIST - k appends amortise one compaction
    // rings: 4 x 16 KB of queries + 4 x 16 KB of corpus (128 KB), next to 64 KB of lists and 16 KB of chunk scratch;
    // the rest of the SM's 228 KB stays L1.  Measured alternatives (profiles/r1_k1_epilogue_history.md, "rings"):
    // 3+5 / 3+6 starve the tensor-bound workload (-18 %), 4+5 and 5+4 change nothing on either regime.
    // AHR_TC_RING=45 keeps the 4+5 build for experiments.
    // AHR_TC_RING = 44 | 45 | 36 forces one build.  Default: a single query tile (<= 256 queries) streams the corpus from
    // HBM while the same 12-16 query k-blocks come out of L2 again and again -- one more corpus slot takes the MMA issuer's
    // wait for data from 1.76M to 1.02M of 5.3M cycles (profiles/r2_k1_q256_wait_counters.md); with several query tiles the
    // corpus is mostly L2 traffic too and the extra 16 KB of shared memory only costs L1 (c3: 63 -> 69 ms).
    static const int ring_env = getenv("AHR_TC_RING") ? atoi(getenv("AHR_TC_RING")) : 0;
    const int ring = ring_env ? ring_env : (p.n_qtiles == 1 ? 45 : 44);
    if (ring == 45) return tc2_launch_t<64, 4, 5, false>(map_q, map_x, map_q_lo, mx_lo, p, n_clusters, stream);
    if (ring == 36) return tc2_launch_t<64, 3, 6, false>(map_q, map_x, map_q_lo, mx_lo, p, n_clusters, stream);
    return tc2_launch_t<64, 4, 4, false>(map_q, map_x, map_q_lo, mx_lo, p, n_clusters, stream);
  }
  return tc2_launch_t<128, 2, 2, false>(map_q, map_x, map_q_lo, mx_lo, p, n_clusters, stream);
}

}  // namespace ahr
