// topk_tc.cu -- K1: fused bf16 similarity GEMM (tcgen05, TMEM accumulators, TMA-staged tiles) with a
// top-k epilogue straight out of TMEM.  The [nq, n_rows] score matrix never touches HBM.
//
// Replaces, for a whole query batch, the per-query Collection.query call of the reference
// (src/ah_rag/graph/hierarchical_graph.py:714).
//
// Decomposition
//   query tile  : 128 queries  = UMMA M (one TMEM lane per query)
//   corpus tile : 256 rows     = UMMA N (one TMEM column per row), fp32 accumulator 128 x 256
//   k-block     : 64 bf16      = one 128-byte swizzle row; 4 x tcgen05.mma (K = 16) per k-block
//   item        : (slab, query tile).  The corpus is cut into S slabs of consecutive tiles; item index
//                 = slab * n_qtiles + qtile, CTA b runs items b, b + grid, ...  Slab-major order makes the
//                 CTAs that run at the same time sweep the SAME few slabs with different query tiles, so
//                 a corpus tile is fetched from HBM once and hit in L2 by the other query tiles.
//   per item the epilogue keeps one sorted k-list per query row in shared memory (threshold filter in
//   registers, insertion on the rare hit) and flushes it to part[q][slab][k] at the end of the item;
//   select.cu then reduces the S lists of a query to the final top-k.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer (one
// lane), warps 2..5 = epilogue (warp w reads TMEM lanes 32*(w%4) .. +31: thread = query row).
// TMEM: 512 columns = 2 accumulator stages x 256, so the epilogue drains tile i while tile i+1 is computed.
#include "common.cuh"
#include "kernels.h"
#include "tc_shared.cuh"

namespace ahr {

template <int KLIST, int TC_STAGES, bool DBG>
__global__ void __launch_bounds__(TC_THREADS, 1)
topk_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled operand tiles need 1024-byte alignment
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *tiles = smem;
  unsigned long long *lists = reinterpret_cast<unsigned long long *>(tiles + TC_STAGES * TC_STAGE_BYTES);  // [128][KLIST] keys
  float *scratch = reinterpret_cast<float *>(lists + KLIST * TC_BM);  // [4 warps][32 columns][32 lanes] chunk scratch
  uint64_t *bars = reinterpret_cast<uint64_t *>(scratch + 4 * 32 * 32);
  uint64_t *full = bars;                  // [STAGES] TMA -> MMA
  uint64_t *empty = bars + TC_STAGES;     // [STAGES] MMA -> TMA
  uint64_t *tfull = bars + 2 * TC_STAGES; // [2] MMA -> epilogue
  uint64_t *tempty = tfull + 2;           // [2] epilogue -> MMA
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_items = p.n_slabs * p.n_qtiles;
  const int n_kblocks = (p.dim + TC_BK - 1) / TC_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_q);
    prefetch_tensormap(&map_x);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < TC_STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], 4);  // one arrival per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int slab = item / p.n_qtiles, qt = item % p.n_qtiles;
        int64_t t0, t1;
        slab_range(p, slab, t0, t1);
        for (int64_t ct = t0; ct < t1; ++ct) {
          for (int kb = 0; kb < n_kblocks; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t *a_dst = tiles + stage * TC_STAGE_BYTES;
            uint8_t *b_dst = a_dst + TC_A_BYTES;
            mbar_arrive_expect_tx(&full[stage], TC_STAGE_BYTES);
            // queries are re-read by every corpus tile: keep them in L2; the corpus streams through
            tma_load_2d(a_dst, &map_q, &full[stage], kb * TC_BK, qt * TC_BM, kEvictLast);
            tma_load_2d(b_dst, &map_x, &full[stage], kb * TC_BK, (int32_t)(ct * p.tile_stride * TC_BN), kEvictNormal);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();  // reconverge before the block/cluster-wide barrier below
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(TC_BM, TC_BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int slab = item / p.n_qtiles;
        int64_t t0, t1;
        slab_range(p, slab, t0, t1);
        for (int64_t ct = t0; ct < t1; ++ct) {
          mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator stage
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TC_BN;
          for (int kb = 0; kb < n_kblocks; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(tiles + stage * TC_STAGE_BYTES);
            const uint64_t a_desc = make_smem_desc_sw128(a_addr);
            const uint64_t b_desc = make_smem_desc_sw128(a_addr + TC_A_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte address units
              umma_bf16_ss(d_tmem, a_desc + (uint64_t)(k4 * 2), b_desc + (uint64_t)(k4 * 2), idesc,
                           (kb | k4) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[stage]);  // smem stage reusable once these MMAs have read it
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tfull[acc]);  // accumulator complete -> epilogue
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> registers -> per-row top-k =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access; thread = query row = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t lists_warp = smem_u32(lists) + (uint32_t)(quad * 32) * (KLIST * 8);
    const uint32_t scratch_warp = smem_u32(scratch) + (uint32_t)quad * (32 * 32 * 4);
    const int k = p.k;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int slab = item / p.n_qtiles, qt = item % p.n_qtiles;
      int64_t t0, t1;
      slab_range(p, slab, t0, t1);
      const int64_t q_first = (int64_t)qt * TC_BM + quad * 32;
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
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        thr = shared_bound_apply(bound_bits, thr);
        const int64_t col0 = ct * p.tile_stride * TC_BN;
        const int ncols = (int)min((int64_t)TC_BN, p.n_rows - col0);
        uint64_t *release_bar = &tempty[acc];
        epilogue_scan_tile<KLIST, TC_BN, DBG>(t_lane + (uint32_t)(acc * TC_BN), ncols, row_valid, k, col0, q_global, p,
                                              lists_warp, scratch_warp, thr, cnt, nullptr, bound_slot, p.bound_m, best, age == 0, [&]() {
                                                tc_fence_before();
                                                __syncwarp();
                                                if (lane == 0) mbar_arrive(release_bar);
                                              });
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      lists_flush<KLIST>(lists_warp, k, cnt, q_first, p, slab);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// [rows, dim] bf16 row-major -> tiled map with box {64, box_rows}, 128-byte swizzle, zero OOB fill
int make_bf16_tensor_map(CUtensorMap *map, const void *dev_ptr, int64_t rows, int32_t dim, int32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return AHR_E_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(dev_ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld dim=%d box_rows=%d ptr=%p)", (int)r,
              (long long)rows, dim, box_rows, dev_ptr);
    return AHR_E_CUDA;
  }
  return AHR_OK;
}

int tc_make_corpus_map(CUtensorMap *map, const void *dev_rows, int64_t n_rows, int32_t dim) {
  return make_bf16_tensor_map(map, dev_rows, n_rows, dim, TC_BN);
}

// number of corpus slabs: minimise the makespan ceil(items / grid) * ceil(tiles / S) with a mild
// preference for few slabs (each slab restarts the per-row lists -> more insertions, more partials)
int tc_choose_slabs(int32_t n_qtiles, int64_t n_ctiles, int grid) {
  int best = 1;
  double best_cost = 1e300;
  const int smax = (int)std::min<int64_t>(n_ctiles, 160);
  for (int s = 1; s <= smax; ++s) {
    const int64_t items = (int64_t)n_qtiles * s;
    const int64_t rounds = (items + grid - 1) / grid;
    const int64_t tiles = (n_ctiles + s - 1) / s;
    const double cost = (double)rounds * (double)tiles * (1.0 + 0.002 * s);
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

template <int KLIST, int STAGES>
static constexpr size_t tc_smem_bytes() {
  return 1024 + (size_t)STAGES * TC_STAGE_BYTES + (size_t)KLIST * TC_BM * 8 + 4 * 32 * 32 * 4 + 128;
}

template <int KLIST, int STAGES, bool DBG>
static int tc_launch_t(const CUtensorMap &map_q, const CUtensorMap &map_x, const TcParams &p, int grid,
                       cudaStream_t stream) {
  static std::atomic<uint64_t> configured{0};  // bit d: attribute set on device d (function attributes are per device)
  constexpr size_t smem = tc_smem_bytes<KLIST, STAGES>();
  static_assert(smem <= 232448, "exceeds the 227 KB shared memory of an sm_100 CTA");
  int dev = 0;
  AHR_CUDA_CHECK(cudaGetDevice(&dev));
  if (!((configured.load(std::memory_order_relaxed) >> (dev & 63)) & 1)) {
    AHR_CUDA_CHECK(cudaFuncSetAttribute(topk_tc_kernel<KLIST, STAGES, DBG>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
  }
  topk_tc_kernel<KLIST, STAGES, DBG><<<grid, TC_THREADS, smem, stream>>>(map_q, map_x, p);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// Launch K1.  part_* must hold nq * n_slabs * k entries.  k <= 128.
int tc_topk_partials(const CUtensorMap &map_x, const void *dev_q, int64_t n_rows, int32_t dim, int32_t nq, int32_t k,
                     int32_t n_slabs, int grid, float *part_scores, int32_t *part_ids, float *dbg_scores,
                     const TcExtra &ex, cudaStream_t stream) {
  AHR_REQUIRE(k >= 1 && k <= 120, "tensor-core top-k supports 1 <= k <= 120 (got %d)", k);
  AHR_REQUIRE(dim % 8 == 0, "bf16 index needs dim %% 8 == 0 for TMA (got %d)", dim);
  AHR_REQUIRE(n_rows < (int64_t)INT32_MAX - TC_BN, "rows per index shard must fit int32 (got %lld)", (long long)n_rows);
  CUtensorMap map_q;
  int rc = make_bf16_tensor_map(&map_q, dev_q, std::max(ex.q_rows, nq), dim, TC_BM);
  if (rc != AHR_OK) return rc;
  TcParams p;
  p.n_rows = n_rows;
  p.tile_stride = std::max(1, ex.tile_stride);
  p.n_ctiles = ((n_rows + TC_BN - 1) / TC_BN + p.tile_stride - 1) / p.tile_stride;
  p.thr_src = ex.thr_src;
  p.thr_stride = ex.thr_stride;
  p.thr_shared = n_slabs > 1 ? ex.thr_shared : nullptr;  // one slab: its own threshold is the bound
  p.q_rows = std::max(ex.q_rows, nq);
  p.bound_m = (k + n_slabs - 1) / n_slabs;
  p.list_filter = tc_list_filter_default();
  p.qtile_major = 0;
  p.partial_m = 0;
  p.nq = nq;
  p.dim = dim;
  p.n_qtiles = (nq + TC_BM - 1) / TC_BM;
  p.n_slabs = n_slabs;
  p.k = k;
  p.part_scores = part_scores;
  p.part_ids = part_ids;
  p.dbg_scores = dbg_scores;
  p.stats = nullptr;
  p.n_seg = 1;
  const int items = p.n_slabs * p.n_qtiles;
  if (grid > items) grid = items;
  if (dbg_scores) {
    AHR_REQUIRE(k <= 56, "debug score dump supports k <= 56");
    return tc_launch_t<64, 3, true>(map_q, map_x, p, grid, stream);
  }
  AHR_REQUIRE(k <= 56, "the single-CTA variant (AHR_TC_VARIANT=1) supports k <= 56 (got %d)", k);
  return tc_launch_t<64, 3, false>(map_q, map_x, p, grid, stream);
}

}  // namespace ahr
