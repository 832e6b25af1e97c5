// select.cu -- per-query reduction of candidate (score, id) lists to the top k, ordered by
// (score desc, id asc).  Used for (a) merging the per-slab lists the fused GEMM epilogue emits,
// (b) the cross-GPU k-way merge after the candidate all-gather, (c) selection over exact-score chunks.
// The comparator is a total order (ids are unique), so the result does not depend on how candidates
// were partitioned or in which order they arrive: fp32-exact results are identical for any sharding.
#include "common.cuh"
#include "kernels.h"

namespace ahr {

constexpr int SEL_THREADS = 256;
constexpr int SEL_N = 2048;  // sort capacity; k <= SELECT_MAX_K = SEL_N / 2

struct SelSmem {
  float s[SEL_N];
  int64_t id[SEL_N];
  int fill;
  int have_thr;
  float thr_s;
  int64_t thr_id;
};

// bitonic sort of the first P (power of two) slots, best candidate first
__device__ void sel_sort(SelSmem &sm, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < P / 2; t += SEL_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc_block = (lo & size) == 0;  // this block sorts "best first"
        const float s0 = sm.s[lo], s1 = sm.s[hi];
        const int64_t i0 = sm.id[lo], i1 = sm.id[hi];
        const bool hi_first = ranks_before(s1, i1, s0, i0);
        if (hi_first == desc_block) {
          sm.s[lo] = s1; sm.id[lo] = i1;
          sm.s[hi] = s0; sm.id[hi] = i0;
        }
      }
    }
  }
  __syncthreads();
}

// sort what has been collected, keep the best k, refresh the admission threshold
__device__ void sel_reduce(SelSmem &sm, int k) {
  __syncthreads();
  const int fill = sm.fill;
  int P = 32;
  while (P < fill) P <<= 1;
  for (int t = fill + threadIdx.x; t < P; t += SEL_THREADS) {
    sm.s[t] = -INFINITY;
    sm.id[t] = INT64_MAX;
  }
  sel_sort(sm, P);
  if (threadIdx.x == 0) {
    const int kept = fill < k ? fill : k;
    sm.fill = kept;
    if (kept == k) {
      sm.have_thr = 1;
      sm.thr_s = sm.s[k - 1];
      sm.thr_id = sm.id[k - 1];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(SEL_THREADS)
select_topk_kernel(const float *__restrict__ cand_scores, const int32_t *__restrict__ ids32,
                   const int64_t *__restrict__ ids64, int64_t stride, int64_t n_cand, int64_t inner,
                   int64_t stride_outer, int64_t stride_outer_ids, int32_t k, int64_t id_offset, int carry_in,
                   float *__restrict__ out_scores, int64_t *__restrict__ out_ids) {
  __shared__ SelSmem sm;
  const int64_t q = blockIdx.x;
  if (threadIdx.x == 0) {
    sm.fill = 0;
    sm.have_thr = 0;
  }
  __syncthreads();
  if (carry_in) {
    for (int j = threadIdx.x; j < k; j += SEL_THREADS) {
      const int64_t id = out_ids[q * k + j];
      if (id >= 0) {
        const int slot = atomicAdd(&sm.fill, 1);
        sm.s[slot] = out_scores[q * k + j];
        sm.id[slot] = id;
      }
    }
  }
  // candidate i of query q lives at q * stride + (i / inner) * stride_outer + (i % inner)
  for (int64_t base = 0; base < n_cand; base += SEL_THREADS) {
    __syncthreads();
    const int fill_now = sm.fill;  // read between two barriers: identical in every thread
    __syncthreads();
    if (fill_now + SEL_THREADS > SEL_N) sel_reduce(sm, k);
    const int64_t i = base + threadIdx.x;
    if (i < n_cand) {
      const int64_t at = q * stride + (i / inner) * stride_outer + (i % inner);
      const int64_t at_id = q * stride + (i / inner) * stride_outer_ids + (i % inner);
      const float s = cand_scores[at];
      int64_t id = ids32 ? (int64_t)ids32[at_id] : (ids64 ? ids64[at_id] : i);
      bool ok = id >= 0 && s == s;  // skip empty slots and NaN scores
      id += id_offset;
      if (ok && sm.have_thr) ok = ranks_before(s, id, sm.thr_s, sm.thr_id);
      if (ok) {
        const int slot = atomicAdd(&sm.fill, 1);
        AHR_DEV_ASSERT(slot >= 0 && slot < SEL_N);
        sm.s[slot] = s;
        sm.id[slot] = id;
      }
    }
  }
  sel_reduce(sm, k);
  const int kept = sm.fill;
  for (int j = threadIdx.x; j < k; j += SEL_THREADS) {
    out_scores[q * k + j] = j < kept ? sm.s[j] : -INFINITY;
    out_ids[q * k + j] = j < kept ? sm.id[j] : -1;
  }
}

// ---------------------------------------------------------------------------------------------
// Merge of a few SORTED lists (the slab lists K1 flushes, the per-GPU lists of the exchange): rank counting.
//
// Every list is already ordered by (score desc, id asc) with its empty slots at the end, and the order is total, so the
// final position of candidate (list g, slot j) is  j + sum over the other lists of the number of entries ranking before
// it -- a binary search per other list, no sort, one barrier.  The bitonic selection above needs ~45 barrier-separated
// passes over 512 slots to merge 8 x 50 candidates (0.136 ms for 4096 queries at 8 GPUs, profiles/r1_comm_stats_n8.txt).
//
// PUSH: the block also stores its query's k results into slot `rank` of every PEER's exchange buffer over NVLink and the
// last block of the grid publishes the epoch flags (what push_kernel did in a launch of its own, exchange.cu).
// ---------------------------------------------------------------------------------------------
constexpr int MRG_THREADS = 128;
constexpr int MRG_MAX_LISTS = 16;
constexpr int MRG_MAX_CAND = 2048;

// dynamic shared memory: id[n_cand] res_id[k] (int64) | s[n_cand] res_s[k] (float) | len[MRG_MAX_LISTS] n_valid last (int)
__host__ __device__ inline size_t merge_smem_bytes(int n_cand, int k) {
  return (size_t)(n_cand + k) * 12 + (MRG_MAX_LISTS + 2) * 4 + 16;
}
struct MergeSmem {
  int64_t *id, *res_id;
  float *s, *res_s;
  int *len;
  int &n_valid;
  int &last;
  __device__ MergeSmem(uint8_t *raw, int n_cand, int k)
      : id(reinterpret_cast<int64_t *>(raw)), res_id(id + n_cand), s(reinterpret_cast<float *>(res_id + k)), res_s(s + n_cand),
        len(reinterpret_cast<int *>(res_s + k)), n_valid(len[MRG_MAX_LISTS]), last(len[MRG_MAX_LISTS + 1]) {}
};

__device__ __forceinline__ void st_release_sys_u64_sel(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

template <bool PUSH>
__global__ void __launch_bounds__(MRG_THREADS)
merge_sorted_lists_kernel(const float *__restrict__ cand_scores, const int32_t *__restrict__ ids32,
                          const int64_t *__restrict__ ids64, int64_t stride, int32_t n_lists, int32_t k_in,
                          int64_t stride_list, int64_t stride_list_ids, int32_t k, int64_t id_offset,
                          float *__restrict__ out_scores, int64_t *__restrict__ out_ids, PeerPush push) {
  extern __shared__ __align__(16) uint8_t mrg_raw[];
  const int64_t q = blockIdx.x;
  const int n_cand = n_lists * k_in;
  MergeSmem sm(mrg_raw, n_cand, k);
  // ---- stage the lists (list g, slot j at q * stride + g * stride_list + j) ----
  for (int c = threadIdx.x; c < n_cand; c += MRG_THREADS) {
    const int g = c / k_in, j = c - g * k_in;
    const float sc = cand_scores[q * stride + g * stride_list + j];
    const int64_t at_id = q * stride + g * stride_list_ids + j;
    const int64_t id = ids32 ? (int64_t)ids32[at_id] : ids64[at_id];
    const bool ok = id >= 0 && sc == sc;
    sm.s[c] = ok ? sc : -INFINITY;
    sm.id[c] = ok ? id + id_offset : INT64_MAX;  // empties rank after everything
  }
  __syncthreads();
  if (threadIdx.x < n_lists) {  // valid entries come first in every list
    const int g = threadIdx.x;
    int lo = 0, hi = k_in;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sm.id[g * k_in + mid] != INT64_MAX) lo = mid + 1; else hi = mid;
    }
    sm.len[g] = lo;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    for (int g = 0; g < n_lists; ++g) n += sm.len[g];
    sm.n_valid = n;
  }
  // ---- rank of every valid candidate ----
  for (int c = threadIdx.x; c < n_cand; c += MRG_THREADS) {
    const int g = c / k_in, j = c - g * k_in;
    if (j >= sm.len[g]) continue;
    const float sc = sm.s[c];
    const int64_t id = sm.id[c];
    int rank = j;
    for (int o = 0; o < n_lists && rank < k; ++o) {
      if (o == g) continue;
      const float *ls = sm.s + o * k_in;
      const int64_t *li = sm.id + o * k_in;
      int lo = 0, hi = sm.len[o];  // number of entries of list o that rank before (sc, id)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ranks_before(ls[mid], li[mid], sc, id)) lo = mid + 1; else hi = mid;
      }
      rank += lo;
    }
    AHR_DEV_ASSERT(rank >= 0);
    if (rank < k) { sm.res_s[rank] = sc; sm.res_id[rank] = id; }
  }
  __syncthreads();
  const int kept = sm.n_valid < k ? sm.n_valid : k;
  for (int j = threadIdx.x; j < k; j += MRG_THREADS) {
    const float sc = j < kept ? sm.res_s[j] : -INFINITY;
    const int64_t id = j < kept ? sm.res_id[j] : -1;
    out_scores[q * k + j] = sc;
    out_ids[q * k + j] = id;
    if (PUSH) {
      for (int pi = 0; pi < push.world - 1; ++pi) {  // slot `rank` of every peer, over NVLink (coalesced per warp)
        const int peer = pi < push.rank ? pi : pi + 1;
        uint8_t *slot = push.base[peer] + push.slot_off;
        reinterpret_cast<float *>(slot)[q * k + j] = sc;
        reinterpret_cast<int64_t *>(slot + push.scores_bytes)[q * k + j] = id;
      }
    }
  }
  if (PUSH) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) sm.last = atomicAdd(push.done_counter, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!sm.last) return;
    __threadfence_system();
    if ((int)threadIdx.x < push.world && (int)threadIdx.x != push.rank)
      st_release_sys_u64_sel(reinterpret_cast<uint64_t *>(push.base[threadIdx.x] + push.flag_off) + push.rank, push.epoch);
    if (threadIdx.x == 0) *push.done_counter = 0;  // ready for the next launch (stream-ordered)
  }
}

bool merge_sorted_applicable(int32_t n_lists, int32_t k_in, int32_t k) {
  return n_lists >= 1 && n_lists <= MRG_MAX_LISTS && (int64_t)n_lists * k_in <= MRG_MAX_CAND && k <= SELECT_MAX_K;
}

int merge_sorted_lists(const float *cand_scores, const int32_t *cand_ids32, const int64_t *cand_ids64, int64_t stride,
                       int32_t n_lists, int32_t k_in, int64_t stride_list, int64_t stride_list_ids, int32_t nq, int32_t k,
                       int64_t id_offset, float *out_scores, int64_t *out_ids, const PeerPush *push, cudaStream_t stream) {
  AHR_REQUIRE(merge_sorted_applicable(n_lists, k_in, k), "merge_sorted_lists: %d lists of %d entries exceed the kernel's capacity",
              n_lists, k_in);
  AHR_REQUIRE((cand_ids32 != nullptr) != (cand_ids64 != nullptr), "merge_sorted_lists: exactly one id array");
  if (nq <= 0) return AHR_OK;
  const size_t smem = merge_smem_bytes(n_lists * k_in, k);  // <= 36.2 KB at the capacity limits: no opt-in needed
  if (push != nullptr && push->world > 1)
    merge_sorted_lists_kernel<true><<<nq, MRG_THREADS, smem, stream>>>(cand_scores, cand_ids32, cand_ids64, stride, n_lists, k_in,
                                                                     stride_list, stride_list_ids, k, id_offset, out_scores, out_ids, *push);
  else
    merge_sorted_lists_kernel<false><<<nq, MRG_THREADS, smem, stream>>>(cand_scores, cand_ids32, cand_ids64, stride, n_lists, k_in,
                                                                      stride_list, stride_list_ids, k, id_offset, out_scores, out_ids, PeerPush{});
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

int select_topk_strided2(const float *cand_scores, const int32_t *cand_ids32, const int64_t *cand_ids64, int64_t stride,
                         int64_t n_cand, int64_t inner, int64_t stride_outer, int64_t stride_outer_ids, int32_t nq,
                         int32_t k, int64_t id_offset, bool carry_in, float *out_scores, int64_t *out_ids,
                         cudaStream_t stream) {
  AHR_REQUIRE(k >= 1 && k <= SELECT_MAX_K, "k must be in [1, %d] (got %d)", SELECT_MAX_K, k);
  if (nq <= 0) return AHR_OK;
  if (inner <= 0) inner = 1;
  select_topk_kernel<<<nq, SEL_THREADS, 0, stream>>>(cand_scores, cand_ids32, cand_ids64, stride, n_cand, inner,
                                                     stride_outer, stride_outer_ids, k, id_offset, carry_in ? 1 : 0,
                                                     out_scores, out_ids);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

int select_topk_strided(const float *cand_scores, const int32_t *cand_ids32, const int64_t *cand_ids64, int64_t stride,
                        int64_t n_cand, int64_t inner, int64_t stride_outer, int32_t nq, int32_t k, int64_t id_offset,
                        bool carry_in, float *out_scores, int64_t *out_ids, cudaStream_t stream) {
  return select_topk_strided2(cand_scores, cand_ids32, cand_ids64, stride, n_cand, inner, stride_outer, stride_outer, nq,
                              k, id_offset, carry_in, out_scores, out_ids, stream);
}

int select_topk(const float *cand_scores, const int32_t *cand_ids32, const int64_t *cand_ids64, int64_t stride,
                int64_t n_cand, int32_t nq, int32_t k, int64_t id_offset, bool carry_in, float *out_scores,
                int64_t *out_ids, cudaStream_t stream) {
  // contiguous candidates: a single "outer" block
  return select_topk_strided(cand_scores, cand_ids32, cand_ids64, stride, n_cand, n_cand > 0 ? n_cand : 1, 0, nq, k,
                             id_offset, carry_in, out_scores, out_ids, stream);
}

}  // namespace ahr
