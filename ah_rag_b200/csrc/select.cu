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
