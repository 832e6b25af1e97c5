// topk_exact.cu -- K1x: the fp32-exact mode.
//
// canonical score  s(q, x) = (float) sum_{j ascending} (double) q[j] * (double) x[j]
// (a product of two fp32/bf16 values is exact in fp64, so DFMA == multiply-then-add and the fixed order
// makes the score bit-identical to the CPU oracle, oracle/exact_topk.c).  Ranking: (s desc, id asc).
//
// Two users:
//   * brute force: exact_scores over row chunks + select.cu (fp32 indexes, and the fallback below);
//   * bf16 indexes: the tensor-core kernel proposes kc > k candidates per query, exact_rescore scores
//     them canonically and exact_certify proves the candidate set contains the exact top-k
//     (k-th exact score  >  kc-th tensor score + error bound); queries that cannot be certified are
//     re-run through the brute-force scan.  Either way the returned ids are the exact ones.
#include "common.cuh"
#include "kernels.h"

namespace ahr {

template <typename T>
__device__ __forceinline__ double load_as_double(const T *p, int64_t i);
template <>
__device__ __forceinline__ double load_as_double<float>(const float *p, int64_t i) { return (double)p[i]; }
template <>
__device__ __forceinline__ double load_as_double<uint16_t>(const uint16_t *p, int64_t i) {
  return (double)bf16_bits_to_float(p[i]);
}

// ---------------------------------------------------------------------------------------------
// exact score tile: 64 rows x 16 queries per CTA, feature dimension walked in ascending order
// ---------------------------------------------------------------------------------------------
constexpr int EX_ROWS = 64;
constexpr int EX_QS = 16;
constexpr int EX_DCH = 32;
constexpr int EX_THREADS = 128;

template <typename T>
__global__ void __launch_bounds__(EX_THREADS)
exact_scores_kernel(const T *__restrict__ X, const T *__restrict__ Q, int32_t dim, int64_t row0, int64_t n_chunk,
                    const int32_t *__restrict__ q_sel, int32_t nq_sel, float *__restrict__ out) {
  __shared__ float xs[EX_ROWS][EX_DCH + 1];
  __shared__ float qs[EX_QS][EX_DCH];
  const int tx = threadIdx.x % EX_ROWS;  // row inside the tile
  const int ty = threadIdx.x / EX_ROWS;  // 0..1 -> queries ty*8 .. ty*8+7
  const int64_t r_base = (int64_t)blockIdx.x * EX_ROWS;
  const int q_base = blockIdx.y * EX_QS;
  double acc[EX_QS / 2];
#pragma unroll
  for (int i = 0; i < EX_QS / 2; ++i) acc[i] = 0.0;

  for (int d0 = 0; d0 < dim; d0 += EX_DCH) {
    __syncthreads();
    for (int e = threadIdx.x; e < EX_ROWS * EX_DCH; e += EX_THREADS) {
      const int r = e / EX_DCH, c = e % EX_DCH;
      const int64_t rr = r_base + r;
      float v = 0.f;
      if (rr < n_chunk && d0 + c < dim) v = (float)load_as_double<T>(X, (row0 + rr) * dim + d0 + c);
      xs[r][c] = v;
    }
    for (int e = threadIdx.x; e < EX_QS * EX_DCH; e += EX_THREADS) {
      const int qi = e / EX_DCH, c = e % EX_DCH;
      float v = 0.f;
      if (q_base + qi < nq_sel && d0 + c < dim) {
        const int64_t qrow = q_sel ? q_sel[q_base + qi] : (q_base + qi);
        v = (float)load_as_double<T>(Q, qrow * dim + d0 + c);
      }
      qs[qi][c] = v;
    }
    __syncthreads();
    const int dlim = min(EX_DCH, dim - d0);
    for (int c = 0; c < dlim; ++c) {  // ascending feature index == canonical order
      const double xv = (double)xs[tx][c];
#pragma unroll
      for (int i = 0; i < EX_QS / 2; ++i) acc[i] = __fma_rn((double)qs[ty * (EX_QS / 2) + i][c], xv, acc[i]);
    }
  }
  const int64_t rr = r_base + tx;
  if (rr < n_chunk) {
#pragma unroll
    for (int i = 0; i < EX_QS / 2; ++i) {
      const int qi = q_base + ty * (EX_QS / 2) + i;
      if (qi < nq_sel) out[(int64_t)qi * n_chunk + rr] = (float)acc[i];
    }
  }
}

int exact_scores(const void *dev_rows, const void *dev_q, int dtype, int32_t dim, int64_t row0, int64_t n_chunk,
                 const int32_t *q_sel, int32_t nq_sel, float *out, cudaStream_t stream) {
  if (n_chunk <= 0 || nq_sel <= 0) return AHR_OK;
  dim3 grid((unsigned)((n_chunk + EX_ROWS - 1) / EX_ROWS), (unsigned)((nq_sel + EX_QS - 1) / EX_QS));
  if (dtype == AHR_DTYPE_F32)
    exact_scores_kernel<float><<<grid, EX_THREADS, 0, stream>>>((const float *)dev_rows, (const float *)dev_q, dim,
                                                                 row0, n_chunk, q_sel, nq_sel, out);
  else
    exact_scores_kernel<uint16_t><<<grid, EX_THREADS, 0, stream>>>((const uint16_t *)dev_rows, (const uint16_t *)dev_q,
                                                                    dim, row0, n_chunk, q_sel, nq_sel, out);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// ---------------------------------------------------------------------------------------------
// canonical re-scoring of candidates: one thread per (query, candidate), sequential over features
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void rescore_kernel(const T *__restrict__ X, const T *__restrict__ Q, int32_t dim, int32_t nq, int32_t m,
                               const int64_t *__restrict__ cand, int64_t id_base, float *__restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nq * m) return;
  const int64_t qi = t / m;
  const int64_t id = cand[t];
  if (id < 0) {
    out[t] = -INFINITY;
    return;
  }
  const T *x = X + (id - id_base) * dim;
  const T *q = Q + qi * dim;
  double acc = 0.0;
  for (int j = 0; j < dim; ++j) acc = __fma_rn(load_as_double<T>(q, j), load_as_double<T>(x, j), acc);
  out[t] = (float)acc;
}

int exact_rescore(const void *dev_rows, const void *dev_q, int dtype, int32_t dim, int32_t nq, int32_t m,
                  const int64_t *cand_ids, int64_t id_base, float *out_scores, cudaStream_t stream) {
  const int64_t total = (int64_t)nq * m;
  if (total <= 0) return AHR_OK;
  const int threads = 128;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
  if (dtype == AHR_DTYPE_F32)
    rescore_kernel<float><<<blocks, threads, 0, stream>>>((const float *)dev_rows, (const float *)dev_q, dim, nq, m,
                                                          cand_ids, id_base, out_scores);
  else
    rescore_kernel<uint16_t><<<blocks, threads, 0, stream>>>((const uint16_t *)dev_rows, (const uint16_t *)dev_q, dim,
                                                             nq, m, cand_ids, id_base, out_scores);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// ---------------------------------------------------------------------------------------------
// certification: one warp per query.
//   tc_scores/tc_ids [nq, kc]  tensor-core candidates sorted by tensor score desc (id < 0 = empty slot)
//   ex_scores        [nq, kc]  their canonical scores
// Every row outside the candidate set has tensor score <= t_last (the kc-th), hence canonical score
// <= t_last + eps.  If the k-th best canonical score among the candidates is > t_last + eps, the exact
// top-k is inside the candidate set.  A list with an empty slot holds every row of the shard: trivially ok.
// ---------------------------------------------------------------------------------------------
constexpr int CERT_MAX_KC = 128;

__global__ void certify_kernel(const float *__restrict__ tc_scores, const int64_t *__restrict__ tc_ids,
                               const float *__restrict__ ex_scores, int32_t nq, int32_t kc, int32_t k,
                               const float *__restrict__ eps, float *__restrict__ out_scores,
                               int64_t *__restrict__ out_ids, int32_t *__restrict__ fail_flags,
                               int32_t *__restrict__ fail_count, int32_t *__restrict__ fail_list) {
  __shared__ float ss[4][CERT_MAX_KC];
  __shared__ int64_t si[4][CERT_MAX_KC];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * 4 + w;
  if (q >= nq) return;
  bool complete = false;  // candidate list holds the whole shard
  for (int j = lane; j < kc; j += 32) {
    const int64_t id = tc_ids[(int64_t)q * kc + j];
    ss[w][j] = id >= 0 ? ex_scores[(int64_t)q * kc + j] : -INFINITY;
    si[w][j] = id >= 0 ? id : INT64_MAX;
    if (id < 0) complete = true;
  }
  complete = __any_sync(0xffffffffu, complete);
  __syncwarp();
  // rank of every candidate under (exact score desc, id asc); kc <= 128 -> O(kc^2) is trivial
  int kept = 0;
  float kth = -INFINITY;
  for (int j = lane; j < kc; j += 32) {
    const float s = ss[w][j];
    const int64_t id = si[w][j];
    if (id == INT64_MAX) continue;
    int rank = 0;
    for (int t = 0; t < kc; ++t)
      if (t != j && si[w][t] != INT64_MAX && ranks_before(ss[w][t], si[w][t], s, id)) ++rank;
    if (rank < k) {
      out_scores[(int64_t)q * k + rank] = s;
      out_ids[(int64_t)q * k + rank] = id;
      ++kept;
      if (rank == k - 1) kth = s;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    kept += __shfl_xor_sync(0xffffffffu, kept, o);
    kth = fmaxf(kth, __shfl_xor_sync(0xffffffffu, kth, o));
  }
  for (int j = kept + lane; j < k; j += 32) {
    out_scores[(int64_t)q * k + j] = -INFINITY;
    out_ids[(int64_t)q * k + j] = -1;
  }
  if (lane == 0) {
    bool ok = complete;
    if (!ok && kept == k) {
      const float t_last = tc_scores[(int64_t)q * kc + kc - 1];
      ok = kth > t_last + eps[q];
    }
    fail_flags[q] = ok ? 0 : 1;
    if (!ok) fail_list[atomicAdd(fail_count, 1)] = q;
  }
}

int exact_certify(const float *tc_scores, const int64_t *tc_ids, const float *ex_scores, int32_t nq, int32_t kc,
                  int32_t k, const float *eps_per_query, float *out_scores, int64_t *out_ids, int32_t *fail_flags,
                  int32_t *fail_count, int32_t *fail_list, cudaStream_t stream) {
  AHR_REQUIRE(kc <= CERT_MAX_KC && k <= kc, "certify: need k <= kc <= %d (k=%d kc=%d)", CERT_MAX_KC, k, kc);
  if (nq <= 0) return AHR_OK;
  certify_kernel<<<(nq + 3) / 4, 128, 0, stream>>>(tc_scores, tc_ids, ex_scores, nq, kc, k, eps_per_query, out_scores,
                                                   out_ids, fail_flags, fail_count, fail_list);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// ---------------------------------------------------------------------------------------------
// norms (error-bound scaling) and small gather/scatter helpers
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void row_norm_max_kernel(const T *__restrict__ X, int64_t n_rows, int32_t dim, float *__restrict__ out_max) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float best = 0.f;
  for (int64_t r = warp_id; r < n_rows; r += n_warps) {
    float acc = 0.f;
    for (int j = lane; j < dim; j += 32) {
      const float v = (float)load_as_double<T>(X, r * dim + j);
      acc += v * v;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    best = fmaxf(best, acc);
  }
  if (lane == 0) atomicMax(reinterpret_cast<int *>(out_max), __float_as_int(sqrtf(best)));  // non-negative floats order as ints
}

int row_norms_max(const void *dev_rows, int dtype, int64_t n_rows, int32_t dim, float *dev_out_max, cudaStream_t stream) {
  AHR_CUDA_CHECK(cudaMemsetAsync(dev_out_max, 0, sizeof(float), stream));
  if (n_rows <= 0) return AHR_OK;
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((n_rows + 7) / 8, 148 * 8);
  if (dtype == AHR_DTYPE_F32)
    row_norm_max_kernel<float><<<blocks, threads, 0, stream>>>((const float *)dev_rows, n_rows, dim, dev_out_max);
  else
    row_norm_max_kernel<uint16_t><<<blocks, threads, 0, stream>>>((const uint16_t *)dev_rows, n_rows, dim, dev_out_max);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// eps[q] = eps_rel * ||q|| * max_row_norm  (bound on |tensor score - canonical score|)
template <typename T>
__global__ void query_eps_kernel(const T *__restrict__ Q, int32_t nq, int32_t dim, const float *__restrict__ max_row_norm,
                                 float eps_rel, float *__restrict__ eps) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  float acc = 0.f;
  for (int j = lane; j < dim; j += 32) {
    const float v = (float)load_as_double<T>(Q, (int64_t)q * dim + j);
    acc += v * v;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) eps[q] = eps_rel * fmaxf(sqrtf(acc) * max_row_norm[0], 1e-30f);
}

int query_eps(const void *dev_q, int dtype, int32_t nq, int32_t dim, const float *dev_max_row_norm, float eps_rel,
              float *dev_eps, cudaStream_t stream) {
  if (nq <= 0) return AHR_OK;
  const int threads = 128;
  const int blocks = (nq * 32 + threads - 1) / threads;
  if (dtype == AHR_DTYPE_F32)
    query_eps_kernel<float><<<blocks, threads, 0, stream>>>((const float *)dev_q, nq, dim, dev_max_row_norm, eps_rel, dev_eps);
  else
    query_eps_kernel<uint16_t><<<blocks, threads, 0, stream>>>((const uint16_t *)dev_q, nq, dim, dev_max_row_norm, eps_rel, dev_eps);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// fp32 -> two bf16 halves: hi = rn_bf16(x), lo = rn_bf16(x - hi); |x - hi - lo| <= 2^-18 |x|
__global__ void split_hi_lo_kernel(const float *__restrict__ src, int64_t n, __nv_bfloat16 *__restrict__ hi,
                                   __nv_bfloat16 *__restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float x = src[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

int split_f32_hi_lo(const float *src, int64_t n_elems, void *hi, void *lo, cudaStream_t stream) {
  if (n_elems <= 0) return AHR_OK;
  const int blocks = (int)std::min<int64_t>((n_elems + 255) / 256, 148 * 16);
  split_hi_lo_kernel<<<blocks, 256, 0, stream>>>(src, n_elems, reinterpret_cast<__nv_bfloat16 *>(hi),
                                                 reinterpret_cast<__nv_bfloat16 *>(lo));
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// scatter == false: dst[i] = src[sel[i]] ; scatter == true: dst[sel[i]] = src[i]   (rows of k entries)
__global__ void gather_rows_kernel(const float *__restrict__ src_s, const int64_t *__restrict__ src_i,
                                   const int32_t *__restrict__ sel, int32_t n_sel, int32_t k, float *__restrict__ dst_s,
                                   int64_t *__restrict__ dst_i, int scatter) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_sel * k) return;
  const int64_t i = t / k, j = t % k;
  const int64_t a = scatter ? i : sel[i], b = scatter ? sel[i] : i;
  dst_s[b * k + j] = src_s[a * k + j];
  dst_i[b * k + j] = src_i[a * k + j];
}

int gather_rows_i64(const float *src_s, const int64_t *src_i, const int32_t *sel, int32_t n_sel, int32_t k,
                    float *dst_s, int64_t *dst_i, bool scatter, cudaStream_t stream) {
  const int64_t total = (int64_t)n_sel * k;
  if (total <= 0) return AHR_OK;
  gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src_s, src_i, sel, n_sel, k, dst_s, dst_i,
                                                                           scatter ? 1 : 0);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// out[i] = src[sel[i]] for rows of row_words 32-bit words (query sub-batches of the tiered exact mode)
__global__ void gather_words_kernel(const uint32_t *__restrict__ src, int64_t row_words, const int32_t *__restrict__ sel,
                                    int32_t n_sel, uint32_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_sel * row_words) return;
  const int64_t r = i / row_words, c = i - r * row_words;
  out[i] = src[(int64_t)sel[r] * row_words + c];
}

int gather_query_rows(const void *src, int64_t row_bytes, const int32_t *sel, int32_t n_sel, void *out, cudaStream_t stream) {
  AHR_REQUIRE(row_bytes % 4 == 0, "gather_query_rows: rows must be a whole number of 32-bit words");
  const int64_t total = (int64_t)n_sel * (row_bytes / 4);
  if (total <= 0) return AHR_OK;
  gather_words_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const uint32_t *>(src), row_bytes / 4,
                                                                           sel, n_sel, reinterpret_cast<uint32_t *>(out));
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

}  // namespace ahr
