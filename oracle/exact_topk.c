/*
 * oracle/exact_topk.c -- CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * Exact brute-force restatement of the nearest-neighbour query the reference issues at
 *   src/ah_rag/graph/hierarchical_graph.py:714   coll.query(query_embeddings=[q], n_results=top_k)
 * against a collection created with {"hnsw:space": "cosine"} (hierarchical_graph.py:611) whose rows
 * were L2-normalised by the producer (hierarchical_graph.py:655-656, normalize_embeddings=True):
 *   distance(q, x) = 1 - <q, x>,  results in ascending distance.
 * chromadb / chroma-hnswlib are not vendored or pinned by the reference (requirements.txt:1), so this
 * is "parity unpinned" at that seam; the contract we pin instead (BASELINE.json north_star):
 *
 *   canonical score  s(q, x) = (float) sum_{j = 0 .. d-1, ascending} (double) q[j] * (double) x[j]
 *   ranking          (s descending, row id ascending), first k
 *
 * A product of two fp32 (or bf16) values is exact in fp64, so fused and unfused multiply-add agree and
 * the fixed summation order makes s bit-reproducible on any IEEE-754 machine, CPU or GPU.
 *
 * Build:  make -C oracle      (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, sizeof f);
    return f;
}

/* dtype: 0 = fp32 rows, 1 = bf16 rows (raw uint16).  Queries use the same dtype. */
static inline double canon_dot(const void *q, const void *x, int32_t d, int dtype) {
    double acc = 0.0;
    if (dtype == 0) {
        const float *a = (const float *)q, *b = (const float *)x;
        for (int32_t j = 0; j < d; ++j) acc += (double)a[j] * (double)b[j];
    } else {
        const uint16_t *a = (const uint16_t *)q, *b = (const uint16_t *)x;
        for (int32_t j = 0; j < d; ++j) acc += (double)bf16_to_f32(a[j]) * (double)bf16_to_f32(b[j]);
    }
    return acc;
}

/* canon_dot of four consecutive rows x, x + d, x + 2d, x + 3d at once (same per-pair summation order) */
static inline void canon_dot4(const void *q, const void *x, int32_t d, int dtype, double acc[4]) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (dtype == 0) {
        const float *a = (const float *)q, *b = (const float *)x;
        for (int32_t j = 0; j < d; ++j) {
            const double qj = (double)a[j];
            a0 += qj * (double)b[j];
            a1 += qj * (double)b[(size_t)d + j];
            a2 += qj * (double)b[2 * (size_t)d + j];
            a3 += qj * (double)b[3 * (size_t)d + j];
        }
    } else {
        const uint16_t *a = (const uint16_t *)q, *b = (const uint16_t *)x;
        for (int32_t j = 0; j < d; ++j) {
            const double qj = (double)bf16_to_f32(a[j]);
            a0 += qj * (double)bf16_to_f32(b[j]);
            a1 += qj * (double)bf16_to_f32(b[(size_t)d + j]);
            a2 += qj * (double)bf16_to_f32(b[2 * (size_t)d + j]);
            a3 += qj * (double)bf16_to_f32(b[3 * (size_t)d + j]);
        }
    }
    acc[0] = a0; acc[1] = a1; acc[2] = a2; acc[3] = a3;
}

/* returns 1 when (s1,id1) ranks strictly before (s2,id2) */
static inline int ranks_before(float s1, int64_t id1, float s2, int64_t id2) {
    return (s1 > s2) || (s1 == s2 && id1 < id2);
}

/*
 * Canonical scores of every (query, row) pair: out[nq * n], row-major by query.  Small inputs only.
 */
int ahro_canonical_scores(const void *X, const void *Q, int64_t n, int32_t d, int32_t nq, int dtype,
                          float *out) {
    const size_t esz = dtype == 0 ? 4 : 2;
#pragma omp parallel for schedule(static)
    for (int32_t qi = 0; qi < nq; ++qi) {
        const char *q = (const char *)Q + (size_t)qi * d * esz;
        for (int64_t r = 0; r < n; ++r)
            out[(size_t)qi * n + r] = (float)canon_dot(q, (const char *)X + (size_t)r * d * esz, d, dtype);
    }
    return 0;
}

/*
 * Exact top-k of every query over rows [0, n); ids are id_base + row.  Lists shorter than k
 * (n < k) are padded with score = -inf, id = -1.  Output sorted by (score desc, id asc).
 */
int ahro_exact_topk(const void *X, const void *Q, int64_t n, int32_t d, int32_t nq, int dtype, int32_t k,
                    int64_t id_base, float *out_scores, int64_t *out_ids) {
    if (k <= 0 || d <= 0 || nq < 0 || n < 0) return -1;
    const size_t esz = dtype == 0 ? 4 : 2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int32_t qi = 0; qi < nq; ++qi) {
        const char *q = (const char *)Q + (size_t)qi * d * esz;
        float *bs = out_scores + (size_t)qi * k;
        int64_t *bi = out_ids + (size_t)qi * k;
        int32_t cnt = 0;
        for (int64_t r = 0; r < n; ++r) {
            float s = (float)canon_dot(q, (const char *)X + (size_t)r * d * esz, d, dtype);
            int64_t id = id_base + r;
            if (cnt == k && !ranks_before(s, id, bs[k - 1], bi[k - 1])) continue;
            int32_t p = cnt < k ? cnt : k - 1;
            while (p > 0 && ranks_before(s, id, bs[p - 1], bi[p - 1])) {
                bs[p] = bs[p - 1];
                bi[p] = bi[p - 1];
                --p;
            }
            bs[p] = s;
            bi[p] = id;
            if (cnt < k) ++cnt;
        }
        for (int32_t p = cnt; p < k; ++p) {
            bs[p] = -INFINITY;
            bi[p] = -1;
        }
    }
    return 0;
}

/*
 * Same result as ahro_exact_topk, parallel over ROW BLOCKS instead of queries (few queries, many rows: the sampled
 * parity check of bench.py over the full corpus).  Every block keeps its own list per query with the same comparator;
 * the lists are merged with it afterwards.  (score desc, id asc) is a total order, so the result does not depend on the
 * blocking -- tests/test_oracle.py checks it equals ahro_exact_topk bit for bit.
 */
static void list_insert(float *bs, int64_t *bi, int32_t *cnt, int32_t k, float s, int64_t id) {
    if (*cnt == k && !ranks_before(s, id, bs[k - 1], bi[k - 1])) return;
    int32_t p = *cnt < k ? *cnt : k - 1;
    while (p > 0 && ranks_before(s, id, bs[p - 1], bi[p - 1])) {
        bs[p] = bs[p - 1];
        bi[p] = bi[p - 1];
        --p;
    }
    bs[p] = s;
    bi[p] = id;
    if (*cnt < k) ++*cnt;
}

int ahro_exact_topk_rowpar(const void *X, const void *Q, int64_t n, int32_t d, int32_t nq, int dtype, int32_t k,
                           int64_t id_base, int32_t n_threads, float *out_scores, int64_t *out_ids) {
    if (k <= 0 || d <= 0 || nq < 0 || n < 0) return -1;
    const size_t esz = dtype == 0 ? 4 : 2;
    const int64_t block = 4096;
    const int64_t n_blocks = (n + block - 1) / block;
    int64_t n_lists = n_threads > 0 ? n_threads : 1;
    if (n_lists > n_blocks) n_lists = n_blocks > 0 ? n_blocks : 1;
    float *ls = (float *)malloc((size_t)n_lists * nq * k * sizeof(float));
    int64_t *li = (int64_t *)malloc((size_t)n_lists * nq * k * sizeof(int64_t));
    int32_t *lc = (int32_t *)calloc((size_t)n_lists * nq, sizeof(int32_t));
    if (!ls || !li || !lc) { free(ls); free(li); free(lc); return -2; }
#pragma omp parallel for schedule(static) num_threads((int)n_lists)
    for (int64_t t = 0; t < n_lists; ++t) {
        /* list t owns blocks t, t + n_lists, ... */
        for (int64_t b = t; b < n_blocks; b += n_lists) {
            const int64_t r0 = b * block, r1 = r0 + block < n ? r0 + block : n;
            for (int32_t qi = 0; qi < nq; ++qi) {
                const char *q = (const char *)Q + (size_t)qi * d * esz;
                float *bs = ls + ((size_t)t * nq + qi) * k;
                int64_t *bi = li + ((size_t)t * nq + qi) * k;
                int32_t *cnt = lc + (size_t)t * nq + qi;
                int64_t r = r0;
                for (; r + 4 <= r1; r += 4) {
                    /* four rows side by side: four independent accumulation chains, each still j-ascending */
                    double acc[4];
                    canon_dot4(q, (const char *)X + (size_t)r * d * esz, d, dtype, acc);
                    for (int u = 0; u < 4; ++u) list_insert(bs, bi, cnt, k, (float)acc[u], id_base + r + u);
                }
                for (; r < r1; ++r)
                    list_insert(bs, bi, cnt, k, (float)canon_dot(q, (const char *)X + (size_t)r * d * esz, d, dtype),
                                id_base + r);
            }
        }
    }
    for (int32_t qi = 0; qi < nq; ++qi) {
        float *bs = out_scores + (size_t)qi * k;
        int64_t *bi = out_ids + (size_t)qi * k;
        int32_t cnt = 0;
        for (int64_t t = 0; t < n_lists; ++t) {
            const int32_t c = lc[(size_t)t * nq + qi];
            for (int32_t j = 0; j < c; ++j)
                list_insert(bs, bi, &cnt, k, ls[((size_t)t * nq + qi) * k + j], li[((size_t)t * nq + qi) * k + j]);
        }
        for (int32_t p = cnt; p < k; ++p) {
            bs[p] = -INFINITY;
            bi[p] = -1;
        }
    }
    free(ls); free(li); free(lc);
    return 0;
}

/*
 * Canonical re-scoring of given (query, row) candidate pairs: cand_ids[nq * m] (row ids, -1 = skip).
 */
int ahro_rescore(const void *X, const void *Q, int32_t d, int32_t nq, int dtype, int32_t m,
                 const int64_t *cand_ids, float *out_scores) {
    const size_t esz = dtype == 0 ? 4 : 2;
#pragma omp parallel for schedule(static)
    for (int32_t qi = 0; qi < nq; ++qi) {
        const char *q = (const char *)Q + (size_t)qi * d * esz;
        for (int32_t c = 0; c < m; ++c) {
            int64_t r = cand_ids[(size_t)qi * m + c];
            out_scores[(size_t)qi * m + c] =
                r < 0 ? -INFINITY : (float)canon_dot(q, (const char *)X + (size_t)r * d * esz, d, dtype);
        }
    }
    return 0;
}

/*
 * Python's round(x, 4) on a double (CPython floatobject.c double_round: correctly rounded decimal,
 * ties-to-even on the exact binary value, then nearest double) restated with exact arithmetic:
 * x * 1e4 = p + e exactly (fma residual), n = round-half-even(p + e), result = n / 1e4 (correctly
 * rounded division == strtod of the decimal string).  Mirrors the device routine; checked against
 * CPython in tests/test_oracle.py.
 */
double ahro_round4(double x) {
    if (!isfinite(x) || fabs(x) >= 4.0e11) return x;
    double p = x * 1e4;
    double e = fma(x, 1e4, -p);
    double r = rint(p);          /* ties-to-even on the ROUNDED product p */
    double d0 = p - r;           /* exact, in [-0.5, 0.5] */
    /* p and every half-integer are multiples of ulp(p) and |e| <= ulp(p)/2, so the exact product can
       only sit on the other side of a rounding boundary when p is exactly on it; e breaks that tie. */
    if (d0 == 0.5 && e > 0.0) r += 1.0;
    else if (d0 == -0.5 && e < 0.0) r -= 1.0;
    return r / 1e4;
}

void ahro_round4_array(const double *x, int64_t n, double *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = ahro_round4(x[i]);
}
