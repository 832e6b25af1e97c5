// graph.cu -- K2: hop expansion + score aggregation over the device-resident CSR of the hierarchical
// knowledge graph.  Integer / fp64 work, bit-exact against the reference:
//
//   anchor_rerank_kernel : src/ah_rag/graph/hierarchical_graph.py:719-789 for a batch of queries
//                          (seed -> 1-hop belongs_to expansion with first-writer-wins dedup -> linear
//                          rerank in fp64 -> Python round(., 4) -> stable sort on the rounded score)
//   expand_kernel        : src/ah_rag/agent/environment.py:393-480 for a batch of environments
//                          (ordered concat of typed neighbour lists, stop rule on the raw length,
//                          order-preserving dedup, cut to limit; expand_related walks 2 hops)
//
// The CSR holds, per (edge type, direction), neighbours in the adjacency iteration order of the LOADED
// NetworkX graph, so "first occurrence in reference order" == "lowest position in the generated list".
#include "common.cuh"
#include "kernels.h"
#include "graph_shared.cuh"

namespace ahr {

// Python's round(x, 4): correctly rounded decimal of the exact binary value (ties-to-even), then the
// nearest double.  x * 1e4 = p + e exactly (FMA residual); p and every half-integer are multiples of
// ulp(p) and |e| <= ulp(p)/2, so e only matters when p sits exactly on a rounding boundary.
// Twin of ahro_round4 in oracle/exact_topk.c (checked against CPython there).
__device__ __forceinline__ double round4_py(double x) {
  if (!isfinite(x) || fabs(x) >= 4.0e11) return x;
  const double p = __dmul_rn(x, 1e4);
  const double e = __fma_rn(x, 1e4, -p);
  double r = rint(p);
  const double d0 = __dsub_rn(p, r);
  if (d0 == 0.5 && e > 0.0) r += 1.0;
  else if (d0 == -0.5 && e < 0.0) r -= 1.0;
  return __ddiv_rn(r, 1e4);
}

__global__ void round4_kernel(const double *__restrict__ x, int64_t n, double *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = round4_py(x[i]);
}

__global__ void pack_nodes_kernel(int64_t n, const uint8_t *__restrict__ node_type, const uint8_t *__restrict__ level,
                                  const double *__restrict__ judge, const double *__restrict__ judge_term,
                                  const double *__restrict__ conf, NodeRec *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  NodeRec r;
  r.judge = judge[i]; r.judge_term = judge_term[i]; r.conf = conf[i];
  r.type = node_type[i]; r.level = level[i];
  for (int j = 0; j < 6; ++j) r.pad[j] = 0;
  out[i] = r;
}

int graph_pack_nodes(int64_t n_nodes, const uint8_t *node_type, const uint8_t *level, const double *judge,
                     const double *judge_term, const double *conf, NodeRec *out, cudaStream_t stream) {
  if (n_nodes <= 0) return AHR_OK;
  pack_nodes_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, stream>>>(n_nodes, node_type, level, judge, judge_term, conf, out);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

int debug_round4(const double *x, int64_t n, double *out, cudaStream_t stream) {
  if (n <= 0) return AHR_OK;
  round4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, n, out);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// ---------------------------------------------------------------------------------------------
// anchor expansion + rerank: one CTA per query
// ---------------------------------------------------------------------------------------------
constexpr int AR_THREADS = 128;
constexpr int AR_SMEM_CAP = 1024;  // entries kept in shared memory; larger maps use the global workspace

enum : uint8_t { EK_ABSENT = 0, EK_SEED = AHR_EK_SEED, EK_PARENT = AHR_EK_PARENT, EK_MEMBER = AHR_EK_MEMBER };

struct AnchorArgs {
  GraphView g;
  int32_t nq, n_seeds, top_k, fan, cap;  // fan = max(parent_cap, member_top_m); cap = n_seeds * (1 + fan)
  ahr_rerank_params prm;
  const int32_t *seed_nodes;
  const float *seed_scores;
  int32_t *out_nodes;
  double *out_sem;
  double *out_score;
  int32_t *out_count;
  int32_t *exp_nodes;
  double *exp_sem;
  uint8_t *exp_kind;
  int32_t *exp_count;
  double *seed_sem;
  uint8_t *workspace;  // per query: cap * (4 + 8 + 8 + 1 + 1) bytes when cap > AR_SMEM_CAP
};

__host__ __device__ inline size_t anchor_ws_per_query(int cap) { return ((size_t)cap * 22 + 15) & ~(size_t)15; }

// General version: any capacity (the expanded map lives in the global workspace when it outgrows shared memory).
__global__ void __launch_bounds__(AR_THREADS) anchor_rerank_big_kernel(AnchorArgs a) {
  __shared__ double sm_sem[AR_SMEM_CAP];
  __shared__ double sm_rs[AR_SMEM_CAP];
  __shared__ int32_t sm_node[AR_SMEM_CAP];
  __shared__ uint8_t sm_kind[AR_SMEM_CAP];
  __shared__ uint8_t sm_flag[AR_SMEM_CAP];
  __shared__ int sm_counts[2];

  const int q = blockIdx.x;
  const int cap = a.cap, fan = a.fan, per = 1 + fan;
  double *e_sem, *e_rs;
  int32_t *e_node;
  uint8_t *e_kind, *e_flag;  // flag: bit0 = first occurrence, bit1 = passes filters
  if (cap <= AR_SMEM_CAP) {
    e_sem = sm_sem; e_rs = sm_rs; e_node = sm_node; e_kind = sm_kind; e_flag = sm_flag;
  } else {
    uint8_t *w = a.workspace + (size_t)q * anchor_ws_per_query(cap);
    e_sem = reinterpret_cast<double *>(w);
    e_rs = e_sem + cap;
    e_node = reinterpret_cast<int32_t *>(e_rs + cap);
    e_kind = reinterpret_cast<uint8_t *>(e_node + cap);
    e_flag = e_kind + cap;
  }
  const GraphView &g = a.g;
  const int32_t *seeds = a.seed_nodes + (int64_t)q * a.n_seeds;
  const float *sscore = a.seed_scores + (int64_t)q * a.n_seeds;

  // ---- A. generate the candidate list in reference order: seed r, then its expansions ----
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    const int r = e / per, j = e % per;
    int32_t v = seeds[r];
    if (v >= g.n_nodes) v = -1;
    int32_t node = -1;
    double sem = 0.0;
    uint8_t kind = EK_ABSENT;
    if (v >= 0) {
      // hnswlib's cosine space reports dist = 1.0f - <q,x> (fp32); the reference then takes
      // semantic = 1.0 - float(dist) in fp64 (hierarchical_graph.py:723)
      const double sem_r = 1.0 - (double)__fsub_rn(1.0f, sscore[r]);
      if (j == 0) {
        node = v; sem = sem_r; kind = EK_SEED;
        if (a.seed_sem) a.seed_sem[(int64_t)q * a.n_seeds + r] = sem_r;
      } else {
        const uint8_t t = g.rec[v].type;
        if (t == AHR_NT_ENTITY && j - 1 < a.prm.parent_cap) {  // first 2 out-belongs_to parents, x0.9 (:733-736)
          const int64_t p0 = g.row_ptr[AHR_CSR_BEL_OUT][v], p1 = g.row_ptr[AHR_CSR_BEL_OUT][v + 1];
          if (p0 + (j - 1) < p1) {
            node = g.col_idx[AHR_CSR_BEL_OUT][p0 + (j - 1)];
            sem = __dmul_rn(sem_r, 0.9);
            kind = EK_PARENT;
          }
        } else if (t == AHR_NT_SUMMARY && j - 1 < a.prm.member_top_m) {  // first m in-belongs_to members, x0.85 (:737-740)
          const int64_t p0 = g.row_ptr[AHR_CSR_BEL_IN][v], p1 = g.row_ptr[AHR_CSR_BEL_IN][v + 1];
          if (p0 + (j - 1) < p1) {
            node = g.col_idx[AHR_CSR_BEL_IN][p0 + (j - 1)];
            sem = __dmul_rn(sem_r, 0.85);
            kind = EK_MEMBER;
          }
        }
      }
    } else if (j == 0 && a.seed_sem) {
      a.seed_sem[(int64_t)q * a.n_seeds + r] = 0.0;
    }
    e_node[e] = node; e_sem[e] = sem; e_kind[e] = kind; e_flag[e] = 0;
  }
  if (threadIdx.x < 2) sm_counts[threadIdx.x] = 0;
  __syncthreads();

  // ---- B. first-writer-wins dedup; a seed always owns the value of its node (:732) ----
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    const int32_t node = e_node[e];
    if (node < 0) continue;
    bool first = true;
    for (int t = 0; t < e; ++t)
      if (e_node[t] == node) { first = false; break; }
    if (!first) continue;
    // the LAST seed entry carrying this node overwrites the value (dict assignment), position stays
    int seed_e = -1;
    for (int t = 0; t < cap; t += per)
      if (e_node[t] == node) seed_e = t;
    e_flag[e] = 1 | (seed_e >= 0 ? 4 : 0);
    if (seed_e >= 0) e_rs[e] = e_sem[seed_e]; else e_rs[e] = e_sem[e];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cap; e += AR_THREADS)
    if (e_flag[e] & 1) {
      e_sem[e] = e_rs[e];
      if (e_flag[e] & 4) e_kind[e] = EK_SEED;
    }
  __syncthreads();

  // ---- expanded map in insertion order (return_cluster=True) ----
  if (a.exp_nodes) {
    for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
      if (!(e_flag[e] & 1)) continue;
      int pos = 0;
      for (int t = 0; t < e; ++t) pos += e_flag[t] & 1;
      a.exp_nodes[(int64_t)q * cap + pos] = e_node[e];
      a.exp_sem[(int64_t)q * cap + pos] = e_sem[e];
      if (a.exp_kind) a.exp_kind[(int64_t)q * cap + pos] = e_kind[e];
      atomicAdd(&sm_counts[0], 1);
    }
  }

  // ---- C. filters + linear rerank in fp64, same operation order as :764-768, no FMA contraction ----
  const ahr_rerank_params &P = a.prm;
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    if (!(e_flag[e] & 1)) continue;
    const int32_t node = e_node[e];
    const NodeRec rec = g.rec[node];  // one sector: type, judge, judge_term, conf
    const uint8_t t = rec.type;
    const double judge = rec.judge, conf = rec.conf;
    const bool has_j = judge == judge, has_c = conf == conf;  // NaN encodes None
    bool ok = true;
    if (P.type_mask != 0 && !((P.type_mask >> t) & 1)) ok = false;                       // :756-757
    if (P.has_judge_min && (!has_j || judge < P.judge_min)) ok = false;                  // :759-760
    if (P.has_conf_min && (!has_c || conf < P.conf_min)) ok = false;                     // :761-762
    if (!ok) continue;
    const double jt = has_j ? rec.judge_term : 0.0;
    const double ct = has_c ? __ddiv_rn(conf, 10.0) : 0.0;
    double score = __dmul_rn(P.alpha, e_sem[e]);
    score = __dadd_rn(score, __dmul_rn(P.beta, jt));
    score = __dadd_rn(score, __dmul_rn(P.gamma, ct));
    score = __dadd_rn(score, __dmul_rn(P.delta, P.boost[t & 3]));
    e_rs[e] = round4_py(score);
    e_flag[e] |= 2;
  }
  __syncthreads();

  // ---- D. stable descending sort on the ROUNDED score (:781), keep top_k ----
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    if (!(e_flag[e] & 2)) continue;
    const double rs = e_rs[e];
    int rank = 0;
    for (int t = 0; t < cap; ++t) {
      if (!(e_flag[t] & 2) || t == e) continue;
      const double o = e_rs[t];
      if (o > rs || (o == rs && t < e)) ++rank;
    }
    atomicAdd(&sm_counts[1], 1);
    if (rank < a.top_k) {
      const int64_t o = (int64_t)q * a.top_k + rank;
      a.out_nodes[o] = e_node[e];
      a.out_sem[o] = round4_py(e_sem[e]);
      a.out_score[o] = rs;
    }
  }
  __syncthreads();
  const int n_valid = sm_counts[1];
  const int n_out = n_valid < a.top_k ? n_valid : a.top_k;
  for (int j = n_out + threadIdx.x; j < a.top_k; j += AR_THREADS) {
    const int64_t o = (int64_t)q * a.top_k + j;
    a.out_nodes[o] = -1; a.out_sem[o] = 0.0; a.out_score[o] = 0.0;
  }
  if (threadIdx.x == 0) {
    a.out_count[q] = n_out;
    if (a.exp_count) a.exp_count[q] = sm_counts[0];
  }
}

// Block-wide exclusive prefix sum of one int per thread (AR_THREADS = 4 warps); returns the block total in `total`.
__device__ __forceinline__ int block_exclusive_scan(int v, int *sm_warp /*[4]*/, int &total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, d);
    if (lane >= d) x += y;
  }
  __syncthreads();  // sm_warp may still be read from a previous scan
  if (lane == 31) sm_warp[w] = x;
  __syncthreads();
  int base = 0;
  total = 0;
#pragma unroll
  for (int i = 0; i < AR_THREADS / 32; ++i) {
    const int t = sm_warp[i];
    if (i < w) base += t;
    total += t;
  }
  return base + x - v;
}

// Fast path (expanded map fits shared memory: cap <= AR_SMEM_CAP, i.e. every configuration the reference's callers
// use).  Same results as anchor_rerank_big_kernel, O(cap) work for the map instead of O(cap^2):
//   dedup        open-addressing hash table in shared memory: first occurrence = atomicMin of the entry index,
//                "a seed owns its node's value" = atomicMax over the seed entries carrying the node
//   positions    block prefix sums (insertion order of the expanded map, order of the filtered list)
//   stable sort  rank counting over the COMPACTED filtered list (the only O(n^2) step left, n <= cap)
// The O(cap^2) dedup/position loops of the first version made the kernel instruction-bound: 117 k warp instructions per
// query at 50 seeds (profiles/r1_k2_c5_ncu.md).
__global__ void __launch_bounds__(AR_THREADS) anchor_rerank_kernel(AnchorArgs a) {
  constexpr int HASH_MAX = 2 * AR_SMEM_CAP;
  __shared__ double sm_sem[AR_SMEM_CAP];
  __shared__ double sm_rs[AR_SMEM_CAP];
  __shared__ int32_t sm_node[AR_SMEM_CAP];
  __shared__ uint8_t sm_kind[AR_SMEM_CAP];
  __shared__ uint8_t sm_flag[AR_SMEM_CAP];
  __shared__ int32_t h_key[HASH_MAX];                       // node id, -1 = free
  __shared__ __align__(8) int32_t h_val[2 * HASH_MAX];      // [0, HASH_MAX): first entry, [HASH_MAX, ..): last seed entry
  __shared__ int sm_warp[AR_THREADS / 32];
  // after the dedup the hash values are dead: the compacted filtered list lives there
  double *c_rs = reinterpret_cast<double *>(h_val);         // [AR_SMEM_CAP] rounded scores, 8 KB
  int32_t *c_e = h_val + 2 * AR_SMEM_CAP;                   // [AR_SMEM_CAP] entry index, 4 KB
  int32_t *h_first = h_val, *h_seed = h_val + HASH_MAX;

  const int q = blockIdx.x;
  const int cap = a.cap, fan = a.fan, per = 1 + fan;
  double *e_sem = sm_sem, *e_rs = sm_rs;
  int32_t *e_node = sm_node;
  uint8_t *e_kind = sm_kind, *e_flag = sm_flag;
  const GraphView &g = a.g;
  const int32_t *seeds = a.seed_nodes + (int64_t)q * a.n_seeds;
  const float *sscore = a.seed_scores + (int64_t)q * a.n_seeds;
  int hsize = 64;
  while (hsize < 2 * cap) hsize <<= 1;  // <= HASH_MAX because cap <= AR_SMEM_CAP
  const int hmask = hsize - 1;
  for (int i = threadIdx.x; i < hsize; i += AR_THREADS) {
    h_key[i] = -1;
    h_first[i] = 0x7fffffff;
    h_seed[i] = -1;
  }

  // ---- A. generate the candidate list in reference order: seed r, then its expansions ----
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    const int r = e / per, j = e % per;
    int32_t v = seeds[r];
    if (v >= g.n_nodes) v = -1;
    int32_t node = -1;
    double sem = 0.0;
    uint8_t kind = EK_ABSENT;
    if (v >= 0) {
      // hnswlib's cosine space reports dist = 1.0f - <q,x> (fp32); the reference then takes
      // semantic = 1.0 - float(dist) in fp64 (hierarchical_graph.py:723)
      const double sem_r = 1.0 - (double)__fsub_rn(1.0f, sscore[r]);
      if (j == 0) {
        node = v; sem = sem_r; kind = EK_SEED;
        if (a.seed_sem) a.seed_sem[(int64_t)q * a.n_seeds + r] = sem_r;
      } else {
        const uint8_t t = g.rec[v].type;
        if (t == AHR_NT_ENTITY && j - 1 < a.prm.parent_cap) {  // first 2 out-belongs_to parents, x0.9 (:733-736)
          const int64_t p0 = g.row_ptr[AHR_CSR_BEL_OUT][v], p1 = g.row_ptr[AHR_CSR_BEL_OUT][v + 1];
          if (p0 + (j - 1) < p1) {
            node = g.col_idx[AHR_CSR_BEL_OUT][p0 + (j - 1)];
            sem = __dmul_rn(sem_r, 0.9);
            kind = EK_PARENT;
          }
        } else if (t == AHR_NT_SUMMARY && j - 1 < a.prm.member_top_m) {  // first m in-belongs_to members, x0.85 (:737-740)
          const int64_t p0 = g.row_ptr[AHR_CSR_BEL_IN][v], p1 = g.row_ptr[AHR_CSR_BEL_IN][v + 1];
          if (p0 + (j - 1) < p1) {
            node = g.col_idx[AHR_CSR_BEL_IN][p0 + (j - 1)];
            sem = __dmul_rn(sem_r, 0.85);
            kind = EK_MEMBER;
          }
        }
      }
    } else if (j == 0 && a.seed_sem) {
      a.seed_sem[(int64_t)q * a.n_seeds + r] = 0.0;
    }
    e_node[e] = node; e_sem[e] = sem; e_kind[e] = kind; e_flag[e] = 0;
  }
  __syncthreads();

  // ---- B. first-writer-wins dedup; a seed always owns the value of its node (:732) ----
  auto probe = [&](int32_t node, bool insert) {
    int slot = (int)(((uint32_t)node * 2654435761u) >> 7) & hmask;
    for (;;) {
      const int32_t k = insert ? atomicCAS(&h_key[slot], -1, node) : h_key[slot];
      if (k == node || (insert && k == -1)) return slot;
      slot = (slot + 1) & hmask;
    }
  };
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    const int32_t node = e_node[e];
    if (node < 0) continue;
    const int slot = probe(node, true);
    atomicMin(&h_first[slot], e);
    if (e % per == 0) atomicMax(&h_seed[slot], e);  // the LAST seed entry carrying this node overwrites the value
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    const int32_t node = e_node[e];
    if (node < 0) continue;
    const int slot = probe(node, false);
    if (h_first[slot] != e) continue;
    const int seed_e = h_seed[slot];
    e_flag[e] = 1 | (seed_e >= 0 ? 4 : 0);
    e_rs[e] = seed_e >= 0 ? e_sem[seed_e] : e_sem[e];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < cap; e += AR_THREADS)
    if (e_flag[e] & 1) {
      e_sem[e] = e_rs[e];
      if (e_flag[e] & 4) e_kind[e] = EK_SEED;
    }
  __syncthreads();

  // each thread owns a contiguous run of entries for the two order-preserving compactions below
  const int chunk = (cap + AR_THREADS - 1) / AR_THREADS;
  const int e_lo = min(cap, (int)threadIdx.x * chunk), e_hi = min(cap, e_lo + chunk);

  // ---- expanded map in insertion order (return_cluster=True) ----
  if (a.exp_nodes) {
    int mine = 0, total;
    for (int e = e_lo; e < e_hi; ++e) mine += e_flag[e] & 1;
    int pos = block_exclusive_scan(mine, sm_warp, total);
    for (int e = e_lo; e < e_hi; ++e) {
      if (!(e_flag[e] & 1)) continue;
      a.exp_nodes[(int64_t)q * cap + pos] = e_node[e];
      a.exp_sem[(int64_t)q * cap + pos] = e_sem[e];
      if (a.exp_kind) a.exp_kind[(int64_t)q * cap + pos] = e_kind[e];
      ++pos;
    }
    if (threadIdx.x == 0 && a.exp_count) a.exp_count[q] = total;
  } else if (threadIdx.x == 0 && a.exp_count) {
    a.exp_count[q] = 0;
  }

  // ---- C. filters + linear rerank in fp64, same operation order as :764-768, no FMA contraction ----
  const ahr_rerank_params &P = a.prm;
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    if (!(e_flag[e] & 1)) continue;
    const int32_t node = e_node[e];
    const NodeRec rec = g.rec[node];  // one sector: type, judge, judge_term, conf
    const uint8_t t = rec.type;
    const double judge = rec.judge, conf = rec.conf;
    const bool has_j = judge == judge, has_c = conf == conf;  // NaN encodes None
    bool ok = true;
    if (P.type_mask != 0 && !((P.type_mask >> t) & 1)) ok = false;                       // :756-757
    if (P.has_judge_min && (!has_j || judge < P.judge_min)) ok = false;                  // :759-760
    if (P.has_conf_min && (!has_c || conf < P.conf_min)) ok = false;                     // :761-762
    if (!ok) continue;
    const double jt = has_j ? rec.judge_term : 0.0;
    const double ct = has_c ? __ddiv_rn(conf, 10.0) : 0.0;
    double score = __dmul_rn(P.alpha, e_sem[e]);
    score = __dadd_rn(score, __dmul_rn(P.beta, jt));
    score = __dadd_rn(score, __dmul_rn(P.gamma, ct));
    score = __dadd_rn(score, __dmul_rn(P.delta, P.boost[t & 3]));
    e_rs[e] = round4_py(score);
    e_flag[e] |= 2;
  }
  __syncthreads();

  // ---- D. stable descending sort on the ROUNDED score (:781), keep top_k ----
  int mine = 0, n_valid;
  for (int e = e_lo; e < e_hi; ++e) mine += (e_flag[e] >> 1) & 1;
  int pos = block_exclusive_scan(mine, sm_warp, n_valid);  // its barriers also retire the last reads of the hash
  for (int e = e_lo; e < e_hi; ++e) {
    if (!(e_flag[e] & 2)) continue;
    c_rs[pos] = e_rs[e];
    c_e[pos] = e;
    ++pos;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_valid; i += AR_THREADS) {
    const double rs = c_rs[i];
    int rank = 0;
#pragma unroll 4
    for (int t = 0; t < n_valid; ++t) {
      const double o = c_rs[t];
      rank += (o > rs || (o == rs && t < i)) ? 1 : 0;  // list order = entry order: ties keep insertion order
    }
    if (rank < a.top_k) {
      const int e = c_e[i];
      const int64_t o = (int64_t)q * a.top_k + rank;
      a.out_nodes[o] = e_node[e];
      a.out_sem[o] = round4_py(e_sem[e]);
      a.out_score[o] = rs;
    }
  }
  const int n_out = n_valid < a.top_k ? n_valid : a.top_k;
  for (int j = n_out + threadIdx.x; j < a.top_k; j += AR_THREADS) {
    const int64_t o = (int64_t)q * a.top_k + j;
    a.out_nodes[o] = -1; a.out_sem[o] = 0.0; a.out_score[o] = 0.0;
  }
  if (threadIdx.x == 0) a.out_count[q] = n_out;
}

size_t graph_anchor_workspace_bytes(int32_t nq, int32_t n_seeds, const ahr_rerank_params &prm) {
  const int fan = std::max(prm.parent_cap, prm.member_top_m);
  const int64_t cap = (int64_t)n_seeds * (1 + std::max(fan, 0));
  if (cap <= AR_SMEM_CAP) return 0;
  return (size_t)nq * anchor_ws_per_query((int)cap);
}

int graph_anchor_rerank(const GraphView &g, int32_t nq, int32_t n_seeds, const int32_t *seed_nodes,
                        const float *seed_scores, const ahr_rerank_params &prm, int32_t top_k, int32_t *out_nodes,
                        double *out_sem, double *out_score, int32_t *out_count, int32_t *exp_nodes, double *exp_sem,
                        uint8_t *exp_kind, int32_t *exp_count, double *seed_sem, void *workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
  AHR_REQUIRE(nq >= 0 && n_seeds >= 1 && top_k >= 1, "anchor_rerank: bad sizes nq=%d n_seeds=%d top_k=%d", nq, n_seeds, top_k);
  AHR_REQUIRE(prm.parent_cap >= 0 && prm.member_top_m >= 0, "anchor_rerank: negative caps");
  if (nq == 0) return AHR_OK;
  AnchorArgs a;
  a.g = g;
  a.nq = nq; a.n_seeds = n_seeds; a.top_k = top_k;
  a.fan = std::max(prm.parent_cap, prm.member_top_m);
  const int64_t cap = (int64_t)n_seeds * (1 + a.fan);
  AHR_REQUIRE(cap <= (1 << 20), "anchor_rerank: expanded map capacity %lld too large", (long long)cap);
  a.cap = (int)cap;
  a.prm = prm;
  a.seed_nodes = seed_nodes; a.seed_scores = seed_scores;
  a.out_nodes = out_nodes; a.out_sem = out_sem; a.out_score = out_score; a.out_count = out_count;
  a.exp_nodes = exp_nodes; a.exp_sem = exp_sem; a.exp_count = exp_count; a.seed_sem = seed_sem;
  a.exp_kind = exp_kind;
  a.workspace = reinterpret_cast<uint8_t *>(workspace);
  if (cap > AR_SMEM_CAP)
    AHR_REQUIRE(workspace && workspace_bytes >= (size_t)nq * anchor_ws_per_query((int)cap), "anchor_rerank: workspace too small");
  if (cap <= AR_SMEM_CAP) anchor_rerank_kernel<<<nq, AR_THREADS, 0, stream>>>(a);
  else anchor_rerank_big_kernel<<<nq, AR_THREADS, 0, stream>>>(a);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// ---------------------------------------------------------------------------------------------
// frontier expansion: one warp per environment
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EXP_WARPS * 32)
expand_kernel(GraphView g, int32_t n_env, const int32_t *__restrict__ ops, const int32_t *__restrict__ in_ids,
              const int32_t *__restrict__ in_offsets, int32_t limit, int hash_slots, int32_t *__restrict__ out_ids,
              int32_t *__restrict__ out_counts) {
  extern __shared__ int32_t sm_out[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * EXP_WARPS + w;
  if (env >= n_env) return;
  ExpandState st;
  st.out = sm_out + (size_t)w * limit;
  st.count = 0; st.raw = 0; st.limit = limit;
  if (hash_slots > 0) {  // large limits: a hash set instead of scanning the list for every candidate
    st.hash = sm_out + (size_t)EXP_WARPS * limit + (size_t)w * hash_slots;
    st.hmask = hash_slots - 1;
    for (int j = lane; j < hash_slots; j += 32) st.hash[j] = -1;
    __syncwarp();
  }
  expand_walk(st, g, ops[env], in_ids + in_offsets[env], in_offsets[env + 1] - in_offsets[env], lane);
  for (int j = lane; j < limit; j += 32) out_ids[(int64_t)env * limit + j] = j < st.count ? st.out[j] : -1;
  if (lane == 0) out_counts[env] = st.count;
}

int graph_expand(const GraphView &g, int32_t n_env, const int32_t *ops, const int32_t *in_ids,
                 const int32_t *in_offsets, int32_t limit, int32_t *out_ids, int32_t *out_counts, cudaStream_t stream) {
  AHR_REQUIRE(limit >= 1 && limit <= 2048, "expand: limit must be in [1, 2048] (got %d)", limit);
  if (n_env <= 0) return AHR_OK;
  // the list can hold `limit` values but a batch of up to 32 is inserted before the cut: 2 * (limit + 32) slots, power of 2
  int hash_slots = 0;
  if (limit > 64) {
    hash_slots = 128;
    while (hash_slots < 2 * (limit + 32)) hash_slots <<= 1;
  }
  const size_t smem = (size_t)EXP_WARPS * (limit + hash_slots) * sizeof(int32_t);
  if (smem > 48 * 1024) {
    static std::atomic<uint64_t> configured{0};
    int dev = 0;
    AHR_CUDA_CHECK(cudaGetDevice(&dev));
    if (!((configured.load(std::memory_order_relaxed) >> (dev & 63)) & 1)) {
      AHR_CUDA_CHECK(cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 168 * 1024));
      configured.fetch_or(1ull << (dev & 63), std::memory_order_relaxed);
    }
  }
  expand_kernel<<<(n_env + EXP_WARPS - 1) / EXP_WARPS, EXP_WARPS * 32, smem, stream>>>(g, n_env, ops, in_ids, in_offsets,
                                                                                      limit, hash_slots, out_ids, out_counts);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

}  // namespace ahr
