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
                                  const double *__restrict__ conf, const int64_t *__restrict__ rp_bel_out,
                                  const int32_t *__restrict__ ci_bel_out, const int64_t *__restrict__ rp_bel_in,
                                  const int32_t *__restrict__ ci_bel_in, NodeLine *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  NodeLine r;
  r.head.judge = judge[i]; r.head.judge_term = judge_term[i]; r.head.conf = conf[i];
  r.head.type = node_type[i]; r.head.level = level[i];
  r.head.pad[0] = r.head.pad[1] = 0;
  // anchor list: what a seed of this type expands into (hierarchical_graph.py:733-740)
  const int64_t *rp = r.head.type == AHR_NT_ENTITY ? rp_bel_out : (r.head.type == AHR_NT_SUMMARY ? rp_bel_in : nullptr);
  const int32_t *ci = r.head.type == AHR_NT_ENTITY ? ci_bel_out : ci_bel_in;
  int64_t p0 = 0, deg = 0;
  if (rp != nullptr) { p0 = rp[i]; deg = rp[i + 1] - p0; }
  r.head.anchor_deg = (int32_t)(deg > INT32_MAX ? INT32_MAX : deg);
  for (int j = 0; j < NODE_ANCHOR_INLINE; ++j) r.anchor[j] = j < deg ? ci[p0 + j] : -1;
  out[i] = r;
}

int graph_pack_nodes(int64_t n_nodes, const uint8_t *node_type, const uint8_t *level, const double *judge,
                     const double *judge_term, const double *conf, const int64_t *const row_ptr[AHR_N_CSR],
                     const int32_t *const col_idx[AHR_N_CSR], NodeLine *out, cudaStream_t stream) {
  if (n_nodes <= 0) return AHR_OK;
  pack_nodes_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, stream>>>(
      n_nodes, node_type, level, judge, judge_term, conf, row_ptr[AHR_CSR_BEL_OUT], col_idx[AHR_CSR_BEL_OUT],
      row_ptr[AHR_CSR_BEL_IN], col_idx[AHR_CSR_BEL_IN], out);
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
constexpr int AR_SMEM_CAP = 896;  // entries kept in shared memory (47.9 KB at 896: no opt-in); larger maps use the global workspace

enum : uint8_t { EK_ABSENT = 0, EK_SEED = AHR_EK_SEED, EK_PARENT = AHR_EK_PARENT, EK_MEMBER = AHR_EK_MEMBER };

struct AnchorArgs {
  GraphView g;
  int32_t nq, n_seeds, top_k, fan, cap;  // fan = max(parent_cap, member_top_m); cap = n_seeds * (1 + fan)
  int32_t hsize;                         // hash slots of the large-map variant (power of two >= 2 * cap)
  ahr_rerank_params prm;
  const int32_t *seed_nodes;             // dense node ids, or
  const int64_t *seed_rows;              // index rows as ahr_topk returns them (id_base = 0) + the row -> node table
  const int32_t *row_to_node;
  int64_t n_rows;
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
  uint8_t *workspace;  // per query anchor_ws_per_query(cap, hsize) bytes when cap > AR_SMEM_CAP
  AnchorGathered src;  // g1 != nullptr: the graph is sharded by node range, lines come as gathered stage records
};

__device__ __forceinline__ int anchor_owner(const AnchorGathered &s, int32_t v) {
  int r = 0;
  while (r + 1 < s.world && v >= s.bounds[r + 1]) ++r;
  return r;
}
// attributes of a record: [.., judge(2), judge_term(2), conf(2)] at int offset 4, 8-byte aligned
__device__ __forceinline__ NodeHead anchor_head_of(const int32_t *rec) {
  NodeHead h;
  const double *d = reinterpret_cast<const double *>(rec + 4);
  h.judge = d[0]; h.judge_term = d[1]; h.conf = d[2];
  h.anchor_deg = 0;
  h.type = (uint8_t)rec[1];
  h.level = (uint8_t)rec[3];
  h.pad[0] = h.pad[1] = 0;
  return h;
}

// large maps: sem[cap] rs[cap] c_rs[cap] (fp64) | node[cap] c_e[cap] h_key[hsize] h_first[hsize] h_seed[hsize] (int32) |
// kind[cap] flag[cap] (u8)
__host__ __device__ inline size_t anchor_ws_per_query(int cap, int hsize) {
  return (((size_t)cap * (3 * 8 + 2 * 4 + 2) + (size_t)hsize * 12) + 15) & ~(size_t)15;
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

// One kernel for every map size.  BIG = false: the expanded map (cap <= AR_SMEM_CAP entries, i.e. every configuration the
// reference's callers use) and its hash table live in shared memory; BIG = true: same code over the global workspace.
//   generation   entry (r, j): seed r (j = 0) or its j-th expansion.  ONE 64-byte line per seed (NodeLine: attributes + the
//                first 8 entries of the list a seed of that type expands into) replaces the chain
//                seed -> attributes -> row_ptr -> col_idx (r1: 63.5 MB of DRAM sectors for ~10 MB of useful bytes,
//                profiles/r1_k2_ncu.md); the CSR is only touched for expansions beyond the 8 inline entries.
//   dedup        open-addressing hash table: first occurrence = atomicMin of the entry index, "a seed owns its node's
//                value" (:732) = atomicMax over the seed entries carrying the node
//   positions    block prefix sums (insertion order of the expanded map, order of the filtered list)
//   rerank       fp64, the reference's operation order, one 32-byte sector (NodeHead) per emitted node
//   stable sort  rank counting over the COMPACTED filtered list (the only O(n^2) step, n <= cap)
// shared-memory footprint of the small-map variant: sem[cap] rs[cap] (fp64) | hval[2 * hsize] node[cap] hkey[hsize] (int32)
// | kind[cap] flag[cap] (u8).  Sized by the ACTUAL capacity: 50 seeds x 6 = 300 entries need 19 KB, not the 47 KB of a
// 1024-entry map, so 11 CTAs share an SM instead of 4 (the kernel is latency-bound on its dependent line loads).
__host__ __device__ inline size_t anchor_smem_bytes(int cap, int hsize) {
  // ... + the 32-bit sort keys of the compacted list (16-byte aligned, padded to whole int4 loads)
  return (((size_t)cap * (8 + 8 + 4 + 1 + 1) + (size_t)hsize * 12 + 15) & ~(size_t)15) + (size_t)(cap + 4) * 4 + 16;
}

template <bool BIG>
__global__ void __launch_bounds__(AR_THREADS) anchor_rerank_kernel(AnchorArgs a) {
  extern __shared__ __align__(16) uint8_t ar_smem[];
  __shared__ int sm_warp[AR_THREADS / 32];

  const int q = blockIdx.x;
  const int cap = a.cap, fan = a.fan, per = 1 + fan;
  const int hsize = a.hsize;
  double *e_sem, *e_rs, *c_rs;
  int32_t *e_node, *c_e, *h_key, *h_first, *h_seed, *c_key = nullptr;
  uint8_t *e_kind, *e_flag;  // flag: bit0 = first occurrence, bit1 = passes filters, bit2 = owned by a seed
  if (BIG) {
    uint8_t *w = a.workspace + (size_t)q * anchor_ws_per_query(cap, hsize);
    e_sem = reinterpret_cast<double *>(w);
    e_rs = e_sem + cap;
    c_rs = e_rs + cap;
    e_node = reinterpret_cast<int32_t *>(c_rs + cap);
    c_e = e_node + cap;
    h_key = c_e + cap;
    h_first = h_key + hsize;
    h_seed = h_first + hsize;
    e_kind = reinterpret_cast<uint8_t *>(h_seed + hsize);
    e_flag = e_kind + cap;
  } else {
    e_sem = reinterpret_cast<double *>(ar_smem);
    e_rs = e_sem + cap;
    h_first = reinterpret_cast<int32_t *>(e_rs + cap);  // [hsize] first entry carrying the node
    h_seed = h_first + hsize;                           // [hsize] last seed entry carrying the node
    e_node = h_seed + hsize;
    h_key = e_node + cap;                               // [hsize] node id, -1 = free
    e_kind = reinterpret_cast<uint8_t *>(h_key + hsize);
    e_flag = e_kind + cap;
    // after the dedup the hash values are dead: the compacted filtered list lives there (12 * cap <= 8 * hsize bytes)
    c_rs = reinterpret_cast<double *>(h_first);
    c_e = h_first + 2 * cap;
    c_key = reinterpret_cast<int32_t *>(ar_smem + ((((size_t)cap * (8 + 8 + 4 + 1 + 1) + (size_t)hsize * 12) + 15) & ~(size_t)15));
  }
  const GraphView &g = a.g;
  const float *sscore = a.seed_scores + (int64_t)q * a.n_seeds;
  const int hmask = hsize - 1;
  for (int i = threadIdx.x; i < hsize; i += AR_THREADS) {
    h_key[i] = -1;
    h_first[i] = 0x7fffffff;
    h_seed[i] = -1;
  }

  // ---- A. generate the candidate list in reference order: seed r, then its expansions ----
  for (int e = threadIdx.x; e < cap; e += AR_THREADS) {
    const int r = e / per, j = e % per;
    int32_t v;
    if (a.seed_rows != nullptr) {  // index row -> dense node id (the Chroma id -> G.nodes lookup of :731)
      const int64_t row = a.seed_rows[(int64_t)q * a.n_seeds + r];
      v = (row >= 0 && row < a.n_rows) ? a.row_to_node[row] : -1;
    } else {
      v = a.seed_nodes[(int64_t)q * a.n_seeds + r];
    }
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
      } else if (a.src.g1 != nullptr) {
        // sharded graph: the owner of v sent its type and the capped head of its anchor list (graph_shard.cu, stage 1)
        const int32_t *r1 = a.src.g1 + (((int64_t)anchor_owner(a.src, v) * a.nq + q) * a.n_seeds + r) * a.src.rec1;
        const int t = r1[1], i = j - 1;
        if (i < r1[2]) {  // r1[2] = min(cap of this node type, degree)
          node = r1[ANCHOR_REC_HEAD + i];
          sem = __dmul_rn(sem_r, t == AHR_NT_ENTITY ? 0.9 : 0.85);
          kind = t == AHR_NT_ENTITY ? EK_PARENT : EK_MEMBER;
        }
      } else {
        const NodeLine *L = g.line + v;
        const uint8_t t = L->head.type;
        // entity: first parent_cap out-belongs_to parents, x0.9 (:733-736); summary: first m members, x0.85 (:737-740)
        const int want = t == AHR_NT_ENTITY ? a.prm.parent_cap : (t == AHR_NT_SUMMARY ? a.prm.member_top_m : 0);
        const int i = j - 1;
        if (i < want && i < L->head.anchor_deg) {
          if (i < NODE_ANCHOR_INLINE) {
            node = L->anchor[i];
          } else {
            const int slot = t == AHR_NT_ENTITY ? AHR_CSR_BEL_OUT : AHR_CSR_BEL_IN;
            node = g.col_idx[slot][g.row_ptr[slot][v] + i];
          }
          sem = __dmul_rn(sem_r, t == AHR_NT_ENTITY ? 0.9 : 0.85);
          kind = t == AHR_NT_ENTITY ? EK_PARENT : EK_MEMBER;
        }
      }
    } else if (j == 0 && a.seed_sem) {
      a.seed_sem[(int64_t)q * a.n_seeds + r] = 0.0;
    }
    AHR_DEV_ASSERT(node < g.n_nodes);
    e_node[e] = node; e_sem[e] = sem; e_kind[e] = kind; e_flag[e] = 0;
  }
  __syncthreads();

  // ---- B. first-writer-wins dedup; a seed always owns the value of its node (:732) ----
  auto probe = [&](int32_t node, bool insert) {
    int slot = (int)(((uint32_t)node * 2654435761u) >> 7) & hmask;
    int probes = 0;
    for (;;) {
      AHR_DEV_ASSERT(slot >= 0 && slot < hsize && ++probes <= hsize);
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
      AHR_DEV_ASSERT(pos >= 0 && pos < cap);
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
    NodeHead rec;
    if (a.src.g1 != nullptr) {  // sharded graph: the attributes travelled with the stage records
      const int r = e / per, j = e % per;
      rec = j == 0 ? anchor_head_of(a.src.g1 + (((int64_t)anchor_owner(a.src, node) * a.nq + q) * a.n_seeds + r) * a.src.rec1)
                   : anchor_head_of(a.src.g2 + ((((int64_t)anchor_owner(a.src, node) * a.nq + q) * a.n_seeds + r) * fan + (j - 1)) * ANCHOR_REC_HEAD);
    } else {
      rec = g.line[node].head;  // one sector: type, judge, judge_term, conf
    }
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
  // Rank counting over the compacted list (n = 100-300 entries, the kernel's only O(n^2) step: 53 % of its instructions
  // when the comparison was (double >, double ==, index <), profiles/r2_k2_ncu.md).  A rounded score is r / 1e4 with r an
  // integer; while every |r| < 2^20 -- scores below 104, i.e. always with weights like the reference's -- the whole sort
  // key (r descending, list position ascending) fits one 32-bit integer and a comparison is one ISETP on an int4 of keys;
  // anything else (huge weights, NaN) takes the general comparison.  Same total order either way (-0.0 == 0.0 included).
  int mine = 0, n_valid;
  for (int e = e_lo; e < e_hi; ++e) mine += (e_flag[e] >> 1) & 1;
  int pos = block_exclusive_scan(mine, sm_warp, n_valid);  // its barriers also retire the last reads of the hash
  bool regular = !BIG;
  for (int e = e_lo; e < e_hi; ++e) {
    if (!(e_flag[e] & 2)) continue;
    AHR_DEV_ASSERT(pos >= 0 && pos < cap);
    const double rs = e_rs[e];
    c_rs[pos] = rs;
    c_e[pos] = e;
    if (!BIG) {
      const double r = rint(__dmul_rn(rs, 1e4));  // exact: |fl(r / 1e4) * 1e4 - r| <= |r| * 2^-52
      const bool fits = fabs(r) < 1048576.0;      // (false for NaN)
      regular = regular && fits;
      c_key[pos] = fits ? ((((int)r + 1048576) << 10) | (AR_SMEM_CAP - 1 - pos)) : 0;
    }
    ++pos;
  }
  if (!BIG)
    for (int p = n_valid + (int)threadIdx.x; p < ((n_valid + 3) & ~3); p += AR_THREADS) c_key[p] = INT32_MIN;  // never greater
  const bool fast = __syncthreads_and(regular) != 0;
  for (int i = threadIdx.x; i < n_valid; i += AR_THREADS) {
    const double rs = c_rs[i];
    int rank = 0;
    if (fast) {
      const int ki = c_key[i];
#pragma unroll 2
      for (int t = 0; t < n_valid; t += 4) {
        const int4 kk = *reinterpret_cast<const int4 *>(c_key + t);
        rank += (kk.x > ki) + (kk.y > ki) + (kk.z > ki) + (kk.w > ki);
      }
    } else {
#pragma unroll 4
      for (int t = 0; t < n_valid; ++t) {
        const double o = c_rs[t];
        rank += (o > rs || (o == rs && t < i)) ? 1 : 0;  // list order = entry order: ties keep insertion order
      }
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

static int anchor_hash_slots(int64_t cap) {
  int h = 64;
  while (h < 2 * cap) h <<= 1;
  return h;
}

size_t graph_anchor_workspace_bytes(int32_t nq, int32_t n_seeds, const ahr_rerank_params &prm) {
  const int fan = std::max(prm.parent_cap, prm.member_top_m);
  const int64_t cap = (int64_t)n_seeds * (1 + std::max(fan, 0));
  if (cap <= AR_SMEM_CAP) return 0;
  return (size_t)nq * anchor_ws_per_query((int)cap, anchor_hash_slots(cap));
}

int graph_anchor_rerank(const GraphView &g, int32_t nq, int32_t n_seeds, const int32_t *seed_nodes,
                        const int64_t *seed_rows, const int32_t *row_to_node, int64_t n_rows,
                        const float *seed_scores, const ahr_rerank_params &prm, int32_t top_k, int32_t *out_nodes,
                        double *out_sem, double *out_score, int32_t *out_count, int32_t *exp_nodes, double *exp_sem,
                        uint8_t *exp_kind, int32_t *exp_count, double *seed_sem, void *workspace,
                        size_t workspace_bytes, cudaStream_t stream, const AnchorGathered *gathered) {
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
  a.hsize = anchor_hash_slots(cap);
  a.prm = prm;
  a.seed_nodes = seed_nodes; a.seed_rows = seed_rows; a.row_to_node = row_to_node; a.n_rows = n_rows;
  a.seed_scores = seed_scores;
  a.out_nodes = out_nodes; a.out_sem = out_sem; a.out_score = out_score; a.out_count = out_count;
  a.exp_nodes = exp_nodes; a.exp_sem = exp_sem; a.exp_count = exp_count; a.seed_sem = seed_sem;
  a.exp_kind = exp_kind;
  a.workspace = reinterpret_cast<uint8_t *>(workspace);
  a.src = gathered ? *gathered : AnchorGathered{};
  if (cap > AR_SMEM_CAP) {
    AHR_REQUIRE(workspace && workspace_bytes >= (size_t)nq * anchor_ws_per_query((int)cap, a.hsize), "anchor_rerank: workspace too small");
    anchor_rerank_kernel<true><<<nq, AR_THREADS, 0, stream>>>(a);
  } else {
    anchor_rerank_kernel<false><<<nq, AR_THREADS, anchor_smem_bytes(a.cap, a.hsize), stream>>>(a);  // <= 47.9 KB
  }
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

// ---------------------------------------------------------------------------------------------
// frontier expansion, reference-sized limits (<= 64): one warp per environment
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(EXP_WARPS * 32)
expand_kernel(GraphView g, int32_t n_env, const int32_t *__restrict__ ops, const int32_t *__restrict__ in_ids,
              const int32_t *__restrict__ in_offsets, int32_t limit, int32_t *__restrict__ out_ids,
              int32_t *__restrict__ out_counts) {
  extern __shared__ int32_t sm_out[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * EXP_WARPS + w;
  if (env >= n_env) return;
  ExpandState st;
  st.out = sm_out + (size_t)w * limit;
  st.count = 0; st.raw = 0; st.limit = limit;
  expand_walk(st, g, ops[env], in_ids + in_offsets[env], in_offsets[env + 1] - in_offsets[env], lane);
  for (int j = lane; j < limit; j += 32) out_ids[(int64_t)env * limit + j] = j < st.count ? st.out[j] : -1;
  if (lane == 0) out_counts[env] = st.count;
}

// ---------------------------------------------------------------------------------------------
// frontier expansion, large limits (65-2048): one WARP per environment, hash-set dedup WITHOUT atomics.
//
// History (profiles/r2_k2_ncu.md): the r1 warp walker scanned / probed per batch with two votes and a match (0.52 ms for
// the 273 MB bandwidth view on c5); the first r2 version was edge-parallel over a 64-thread CTA with one 64-bit
// compare-and-swap per neighbour: 0.38 ms -- 46 % of all instructions in the probe loop around the CAS, and shared-memory
// atomics are the slow path of the LSU (B300_MICROARCH.md "ATOMS": ~2 cycles per lane).  Plain shared loads / stores are
// far cheaper on that pipe, and one warp owning the table needs no atomicity, only agreement:
//   per batch of 32 consecutive positions of the reference's concatenation (coalesced col_idx loads, U batches in flight)
//   probe    every lane walks its value's probe sequence with plain loads until it meets the value (seen in an earlier
//            batch: duplicate) or an empty slot;
//   claim    lanes at an empty slot store their value, __syncwarp, read back: the lane whose value stands has inserted
//            it, the others (a different value won the slot) probe on.  Lanes carrying the SAME value see the same table
//            and therefore travel together: they all "win"; at the end of the batch a byte written to owner[slot] by
//            the inserting lanes tells them apart, and only then (never on a DiGraph's duplicate-free adjacency rows,
//            but the ABI accepts any CSR) a __match_any -- 64 cycles of a shared unit per use, tools/ub/ub_match.cu --
//            makes the lowest lane the first occurrence;
//   emit     ballot of the fresh lanes -> output index = emitted so far + prefix popcount; ids go straight to global.
// Batches of one warp are sequential, so "first occurrence" across batches needs no positions in the table: 4 bytes of
// value + 1 byte of owner per slot (5 KB per environment at limit 512: ~44 environments in flight per SM).
// The stop rule of the reference ("un-deduplicated length >= limit after finishing an input node", environment.py:404-406)
// and the cut at `limit` unique values are those of expand_walk (graph_shared.cuh).
// ---------------------------------------------------------------------------------------------
constexpr int EW_UNROLL = 4;  // batches of 32 neighbours loaded ahead of the inserts

__host__ __device__ inline int ew_table_slots(int values) {  // open addressing at a load factor <= 0.75
  int t = 64;
  while (3 * t < 4 * values) t <<= 1;
  return t;
}
// environments (warps) per CTA: 20 KB of tables per CTA at limit 512, 40 KB at limits 1024 / 2048 (no opt-in needed)
__host__ __device__ inline int ew_warps(int limit) { return limit <= 1024 ? 4 : 2; }
__host__ __device__ inline size_t ew_smem_bytes(int limit) { return (size_t)ew_warps(limit) * ew_table_slots(limit + 32) * 5; }

struct WideState {
  uint32_t tbase;   // shared-memory byte address of table[T] (int32 value, -1 = empty)
  uint32_t obase;   // shared-memory byte address of owner[T] (u8: lane that inserted the value; meaningful within the batch)
  uint32_t mask4;   // (T - 1) * 4: byte offsets into the table wrap with one AND
  int tshift2;      // byte offset of the home slot = (hash >> tshift2) & mask4
  int32_t *out;     // this environment's row of out_ids (global)
  int count, raw, limit;
};

// 32-bit shared-window addresses + ld/st.shared: through a generic pointer held in a struct every access paid
// S2R SR_CgaCtaId + LEA + IMAD for the window base (70 instructions per claim pass in the first version's SASS)
__device__ __forceinline__ int32_t ew_lds(uint32_t a) {
  int32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void ew_sts(uint32_t a, int32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ew_lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void ew_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// predicated forms: a load / store inside an `if` costs a BSSY / BRA / BSYNC triple per use in the claim pass
__device__ __forceinline__ int32_t ew_lds_if(bool p, uint32_t a, int32_t otherwise) {
  int32_t v = otherwise;
  asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; @p ld.shared.b32 %0, [%1]; }" : "+r"(v) : "r"(a), "r"((int)p) : "memory");
  return v;
}
__device__ __forceinline__ void ew_sts_if(bool p, uint32_t a, int32_t v) {
  asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; @p st.shared.b32 [%0], %1; }" ::"r"(a), "r"(v), "r"((int)p) : "memory");
}

// one batch of 32 positions (valid = val >= 0 and not the excluded node); returns true once `limit` unique values are out
__device__ __forceinline__ bool wide_push(WideState &st, int32_t val, int32_t exclude, int lane) {
  constexpr unsigned FULL = 0xffffffffu;
  const bool valid = val >= 0 && val != exclude;
  const unsigned vmask = __ballot_sync(FULL, valid);
  if (vmask == 0) return false;
  st.raw += __popc(vmask);
  // double hashing: linear probing piles every lane that enters a cluster onto the cluster's one free slot; an odd
  // per-value stride keeps different values on different sequences (measured: 1.37 claim passes, 1.9 extra probes a batch)
  const uint32_t h = (uint32_t)val * 2654435761u;
  const uint32_t step4 = ((h >> 5) & ~3u) | 4u;  // odd slot stride, in bytes
  uint32_t off = (h >> st.tshift2) & st.mask4;
  bool looking = valid, ins = false;  // looking for its slot | inserted by this lane
  do {
    int32_t e = ew_lds_if(looking, st.tbase + off, val);  // (lanes that are not looking fall through the loop)
    int probes = 0;
    while (e != -1 && e != val) {
      AHR_DEV_ASSERT(++probes <= (int)(st.mask4 >> 2) && off <= st.mask4 && (off & 3u) == 0);  // the table is never full
      off = (off + step4) & st.mask4;
      e = ew_lds(st.tbase + off);
    }
    (void)probes;
    looking = looking && e != val;  // found: inserted by an earlier batch, a duplicate
    __syncwarp();  // every probe of this pass (and every read-back of the last one) is done before anyone writes
    ew_sts_if(looking, st.tbase + off, val);
    __syncwarp();
    const bool won = ew_lds_if(looking, st.tbase + off, ~val) == val;  // lanes carrying the same value all win together
    ins = ins || won;
    if (looking && !won) off = (off + step4) & st.mask4;  // another value took the slot
    looking = looking && !won;
  } while (__any_sync(FULL, looking));
  // the same value in several lanes of this batch: they share the slot, the owner byte tells them apart
  const uint32_t oa = st.obase + (off >> 2);
  if (ins) ew_sts8(oa, (uint32_t)lane);
  __syncwarp();
  const bool tied = ins && ew_lds8(oa) != (uint32_t)lane;
  bool fresh = ins;
  if (__any_sync(FULL, tied)) {  // first occurrence = lowest lane of the group that inserted the value
    const unsigned grp = __match_any_sync(FULL, ins ? val : (int32_t)(-2 - lane));
    fresh = ins && (__ffs(grp) - 1) == lane;
  }
  const unsigned nmask = __ballot_sync(FULL, fresh);
  const int pos = st.count + __popc(nmask & ((1u << lane) - 1u));
  AHR_DEV_ASSERT(pos >= 0);
  if (fresh && pos < st.limit) st.out[pos] = val;
  st.count = min(st.limit, st.count + __popc(nmask));
  return st.count >= st.limit;
}

// one CSR row appended to the concatenation, optionally preceded by the single value `literal` (the hyperedge itself,
// environment.py:459) and without the participants equal to `exclude`
__device__ __forceinline__ bool wide_segment(WideState &st, const GraphView &g, const int64_t *__restrict__ row_ptr,
                                             const int32_t *__restrict__ col_idx, int32_t v, int32_t exclude,
                                             int32_t literal, int lane) {
  const int64_t p0 = row_ptr[v], p1 = row_ptr[v + 1];
  const int32_t *__restrict__ src = col_idx + p0;
  bool lit_here = literal >= 0 && lane == 0;  // lane 0 of the first batch carries the literal, not a column entry
  if (literal >= 0) --src;
  for (int64_t left = p1 - p0 + (literal >= 0 ? 1 : 0); left > 0; left -= 32 * EW_UNROLL, src += 32 * EW_UNROLL) {
    const int rem = (int)min(left, (int64_t)(32 * EW_UNROLL));
    int32_t val[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {  // the loads first (coalesced: consecutive lanes, consecutive neighbours) ...
      const int li = u * 32 + lane;
      val[u] = (li < rem && !(u == 0 && lit_here)) ? src[li] : -1;
    }
    const int32_t excl0 = lit_here ? -1 : exclude;  // the literal is not subject to the exclusion
    if (lit_here) val[0] = literal;
    lit_here = false;
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {  // ... then the inserts, in order
      if (u * 32 >= rem) break;
      AHR_DEV_ASSERT(val[u] < g.n_nodes && (val[u] >= 0 || u * 32 + lane >= rem));
      if (wide_push(st, val[u], u == 0 ? excl0 : exclude, lane)) return true;
    }
  }
  return false;
}

__global__ void __launch_bounds__(128)
expand_wide_kernel(GraphView g, int32_t n_env, const int32_t *__restrict__ ops, const int32_t *__restrict__ in_ids,
                   const int32_t *__restrict__ in_offsets, int32_t limit, int32_t T, int32_t *__restrict__ out_ids,
                   int32_t *__restrict__ out_counts) {
  extern __shared__ __align__(16) uint8_t ew_smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int env = blockIdx.x * n_warps + w;
  if (env >= n_env) return;  // warps are independent: no block-wide barrier below
  WideState st;
  int32_t *table = reinterpret_cast<int32_t *>(ew_smem) + (size_t)w * T;
  st.tbase = (uint32_t)__cvta_generic_to_shared(table);
  st.obase = (uint32_t)__cvta_generic_to_shared(ew_smem + (size_t)n_warps * T * 4 + (size_t)w * T);
  st.out = out_ids + (int64_t)env * limit;
  asm volatile("" : "+l"(st.out));  // keep the row pointer in registers (it was re-derived from env * limit at every store)
  st.mask4 = ((uint32_t)T - 1u) << 2;
  st.tshift2 = 32 - (31 - __clz(T)) - 2;  // T >= 64: at least 6 hash bits, the shift stays positive
  st.count = 0; st.raw = 0; st.limit = limit;
  const int op = ops[env];
  const int32_t *ids = in_ids + in_offsets[env];
  const int n_ids = in_offsets[env + 1] - in_offsets[env];
  for (int i = lane * 4; i < T; i += 128) *reinterpret_cast<int4 *>(table + i) = make_int4(-1, -1, -1, -1);
  __syncwarp();
  // pull a neighbour list into L2 while the list before it is being inserted (one line per lane and step)
  auto prefetch = [&](int slot_csr, int32_t v) {
    const int64_t p0 = g.row_ptr[slot_csr][v], p1 = g.row_ptr[slot_csr][v + 1];
    const char *b = reinterpret_cast<const char *>(g.col_idx[slot_csr] + p0);
    const int64_t bytes = min((p1 - p0) * 4, (int64_t)(limit + 32) * 4);
    for (int64_t o = (int64_t)lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + o));
  };
  bool full = false;
  for (int idx = 0; idx < n_ids && !full; ++idx) {
    const int32_t v = ids[idx];
    if (v < 0 || v >= g.n_nodes) continue;  // `if nid not in G: continue`
    // The segments of v, walked by ONE copy of the segment code (five inlined copies x EW_UNROLL pushes were 157 KB of SASS:
    // instruction-cache misses showed up as no_instruction / branch_resolving stalls):
    //   parents: belongs_to out | children: belongs_to in | related: related_to out, related_to in, then for an entity every
    //   hyperedge h of v followed by the other participants of h (environment.py:456-467)
    int64_t h0 = 0, n_seg = 1;
    if (op != AHR_OP_PARENTS && op != AHR_OP_CHILDREN) {
      prefetch(AHR_CSR_REL_IN, v);
      n_seg = 2;
      if (g.line[v].head.type == AHR_NT_ENTITY) {
        h0 = g.row_ptr[AHR_CSR_PART_OUT][v];
        n_seg += g.row_ptr[AHR_CSR_PART_OUT][v + 1] - h0;
      }
    }
    for (int64_t sg = 0; sg < n_seg && !full; ++sg) {
      // (pointers picked by selects: indexing the kernel parameter with a runtime slot would copy it to local memory)
      const int64_t *rp = op == AHR_OP_PARENTS ? g.row_ptr[AHR_CSR_BEL_OUT] : (op == AHR_OP_CHILDREN ? g.row_ptr[AHR_CSR_BEL_IN] : (sg == 0 ? g.row_ptr[AHR_CSR_REL_OUT] : g.row_ptr[AHR_CSR_REL_IN]));
      const int32_t *ci = op == AHR_OP_PARENTS ? g.col_idx[AHR_CSR_BEL_OUT] : (op == AHR_OP_CHILDREN ? g.col_idx[AHR_CSR_BEL_IN] : (sg == 0 ? g.col_idx[AHR_CSR_REL_OUT] : g.col_idx[AHR_CSR_REL_IN]));
      int32_t node = v, exclude = -1, literal = -1;
      if (sg >= 2) {
        const int32_t h = g.col_idx[AHR_CSR_PART_OUT][h0 + sg - 2];
        AHR_DEV_ASSERT(h >= 0 && h < g.n_nodes);
        rp = g.row_ptr[AHR_CSR_PART_IN]; ci = g.col_idx[AHR_CSR_PART_IN]; node = h; exclude = v; literal = h;
      }
      full = wide_segment(st, g, rp, ci, node, exclude, literal, lane);
    }
    if (st.raw >= limit) break;  // stop rule on the UN-deduplicated length, checked after finishing v
  }
  for (int j = st.count + lane; j < limit; j += 32) st.out[j] = -1;
  if (lane == 0) out_counts[env] = st.count;
}

int graph_expand(const GraphView &g, int32_t n_env, const int32_t *ops, const int32_t *in_ids,
                 const int32_t *in_offsets, int32_t limit, int32_t *out_ids, int32_t *out_counts, cudaStream_t stream) {
  AHR_REQUIRE(limit >= 1 && limit <= 2048, "expand: limit must be in [1, 2048] (got %d)", limit);
  if (n_env <= 0) return AHR_OK;
  if (limit <= 64) {
    const size_t smem = (size_t)EXP_WARPS * limit * sizeof(int32_t);
    expand_kernel<<<(n_env + EXP_WARPS - 1) / EXP_WARPS, EXP_WARPS * 32, smem, stream>>>(g, n_env, ops, in_ids, in_offsets,
                                                                                        limit, out_ids, out_counts);
    AHR_KERNEL_CHECK();
    return AHR_OK;
  }
  const int warps = ew_warps(limit);
  expand_wide_kernel<<<(n_env + warps - 1) / warps, warps * 32, ew_smem_bytes(limit), stream>>>(
      g, n_env, ops, in_ids, in_offsets, limit, ew_table_slots(limit + 32), out_ids, out_counts);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

}  // namespace ahr
