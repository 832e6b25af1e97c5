// graph_shard.cu -- frontier expansion over a CSR SHARDED BY NODE RANGE across the GPUs of a box (SURVEY section 8e,
// option B; BASELINE north star: "the corpus index and the CSR shard by node range").
//
// Rank r owns the adjacency rows (all six typed CSRs) and the node types of the dense ids [lo_r, hi_r).  The inputs of a
// batch (ops, input ids) are replicated, like the queries of the sharded index.  One expand_* call of the reference
// (src/ah_rag/agent/environment.py:393-480) -- ordered concatenation of the typed neighbour lists of the input nodes,
// stop rule on the un-deduplicated length after each input node, order-preserving dedup, cut to `limit` -- becomes
//
//   stage 1 (owner of input node v)    the order-preserving unique prefix of v's first-level list (parents / children /
//                                      related out ++ in), its un-deduplicated length, and for expand_related on an
//                                      entity the first `limit` hyperedges of v
//   exchange 1                         all-gather of the fixed-size stage-1 records
//   stage 2 (owner of hyperedge h)     the unique prefix of h's other participants and their count   (2-hop walk, :456-467)
//   exchange 2                         all-gather of the stage-2 records
//   merge (every rank)                 per environment, in input order: append the prefixes, test the stop rule, dedup, cut
//
// Why bounded prefixes are enough (bit-exactness): the output holds at most `limit` values.  When input node B is reached
// the nodes before it contributed fewer than `limit` elements (otherwise the stop rule had ended the walk), so at most
// limit - 1 of B's values are duplicates of earlier ones: the first cap = 2 * limit - 1 UNIQUE values of B contain every
// value B can add.  The stop rule only asks whether the un-deduplicated length reached `limit`, so lengths travel
// saturated at `limit`; a prefix that filled up has seen >= cap >= limit elements.  Every hyperedge adds at least itself,
// so the first `limit` hyperedges of an entity decide both the output and the stop rule.
#include "common.cuh"
#include "graph_shared.cuh"
#include "kernels.h"

namespace ahr {

struct ShardView {
  int64_t n_nodes;  // total
  int64_t lo, hi;   // owned dense ids
  const int64_t *row_ptr[AHR_N_CSR];  // local: row v - lo
  const int32_t *col_idx[AHR_N_CSR];
  const uint8_t *node_type;           // local
  // local attribute slices (anchor expansion over the sharded graph; optional)
  const uint8_t *level = nullptr;
  const double *judge = nullptr, *judge_term = nullptr, *conf = nullptr;
};

// record geometry shared by the three kernels and the host wrappers
struct ShardGeom {
  int limit, cap, max_ids;
  __host__ __device__ int rec1() const { return 4 + cap + limit; }  // [raw_sat, flags, n_c, n_h] c[cap] h[limit]
  __host__ __device__ int rec2() const { return 4 + cap; }          // [raw_sat, present, n_p, 0] p[cap]
};
constexpr int SH_PRESENT = 1, SH_ENTITY = 2;
constexpr int SH_WARPS = 4;

__device__ __forceinline__ bool shard_segment(ExpandState &st, const ShardView &g, int slot, int32_t v, int32_t exclude, int lane) {
  const int64_t r = v - g.lo;
  const int64_t p0 = g.row_ptr[slot][r], p1 = g.row_ptr[slot][r + 1];
  for (int64_t b = p0; b < p1; b += 32) {
    const int64_t i = b + lane;
    int32_t val = -1;
    bool valid = i < p1;
    if (valid) {
      val = g.col_idx[slot][i];
      AHR_DEV_ASSERT(val >= 0 && val < g.n_nodes);
      valid = val != exclude;
    }
    if (expand_push(st, val, valid, lane)) return true;
  }
  return false;
}

// stage 1: one warp per (environment, input position)
__global__ void __launch_bounds__(SH_WARPS * 32)
shard_stage1_kernel(ShardView g, ShardGeom geo, int32_t n_env, const int32_t *__restrict__ ops, const int32_t *__restrict__ in_ids,
                    const int32_t *__restrict__ in_offsets, int32_t *__restrict__ out) {
  extern __shared__ int32_t sm_list[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * SH_WARPS + w;
  if (item >= (int64_t)n_env * geo.max_ids) return;
  const int env = (int)(item / geo.max_ids), idx = (int)(item % geo.max_ids);
  int32_t *rec = out + item * geo.rec1();
  const int n_ids = in_offsets[env + 1] - in_offsets[env];
  const int32_t v = idx < n_ids ? in_ids[in_offsets[env] + idx] : -1;
  if (v < g.lo || v >= g.hi) {  // not ours (or no such input): an absent record
    if (lane < 4) rec[lane] = 0;
    return;
  }
  const int op = ops[env];
  ExpandState st;
  st.out = sm_list + (size_t)w * geo.cap;
  st.count = 0; st.raw = 0; st.limit = geo.cap;
  bool full;
  if (op == AHR_OP_PARENTS) full = shard_segment(st, g, AHR_CSR_BEL_OUT, v, -1, lane);
  else if (op == AHR_OP_CHILDREN) full = shard_segment(st, g, AHR_CSR_BEL_IN, v, -1, lane);
  else {
    full = shard_segment(st, g, AHR_CSR_REL_OUT, v, -1, lane);
    if (!full) full = shard_segment(st, g, AHR_CSR_REL_IN, v, -1, lane);
  }
  (void)full;
  const bool entity = g.node_type[v - g.lo] == AHR_NT_ENTITY;
  int n_h = 0;
  if (op == AHR_OP_RELATED && entity) {
    const int64_t r = v - g.lo;
    const int64_t h0 = g.row_ptr[AHR_CSR_PART_OUT][r], h1 = g.row_ptr[AHR_CSR_PART_OUT][r + 1];
    n_h = (int)min((int64_t)geo.limit, h1 - h0);
    for (int j = lane; j < n_h; j += 32) rec[4 + geo.cap + j] = g.col_idx[AHR_CSR_PART_OUT][h0 + j];
  }
  __syncwarp();
  for (int j = lane; j < st.count; j += 32) rec[4 + j] = st.out[j];
  if (lane == 0) {
    rec[0] = min(st.raw, geo.limit);  // a prefix that filled up saw >= cap >= limit elements
    rec[1] = SH_PRESENT | (entity ? SH_ENTITY : 0);
    rec[2] = st.count;
    rec[3] = n_h;
  }
}

__device__ __forceinline__ int shard_owner(const int64_t *bounds, int world, int32_t v) {
  int r = 0;
  while (r + 1 < world && v >= bounds[r + 1]) ++r;
  return r;
}

// stage 2: one warp per (environment, input position, hyperedge slot); reads the gathered stage-1 records
__global__ void __launch_bounds__(SH_WARPS * 32)
shard_stage2_kernel(ShardView g, ShardGeom geo, int32_t n_env, int32_t world, const int64_t *__restrict__ bounds,
                    const int32_t *__restrict__ ops, const int32_t *__restrict__ in_ids, const int32_t *__restrict__ in_offsets,
                    const int32_t *__restrict__ gathered1, int32_t *__restrict__ out) {
  extern __shared__ int32_t sm_list[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * SH_WARPS + w;
  const int64_t n_items = (int64_t)n_env * geo.max_ids * geo.limit;
  if (item >= n_items) return;
  const int j = (int)(item % geo.limit);
  const int64_t ei = item / geo.limit;  // (env, idx)
  const int env = (int)(ei / geo.max_ids), idx = (int)(ei % geo.max_ids);
  int32_t *rec = out + item * geo.rec2();
  const int n_ids = in_offsets[env + 1] - in_offsets[env];
  const int32_t v = idx < n_ids ? in_ids[in_offsets[env] + idx] : -1;
  int32_t h = -1;
  if (ops[env] == AHR_OP_RELATED && v >= 0 && v < g.n_nodes) {
    const int32_t *r1 = gathered1 + ((int64_t)shard_owner(bounds, world, v) * n_env * geo.max_ids + ei) * geo.rec1();
    if ((r1[1] & SH_ENTITY) && j < r1[3]) h = r1[4 + geo.cap + j];
  }
  if (h < g.lo || h >= g.hi) {
    if (lane < 4) rec[lane] = 0;
    return;
  }
  ExpandState st;
  st.out = sm_list + (size_t)w * geo.cap;
  st.count = 0; st.raw = 0; st.limit = geo.cap;
  shard_segment(st, g, AHR_CSR_PART_IN, h, v, lane);
  for (int t = lane; t < st.count; t += 32) rec[4 + t] = st.out[t];
  if (lane == 0) {
    rec[0] = min(st.raw, geo.limit);
    rec[1] = SH_PRESENT;
    rec[2] = st.count;
    rec[3] = 0;
  }
}

// merge: one warp per environment, over the gathered records of both stages
__global__ void __launch_bounds__(SH_WARPS * 32)
shard_merge_kernel(ShardGeom geo, int64_t n_nodes, int32_t n_env, int32_t world, const int64_t *__restrict__ bounds,
                   const int32_t *__restrict__ ops, const int32_t *__restrict__ in_ids, const int32_t *__restrict__ in_offsets,
                   const int32_t *__restrict__ gathered1, const int32_t *__restrict__ gathered2, int32_t *__restrict__ out_ids,
                   int32_t *__restrict__ out_counts) {
  extern __shared__ int32_t sm_list[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * SH_WARPS + w;
  if (env >= n_env) return;
  ExpandState st;
  st.out = sm_list + (size_t)w * geo.limit;
  st.count = 0; st.raw = 0; st.limit = geo.limit;
  const int op = ops[env];
  const int n_ids = min(in_offsets[env + 1] - in_offsets[env], geo.max_ids);
  int raw = 0;
  bool full = false;
  auto push_list = [&](const int32_t *vals, int n) {
    for (int b = 0; b < n && !full; b += 32) {
      const int i = b + lane;
      full = expand_push(st, i < n ? vals[i] : -1, i < n, lane);
    }
  };
  for (int idx = 0; idx < n_ids && !full; ++idx) {
    const int32_t v = in_ids[in_offsets[env] + idx];
    if (v < 0 || v >= n_nodes) continue;  // `if nid not in G: continue`
    const int64_t ei = (int64_t)env * geo.max_ids + idx;
    const int32_t *r1 = gathered1 + ((int64_t)shard_owner(bounds, world, v) * n_env * geo.max_ids + ei) * geo.rec1();
    AHR_DEV_ASSERT(r1[2] >= 0 && r1[2] <= geo.cap && r1[3] >= 0 && r1[3] <= geo.limit);
    push_list(r1 + 4, r1[2]);
    raw += r1[0];
    if (op == AHR_OP_RELATED && (r1[1] & SH_ENTITY)) {
      for (int j = 0; j < r1[3] && !full; ++j) {
        const int32_t h = r1[4 + geo.cap + j];
        full = expand_push(st, h, lane == 0, lane);
        raw += 1;
        const int32_t *r2 = gathered2 + ((int64_t)shard_owner(bounds, world, h) * n_env * geo.max_ids * geo.limit + ei * geo.limit + j) * geo.rec2();
        if (!full) push_list(r2 + 4, r2[2]);
        raw += r2[0];
      }
    }
    if (raw >= geo.limit) break;  // stop rule on the un-deduplicated length (saturated terms keep the comparison exact)
  }
  for (int t = lane; t < geo.limit; t += 32) out_ids[(int64_t)env * geo.limit + t] = t < st.count ? st.out[t] : -1;
  if (lane == 0) out_counts[env] = st.count;
}

// ---------------------------------------------------------------------------------------------
// Anchor expansion + rerank (hierarchical_graph.py:719-789) over the sharded graph: the two stages that replace the
// NodeLine reads of anchor_rerank_kernel (graph.cu); the rerank itself is that same kernel, fed with the gathered records.
//   stage 1 (owner of seed v)        type, attributes, and the first parent_cap parents (entity) / member_top_m members
//                                    (summary) of v, in adjacency order
//   stage 2 (owner of neighbour u)   type and attributes of every node stage 1 emitted
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void shard_write_head(int32_t *rec, const ShardView &g, int32_t v, int n_nbr) {
  const int64_t r = v - g.lo;
  rec[0] = SH_PRESENT;
  rec[1] = g.node_type[r];
  rec[2] = n_nbr;
  rec[3] = g.level[r];
  double *d = reinterpret_cast<double *>(rec + 4);
  d[0] = g.judge[r]; d[1] = g.judge_term[r]; d[2] = g.conf[r];
}

__global__ void shard_anchor_stage1_kernel(ShardView g, int32_t nq, int32_t n_seeds, int32_t fan, int32_t rec1, int32_t parent_cap,
                                           int32_t member_top_m, const int32_t *__restrict__ seeds, int32_t *__restrict__ out) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (int64_t)nq * n_seeds) return;
  int32_t *rec = out + item * rec1;
  const int32_t v = seeds[item];
  if (v < g.lo || v >= g.hi) {
    rec[0] = 0; rec[1] = AHR_NT_OTHER; rec[2] = 0; rec[3] = 0;
    return;
  }
  const int64_t r = v - g.lo;
  const int t = g.node_type[r];
  const int slot = t == AHR_NT_ENTITY ? AHR_CSR_BEL_OUT : AHR_CSR_BEL_IN;
  const int want = t == AHR_NT_ENTITY ? parent_cap : (t == AHR_NT_SUMMARY ? member_top_m : 0);
  const int64_t p0 = g.row_ptr[slot][r], p1 = g.row_ptr[slot][r + 1];
  const int n_nbr = (int)min((int64_t)min(want, fan), p1 - p0);
  for (int j = 0; j < n_nbr; ++j) rec[ANCHOR_REC_HEAD + j] = g.col_idx[slot][p0 + j];
  shard_write_head(rec, g, v, n_nbr);
}

__global__ void shard_anchor_stage2_kernel(ShardView g, int32_t nq, int32_t n_seeds, int32_t fan, int32_t rec1, int32_t world,
                                           const int64_t *__restrict__ bounds, const int32_t *__restrict__ seeds,
                                           const int32_t *__restrict__ gathered1, int32_t *__restrict__ out) {
  const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= (int64_t)nq * n_seeds * fan) return;
  const int j = (int)(item % fan);
  const int64_t qr = item / fan;  // (query, seed)
  int32_t *rec = out + item * ANCHOR_REC_HEAD;
  const int32_t v = seeds[qr];
  int32_t u = -1;
  if (v >= 0 && v < g.n_nodes) {
    const int32_t *r1 = gathered1 + ((int64_t)shard_owner(bounds, world, v) * nq * n_seeds + qr) * rec1;
    if (j < r1[2]) u = r1[ANCHOR_REC_HEAD + j];
  }
  if (u < g.lo || u >= g.hi) {
    rec[0] = 0; rec[1] = AHR_NT_OTHER; rec[2] = 0; rec[3] = 0;
    return;
  }
  shard_write_head(rec, g, u, 0);
}

}  // namespace ahr

using namespace ahr;

struct ahr_graph_shard {
  int device = 0;
  ShardView view;
  DevBuf ws;  // workspace of the rerank over large expanded maps
};

static int anchor_rec1_ints(int32_t fan) { return ANCHOR_REC_HEAD + fan + (fan & 1); }  // even: the doubles stay 8-byte aligned

static ShardGeom make_geom(int32_t limit, int32_t max_ids) {
  ShardGeom g;
  g.limit = limit;
  g.cap = 2 * limit - 1;
  g.max_ids = max_ids;
  return g;
}

extern "C" {

int ahr_graph_shard_create(ahr_graph_shard **out, int device, int64_t n_nodes_total, int64_t lo, int64_t hi,
                           const int64_t *const dev_row_ptr_local[AHR_N_CSR], const int32_t *const dev_col_idx_local[AHR_N_CSR],
                           const uint8_t *dev_node_type_local) {
  AHR_REQUIRE(out && dev_row_ptr_local && dev_col_idx_local, "ahr_graph_shard_create: NULL argument");
  AHR_REQUIRE(n_nodes_total >= 0 && n_nodes_total < (int64_t)INT32_MAX && lo >= 0 && lo <= hi && hi <= n_nodes_total,
              "ahr_graph_shard_create: bad range [%lld, %lld) of %lld", (long long)lo, (long long)hi, (long long)n_nodes_total);
  for (int s = 0; s < AHR_N_CSR; ++s) AHR_REQUIRE(dev_row_ptr_local[s] != nullptr, "ahr_graph_shard_create: row_ptr[%d] is NULL", s);
  AHR_REQUIRE(hi == lo || dev_node_type_local, "ahr_graph_shard_create: NULL node types");
  ahr_graph_shard *g = new ahr_graph_shard();
  g->device = device;
  g->view.n_nodes = n_nodes_total;
  g->view.lo = lo;
  g->view.hi = hi;
  for (int s = 0; s < AHR_N_CSR; ++s) {
    g->view.row_ptr[s] = dev_row_ptr_local[s];
    g->view.col_idx[s] = dev_col_idx_local[s];
  }
  g->view.node_type = dev_node_type_local;
  *out = g;
  return AHR_OK;
}

int ahr_graph_shard_destroy(ahr_graph_shard *g) {
  if (!g) return AHR_OK;
  DeviceGuard guard(g->device);
  g->ws.release();
  delete g;
  return AHR_OK;
}

int ahr_graph_shard_set_attrs(ahr_graph_shard *g, const uint8_t *dev_level_local, const double *dev_judge_local,
                              const double *dev_judge_term_local, const double *dev_conf_local) {
  AHR_REQUIRE(g != nullptr, "ahr_graph_shard_set_attrs: shard is NULL");
  AHR_REQUIRE(g->view.hi == g->view.lo || (dev_level_local && dev_judge_local && dev_judge_term_local && dev_conf_local),
              "ahr_graph_shard_set_attrs: NULL attribute array");
  g->view.level = dev_level_local;
  g->view.judge = dev_judge_local;
  g->view.judge_term = dev_judge_term_local;
  g->view.conf = dev_conf_local;
  return AHR_OK;
}

int64_t ahr_anchor_shard_record_ints(int32_t stage, int32_t n_seeds, const ahr_rerank_params *params) {
  if (!params || n_seeds < 1 || params->parent_cap < 0 || params->member_top_m < 0) return -1;
  const int fan = std::max(params->parent_cap, params->member_top_m);
  return stage == 1 ? (int64_t)n_seeds * anchor_rec1_ints(fan) : (stage == 2 ? (int64_t)n_seeds * fan * ANCHOR_REC_HEAD : -1);
}

int ahr_anchor_shard_stage1(ahr_graph_shard *g, int32_t nq, int32_t n_seeds, const int32_t *dev_seed_nodes,
                            const ahr_rerank_params *params, int32_t *dev_out_records, void *stream) {
  AHR_REQUIRE(g && params, "ahr_anchor_shard_stage1: NULL handle or params");
  AHR_REQUIRE(g->view.hi == g->view.lo || g->view.judge != nullptr, "ahr_anchor_shard_stage1: call ahr_graph_shard_set_attrs first");
  AHR_REQUIRE(nq >= 0 && n_seeds >= 1 && params->parent_cap >= 0 && params->member_top_m >= 0, "ahr_anchor_shard_stage1: bad sizes");
  if (nq == 0) return AHR_OK;
  AHR_REQUIRE(dev_seed_nodes && dev_out_records, "ahr_anchor_shard_stage1: NULL buffer");
  DeviceGuard guard(g->device);
  const int fan = std::max(params->parent_cap, params->member_top_m);
  const int64_t items = (int64_t)nq * n_seeds;
  shard_anchor_stage1_kernel<<<(unsigned)((items + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      g->view, nq, n_seeds, fan, anchor_rec1_ints(fan), params->parent_cap, params->member_top_m, dev_seed_nodes, dev_out_records);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

int ahr_anchor_shard_stage2(ahr_graph_shard *g, int32_t nq, int32_t n_seeds, int32_t world, const int64_t *dev_bounds,
                            const int32_t *dev_seed_nodes, const ahr_rerank_params *params, const int32_t *dev_gathered1,
                            int32_t *dev_out_records, void *stream) {
  AHR_REQUIRE(g && params, "ahr_anchor_shard_stage2: NULL handle or params");
  AHR_REQUIRE(g->view.hi == g->view.lo || g->view.judge != nullptr, "ahr_anchor_shard_stage2: call ahr_graph_shard_set_attrs first");
  AHR_REQUIRE(nq >= 0 && n_seeds >= 1 && world >= 1, "ahr_anchor_shard_stage2: bad sizes");
  const int fan = std::max(params->parent_cap, params->member_top_m);
  if (nq == 0 || fan == 0) return AHR_OK;
  AHR_REQUIRE(dev_bounds && dev_seed_nodes && dev_gathered1 && dev_out_records, "ahr_anchor_shard_stage2: NULL buffer");
  DeviceGuard guard(g->device);
  const int64_t items = (int64_t)nq * n_seeds * fan;
  shard_anchor_stage2_kernel<<<(unsigned)((items + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      g->view, nq, n_seeds, fan, anchor_rec1_ints(fan), world, dev_bounds, dev_seed_nodes, dev_gathered1, dev_out_records);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

int ahr_anchor_shard_merge(ahr_graph_shard *g, int32_t nq, int32_t n_seeds, int32_t world, const int64_t *dev_bounds,
                           const int32_t *dev_seed_nodes, const float *dev_seed_scores, const ahr_rerank_params *params,
                           int32_t top_k, const int32_t *dev_gathered1, const int32_t *dev_gathered2, int32_t *dev_out_nodes,
                           double *dev_out_semantic, double *dev_out_score, int32_t *dev_out_count, int32_t *dev_exp_nodes,
                           double *dev_exp_semantic, uint8_t *dev_exp_kind, int32_t *dev_exp_count, double *dev_seed_semantic,
                           void *stream) {
  AHR_REQUIRE(g && params, "ahr_anchor_shard_merge: NULL handle or params");
  if (nq == 0) return AHR_OK;
  AHR_REQUIRE(world >= 1 && dev_bounds && dev_seed_nodes && dev_seed_scores && dev_gathered1 && dev_out_nodes && dev_out_semantic &&
                  dev_out_score && dev_out_count, "ahr_anchor_shard_merge: NULL buffer");
  AHR_REQUIRE(!dev_exp_nodes || dev_exp_semantic, "ahr_anchor_shard_merge: exp_nodes needs exp_semantic");
  const int fan = std::max(params->parent_cap, params->member_top_m);
  AHR_REQUIRE(fan == 0 || dev_gathered2, "ahr_anchor_shard_merge: stage-2 records are NULL");
  DeviceGuard guard(g->device);
  const size_t need = graph_anchor_workspace_bytes(nq, n_seeds, *params);
  if (need) {
    int rc = g->ws.reserve(need);
    if (rc != AHR_OK) return rc;
  }
  GraphView view{};
  view.n_nodes = g->view.n_nodes;  // lines / CSR are not read in gathered mode
  AnchorGathered src;
  src.g1 = dev_gathered1; src.g2 = dev_gathered2; src.bounds = dev_bounds; src.world = world; src.rec1 = anchor_rec1_ints(fan);
  return graph_anchor_rerank(view, nq, n_seeds, dev_seed_nodes, nullptr, nullptr, 0, dev_seed_scores, *params, top_k, dev_out_nodes,
                             dev_out_semantic, dev_out_score, dev_out_count, dev_exp_nodes, dev_exp_semantic, dev_exp_kind,
                             dev_exp_count, dev_seed_semantic, g->ws.ptr, g->ws.bytes, reinterpret_cast<cudaStream_t>(stream), &src);
}

int64_t ahr_expand_shard_record_ints(int32_t stage, int32_t limit, int32_t max_ids) {
  if (limit < 1 || limit > 32 || max_ids < 1) return -1;
  const ShardGeom geo = make_geom(limit, max_ids);
  return stage == 1 ? (int64_t)max_ids * geo.rec1() : (stage == 2 ? (int64_t)max_ids * limit * geo.rec2() : -1);
}

static int check_shard_args(int32_t n_env, int32_t limit, int32_t max_ids, const char *who) {
  AHR_REQUIRE(n_env >= 0, "%s: negative n_env", who);
  AHR_REQUIRE(limit >= 1 && limit <= 32, "%s: the sharded expansion supports limits up to 32 (got %d; the reference uses 10)", who, limit);
  AHR_REQUIRE(max_ids >= 1 && max_ids <= 64, "%s: max_ids must be in [1, 64] (got %d)", who, max_ids);
  return AHR_OK;
}

int ahr_expand_shard_stage1(ahr_graph_shard *g, int32_t n_env, const int32_t *dev_ops, const int32_t *dev_in_ids,
                            const int32_t *dev_in_offsets, int32_t max_ids, int32_t limit, int32_t *dev_out_records,
                            void *stream) {
  AHR_REQUIRE(g != nullptr, "ahr_expand_shard_stage1: shard is NULL");
  int rc = check_shard_args(n_env, limit, max_ids, "ahr_expand_shard_stage1");
  if (rc != AHR_OK) return rc;
  if (n_env == 0) return AHR_OK;
  AHR_REQUIRE(dev_ops && dev_in_offsets && dev_out_records, "ahr_expand_shard_stage1: NULL buffer");
  DeviceGuard guard(g->device);
  const ShardGeom geo = make_geom(limit, max_ids);
  const int64_t items = (int64_t)n_env * max_ids;
  shard_stage1_kernel<<<(unsigned)((items + SH_WARPS - 1) / SH_WARPS), SH_WARPS * 32, (size_t)SH_WARPS * geo.cap * 4,
                        reinterpret_cast<cudaStream_t>(stream)>>>(g->view, geo, n_env, dev_ops, dev_in_ids, dev_in_offsets,
                                                                  dev_out_records);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

int ahr_expand_shard_stage2(ahr_graph_shard *g, int32_t n_env, int32_t world, const int64_t *dev_bounds, const int32_t *dev_ops,
                            const int32_t *dev_in_ids, const int32_t *dev_in_offsets, int32_t max_ids, int32_t limit,
                            const int32_t *dev_gathered1, int32_t *dev_out_records, void *stream) {
  AHR_REQUIRE(g != nullptr, "ahr_expand_shard_stage2: shard is NULL");
  int rc = check_shard_args(n_env, limit, max_ids, "ahr_expand_shard_stage2");
  if (rc != AHR_OK) return rc;
  if (n_env == 0) return AHR_OK;
  AHR_REQUIRE(world >= 1 && dev_bounds && dev_ops && dev_in_offsets && dev_gathered1 && dev_out_records,
              "ahr_expand_shard_stage2: NULL buffer");
  DeviceGuard guard(g->device);
  const ShardGeom geo = make_geom(limit, max_ids);
  const int64_t items = (int64_t)n_env * max_ids * limit;
  shard_stage2_kernel<<<(unsigned)((items + SH_WARPS - 1) / SH_WARPS), SH_WARPS * 32, (size_t)SH_WARPS * geo.cap * 4,
                        reinterpret_cast<cudaStream_t>(stream)>>>(g->view, geo, n_env, world, dev_bounds, dev_ops, dev_in_ids,
                                                                  dev_in_offsets, dev_gathered1, dev_out_records);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

int ahr_expand_shard_merge(int device, int64_t n_nodes_total, int32_t n_env, int32_t world, const int64_t *dev_bounds,
                           const int32_t *dev_ops, const int32_t *dev_in_ids, const int32_t *dev_in_offsets, int32_t max_ids,
                           int32_t limit, const int32_t *dev_gathered1, const int32_t *dev_gathered2, int32_t *dev_out_ids,
                           int32_t *dev_out_counts, void *stream) {
  int rc = check_shard_args(n_env, limit, max_ids, "ahr_expand_shard_merge");
  if (rc != AHR_OK) return rc;
  if (n_env == 0) return AHR_OK;
  AHR_REQUIRE(world >= 1 && dev_bounds && dev_ops && dev_in_offsets && dev_gathered1 && dev_gathered2 && dev_out_ids && dev_out_counts,
              "ahr_expand_shard_merge: NULL buffer");
  DeviceGuard guard(device);
  const ShardGeom geo = make_geom(limit, max_ids);
  shard_merge_kernel<<<(n_env + SH_WARPS - 1) / SH_WARPS, SH_WARPS * 32, (size_t)SH_WARPS * limit * 4,
                       reinterpret_cast<cudaStream_t>(stream)>>>(geo, n_nodes_total, n_env, world, dev_bounds, dev_ops, dev_in_ids,
                                                                 dev_in_offsets, dev_gathered1, dev_gathered2, dev_out_ids,
                                                                 dev_out_counts);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}

}  // extern "C"
