// kernels.h -- internal host-side entry points of the kernel translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/ahrag_b200.h"

namespace ahr {

void set_error(const char *fmt, ...);

// Grow-only device buffer.  Growing must not synchronise the device (cudaFree does): the superseded allocation is RETIRED
// -- kept until the handle is destroyed -- and the new one is at least 1.5x the old, so the retired blocks of a handle add
// up to less than twice its final workspace.  Kernels already enqueued on the old block keep a valid pointer.
struct DevBuf {
  void *ptr = nullptr;
  size_t bytes = 0;
  std::vector<void *> retired;
  int reserve(size_t need) {
    if (need <= bytes) return AHR_OK;
    need = std::max(need, bytes + bytes / 2);
    need = (need + 255) & ~(size_t)255;
    void *fresh = nullptr;
    cudaError_t e = cudaMalloc(&fresh, need);
    if (e != cudaSuccess) {
      cudaGetLastError();
      // out of memory: give the retired blocks back (this does synchronise) and try once more
      for (void *r : retired) cudaFree(r);
      retired.clear();
      e = cudaMalloc(&fresh, need);
    }
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
      cudaGetLastError();
      return AHR_E_NOMEM;
    }
    if (ptr) retired.push_back(ptr);
    ptr = fresh;
    bytes = need;
    return AHR_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    for (void *r : retired) cudaFree(r);
    retired.clear();
    ptr = nullptr;
    bytes = 0;
  }
};


// ---- topk_tc.cu (K1) ----
int tc_make_corpus_map(CUtensorMap *map, const void *dev_rows, int64_t n_rows, int32_t dim);
int tc_choose_slabs(int32_t n_qtiles, int64_t n_ctiles, int grid);
int tc_topk_partials(const CUtensorMap &map_x, const void *dev_q, int64_t n_rows, int32_t dim, int32_t nq, int32_t k,
                     int32_t n_slabs, int grid, float *part_scores, int32_t *part_ids, float *dbg_scores,
                     const struct TcExtra &ex, cudaStream_t stream);

// ---- topk_tc2.cu (K1, CTA-pair variant) ----
int tc2_make_corpus_map(CUtensorMap *map, const void *dev_rows, int64_t n_rows, int32_t dim);
// map_x_lo / dev_q_lo non-null: fp32 operands split into hi + lo bf16 halves (3 operand segments per tile)
int tc2_topk_partials(const CUtensorMap &map_x, const CUtensorMap *map_x_lo, const void *dev_q, const void *dev_q_lo,
                      int64_t n_rows, int32_t dim, int32_t nq, int32_t k, int32_t n_slabs, int n_pairs,
                      float *part_scores, int32_t *part_ids, float *dbg_scores,
                      unsigned long long *stats, const struct TcExtra &ex, cudaStream_t stream);
// fp32 [n, dim] -> hi = bf16(x), lo = bf16(x - hi)
int split_f32_hi_lo(const float *src, int64_t n_elems, void *hi, void *lo, cudaStream_t stream);

// ---- select.cu ----
// Per query: reduce n_cand candidates to the top k by (score desc, id asc).
//   cand_scores[q * stride + i];  ids: cand_ids32 / cand_ids64 (one may be given) or implicit = i;
//   final id = id + id_offset (entries with raw id < 0 are skipped).
//   carry_in: the current contents of out_* (k entries, id < 0 = empty) take part in the selection.
int select_topk(const float *cand_scores, const int32_t *cand_ids32, const int64_t *cand_ids64, int64_t stride,
                int64_t n_cand, int32_t nq, int32_t k, int64_t id_offset, bool carry_in, float *out_scores,
                int64_t *out_ids, cudaStream_t stream);
// general addressing: candidate i of query q at q * stride + (i / inner) * stride_outer + (i % inner)
int select_topk_strided(const float *cand_scores, const int32_t *cand_ids32, const int64_t *cand_ids64, int64_t stride,
                        int64_t n_cand, int64_t inner, int64_t stride_outer, int32_t nq, int32_t k, int64_t id_offset,
                        bool carry_in, float *out_scores, int64_t *out_ids, cudaStream_t stream);
// separate outer strides (in elements) for the score and id arrays
int select_topk_strided2(const float *cand_scores, const int32_t *cand_ids32, const int64_t *cand_ids64, int64_t stride,
                         int64_t n_cand, int64_t inner, int64_t stride_outer, int64_t stride_outer_ids, int32_t nq,
                         int32_t k, int64_t id_offset, bool carry_in, float *out_scores, int64_t *out_ids,
                         cudaStream_t stream);
constexpr int SELECT_MAX_K = 1024;
// Peer destinations of a fused "reduce + push" launch (exchange.cu fills it from an ahr_comm): slot `rank` of every peer's
// exchange buffer receives this rank's [nq, k] result, the last block publishes `epoch` into the peers' flags.
struct PeerPush {
  uint8_t *base[16] = {};   // mapped peer allocations (base[rank] = own)
  int rank = 0, world = 1;
  size_t slot_off = 0;      // byte offset of slot `rank` of the current half inside every allocation
  size_t flag_off = 0;      // byte offset of the current half's flag array
  size_t scores_bytes = 0;  // nq * k * 4: the ids of a record start there
  uint64_t epoch = 0;
  unsigned int *done_counter = nullptr;
};
// Merge of n_lists SORTED lists per query (valid entries first, (score desc, id asc)): list g, slot j of query q at
// q * stride + g * stride_list + j (ids: stride_list_ids).  Optionally pushes the result to the peers (PeerPush).
bool merge_sorted_applicable(int32_t n_lists, int32_t k_in, int32_t k);
int merge_sorted_lists(const float *cand_scores, const int32_t *cand_ids32, const int64_t *cand_ids64, int64_t stride,
                       int32_t n_lists, int32_t k_in, int64_t stride_list, int64_t stride_list_ids, int32_t nq, int32_t k,
                       int64_t id_offset, float *out_scores, int64_t *out_ids, const PeerPush *push, cudaStream_t stream);
// exchange.cu: the stand-alone push of a finished record (shapes the fused reduction does not take, fp32-exact mode)
int peer_push_record(const PeerPush &push, size_t rec_bytes, cudaStream_t stream);
// the bf16-mode top-k of an index with the final reduction pushing to the peers (api.cu; used by ahr_topk_sharded)
int index_topk_bf16(struct ::ahr_index *idx, const void *dev_q, int32_t nq, int32_t k, float *dev_scores, int64_t *dev_ids,
                    const PeerPush *push, cudaStream_t stream);

// ---- topk_exact.cu (K1x) ----
// canonical scores of rows [row0, row0 + n_chunk) for the queries listed in q_sel (or all nq when q_sel == nullptr):
// out[qi * n_chunk + r]
int exact_scores(const void *dev_rows, const void *dev_q, int dtype, int32_t dim, int64_t row0, int64_t n_chunk,
                 const int32_t *q_sel, int32_t nq_sel, float *out, cudaStream_t stream);
int exact_rescore(const void *dev_rows, const void *dev_q, int dtype, int32_t dim, int32_t nq, int32_t m,
                  const int64_t *cand_ids, int64_t id_base, float *out_scores, cudaStream_t stream);
// certification of tensor-core candidate sets (see api.cu)
int exact_certify(const float *tc_scores, const int64_t *tc_ids, const float *ex_scores, int32_t nq, int32_t kc,
                  int32_t k, const float *eps_per_query, float *out_scores, int64_t *out_ids, int32_t *fail_flags,
                  int32_t *fail_count, int32_t *fail_list, cudaStream_t stream);
int row_norms_max(const void *dev_rows, int dtype, int64_t n_rows, int32_t dim, float *dev_out_max, cudaStream_t stream);
int query_eps(const void *dev_q, int dtype, int32_t nq, int32_t dim, const float *dev_max_row_norm, float eps_rel,
              float *dev_eps, cudaStream_t stream);
int gather_rows_i64(const float *src_s, const int64_t *src_i, const int32_t *sel, int32_t n_sel, int32_t k,
                    float *dst_s, int64_t *dst_i, bool scatter, cudaStream_t stream);

int gather_query_rows(const void *src, int64_t row_bytes, const int32_t *sel, int32_t n_sel, void *out, cudaStream_t stream);

// ---- graph.cu (K2) ----
// Node attributes + the head of the node's anchor list packed into ONE 64-byte line per node (library-owned, built by
// ahr_graph_create).  The five separate attribute arrays cost five DRAM sectors per touched node, and a seed's expansion
// walked seed -> attributes -> row_ptr -> col_idx -> attributes: four dependent random sectors (hop 1 on c5: 63.5 MB of
// sector traffic for ~10 MB of useful bytes, profiles/r1_k2_ncu.md).
//   sector 0 (NodeHead)  what the rerank / featurizer read for every emitted node
//   sector 1             the first 8 entries of the list a seed of this type expands into: entity -> belongs_to parents
//                        (G.out_edges order), summary -> members (G.in_edges order); -1 padded.  The reference's caps
//                        (2 parents, member_top_m = 5) never leave the line.
constexpr int NODE_ANCHOR_INLINE = 8;
struct alignas(32) NodeHead {
  double judge;        // NaN = None
  double judge_term;   // 1 / (1 + exp(-judge / 10)) as evaluated by the host (the reference's own np.exp)
  double conf;         // NaN = None
  int32_t anchor_deg;  // full length of the anchor list (0 for hyperedges / unknown types)
  uint8_t type;        // AHR_NT_*
  uint8_t level;
  uint8_t pad[2];
};
struct alignas(64) NodeLine {
  NodeHead head;
  int32_t anchor[NODE_ANCHOR_INLINE];
};
static_assert(sizeof(NodeHead) == 32 && sizeof(NodeLine) == 64, "NodeLine must be two 32-byte sectors");

struct GraphView {
  int64_t n_nodes;
  const int64_t *row_ptr[AHR_N_CSR];
  const int32_t *col_idx[AHR_N_CSR];
  const NodeLine *line;  // [n_nodes], library-owned
};
// Anchor expansion over a graph SHARDED by node range: instead of NodeLines the kernel reads the all-gathered records of
// graph_shard.cu.  Stage 1 (owner of a seed), per (query, seed): [present, type, n_nbr, level, judge(2), judge_term(2),
// conf(2)] + nbr[fan] (+ pad to an even count); stage 2 (owner of an emitted neighbour), per (query, seed, j): the same
// 10-int head.  Buffers are [world][nq][n_seeds][...] in rank order.
constexpr int ANCHOR_REC_HEAD = 10;
struct AnchorGathered {
  const int32_t *g1 = nullptr, *g2 = nullptr;
  const int64_t *bounds = nullptr;  // [world + 1] first node id of every rank
  int32_t world = 1, rec1 = 0;      // ints per stage-1 record
};
// seeds: dense node ids (seed_nodes), or index rows + the row -> node table (seed_rows / row_to_node / n_rows)
int graph_anchor_rerank(const GraphView &g, int32_t nq, int32_t n_seeds, const int32_t *seed_nodes,
                        const int64_t *seed_rows, const int32_t *row_to_node, int64_t n_rows, const float *seed_scores, const ahr_rerank_params &prm, int32_t top_k, int32_t *out_nodes,
                        double *out_sem, double *out_score, int32_t *out_count, int32_t *exp_nodes, double *exp_sem,
                        uint8_t *exp_kind, int32_t *exp_count, double *seed_sem, void *workspace, size_t workspace_bytes,
                        cudaStream_t stream, const AnchorGathered *gathered = nullptr);
size_t graph_anchor_workspace_bytes(int32_t nq, int32_t n_seeds, const ahr_rerank_params &prm);
int graph_expand(const GraphView &g, int32_t n_env, const int32_t *ops, const int32_t *in_ids,
                 const int32_t *in_offsets, int32_t limit, int32_t *out_ids, int32_t *out_counts, cudaStream_t stream);
int debug_round4(const double *x, int64_t n, double *out, cudaStream_t stream);
int graph_pack_nodes(int64_t n_nodes, const uint8_t *node_type, const uint8_t *level, const double *judge,
                     const double *judge_term, const double *conf, const int64_t *const row_ptr[AHR_N_CSR],
                     const int32_t *const col_idx[AHR_N_CSR], NodeLine *out, cudaStream_t stream);

// ---- lca.cu ----
size_t graph_lca_workspace_bytes(int64_t n_nodes, int32_t n_ids);
int graph_lca_closure(const GraphView &g, const int32_t *host_ids, int32_t n_ids, uint8_t *out_flags, int64_t *n_result,
                      void *workspace, cudaStream_t stream);

// ---- env.cu ----
int env_reset(const GraphView &g, const ahr_env_batch &b, const int32_t *nodes, const double *sem, const double *score,
              const int32_t *count, int32_t top_k, int32_t n_seeds, float *feat, uint8_t *mask, cudaStream_t stream);
int env_step(const GraphView &g, const ahr_env_batch &b, const int32_t *actions, int32_t max_steps, double repeat_penalty,
             float *feat, double *rewards, uint8_t *done, uint8_t *mask, cudaStream_t stream);

}  // namespace ahr
