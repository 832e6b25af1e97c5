/*
 * ahrag_b200.h -- C ABI of the B200-native retrieval engine for AH-RAG's hot path.
 *
 * The reference (Yaron9/AH-RAG) is pure Python and exposes no FFI; the seams this library replaces are
 * Python call sites.  Each entry point below names the reference interface it stands in for
 * (paths relative to the reference tree).  See INTEGRATION.md for the ctypes binding a maintainer adds.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types; every function returns 0 on success and a negative
 *     AHR_E_* code otherwise; the message is available from ahr_last_error() (thread-local).
 *   - "dev_" pointers are CUDA device pointers on the handle's device, owned by the caller; the library
 *     never frees them.  Inputs of *_create must stay alive until *_destroy.
 *   - work is enqueued on the caller's CUDA stream (a cudaStream_t passed as void*; NULL = default
 *     stream) and is asynchronous unless the function name ends in _host or the doc says otherwise.
 *   - a handle is not thread-safe; distinct handles are independent.  One process per GPU.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with AHR_E_CUDA.
 */
#ifndef AHRAG_B200_H_
#define AHRAG_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AHR_VERSION 200

/* error codes */
#define AHR_OK 0
#define AHR_E_INVALID (-1)  /* bad argument */
#define AHR_E_CUDA (-2)     /* CUDA runtime / driver failure (message has the CUDA error string) */
#define AHR_E_UNSUPPORTED (-3)
#define AHR_E_NOMEM (-4)

/* row dtypes of an embedding index */
#define AHR_DTYPE_F32 0
#define AHR_DTYPE_BF16 1

/* top-k modes */
#define AHR_MODE_BF16 0   /* tcgen05 bf16 GEMM, fp32 accumulate, fused top-k epilogue (needs a bf16 index) */
#define AHR_MODE_EXACT 1  /* canonical fp32 score (fp64 ordered accumulation), ties -> lower id */

/* edge CSR slots: 3 edge types x 2 directions, adjacency in the LOADED NetworkX iteration order */
#define AHR_CSR_PART_OUT 0 /* participates_in, G.out_edges order  (entity -> hyperedge) */
#define AHR_CSR_PART_IN 1  /* participates_in, G.in_edges order */
#define AHR_CSR_BEL_OUT 2  /* belongs_to out  (child -> parent) */
#define AHR_CSR_BEL_IN 3   /* belongs_to in   (parent <- members) */
#define AHR_CSR_REL_OUT 4  /* related_to out */
#define AHR_CSR_REL_IN 5   /* related_to in */
#define AHR_N_CSR 6

/* node types (uint8 attribute array) */
#define AHR_NT_ENTITY 0
#define AHR_NT_SUMMARY 1
#define AHR_NT_HYPEREDGE 2
#define AHR_NT_OTHER 3 /* node_type missing / unknown (d.get("node_type") is None) */

/* how an entry of the expanded map was inserted (ahr_anchor_rerank, exp_kind) */
#define AHR_EK_SEED 1
#define AHR_EK_PARENT 2
#define AHR_EK_MEMBER 3

/* expand ops */
#define AHR_OP_PARENTS 0  /* GraphEnvironment.expand_parents   src/ah_rag/agent/environment.py:417-438 */
#define AHR_OP_CHILDREN 1 /* GraphEnvironment.expand_children  src/ah_rag/agent/environment.py:393-415 */
#define AHR_OP_RELATED 2  /* GraphEnvironment.expand_related   src/ah_rag/agent/environment.py:440-480 */

typedef struct ahr_index ahr_index;
typedef struct ahr_graph ahr_graph;

const char *ahr_last_error(void);
int ahr_version(void);
/* number of CUDA kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t ahr_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Embedding index.  Replaces the Chroma collection built at
 *   src/ah_rag/graph/hierarchical_graph.py:605-611,660  (PersistentClient / get_or_create_collection
 *   {"hnsw:space": "cosine"} / Collection.upsert)
 * by a device-resident row-major [n_rows, dim] matrix the caller already holds.  Rows are expected
 * L2-normalised by the producer, as the reference does (:655-656); ids are id_base + row.
 * dim must be a multiple of 8 for the tensor-core path (16-byte TMA row pitch); other dims fall back to the
 * brute-force exact scan.  For fp32 rows the library additionally keeps its own hi/lo bf16 split of the matrix
 * (2 x n_rows x dim x 2 bytes) so that fp32 indexes also run on the tensor cores (3 operand segments per tile).
 * ---------------------------------------------------------------------------------------------- */
int ahr_index_create(ahr_index **out, int device, int64_t n_rows, int32_t dim, int dtype,
                     const void *dev_rows, int64_t id_base);
int ahr_index_destroy(ahr_index *idx);

/*
 * Batched nearest-neighbour query.  Replaces, for nq queries at once,
 *   src/ah_rag/graph/hierarchical_graph.py:714   coll.query(query_embeddings=[q], n_results=k)
 * dev_q: [nq, dim] in the index dtype.  Outputs [nq, k]: scores = <q, x> (the reference's
 * `semantic` is 1 - fl32(1 - score), :723) and ids, sorted by (score desc, id asc); lists shorter than k
 * (n_rows < k) are padded with (-inf, -1).
 *   AHR_MODE_BF16 : k <= 120.  scores are the tensor-core fp32 accumulations (fp32 indexes: hi/lo split, ~2^-16).
 *   AHR_MODE_EXACT: canonical scores, bit-reproducible; ids bit-identical to the CPU oracle.  The tensor cores
 *                   propose k + slack candidates, which are re-scored canonically and CERTIFIED (k-th exact score
 *                   above the last candidate's tensor score + error bound); queries that cannot be certified, and
 *                   shapes the tensor path does not take (k > 96, n_rows <= 4096, dim % 8 != 0), run the brute-force
 *                   fp64 scan.  = ahr_topk_exact_begin + ahr_topk_exact_finish: waits once for the certification count
 *                   (and again per fallback tier when some query could not be certified).
 */
int ahr_topk(ahr_index *idx, const void *dev_q, int32_t nq, int32_t k, int mode, float *dev_scores,
             int64_t *dev_ids, void *stream);

/*
 * AHR_MODE_EXACT in two halves, for callers that overlap the query with other work or capture it in a CUDA graph:
 *   begin   enqueues the tensor-core candidates, the canonical rescoring and the certification on `stream` and returns
 *           WITHOUT synchronising (shapes the tensor path does not take: the brute-force scan, also just enqueued);
 *           the number of uncertified queries travels to a pinned host word behind an event;
 *   finish  waits for that event; in the common case (every query certified) that is all it does -- dev_scores / dev_ids
 *           are final in stream order.  Otherwise it runs the wider candidate tier and the brute-force scan for the
 *           uncertified queries (these synchronise `stream`).  One exact query per index may be pending at a time.
 */
int ahr_topk_exact_begin(ahr_index *idx, const void *dev_q, int32_t nq, int32_t k, float *dev_scores, int64_t *dev_ids,
                         void *stream);
int ahr_topk_exact_finish(ahr_index *idx);

/* Same query through HOST buffers: copies host_q to the device, runs ahr_topk, copies results back and
 * synchronises the stream.  This is the call bench.py times for the end-to-end ("e2e") figure. */
int ahr_topk_host(ahr_index *idx, const void *host_q, int32_t nq, int32_t k, int mode, float *host_scores,
                  int64_t *host_ids, void *stream);

/*
 * k-way merge of `parts` candidate lists per query (multi-GPU: the lists all-gathered from the
 * node-range shards).  dev_scores/dev_ids: [parts, nq, k], every list sorted by (score desc, id asc) with its empty
 * entries (id < 0) at the end, exactly as ahr_topk writes it.
 * Output [nq, k] sorted by (score desc, id asc) -- identical for every sharding of the same corpus.
 */
int ahr_merge_topk(int device, const float *dev_scores, const int64_t *dev_ids, int32_t parts, int32_t nq,
                   int32_t k, float *dev_out_scores, int64_t *dev_out_ids, void *stream);

/* Same merge over the PACKED exchange buffer of the sharded path: `parts` consecutive records of
 * record_bytes each, a record = [nq*k fp32 scores][nq*k int64 ids] exactly as one rank contributes it to the
 * all-gather (record_bytes >= nq*k*12, nq*k*4 must be a multiple of 8).  Avoids re-packing copies. */
int ahr_merge_topk_packed(int device, const void *dev_packed, int64_t record_bytes, int32_t parts, int32_t nq,
                          int32_t k, float *dev_out_scores, int64_t *dev_out_ids, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Knowledge-graph CSR.  Replaces the NetworkX adjacency reads of
 *   src/ah_rag/graph/hierarchical_graph.py:441-462 (_neighbors_by_edge_type, get_belongs_to,
 *   get_summary_members) and src/ah_rag/agent/environment.py:401-402,424-425,448-467.
 * Dense node id = position in G.nodes order of the loaded graph.  For each of the AHR_N_CSR slots:
 * row_ptr[n_nodes + 1] (int64) and col_idx (int32), neighbours in adjacency iteration order.
 * Attributes (per node): node_type (AHR_NT_*), level (0 if absent), and fp64 arrays
 *   judge       overall judge score, NaN when absent (judge is None)       hierarchical_graph.py:746-752
 *   judge_term  1/(1+exp(-judge/10)) evaluated by the host exactly as :766 does, 0.0 when absent
 *   conf        confidence, NaN when absent                                :753
 * The CSR arrays are borrowed (they must outlive the handle); the five attribute arrays are copied at create time into
 * one 64-byte line per node together with the first 8 entries of the list a seed of that type expands into (belongs_to
 * parents of an entity / members of a summary), so a seed costs one line instead of a chain of four dependent sectors;
 * the attribute arrays are not referenced afterwards.
 * ---------------------------------------------------------------------------------------------- */
int ahr_graph_create(ahr_graph **out, int device, int64_t n_nodes, const int64_t *const dev_row_ptr[AHR_N_CSR],
                     const int32_t *const dev_col_idx[AHR_N_CSR], const uint8_t *dev_node_type,
                     const uint8_t *dev_level, const double *dev_judge, const double *dev_judge_term,
                     const double *dev_conf);
int ahr_graph_destroy(ahr_graph *g);

typedef struct ahr_rerank_params {
  double alpha, beta, gamma, delta; /* hierarchical_graph.py:683-686 */
  double boost[4];                  /* layer_boost per AHR_NT_* (index 3 = unknown type -> 0.0)  :699 */
  double judge_min, conf_min;       /* thresholds, used when the has_* flag is set              :759-762 */
  int32_t has_judge_min, has_conf_min;
  int32_t type_mask;                /* bit t set = node type t passes; 0 = no type filter        :756-757 */
  int32_t member_top_m;             /* summary -> first m members                                :737 */
  int32_t parent_cap;               /* entity -> first 2 parents                                 :734 */
  int32_t reserved;
} ahr_rerank_params;

/*
 * Anchor expansion + rerank for nq queries.  Replaces, per query,
 *   src/ah_rag/graph/hierarchical_graph.py:719-789  (seeds -> expanded -> reranked)
 * Inputs: seed_nodes [nq, n_seeds] dense node ids in rank order (-1 = no seed), seed_scores [nq, n_seeds]
 * fp32 similarity scores.  cap = n_seeds * (1 + max(parent_cap, member_top_m)) bounds the expanded map.
 * Outputs (any of the exp_* / seed_sem pointers may be NULL):
 *   out_nodes/out_semantic/out_score [nq, top_k], out_count [nq]: the reranked list; semantic and score
 *     are already rounded to 4 decimals exactly as Python's round() does (:769-776), order = stable sort
 *     on the rounded score descending (:781);
 *   exp_nodes/exp_semantic/exp_kind [nq, cap], exp_count [nq]: the `expanded` map in insertion order with
 *     the unrounded fp64 semantic (return_cluster=True, :783-788); kind = AHR_EK_SEED / AHR_EK_PARENT
 *     (inserted by the entity rule, node_type reported as "summary", :736) / AHR_EK_MEMBER (summary rule,
 *     reported as "entity", :740);
 *   seed_semantic [nq, n_seeds]: 1.0 - (double)fl32(1 - score)  (:723).
 */
int ahr_anchor_rerank(ahr_graph *g, int32_t nq, int32_t n_seeds, const int32_t *dev_seed_nodes,
                      const float *dev_seed_scores, const ahr_rerank_params *params, int32_t top_k,
                      int32_t *dev_out_nodes, double *dev_out_semantic, double *dev_out_score,
                      int32_t *dev_out_count, int32_t *dev_exp_nodes, double *dev_exp_semantic,
                      uint8_t *dev_exp_kind, int32_t *dev_exp_count, double *dev_seed_semantic, void *stream);

/*
 * Same call fed straight from ahr_topk: dev_seed_rows [nq, n_seeds] are the int64 ids of an index created with
 * id_base = 0 (index rows, -1 = empty slot), dev_row_to_node [n_rows] maps an index row to its dense node id (-1 = the
 * indexed id is not a node of the graph) -- the lookup hierarchical_graph.py:731 does with G.nodes.get(nid).  Saves the
 * gather / cast launches between the two kernels of a search.
 */
int ahr_anchor_rerank_rows(ahr_graph *g, int32_t nq, int32_t n_seeds, const int64_t *dev_seed_rows,
                           const int32_t *dev_row_to_node, int64_t n_rows, const float *dev_seed_scores,
                           const ahr_rerank_params *params, int32_t top_k, int32_t *dev_out_nodes,
                           double *dev_out_semantic, double *dev_out_score, int32_t *dev_out_count, int32_t *dev_exp_nodes,
                           double *dev_exp_semantic, uint8_t *dev_exp_kind, int32_t *dev_exp_count,
                           double *dev_seed_semantic, void *stream);

/*
 * Batched frontier expansion.  Replaces, for n_env environments at once,
 *   GraphEnvironment.expand_parents / expand_children / expand_related
 *   (src/ah_rag/agent/environment.py:393-480): ordered concatenation of typed neighbour lists of the input
 *   nodes, stop after the input node at which the un-deduplicated length reaches `limit`,
 *   order-preserving dedup, cut to `limit`.
 * dev_ops[n_env] (AHR_OP_*), inputs in CSR form: dev_in_ids[dev_in_offsets[e] .. dev_in_offsets[e+1])
 * (-1 = id not in the graph, skipped like environment.py:398).  Outputs: dev_out_ids [n_env, limit]
 * (padded with -1) and dev_out_counts [n_env].
 */
int ahr_expand(ahr_graph *g, int32_t n_env, const int32_t *dev_ops, const int32_t *dev_in_ids,
               const int32_t *dev_in_offsets, int32_t limit, int32_t *dev_out_ids, int32_t *dev_out_counts,
               void *stream);

/*
 * Closure step of GraphEnvironment.expand_to_lca (src/ah_rag/agent/environment.py:243-299) over the belongs_to CSRs:
 * for every input node the set of nodes below it (transitive members, itself included -- what nx.ancestors yields on
 * child -> parent edges), the intersection of those sets, and of it the nodes without a belongs_to parent inside it.
 * host_ids [n_ids]: dense node ids on the HOST (a handful per call; copied before the call returns).  dev_out_flags
 * [n_nodes] u8: 1 = result node.  The walk is two launches on the caller's stream (all inputs and all BFS levels in one
 * kernel).  n_result may be NULL: the call is then fully asynchronous; otherwise the count of result nodes is read back,
 * which synchronises the stream once.  The caller orders the result ((level or 1, node id), environment.py:283) and checks
 * that every input has a belongs_to edge (nx.ancestors raises otherwise).
 */
int ahr_lca_closure(ahr_graph *g, const int32_t *host_ids, int32_t n_ids, uint8_t *dev_out_flags, int64_t *n_result,
                    void *stream);

/* ------------------------------------------------------------------------------------------------
 * Device-resident batched gym environment.  Replaces, for n_env environments per launch,
 *   AHRAGGymEnv.reset / .step / .get_action_mask        src/ah_rag/agent/gym_env.py:41-125
 *   featurize_observation                                src/ah_rag/agent/featurizer.py:21-56
 *   step_reward                                          src/ah_rag/agent/reward.py:6-26
 * over GraphEnvironment's state rules (src/ah_rag/agent/environment.py:155-241, :301-355, :393-480).
 * All arrays are caller-owned device memory (the Python wrapper allocates them once per batch):
 *   sel_ids [n_env, cap_sel] / sel_cnt [n_env]   selection_set        fr_ids [n_env, cap_fr] / fr_cnt   frontier_set
 *   shown_ids / shown_score / shown_sem [n_env, 10], shown_cnt       the nodes of the current obs["selection"]
 *   n_seeds, obs_step, obs_sel, obs_fr [n_env]                       len(obs["seeds"]), obs["step"], sizes in obs["state"]
 *                                                                    (-1: the observation has no "state", :325-333)
 *   env_steps, gym_steps, last_action [n_env], done [n_env], overflow [1]
 * ---------------------------------------------------------------------------------------------- */
typedef struct ahr_env_batch {
  int32_t n_env, cap_sel, cap_fr, reserved;
  int32_t *sel_ids, *sel_cnt, *fr_ids, *fr_cnt;
  int32_t *shown_ids;
  double *shown_score, *shown_sem;
  int32_t *shown_cnt, *n_seeds, *obs_step, *obs_sel, *obs_fr, *env_steps, *gym_steps, *last_action;
  uint8_t *done;
  int32_t *overflow;
} ahr_env_batch;

/* reset from the output of ahr_anchor_rerank run with top_k (gym: 5): dev_nodes/dev_semantic/dev_score [n_env, top_k],
 * dev_count [n_env]; n_seeds = number of seeds every query had.  Writes features [n_env, 84] fp32 and the action mask
 * [n_env, 6] u8. */
int ahr_env_reset(ahr_graph *g, const ahr_env_batch *env, const int32_t *dev_nodes, const double *dev_semantic,
                  const double *dev_score, const int32_t *dev_count, int32_t top_k, int32_t n_seeds,
                  float *dev_features, uint8_t *dev_mask, void *stream);
/* one gym step per environment: dev_actions [n_env] in 0..5.  Writes features [n_env, 84], rewards [n_env] f64,
 * done [n_env] u8, mask [n_env, 6] u8. */
int ahr_env_step(ahr_graph *g, const ahr_env_batch *env, const int32_t *dev_actions, int32_t max_steps,
                 double repeat_penalty, float *dev_features, double *dev_rewards, uint8_t *dev_done, uint8_t *dev_mask,
                 void *stream);

/* ------------------------------------------------------------------------------------------------
 * Cross-GPU candidate exchange over NVLink peer memory (one process per GPU, one box).
 * The reference has no multi-GPU path; this is the scaling of the Collection.query seam
 * (src/ah_rag/graph/hierarchical_graph.py:714) over an index sharded by node range: every rank's
 * [nq, k] candidate list is stored directly into the peers' exchange buffers, a flag per (half, source) carries
 * the epoch, and the k-way merge runs on every rank.  No collective library call on the data path.
 *
 *   create  -> export (64-byte CUDA IPC handle) -> all ranks swap handles (any transport) -> connect
 *   per batch: ahr_comm_begin gives this rank's slot; run ahr_topk with the slot as its output (global ids);
 *              ahr_comm_exchange_merge pushes the slot to the peers, waits for theirs, merges into out.
 * ---------------------------------------------------------------------------------------------- */
#define AHR_IPC_HANDLE_BYTES 64
typedef struct ahr_comm ahr_comm;
int ahr_comm_create(ahr_comm **out, int device, int rank, int world, int64_t max_record_bytes /* >= nq*k*12 */);
int ahr_comm_export(ahr_comm *c, void *handle_out /* AHR_IPC_HANDLE_BYTES */);
int ahr_comm_connect(ahr_comm *c, const void *all_handles /* world x AHR_IPC_HANDLE_BYTES, rank order */);
int ahr_comm_destroy(ahr_comm *c);
/* start an exchange of [nq, k] lists (nq*k %% 4 == 0): returns device pointers to this rank's slot of the NEXT exchange
 * (the exchange counter only advances in ahr_comm_exchange_merge / ahr_topk_sharded) */
int ahr_comm_begin(ahr_comm *c, int32_t nq, int32_t k, float **dev_slot_scores, int64_t **dev_slot_ids);
/* push the slot to every peer, wait for every peer's slot of this exchange, merge: out [nq, k] identical on all ranks */
int ahr_comm_exchange_merge(ahr_comm *c, int32_t nq, int32_t k, float *dev_out_scores, int64_t *dev_out_ids, void *stream);
/*
 * The whole sharded query in one call: local top-k over this rank's node range (ids = id_base + row), exchange, merge;
 * out [nq, k] identical on every rank and for every shard count.  In AHR_MODE_BF16 the kernel that reduces the slab lists
 * of the local top-k stores its result straight into the peers' buffers and publishes the flags (no separate push
 * launch); the merge is a rank-counting merge of the `world` sorted lists.  Arguments are validated before the exchange
 * counter moves; a rank that fails afterwards poisons its flags so that the peers stop at once.  Every rank must call it
 * with the same nq, k and mode.
 */
int ahr_topk_sharded(ahr_index *idx, ahr_comm *c, const void *dev_q, int32_t nq, int32_t k, int mode, float *dev_out_scores,
                     int64_t *dev_out_ids, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Knowledge-graph CSR SHARDED BY NODE RANGE (one process per GPU): rank r owns the adjacency rows and node types of the
 * dense ids [lo, hi).  The reference has no multi-GPU path; this is GraphEnvironment.expand_parents / expand_children /
 * expand_related (src/ah_rag/agent/environment.py:393-480) over a graph that does not fit, or is not wanted, on every GPU.
 * The batch inputs (ops, input ids in CSR form) are replicated on the ranks, like the queries of the sharded index.
 *
 *   stage1  (owner of an input node)   unique prefix of the node's first-level list, its saturated length, and for
 *                                      expand_related on an entity its first `limit` hyperedges
 *   -- the caller all-gathers the stage-1 records of all ranks (NCCL / gloo / peer copies: any transport) --
 *   stage2  (owner of a hyperedge)     unique prefix of the hyperedge's other participants (the 2-hop walk, :456-467)
 *   -- all-gather of the stage-2 records --
 *   merge   (any rank)                 per environment, in input order: prefixes appended, stop rule, dedup, cut
 *
 * Records are int32: stage 1 n_env x ahr_expand_shard_record_ints(1, limit, max_ids), stage 2 n_env x
 * ahr_expand_shard_record_ints(2, ...); gathered buffers are [world][n_env][...] in rank order.  dev_bounds [world + 1]
 * int64: bounds[r] = lo of rank r, bounds[world] = n_nodes.  limit <= 32 (the reference uses 10), at most max_ids input
 * nodes per environment are read.  Results are bit-identical to ahr_expand on the unsharded graph for every world size
 * provided a participates_in row holds distinct hyperedges (true of any CSR exported from a DiGraph).
 * row_ptr_local[s] has hi - lo + 1 entries (row v - lo), col_idx_local[s] holds GLOBAL dense ids.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ahr_graph_shard ahr_graph_shard;
int ahr_graph_shard_create(ahr_graph_shard **out, int device, int64_t n_nodes_total, int64_t lo, int64_t hi,
                           const int64_t *const dev_row_ptr_local[AHR_N_CSR], const int32_t *const dev_col_idx_local[AHR_N_CSR],
                           const uint8_t *dev_node_type_local);
int ahr_graph_shard_destroy(ahr_graph_shard *g);
int64_t ahr_expand_shard_record_ints(int32_t stage, int32_t limit, int32_t max_ids);
int ahr_expand_shard_stage1(ahr_graph_shard *g, int32_t n_env, const int32_t *dev_ops, const int32_t *dev_in_ids,
                            const int32_t *dev_in_offsets, int32_t max_ids, int32_t limit, int32_t *dev_out_records,
                            void *stream);
int ahr_expand_shard_stage2(ahr_graph_shard *g, int32_t n_env, int32_t world, const int64_t *dev_bounds, const int32_t *dev_ops,
                            const int32_t *dev_in_ids, const int32_t *dev_in_offsets, int32_t max_ids, int32_t limit,
                            const int32_t *dev_gathered1, int32_t *dev_out_records, void *stream);
int ahr_expand_shard_merge(int device, int64_t n_nodes_total, int32_t n_env, int32_t world, const int64_t *dev_bounds,
                           const int32_t *dev_ops, const int32_t *dev_in_ids, const int32_t *dev_in_offsets, int32_t max_ids,
                           int32_t limit, const int32_t *dev_gathered1, const int32_t *dev_gathered2, int32_t *dev_out_ids,
                           int32_t *dev_out_counts, void *stream);

/*
 * Anchor expansion + rerank (ahr_anchor_rerank, hierarchical_graph.py:719-789) over the SHARDED graph.  The seeds
 * (dense node ids, the merged top-k of the sharded index, identical on every rank) are replicated.
 *   ahr_graph_shard_set_attrs   the owned slices of level / judge / judge_term / conf (see ahr_graph_create), borrowed
 *   stage1  (owner of a seed)        its type + attributes + the first parent_cap parents / member_top_m members
 *   -- all-gather --
 *   stage2  (owner of a neighbour)   type + attributes of every node stage 1 emitted
 *   -- all-gather --
 *   merge   (any rank)               the rerank kernel of ahr_anchor_rerank fed with the gathered records: same outputs,
 *                                    bit-identical to the unsharded call for every world size
 * Records are int32, n_q x ahr_anchor_shard_record_ints(stage, n_seeds, params) per rank.
 */
int ahr_graph_shard_set_attrs(ahr_graph_shard *g, const uint8_t *dev_level_local, const double *dev_judge_local,
                              const double *dev_judge_term_local, const double *dev_conf_local);
int64_t ahr_anchor_shard_record_ints(int32_t stage, int32_t n_seeds, const ahr_rerank_params *params);
int ahr_anchor_shard_stage1(ahr_graph_shard *g, int32_t nq, int32_t n_seeds, const int32_t *dev_seed_nodes,
                            const ahr_rerank_params *params, int32_t *dev_out_records, void *stream);
int ahr_anchor_shard_stage2(ahr_graph_shard *g, int32_t nq, int32_t n_seeds, int32_t world, const int64_t *dev_bounds,
                            const int32_t *dev_seed_nodes, const ahr_rerank_params *params, const int32_t *dev_gathered1,
                            int32_t *dev_out_records, void *stream);
int ahr_anchor_shard_merge(ahr_graph_shard *g, int32_t nq, int32_t n_seeds, int32_t world, const int64_t *dev_bounds,
                           const int32_t *dev_seed_nodes, const float *dev_seed_scores, const ahr_rerank_params *params,
                           int32_t top_k, const int32_t *dev_gathered1, const int32_t *dev_gathered2, int32_t *dev_out_nodes,
                           double *dev_out_semantic, double *dev_out_score, int32_t *dev_out_count, int32_t *dev_exp_nodes,
                           double *dev_exp_semantic, uint8_t *dev_exp_kind, int32_t *dev_exp_count, double *dev_seed_semantic,
                           void *stream);

/* ---- diagnostics used by the parity tests (not part of the reference seam) ---------------------- */
/* full [nq, n_rows] score matrix of the tensor-core path (small inputs only) */
int ahr_debug_tc_scores(ahr_index *idx, const void *dev_q, int32_t nq, float *dev_out, void *stream);
/* canonical scores of given candidate rows: cand [nq, m] row ids (-1 -> -inf) */
int ahr_rescore(ahr_index *idx, const void *dev_q, int32_t nq, int32_t m, const int64_t *dev_cand_ids,
                float *dev_out_scores, void *stream);
/* Python round(x, 4) on the device, elementwise (the rerank kernel's rounding routine) */
int ahr_debug_round4(int device, const double *dev_x, int64_t n, double *dev_out, void *stream);
/* statistics of the last AHR_MODE_EXACT call on this index: queries certified by the tensor-core
 * candidate path vs. re-run through the brute-force exact scan */
int ahr_exact_stats(ahr_index *idx, int64_t *n_certified, int64_t *n_fallback);

/* CUDA-event timing of the dominant kernel (the fused GEMM + top-k, K1) on the stream it is launched on.
 * ahr_get_profile waits for the recorded events, returns the summed duration (ms), the number of K1
 * launches since the last call and the slab count of the last launch, and resets the record. */
int ahr_set_profiling(ahr_index *idx, int enable);
int ahr_get_profile(ahr_index *idx, double *k1_ms_total, int64_t *k1_launches, int32_t *n_slabs);

#ifdef __cplusplus
}
#endif
#endif /* AHRAG_B200_H_ */
