// graph_shared.cuh -- warp-cooperative ordered neighbour walk shared by graph.cu (ahr_expand) and env.cu (the
// device-resident batched environment).  Semantics: src/ah_rag/agent/environment.py:393-480.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ahr {

constexpr int EXP_WARPS = 4;

struct ExpandState {
  int32_t *out;  // this warp's list in shared memory
  int count;
  int raw;
  int limit;
};

// append values to the order-preserving unique list; returns true once the list is full
// (limits up to 64: the list is scanned per candidate; larger limits take the hash-set kernel of graph.cu)
__device__ __forceinline__ bool expand_push(ExpandState &st, int32_t val, bool valid, int lane) {
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
  st.raw += __popc(vmask);
  // duplicates inside this batch of 32: the lowest lane carrying the value is the first occurrence
  // (MATCH.ANY costs ~2 cycles of an SM-wide unit per DISTINCT value, tools/ub/ub_match.cu: the lanes without a value share
  // one sentinel instead of 32 different ones)
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? val : (int32_t)-2);
  const bool first_in_batch = (__ffs(peers & vmask) - 1) == lane;
  bool dup = false;
  if (valid)
    for (int t = 0; t < st.count; ++t) dup |= (st.out[t] == val);
  const bool fresh = valid && !dup && first_in_batch;
  const unsigned nmask = __ballot_sync(0xffffffffu, fresh);
  const int pos = st.count + __popc(nmask & ((1u << lane) - 1u));
  AHR_DEV_ASSERT(pos >= 0);
  if (fresh && pos < st.limit) st.out[pos] = val;
  st.count = min(st.limit, st.count + __popc(nmask));
  __syncwarp();
  return st.count >= st.limit;
}

__device__ __forceinline__ bool expand_segment(ExpandState &st, const GraphView &g, int slot, int32_t v, int32_t exclude,
                                               int lane) {
  const int64_t p0 = g.row_ptr[slot][v], p1 = g.row_ptr[slot][v + 1];
  for (int64_t b = p0; b < p1; b += 32) {
    const int64_t i = b + lane;
    int32_t val = -1;
    bool valid = i < p1;
    if (valid) {
      val = g.col_idx[slot][i];  // coalesced: consecutive lanes read consecutive neighbours
      AHR_DEV_ASSERT(val >= 0 && val < g.n_nodes);
      valid = val != exclude;
    }
    if (expand_push(st, val, valid, lane)) return true;
  }
  return false;
}


// The whole expand_* call for one environment: ordered concat of the typed neighbour lists of `ids`, stop rule on the
// un-deduplicated length after each input node, order-preserving dedup, cut to st.limit.  All 32 lanes call.
__device__ __forceinline__ void expand_walk(ExpandState &st, const GraphView &g, int op, const int32_t *ids, int n_ids,
                                            int lane) {
  bool full = false;
  for (int idx = 0; idx < n_ids && !full; ++idx) {
    const int32_t v = ids[idx];
    if (v < 0 || v >= g.n_nodes) continue;  // `if nid not in G: continue`
    if (op == AHR_OP_PARENTS) {
      full = expand_segment(st, g, AHR_CSR_BEL_OUT, v, -1, lane);
    } else if (op == AHR_OP_CHILDREN) {
      full = expand_segment(st, g, AHR_CSR_BEL_IN, v, -1, lane);
    } else {
      full = expand_segment(st, g, AHR_CSR_REL_OUT, v, -1, lane);
      if (!full) full = expand_segment(st, g, AHR_CSR_REL_IN, v, -1, lane);
      if (!full && g.line[v].head.type == AHR_NT_ENTITY) {
        // second hop: every hyperedge h of v, then the other participants of h  (environment.py:456-467)
        const int64_t h0 = g.row_ptr[AHR_CSR_PART_OUT][v], h1 = g.row_ptr[AHR_CSR_PART_OUT][v + 1];
        for (int64_t hi = h0; hi < h1 && !full; ++hi) {
          const int32_t h = g.col_idx[AHR_CSR_PART_OUT][hi];
          full = expand_push(st, h, lane == 0, lane);
          if (!full) full = expand_segment(st, g, AHR_CSR_PART_IN, h, v, lane);
        }
      }
    }
    if (st.raw >= st.limit) break;  // stop rule on the UN-deduplicated length, checked after finishing v
  }
}

}  // namespace ahr
