// lca.cu -- closure step of GraphEnvironment.expand_to_lca (src/ah_rag/agent/environment.py:243-299) on the device CSR.
//
// The reference builds a DiGraph of the belongs_to edges (child -> parent) per call and takes, for every input node,
// nx.ancestors(dag, nid) | {nid} -- on child -> parent edges that is the set of nodes BELOW the input (its transitive
// members) plus the input -- intersects those sets and keeps the members of the intersection that have no belongs_to
// parent inside it.  Here:
//   per input i (tag = i + 1): level-synchronous BFS over the in-belongs_to CSR.  vis[v] = tag marks "reached from
//     input i" (atomicExch: exactly one thread sees the old value and appends v to the next frontier);
//     seen[v] goes tag - 1 -> tag only for nodes already in every previous closure, so after the last input
//     seen[v] == n_ids  <=>  v is in the intersection.  One host round trip per BFS level (the frontier size).
//   final pass: v with seen[v] == n_ids and no out-belongs_to parent p with seen[p] == n_ids.
// The hierarchy is a few levels deep, so a call is a handful of small launches; the CSR is the one already resident
// for the hop kernels (no per-call graph rebuild).
#include "common.cuh"
#include "kernels.h"

namespace ahr {

__global__ void lca_seed_kernel(int32_t *seen, int32_t *vis, int32_t *frontier, int32_t v, int32_t tag) {
  vis[v] = tag;
  if (seen[v] == tag - 1) seen[v] = tag;
  frontier[0] = v;
}

__global__ void __launch_bounds__(256)
lca_level_kernel(GraphView g, const int32_t *__restrict__ frontier_in, int32_t n_in, int32_t *__restrict__ frontier_out,
                 int32_t *__restrict__ count_out, int32_t *__restrict__ seen, int32_t *__restrict__ vis, int32_t tag) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= n_in) return;
  const int32_t u = frontier_in[warp];
  const int64_t p0 = g.row_ptr[AHR_CSR_BEL_IN][u], p1 = g.row_ptr[AHR_CSR_BEL_IN][u + 1];
  for (int64_t i = p0 + lane; i < p1; i += 32) {  // coalesced: consecutive lanes read consecutive members
    const int32_t c = g.col_idx[AHR_CSR_BEL_IN][i];
    if (atomicExch(&vis[c], tag) != tag) {  // first visit from this input
      frontier_out[atomicAdd(count_out, 1)] = c;
      if (seen[c] == tag - 1) seen[c] = tag;
    }
  }
}

__global__ void lca_final_kernel(GraphView g, const int32_t *__restrict__ seen, int32_t n_ids, uint8_t *__restrict__ out_flags,
                                 unsigned long long *__restrict__ n_result) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.n_nodes) return;
  uint8_t keep = 0;
  if (seen[v] == n_ids) {
    keep = 1;
    for (int64_t i = g.row_ptr[AHR_CSR_BEL_OUT][v]; i < g.row_ptr[AHR_CSR_BEL_OUT][v + 1]; ++i)
      if (seen[g.col_idx[AHR_CSR_BEL_OUT][i]] == n_ids) { keep = 0; break; }
    if (keep) atomicAdd(n_result, 1ull);
  }
  out_flags[v] = keep;
}

// workspace: seen, vis, frontier A, frontier B (int32 [n_nodes] each) + 2 counters
size_t graph_lca_workspace_bytes(int64_t n_nodes) { return ((size_t)n_nodes * 16 + 255 + 64) & ~(size_t)255; }

int graph_lca_closure(const GraphView &g, const int32_t *host_ids, int32_t n_ids, uint8_t *out_flags, int64_t *n_result,
                      void *workspace, cudaStream_t stream) {
  const int64_t n = g.n_nodes;
  int32_t *seen = reinterpret_cast<int32_t *>(workspace);
  int32_t *vis = seen + n, *fa = vis + n, *fb = fa + n;
  int32_t *count = fb + n;  // [0]: next frontier size
  unsigned long long *res = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(count + 2) + 7) & ~(uintptr_t)7);
  AHR_CUDA_CHECK(cudaMemsetAsync(seen, 0, (size_t)n * 8, stream));  // seen + vis
  for (int32_t i = 0; i < n_ids; ++i) {
    const int32_t tag = i + 1;
    lca_seed_kernel<<<1, 1, 0, stream>>>(seen, vis, fa, host_ids[i], tag);
    AHR_KERNEL_CHECK();
    int32_t n_front = 1;
    int32_t *fin = fa, *fout = fb;
    while (n_front > 0) {
      AHR_CUDA_CHECK(cudaMemsetAsync(count, 0, 4, stream));
      const int64_t threads = (int64_t)n_front * 32;
      lca_level_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(g, fin, n_front, fout, count, seen, vis, tag);
      AHR_KERNEL_CHECK();
      AHR_CUDA_CHECK(cudaMemcpyAsync(&n_front, count, 4, cudaMemcpyDeviceToHost, stream));
      AHR_CUDA_CHECK(cudaStreamSynchronize(stream));
      int32_t *t = fin; fin = fout; fout = t;
    }
  }
  AHR_CUDA_CHECK(cudaMemsetAsync(res, 0, 8, stream));
  lca_final_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(g, seen, n_ids, out_flags, res);
  AHR_KERNEL_CHECK();
  unsigned long long h = 0;
  AHR_CUDA_CHECK(cudaMemcpyAsync(&h, res, 8, cudaMemcpyDeviceToHost, stream));
  AHR_CUDA_CHECK(cudaStreamSynchronize(stream));
  *n_result = (int64_t)h;
  return AHR_OK;
}

}  // namespace ahr
