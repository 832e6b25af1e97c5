// lca.cu -- closure step of GraphEnvironment.expand_to_lca (src/ah_rag/agent/environment.py:243-299) on the device CSR.
//
// The reference builds a DiGraph of the belongs_to edges (child -> parent) per call and takes, for every input node,
// nx.ancestors(dag, nid) | {nid} -- on child -> parent edges that is the set of nodes BELOW the input (its transitive
// members) plus the input -- intersects those sets and keeps the members of the intersection that have no belongs_to
// parent inside it.  Here:
//   per input i (tag = i + 1): level-synchronous BFS over the in-belongs_to CSR.  vis[v] = tag marks "reached from
//     input i" (atomicExch: exactly one thread sees the old value and appends v to the next frontier);
//     seen[v] goes tag - 1 -> tag only for nodes already in every previous closure, so after the last input
//     seen[v] == n_ids  <=>  v is in the intersection.
//   The whole walk -- every input, every level -- is ONE launch of one CTA that loops over the levels with block
//   barriers (frontier sizes live in shared memory): no host round trip per level, nothing to synchronise.  The
//   hierarchy is a few levels deep and a closure holds at most a few hundred thousand nodes, which 1024 threads walk in
//   well under a millisecond; the CSR is the one already resident for the hop kernels (no per-call graph rebuild).
//   final pass: v with seen[v] == n_ids and no out-belongs_to parent p with seen[p] == n_ids.
#include "common.cuh"
#include "kernels.h"

namespace ahr {

constexpr int LCA_THREADS = 1024;

__global__ void __launch_bounds__(LCA_THREADS)
lca_closure_kernel(GraphView g, const int32_t *__restrict__ ids, int32_t n_ids, int32_t *__restrict__ fa,
                   int32_t *__restrict__ fb, int32_t *__restrict__ seen, int32_t *__restrict__ vis) {
  __shared__ int s_front, s_next;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int N_WARPS = LCA_THREADS / 32;
  for (int32_t i = 0; i < n_ids; ++i) {
    const int32_t tag = i + 1;
    int32_t *fin = fa, *fout = fb;
    if (threadIdx.x == 0) {
      const int32_t v = ids[i];
      vis[v] = tag;
      if (seen[v] == tag - 1) seen[v] = tag;
      fin[0] = v;
      s_front = 1;
      s_next = 0;
    }
    __syncthreads();
    for (;;) {
      const int n_front = s_front;
      if (n_front == 0) break;  // uniform: read between two barriers
      for (int f = warp; f < n_front; f += N_WARPS) {
        const int32_t u = fin[f];
        const int64_t p0 = g.row_ptr[AHR_CSR_BEL_IN][u], p1 = g.row_ptr[AHR_CSR_BEL_IN][u + 1];
        for (int64_t e = p0 + lane; e < p1; e += 32) {  // coalesced: consecutive lanes read consecutive members
          const int32_t c = g.col_idx[AHR_CSR_BEL_IN][e];
          if (atomicExch(&vis[c], tag) != tag) {  // first visit from this input
            fout[atomicAdd(&s_next, 1)] = c;
            if (seen[c] == tag - 1) seen[c] = tag;
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) { s_front = s_next; s_next = 0; }
      int32_t *t = fin; fin = fout; fout = t;
      __syncthreads();
    }
    __syncthreads();
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

// workspace: seen, vis, frontier A, frontier B (int32 [n_nodes] each), the result counter, the input ids
size_t graph_lca_workspace_bytes(int64_t n_nodes, int32_t n_ids) {
  return ((size_t)n_nodes * 16 + 64 + (size_t)n_ids * 4 + 255) & ~(size_t)255;
}

int graph_lca_closure(const GraphView &g, const int32_t *host_ids, int32_t n_ids, uint8_t *out_flags, int64_t *n_result,
                      void *workspace, cudaStream_t stream) {
  const int64_t n = g.n_nodes;
  int32_t *seen = reinterpret_cast<int32_t *>(workspace);
  int32_t *vis = seen + n, *fa = vis + n, *fb = fa + n;
  unsigned long long *res = reinterpret_cast<unsigned long long *>((reinterpret_cast<uintptr_t>(fb + n) + 7) & ~(uintptr_t)7);
  int32_t *dev_ids = reinterpret_cast<int32_t *>(res + 2);
  AHR_CUDA_CHECK(cudaMemsetAsync(seen, 0, (size_t)n * 8, stream));  // seen + vis
  AHR_CUDA_CHECK(cudaMemsetAsync(res, 0, 8, stream));
  // pageable source: the runtime stages the few bytes before it returns, the caller's array is free afterwards
  AHR_CUDA_CHECK(cudaMemcpyAsync(dev_ids, host_ids, (size_t)n_ids * 4, cudaMemcpyHostToDevice, stream));
  lca_closure_kernel<<<1, LCA_THREADS, 0, stream>>>(g, dev_ids, n_ids, fa, fb, seen, vis);
  AHR_KERNEL_CHECK();
  lca_final_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(g, seen, n_ids, out_flags, res);
  AHR_KERNEL_CHECK();
  if (n_result != nullptr) {  // the only host round trip, and only for callers that want the count
    unsigned long long h = 0;
    AHR_CUDA_CHECK(cudaMemcpyAsync(&h, res, 8, cudaMemcpyDeviceToHost, stream));
    AHR_CUDA_CHECK(cudaStreamSynchronize(stream));
    *n_result = (int64_t)h;
  }
  return AHR_OK;
}

}  // namespace ahr
