// microbenchmark: cost of MATCH.ANY vs a vote, per warp instruction, at 4..16 warps per SMSP
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(const int *in, int *out, int iters) {
  int v = in[threadIdx.x + blockIdx.x * blockDim.x];
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) acc += __match_any_sync(0xffffffffu, v + (int)(acc & 1));
    if (MODE == 1) acc += __ballot_sync(0xffffffffu, (v + acc) & 1);
    if (MODE == 2) acc += __match_any_sync(0xffffffffu, (v & 7) + (int)(acc & 1));  // 8 distinct values
    if (MODE == 3) acc += __shfl_xor_sync(0xffffffffu, v + acc, 5);
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("mode %d warps/block %d: %.1f cycles per dependent op\n", MODE, blockDim.x / 32, double(t1 - t0) / iters);
}
int main() {
  int n = 148 * 1024;
  int *in, *out;
  cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4);
  int *h = new int[n];
  for (int i = 0; i < n; ++i) h[i] = (i * 2654435761u) >> 8;
  cudaMemcpy(in, h, n * 4, cudaMemcpyHostToDevice);
  for (int threads : {32, 128, 512, 1024}) {
    k<0><<<148, threads>>>(in, out, 4096); cudaDeviceSynchronize();
    k<2><<<148, threads>>>(in, out, 4096); cudaDeviceSynchronize();
    k<1><<<148, threads>>>(in, out, 4096); cudaDeviceSynchronize();
    k<3><<<148, threads>>>(in, out, 4096); cudaDeviceSynchronize();
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
