// Does a clock read placed right after __syncthreads() wait for the barrier to complete?
// Warp 1 spins ~100k cycles before the barrier; warp 0 stamps the clock before the barrier, right after it, and after a
// dependent shared-memory load that follows.  Build: nvcc -arch=sm_100a -o bar_clock_probe bar_clock_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(long long* out, int spin) {
  __shared__ int s[64];
  s[threadIdx.x] = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x >= 32) {
    while (clock64() - t0 < spin) {}
  }
  long long a = clock64();
  __syncthreads();
  long long b = clock64();
  int v = s[(threadIdx.x + 1) & 63];
  long long c = 0;
  if (v >= 0) c = clock64();  // the branch needs the loaded value: the stamp cannot be taken before the load returns
  if (threadIdx.x == 0) {
    out[0] = a - t0;
    out[1] = b - a;
    out[2] = c - b + (v == 12345);
  }
}
int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (int r = 0; r < 3; ++r) {
    probe<<<1, 64>>>(d, 100000);
    long long h[3];
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("warp 0: before barrier %lld, across barrier %lld, after dependent LDS %lld cycles\n", h[0], h[1], h[2]);
  }
  return 0;
}
