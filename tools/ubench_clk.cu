// ubench_clk.cu -- effective SM clock: clock64 vs globaltimer over a busy FP32 loop on all SMs
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float *out, long long *res, int iters) {
  float a = threadIdx.x * 0.001f, b = 1.0001f;
  unsigned long long g0, g1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
  long long c0 = clock64();
  for (int i = 0; i < iters; ++i) { a = a * b + 0.5f; b = b * 0.99999f + 0.00001f; }
  long long c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
  if (threadIdx.x == 0 && blockIdx.x == 0) { res[0] = c1 - c0; res[1] = (long long)(g1 - g0); }
}
int main() {
  float *out; long long *res, h[2];
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&res, 16);
  for (int r = 0; r < 4; ++r) {
    k<<<148, 512>>>(out, res, 2000000);
    cudaMemcpy(h, res, 16, cudaMemcpyDeviceToHost);
    printf("cycles %lld  ns %lld  -> %.1f MHz\n", h[0], h[1], 1e3 * h[0] / h[1]);
  }
  return 0;
}
