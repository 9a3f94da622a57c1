// ubench_acc2.cu -- variants of the ordered accumulation loop (one warp, lane = slot)
#include <cstdio>
#include <cuda_runtime.h>
#define NP 121
__device__ __forceinline__ double f2d_nb(float f) {   // branch-free: normal or zero only
  const unsigned u = __float_as_uint(f);
  const unsigned a = u & 0x7fffffffu;
  unsigned hi = (((int)u >> 3) & 0x8fffffffu) + 0x38000000u;
  hi = (a == 0u) ? u : hi;
  return __hiloint2double((int)hi, (int)(u << 29));
}
template <int VAR, int UNR>
__global__ void acc_kernel(float *out, long long *cyc, int reps) {
  extern __shared__ __align__(16) unsigned char sm[];
  float4 *rec = (float4 *)sm;           // [32][121]
  double2 *drec = (double2 *)sm;        // [32][121] (same bytes; VAR 2)
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * NP; i += blockDim.x) {
    if (VAR == 2) drec[i] = make_double2(0.3 * (i % 17) - 2., 0.1 * (i % 13) - 1.);
    else rec[i] = make_float4(0.3f * (i % 17) - 2.f, 0.1f * (i % 13) - 1.f, 0.01f * (i % 7), 1.f);
  }
  __syncthreads();
  const double c = -100.5;
  double r0 = 0;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    double h00 = 0, h10 = 0, h11 = 0, h20 = 0, h21 = 0, h22 = 0, h30 = 0, h31 = 0;
#pragma unroll UNR
    for (int p = 0; p < NP; ++p) {
      double ix, iy;
      if (VAR == 0) { const float4 r = rec[lane * NP + p]; ix = (double)r.x; iy = (double)r.y; }
      if (VAR == 1) { const float4 r = rec[lane * NP + p]; ix = f2d_nb(r.x); iy = f2d_nb(r.y); }
      if (VAR == 2) { const double2 r = drec[lane * NP + p]; ix = r.x; iy = r.y; }
      if (VAR == 3) { const float4 r = rec[lane * NP + p]; ix = (double)r.x; iy = c; }      // one conversion
      if (VAR == 4) { const float4 r = rec[lane * NP + p]; ix = __hiloint2double(__float_as_int(r.x), 0); iy = __hiloint2double(__float_as_int(r.y), 0); }  // no conversion work, float loads
      h00 = fma(ix, ix, h00); h10 = fma(iy, ix, h10); h11 = fma(iy, iy, h11);
      h20 = fma(c, ix, h20); h21 = fma(c, iy, h21); h22 = fma(c, c, h22);
      h30 = h30 + ix; h31 = h31 + iy;
    }
    r0 += h00 + h10 + h11 + h20 + h21 + h22 + h30 + h31;
  }
  long long t1 = clock64();
  out[threadIdx.x] = (float)r0;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int VAR, int UNR> void run(const char *name) {
  float *out; long long *cyc, h; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(acc_kernel<VAR, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * NP * 16);
  const int reps = 50;
  acc_kernel<VAR, UNR><<<148, 32, 32 * NP * 16>>>(out, cyc, reps);
  acc_kernel<VAR, UNR><<<148, 32, 32 * NP * 16>>>(out, cyc, reps);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s unroll %2d: %8.0f cycles per pass (%5.1f per step)\n", name, UNR, (double)h / reps, (double)h / reps / NP);
}
int main() {
  run<0, 4>("F2F x2 + 8 D-ops"); run<0, 8>("F2F x2 + 8 D-ops"); run<0, 11>("F2F x2 + 8 D-ops");
  run<1, 4>("bit-convert x2 (branch-free) + 8 D-ops"); run<1, 8>("bit-convert x2 (branch-free) + 8 D-ops");
  run<2, 4>("doubles in smem + 8 D-ops"); run<2, 8>("doubles in smem + 8 D-ops");
  run<3, 4>("F2F x1 + 8 D-ops");
  run<4, 4>("no conversion, float loads + 8 D-ops");
  return 0;
}
