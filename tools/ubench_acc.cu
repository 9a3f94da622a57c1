// ubench_acc.cu -- intrinsic speed of the ordered accumulation loop (one warp, lane = slot) and of the 4x4 solve
#include <cstdio>
#include <cuda_runtime.h>
#include "../pixel_aware_gyro_aided_klt_feature_tracker_b200/csrc/pagk_device.cuh"
#define NP 121
// exact float -> double on the integer pipe (no F2F): rebias the exponent, split the mantissa
__device__ __forceinline__ double f2d_bits(float f) {
  const unsigned u = __float_as_uint(f);
  const unsigned a = u & 0x7fffffffu;
  if ((a - 0x00800000u) >= 0x7f000000u && a != 0u) return (double)f;  // denormal / inf / nan: real conversion
  unsigned hi = (((int)u >> 3) & 0x8fffffffu) + 0x38000000u;
  hi = (a == 0u) ? u : hi;
  return __hiloint2double((int)hi, (int)(u << 29));
}
#ifdef USE_BITS
#define CVT(x) f2d_bits(x)
#else
#define CVT(x) ((double)(x))
#endif
template <int VAR>
__global__ void acc_kernel(float *out, long long *cyc, int reps, int warps_active) {
  extern __shared__ float4 rec[];  // [32][121]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * NP; i += blockDim.x) rec[i] = make_float4(0.3f * (i % 17) - 2.f, 0.1f * (i % 13) - 1.f, 0.01f * (i % 7), 1.f);
  __syncthreads();
  const double c = -100.5;
  double r0 = 0, r1 = 0;
  long long t0 = clock64();
  if (warp < warps_active)
  for (int rep = 0; rep < reps; ++rep) {
    const float4 *p4 = rec + lane * NP;
    if (VAR == 0) {  // 8 H chains, 2 conversions
      double h00 = 0, h10 = 0, h11 = 0, h20 = 0, h21 = 0, h22 = 0, h30 = 0, h31 = 0;
#pragma unroll 4
      for (int p = 0; p < NP; ++p) {
        const float4 r = p4[p];
        const double ix = CVT(r.x), iy = CVT(r.y);
        h00 = fma(ix, ix, h00); h10 = fma(iy, ix, h10); h11 = fma(iy, iy, h11);
        h20 = fma(c, ix, h20); h21 = fma(c, iy, h21); h22 = fma(c, c, h22);
        h30 = h30 + ix; h31 = h31 + iy;
      }
      r0 += h00 + h10 + h11 + h20; r1 += h21 + h22 + h30 + h31;
    } else if (VAR == 1) {  // 4 b chains + cost, 3 conversions
      double b0 = 0, b1 = 0, b2 = 0, b3 = 0; float cost = 0;
#pragma unroll 4
      for (int p = 0; p < NP; ++p) {
        const float4 r = p4[p];
        const double ix = CVT(r.x), iy = CVT(r.y), me = CVT(r.z);
        b0 = fma(ix, me, b0); b1 = fma(iy, me, b1); b2 = fma(c, me, b2); b3 = b3 + me;
        cost = cost + r.z * r.z;
      }
      r0 += b0 + b1; r1 += b2 + b3 + cost;
    } else if (VAR == 2) {  // the 4x4 solve
      double u0, u1, u2, u3;
      const double a = 1000.0 + lane + rep;
      pagk_llt_solve4(a, 10.0, a + 5, -3.0, 2.0, a * 2, 7.0, 1.0, 121 * 0.5, 121.0, 1.0 + rep, 2.0, 3.0, 4.0, u0, u1, u2, u3);
      r0 += u0 + u1; r1 += u2 + u3;
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = (float)(r0 + r1);
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int VAR> void run(const char *name, int threads, int warps_active) {
  float *out; long long *cyc, h; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(acc_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * NP * 16);
  const int reps = 50;
  acc_kernel<VAR><<<148, threads, 32 * NP * 16>>>(out, cyc, reps, warps_active);
  acc_kernel<VAR><<<148, threads, 32 * NP * 16>>>(out, cyc, reps, warps_active);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-34s warps running %d: %8.0f cycles per pass (%5.1f per step)\n", name, warps_active, (double)h / reps, (double)h / reps / NP);
}
int main() {
  run<0>("H chains (8 D-ops, 2 cvt)", 32, 1);
  run<1>("b chains (4 D-ops, 3 cvt, cost)", 32, 1);
  run<2>("llt solve", 32, 1);
  run<0>("H chains, 4 warps same code", 128, 4);
  run<2>("llt solve, 4 warps", 128, 4);
  return 0;
}
