// ubench.cu -- per-SM-sub-partition throughput / latency of the instructions the LK kernel leans on.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define N_ITER 4096
template <int OP, int ILP>
__global__ void k(float *out, float seed, long long *cycles) {
  float f[ILP];
  double d[ILP];
  int n[ILP];
  for (int i = 0; i < ILP; ++i) { f[i] = seed + i + threadIdx.x * 0.001f; d[i] = f[i]; n[i] = (int)f[i]; }
  const double c = seed;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < N_ITER; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (OP == 0) d[i] = fma(d[i], c, c);                        // DFMA
      if (OP == 1) { d[i] = (double)f[i]; f[i] = __int_as_float(__double2hiint(d[i]) + it); }  // F2F.F64.F32 (+ int op)
      if (OP == 2) f[i] = floorf(f[i] * 1.0001f);                  // FRND.FLOOR (+ FMUL)
      if (OP == 3) f[i] = (float)(unsigned char)(n[i] + it) + f[i] * 0.5f, n[i] += 3;  // I2F.U8-ish
      if (OP == 4) { n[i] = (int)f[i]; f[i] = __int_as_float(n[i] ^ 0x3f800000); }     // F2I (+ LOP)
      if (OP == 5) f[i] = f[i] * 1.0001f + 0.5f;                   // FMUL + FADD
      if (OP == 6) d[i] = d[i] + c;                                 // DADD
      if (OP == 7) d[i] = d[i] / c;                                 // DDIV
      if (OP == 8) d[i] = sqrt(d[i]) + c;                           // DSQRT
    }
  }
  const long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < ILP; ++i) acc += f[i] + (float)d[i] + n[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP, int ILP>
void run(const char *name, int warps) {
  float *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  k<OP, ILP><<<148, warps * 32>>>(out, 1.5f, cyc);
  k<OP, ILP><<<148, warps * 32>>>(out, 1.5f, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s ILP %2d warps/SM %2d : %7.2f cycles per op-group per warp, %6.2f cyc/instr/SMSP\n", name, ILP, warps,
         (double)h / N_ITER / ILP, (double)h / N_ITER / ILP / (warps > 4 ? warps / 4.0 : 1.0) * (warps >= 4 ? 1 : 1));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0, 1>("DFMA latency", 1);  run<0, 8>("DFMA", 4);  run<0, 8>("DFMA", 8);
  run<6, 1>("DADD latency", 1);  run<6, 8>("DADD", 4);
  run<1, 1>("F2F.F64.F32+IADD lat", 1); run<1, 8>("F2F.F64.F32+IADD", 4); run<1, 8>("F2F.F64.F32+IADD", 8);
  run<2, 1>("FRND.FLOOR+FMUL lat", 1); run<2, 8>("FRND.FLOOR+FMUL", 4); run<2, 8>("FRND.FLOOR+FMUL", 8);
  run<3, 8>("I2F.U8+...", 4); run<3, 8>("I2F.U8+...", 8);
  run<4, 1>("F2I+LOP lat", 1); run<4, 8>("F2I+LOP", 4); run<4, 8>("F2I+LOP", 8);
  run<5, 1>("FMUL+FADD lat", 1); run<5, 8>("FMUL+FADD", 4); run<5, 8>("FMUL+FADD", 8);
  run<7, 1>("DDIV latency", 1); run<7, 4>("DDIV", 4);
  run<8, 1>("DSQRT+DADD latency", 1); run<8, 4>("DSQRT+DADD", 4);
  return 0;
}
