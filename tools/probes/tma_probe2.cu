// developer probe 2: which tensor-map shapes does a single-thread TMA tile load accept? one variant per process
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void __launch_bounds__(32) probe(const __grid_constant__ CUtensorMap map, unsigned char *out, int bytes, int x, int y, int z) {
  __shared__ __align__(1024) unsigned char raw[8192];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x;
  const unsigned mbar = smem_u32(&bar);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(raw)),
                   "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(raw)),
                   "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(mbar) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mbar), "r"(0) : "memory");
  __syncthreads();
  for (int i = lane; i < bytes; i += 32) out[i] = raw[i];
}
int main(int argc, char **argv) {
  const int v = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 96, H = 40, Z = 4;
  std::vector<unsigned char> h((size_t)W * H * Z);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(1 + (i * 7 + i / W) % 250);
  unsigned char *d, *o;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, 8192);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                         const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap m; memset(&m, 0, sizeof(m));
  int rank = 3, bw = 16, bh = 13, x = 3, y = 2, z = 1; CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; int es = 1;
  if (v == 1) { rank = 2; }
  if (v == 2) { rank = 2; bw = 32; }
  if (v == 3) { rank = 2; dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; es = 4; bw = 4; x = 1; }
  if (v == 4) { bw = 32; }
  if (v == 5) { rank = 2; bw = 64; bh = 8; }
  if (v == 6) { rank = 2; x = 0; y = 0; }
  if (v == 7) { x = -5; y = -2; }
  cuuint64_t dims[3] = {(cuuint64_t)(W / es), (cuuint64_t)(rank == 2 ? H * Z : H), Z}, str[2] = {W, (cuuint64_t)W * H};
  cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, est[3] = {1, 1, 1};
  CUresult r = ((Fn)fp)(&m, dt, rank, d, dims, str, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int bytes = bw * es * bh;
  printf("variant %d: rank %d box %dx%d es %d encode -> %d; ", v, rank, bw, bh, es, (int)r);
  cudaMemset(o, 0xEE, 8192);
  if (rank == 3) probe<3><<<1, 32>>>(m, o, bytes, x, y, z); else probe<2><<<1, 32>>>(m, o, bytes, x, y, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<unsigned char> rr(8192);
  cudaMemcpy(rr.data(), o, 8192, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < bh; ++i) for (int c = 0; c < bw * es; ++c) {
    const int xx = x * es + c, yy = y + i;
    unsigned char want = 0;
    const int zz = rank == 3 ? z : 0;
    if (xx >= 0 && xx < W && yy >= 0 && yy < (rank == 3 ? H : H * Z)) want = h[((size_t)zz * H + yy) * W + xx];
    if (rr[i * bw * es + c] != want) ++bad;
  }
  printf("   %d mismatching bytes of %d\n", bad, bytes);
  return 0;
}
