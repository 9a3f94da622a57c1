// developer probe: 3-D u8 TMA tile loads issued by every lane of a warp with its own descriptor / coordinates
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/probes/tma_probe tools/probes/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
struct Maps { CUtensorMap lv[8]; };
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32) probe(const __grid_constant__ Maps maps, const Maps *gmaps, unsigned char *out, int mode) {
  __shared__ __align__(128) unsigned char raw[32][256];
  __shared__ __align__(8) unsigned long long bar;
  const int lane = threadIdx.x;
  const unsigned mbar = smem_u32(&bar);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int n = mode == 0 ? 1 : 32;
  if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(208 * n) : "memory");
  __syncwarp();
  if (lane < n) {
    const int lv = mode == 2 ? (lane & 1) : 0;
    const int x = lane * 3 - 5, y = lane - 2, z = lane & 3;
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(&raw[lane][0])),
                 "l"(gmaps ? reinterpret_cast<unsigned long long>(&gmaps->lv[lv]) : reinterpret_cast<unsigned long long>(&maps.lv[lv])), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(mbar), "r"(0) : "memory");
  __syncwarp();
  for (int i = lane; i < 32 * 256; i += 32) out[i] = raw[i / 256][i % 256];
}
int main(int argc, char **argv) {
  const int W = 96, H = 40, Z = 4;
  std::vector<unsigned char> h((size_t)W * H * Z * 2);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(1 + (i * 7 + i / W) % 250);
  unsigned char *d, *o;
  cudaMalloc(&d, h.size()); cudaMalloc(&o, 32 * 256);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                         const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  Maps m; memset(&m, 0, sizeof(m));
  for (int l = 0; l < 2; ++l) {
    cuuint64_t dims[3] = {W, H, Z}, str[2] = {W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {16, 13, 1}, es[3] = {1, 1, 1};
    CUresult r = ((Fn)fp)(&m.lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d + (size_t)l * W * H * Z, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode level %d -> %d\n", l, (int)r);
  }
  Maps *gm; cudaMalloc(&gm, sizeof(Maps)); cudaMemcpy(gm, &m, sizeof(Maps), cudaMemcpyHostToDevice);
  for (int mode = (argc > 1 ? 3 : 0); mode < 6; ++mode) {
    cudaMemset(o, 0xEE, 32 * 256);
    const bool useg = mode >= 3;
    printf("descriptor in %s memory\n", useg ? "global" : "param");
    probe<<<1, 32>>>(m, useg ? gm : nullptr, o, mode % 3);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) { if (mode < 3) { cudaDeviceReset(); printf("(context reset not attempted; rerun with global)\n"); } return 1; }
    std::vector<unsigned char> r(32 * 256);
    cudaMemcpy(r.data(), o, r.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    const int n = mode % 3 == 0 ? 1 : 32;
    for (int lane = 0; lane < n; ++lane) {
      const int lv = mode % 3 == 2 ? (lane & 1) : 0, x0 = lane * 3 - 5, y0 = lane - 2, z = lane & 3;
      for (int rr = 0; rr < 13; ++rr) for (int c = 0; c < 16; ++c) {
        const int x = x0 + c, y = y0 + rr;
        unsigned char want = 0;
        if (x >= 0 && x < W && y >= 0 && y < H) want = h[(size_t)lv * W * H * Z + ((size_t)z * H + y) * W + x];
        if (r[lane * 256 + rr * 16 + c] != want) ++bad;
      }
    }
    printf("mode %d: %d mismatching bytes\n", mode, bad);
  }
  return 0;
}
