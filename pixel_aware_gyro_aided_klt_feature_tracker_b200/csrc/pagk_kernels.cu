// pagk_kernels.cu -- the sm_100a kernels of the pixel-aware gyro-aided KLT hot path.
//
//   K1  pagk_pyramid_fused_kernel / pagk_pyramid_general_kernel
//         PatchMatch::CreatePyramids               reference src/patch_match.cpp:61-76
//   K2  pagk_gyro_predict_kernel
//         GyroAidedTracker::GyroPredictFeatures    reference src/gyro_aided_tracker.cpp:118-256
//   K3  pagk_lk_kernel (generic, any patch size)
//         PatchMatch::OpticalFlowMultiLevel + OpticalFlowConsideringIlluminationChange_onePixel
//                                                  reference src/patch_match.cpp:79-142, 167-367
//   K4  pagk_epilogue1_kernel / pagk_epilogue_points, _mean, _filter_kernel
//         DistortPoints / SetMatcher               reference src/patch_match.cpp:370-388, 409-416
//         threshold filter                         reference src/gyro_aided_tracker.cpp:289-336
//
// Compiled with --fmad=false (see pagk_device.cuh for the arithmetic contract).
#include "pagk_device.cuh"
#include "pagk_kernels.h"
#include "pagk_ransac.h"

#include <math_constants.h>

// =================================================================================================
// K1: image pyramid.  cv::resize(INTER_LINEAR) to exactly half size is OpenCV's 2x2 INTER_AREA fast
// path: dst = (a + b + c + d + 2) >> 2 (verified against cv2).  One CTA owns a
// 128 x 128 tile of level 0 and produces that tile of every further level of the exact-2x chain from
// shared memory, so level 0 is read from HBM once and nothing is re-read.  A thread has its four 16-byte
// loads (two row pairs) in flight together -- 64 KB per SM with four CTAs resident, what the HBM latency
// needs --, stores are 8 bytes.  The CTA that owns the last row of a level also writes that level's guard row.
// =================================================================================================
#define PYR_TW 128
#define PYR_TH 128

// Pixel (0, gy) = v has just been produced: write the wrap bytes that mirror it (pagk_device.cuh, PagkLevelGeom):
// column `cols` of the row above when the level is padded, and for the last row its own wrap byte and the guard row's.
__device__ __forceinline__ void pagk_write_wrap(unsigned char *ll, int pitch, int cols, int rows, int gy, unsigned char v) {
  if (gy >= 1 && pitch > cols) ll[(size_t)(gy - 1) * pitch + cols] = v;
  if (gy == rows - 1) {
    if (pitch > cols) ll[(size_t)gy * pitch + cols] = v;
    ll[(size_t)rows * pitch + cols] = v;
  }
}

// r0, r1: four horizontally adjacent pixels of two rows -> (a + b + c + d + 2) >> 2 of the two 2 x 2 blocks, in bytes
// 0 and 2 of the result.  The even and the odd bytes of a word go to 16-bit lanes with one PRMT each (byte 4 of the
// selector is the zero register), the four partial sums are two three-input adds.
__device__ __forceinline__ unsigned int pagk_avg2x2_lanes(unsigned int r0, unsigned int r1) {
  const unsigned int e0 = __byte_perm(r0, 0u, 0x4240), o0 = __byte_perm(r0, 0u, 0x4341);
  const unsigned int e1 = __byte_perm(r1, 0u, 0x4240), o1 = __byte_perm(r1, 0u, 0x4341);
  return (((e0 + o0 + e1) + (o1 + 0x00020002u)) >> 2) & 0x00ff00ffu;
}
// eight pixels of two rows (two words each) -> four output pixels in one word
__device__ __forceinline__ unsigned int pagk_avg8x8(unsigned int a0, unsigned int a1, unsigned int b0, unsigned int b1) {
  return __byte_perm(pagk_avg2x2_lanes(a0, b0), pagk_avg2x2_lanes(a1, b1), 0x6420);
}
__device__ __forceinline__ unsigned int pagk_avg4x8(unsigned int r0, unsigned int r1) {
  // r0, r1: four horizontally adjacent pixels of two rows -> two output pixels in the low 16 bits
  return __byte_perm(pagk_avg2x2_lanes(r0, r1), 0u, 0x4420);
}

// One more level of a tile: the TW x TH tile of level l-1 in shared memory `s` -> its (TW/2) x (TH/2) tile of level l,
// to shared memory `d` and to the image; then recurses to level l+1.  A thread produces four adjacent pixels of a row
// (two 8-byte shared-memory loads, one 4-byte store to each side); tile origins and pitches are multiples of four at
// every fused level, so only a level's ragged right edge falls back to bytes.
template <int TW, int TH>
__device__ __forceinline__ void pagk_pyramid_tile_level(const unsigned char *s, unsigned char *d, unsigned char *img,
                                                        const PagkGeom &g, int l, int n_fused, int tx0, int ty0, int t) {
  constexpr int OW = TW / 2, OH = TH / 2;
  if constexpr (OW >= 4 && OH >= 1) {
    __syncthreads();
    const int colsl = g.lv[l].cols, rowsl = g.lv[l].rows, pl = g.lv[l].pitch;
    unsigned char *ll = img + g.lv[l].offset;
    const int ox0 = tx0 >> l, oy0 = ty0 >> l;
    constexpr int QW = OW / 4;  // groups of four pixels per row
#pragma unroll
    for (int p = t; p < QW * OH; p += 256) {
      const int oq = p % QW, oy = p / QW, ox = oq * 4;  // QW is a power of two: a mask and a shift
      const uint2 r0 = *reinterpret_cast<const uint2 *>(s + (2 * oy) * TW + 2 * ox);
      const uint2 r1 = *reinterpret_cast<const uint2 *>(s + (2 * oy + 1) * TW + 2 * ox);
      const unsigned int v4 = pagk_avg8x8(r0.x, r0.y, r1.x, r1.y);
      *reinterpret_cast<unsigned int *>(d + oy * OW + ox) = v4;
      const int gx = ox0 + ox, gy = oy0 + oy;
      if (gx < colsl && gy < rowsl) {
        unsigned char *dst = ll + (unsigned int)(gy * pl + gx);
        if (gx + 4 <= colsl) {
          *reinterpret_cast<unsigned int *>(dst) = v4;
          if (gy == rowsl - 1) *reinterpret_cast<unsigned int *>(dst + pl) = v4;  // guard row
        } else {
          for (int k = 0; k < 4 && gx + k < colsl; ++k) {
            dst[k] = (unsigned char)(v4 >> (8 * k));
            if (gy == rowsl - 1) dst[pl + k] = (unsigned char)(v4 >> (8 * k));
          }
        }
        if (gx == 0) pagk_write_wrap(ll, pl, colsl, rowsl, gy, (unsigned char)(v4 & 0xffu));
      }
    }
    if (l < n_fused) pagk_pyramid_tile_level<OW, OH>(d, const_cast<unsigned char *>(s), img, g, l + 1, n_fused, tx0, ty0, t);
  }
}

#ifndef PAGK_PYR_MIN_BLOCKS
#define PAGK_PYR_MIN_BLOCKS 5
#endif
// VEC: the width of level 0 is a multiple of 16 (then its pitch is its width and the fused levels are continuous too)
template <bool VEC>
__global__ void __launch_bounds__(256, PAGK_PYR_MIN_BLOCKS) pagk_pyramid_fused_kernel(unsigned char *__restrict__ images, PagkGeom g,
                                                               int n_fused /* last level produced here */,
                                                               int z_stride /* image z lives in slot z * z_stride */) {
  __shared__ __align__(16) unsigned char sbuf[2][(PYR_TW / 2) * (PYR_TH / 2)];
  unsigned char *img = images + (size_t)blockIdx.z * z_stride * g.slot_bytes;
  const int t = threadIdx.x;
  const int cols0 = g.lv[0].cols, rows0 = g.lv[0].rows, p0 = g.lv[0].pitch;
  unsigned char *l0 = img + g.lv[0].offset;
  const int tx0 = blockIdx.x * PYR_TW, ty0 = blockIdx.y * PYR_TH;

  // ---- level 0 -> level 1 (or only the guard row of level 0 when there is a single level)
  const int tx = t & 7, ty = t >> 3;  // 8 threads x 16 px per row pair, 32 row pairs, two of them per thread
  const int x0 = tx0 + tx * 16;
  constexpr int NR = PYR_TH / 64;
  unsigned int a[NR][4], b[NR][4];
  bool in0[NR], in1[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int y0 = ty0 + r * 64 + ty * 2;
    in0[r] = (x0 < cols0) && (y0 < rows0); in1[r] = (x0 < cols0) && (y0 + 1 < rows0);
#pragma unroll
    for (int k = 0; k < 4; ++k) { a[r][k] = 0; b[r][k] = 0; }
    const unsigned char *src = l0 + (unsigned int)(y0 * p0 + x0);
    if (VEC) {
      if (in0[r]) { const uint4 v = *reinterpret_cast<const uint4 *>(src); a[r][0] = v.x; a[r][1] = v.y; a[r][2] = v.z; a[r][3] = v.w; }
      if (in1[r]) { const uint4 v = *reinterpret_cast<const uint4 *>(src + p0); b[r][0] = v.x; b[r][1] = v.y; b[r][2] = v.z; b[r][3] = v.w; }
    } else {
      for (int k = 0; k < 16; ++k) {
        if (in0[r] && x0 + k < cols0) a[r][k >> 2] |= (unsigned int)src[k] << (8 * (k & 3));
        if (in1[r] && x0 + k < cols0) b[r][k >> 2] |= (unsigned int)src[p0 + k] << (8 * (k & 3));
      }
    }
  }
  const int cols1 = g.lv[1].cols, rows1 = g.lv[1].rows, p1 = g.lv[1].pitch;
  unsigned char *l1 = img + g.lv[1].offset;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int y0 = ty0 + r * 64 + ty * 2;
    // explicit wrap bytes of a padded level 0 (the caller's image has a width that is not a multiple of 4)
    if (!VEC && p0 > cols0 && x0 == 0) {
      if (in0[r] && y0 >= 1) l0[(size_t)(y0 - 1) * p0 + cols0] = (unsigned char)(a[r][0] & 0xffu);
      if (in1[r]) l0[(size_t)y0 * p0 + cols0] = (unsigned char)(b[r][0] & 0xffu);
      if (in0[r] && y0 == rows0 - 1) l0[(size_t)y0 * p0 + cols0] = (unsigned char)(a[r][0] & 0xffu);
      if (in1[r] && y0 + 1 == rows0 - 1) l0[(size_t)(y0 + 1) * p0 + cols0] = (unsigned char)(b[r][0] & 0xffu);
    }
    // guard row of level 0 = copy of row rows0-1 (+ one byte)
    if (in0[r] && (y0 == rows0 - 1 || y0 + 1 == rows0 - 1)) {
      const unsigned int *src = (y0 == rows0 - 1) ? a[r] : b[r];
      unsigned char *gr = l0 + (size_t)rows0 * p0 + x0;
      if (VEC) {
        *reinterpret_cast<uint4 *>(gr) = make_uint4(src[0], src[1], src[2], src[3]);
      } else {
        for (int k = 0; k < 16 && x0 + k < cols0; ++k) gr[k] = (unsigned char)(src[k >> 2] >> (8 * (k & 3)));
      }
      if (x0 == 0) l0[(size_t)rows0 * p0 + cols0] = (unsigned char)(src[0] & 0xffu);
    }
    if (n_fused >= 1) {
      uint2 o;
      o.x = pagk_avg8x8(a[r][0], a[r][1], b[r][0], b[r][1]);
      o.y = pagk_avg8x8(a[r][2], a[r][3], b[r][2], b[r][3]);
      *reinterpret_cast<uint2 *>(&sbuf[0][(r * 32 + ty) * (PYR_TW / 2) + tx * 8]) = o;
      const int x1 = x0 >> 1, y1 = y0 >> 1;
      if (y1 < rows1 && x1 < cols1) {
        unsigned char *dst = l1 + (unsigned int)(y1 * p1 + x1);
        if (VEC && x1 + 8 <= cols1) {
          *reinterpret_cast<uint2 *>(dst) = o;
          if (y1 == rows1 - 1) *reinterpret_cast<uint2 *>(dst + p1) = o;
        } else {
          for (int k = 0; k < 8 && x1 + k < cols1; ++k) {
            const unsigned char v = (unsigned char)((k < 4 ? o.x : o.y) >> (8 * (k & 3)));
            dst[k] = v;
            if (y1 == rows1 - 1) dst[p1 + k] = v;
          }
        }
        if (x1 == 0) pagk_write_wrap(l1, p1, cols1, rows1, y1, (unsigned char)(o.x & 0xffu));
      }
    }
  }
  // ---- level l-1 (shared memory) -> level l, l = 2 .. n_fused: tile sizes are compile-time constants
  if (n_fused >= 2) pagk_pyramid_tile_level<PYR_TW / 2, PYR_TH / 2>(sbuf[0], sbuf[1], img, g, 2, n_fused, tx0, ty0, t);
}

// General half-size cv::resize(INTER_LINEAR) for levels whose source has an odd dimension: the
// 11-bit fixed-point bilinear of OpenCV.  One thread per output pixel.
__global__ void __launch_bounds__(256) pagk_pyramid_general_kernel(unsigned char *__restrict__ images, PagkGeom g,
                                                                 int level, int z_stride) {
  unsigned char *img = images + (size_t)blockIdx.z * z_stride * g.slot_bytes;
  const int scols = g.lv[level - 1].cols, srows = g.lv[level - 1].rows, sp = g.lv[level - 1].pitch;
  const int dcols = g.lv[level].cols, drows = g.lv[level].rows, dp = g.lv[level].pitch;
  const unsigned char *src = img + g.lv[level - 1].offset;
  unsigned char *dst = img + g.lv[level].offset;
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  const int dy = blockIdx.y;
  if (dx >= dcols || dy >= drows) return;
  unsigned char v;
  if (scols == 2 * dcols && srows == 2 * drows) {
    const unsigned char *r0 = src + (size_t)(2 * dy) * sp + 2 * dx;
    v = (unsigned char)((r0[0] + r0[1] + r0[sp] + r0[sp + 1] + 2) >> 2);
  } else {
    const double sx = (double)scols / dcols, sy = (double)srows / drows;
    float fx = (float)((dx + 0.5) * sx - 0.5);
    int ix = (int)floorf(fx);
    fx -= (float)ix;
    if (ix < 0) { ix = 0; fx = 0.f; }
    if (ix >= scols - 1) { ix = scols - 1; fx = 0.f; }
    const int a0 = __float2int_rn((1.f - fx) * 2048.f), a1 = __float2int_rn(fx * 2048.f);
    float fy = (float)((dy + 0.5) * sy - 0.5);
    int iy = (int)floorf(fy);
    fy -= (float)iy;
    const int y0 = min(max(iy, 0), srows - 1), y1 = min(max(iy + 1, 0), srows - 1);
    const int b0 = __float2int_rn((1.f - fy) * 2048.f), b1 = __float2int_rn(fy * 2048.f);
    const int ix1 = min(ix + 1, scols - 1);
    const unsigned char *r0 = src + (size_t)y0 * sp, *r1 = src + (size_t)y1 * sp;
    const int t0 = r0[ix] * a0 + r0[ix1] * a1;
    const int t1 = r1[ix] * a0 + r1[ix1] * a1;
    const int o = (((b0 * (t0 >> 4)) >> 16) + ((b1 * (t1 >> 4)) >> 16) + 2) >> 2;
    v = (unsigned char)min(max(o, 0), 255);
  }
  dst[(size_t)dy * dp + dx] = v;
  if (dy == drows - 1) dst[(size_t)drows * dp + dx] = v;
  if (dx == 0) pagk_write_wrap(dst, dp, dcols, drows, dy, v);
}

// =================================================================================================
// K2: pixel-aware gyro prediction, one thread per feature.
// =================================================================================================
__device__ __forceinline__ void pagk_predict_one(const PagkPairConst &c, int method, const float *__restrict__ ntab,
                                                 int width, float2 ref, float2 &pun, float2 &pd) {
  float xn, yn;
  if (ntab) {
    const size_t o = ((size_t)((int)ref.y) * width + (int)ref.x) * 2;
    xn = ntab[o]; yn = ntab[o + 1];
  } else {
    xn = (ref.x - c.cx) * c.fx_inv;
    yn = (ref.y - c.cy) * c.fy_inv;
  }
  float lambda = 1.0f;
  if (method == 1) {
    const float den = (c.r31 * xn + c.r32 * yn) + c.r33;
    lambda = (float)(1.0 / (double)den);
  }
  pun.x = ((c.M[0] * ref.x + c.M[1] * ref.y) + c.M[2]) * lambda;
  pun.y = ((c.M[3] * ref.x + c.M[4] * ref.y) + c.M[5]) * lambda;
  pd = pagk_distort(c, pun);
}

__global__ void __launch_bounds__(128) pagk_gyro_predict_kernel(const PagkPairConst *__restrict__ pcs,
                                                              const float2 *__restrict__ keys_un,
                                                              const float2 *__restrict__ keys, PagkOutPtrs out,
                                                              PagkMode mode, int max_keys, int width, int height,
                                                              const float *__restrict__ ntab_all,
                                                              unsigned long long ntab_stride) {
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const PagkPairConst &c = pcs[pair];
  if (i >= c.n_keys) return;
  const size_t o = (size_t)pair * max_keys + i;
  const float2 ref = keys_un[o];
  float2 z = make_float2(0.f, 0.f);
  float2 pun = z, pd = z, fl = z, cf[4] = {z, z, z, z}, cu[4] = {z, z, z, z}, cd[4] = {z, z, z, z};
  float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned char st = 0;
  if (!mode.gyro_init) {  // eType 5, reference src/gyro_aided_tracker.cpp:264-270
    pun = ref; pd = keys[o]; st = 1;
    A = make_float4(1.f, 0.f, 0.f, 1.f);
    out.pt_gyro_un[o] = z; out.pt_gyro[o] = z;
  } else {
    const float *ntab = c.has_table ? ntab_all + (size_t)pair * ntab_stride : nullptr;
    float2 p, d;
    pagk_predict_one(c, mode.predict_method, ntab, width, ref, p, d);
    const float W = (float)width, H = (float)height;
    const bool ok = !(p.x < 0.f || p.x >= W || p.y < 0.f || p.y >= H) && !(d.x < 0.f || d.x >= W || d.y < 0.f || d.y >= H);
    if (ok) {
      pun = p; pd = d; st = 1;
      fl = make_float2(p.x - ref.x, p.y - ref.y);
      const float hf = (float)mode.half;
      const float bx[4] = {-hf, hf, -hf, hf}, by[4] = {-hf, -hf, hf, hf};
      double s00 = 0, s01 = 0, s10 = 0, s11 = 0;
      double q00[4], q01[4], q10[4], q11[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pagk_predict_one(c, mode.predict_method, ntab, width, make_float2(ref.x + bx[j], ref.y + by[j]), cu[j], cd[j]);
        cf[j] = make_float2(cu[j].x - p.x, cu[j].y - p.y);
        q00[j] = (double)cf[j].x * (double)bx[j]; q01[j] = (double)cf[j].x * (double)by[j];
        q10[j] = (double)cf[j].y * (double)bx[j]; q11[j] = (double)cf[j].y * (double)by[j];
      }
      // C * B^T with OpenCV's double accumulation: s0 = q0, then s0 += (q1 + q2) + q3
      s00 = q00[0] + ((q00[1] + q00[2]) + q00[3]); s01 = q01[0] + ((q01[1] + q01[2]) + q01[3]);
      s10 = q10[0] + ((q10[1] + q10[2]) + q10[3]); s11 = q11[0] + ((q11[1] + q11[2]) + q11[3]);
      const float S00 = (float)s00, S01 = (float)s01, S10 = (float)s10, S11 = (float)s11;
      // times (B B^T)^-1 = [[bb, -0],[-0, bb]] as a float 2x2 small gemm
      const float bb = mode.bb_inv, nz = -0.0f;
      A.x = S00 * bb + S01 * nz; A.y = S00 * nz + S01 * bb;
      A.z = S10 * bb + S11 * nz; A.w = S10 * nz + S11 * bb;
    }
    out.pt_gyro_un[o] = pun; out.pt_gyro[o] = pd;
  }
  out.pt_predict_un[o] = pun; out.pt_predict[o] = pd;
  out.status[o] = st; out.gyro_status[o] = st;
  out.flows[o] = fl; out.affine[o] = A;
#pragma unroll
  for (int j = 0; j < 4; ++j) { out.cflows[o * 4 + j] = cf[j]; out.corners_un[o * 4 + j] = cu[j]; out.corners[o * 4 + j] = cd[j]; }
}

// =================================================================================================
// K3 (generic): coarse-to-fine forward-additive Lucas-Kanade with gain/bias and affine pre-warp.
// One warp per feature, all pyramid levels inside the kernel (features are independent across
// levels; the reference's per-level parallel_for is only a CPU scheduling choice).
//   phase A  lanes = patch pixels: bilinear samples, residual e and gradient (Ix, Iy) -> shared memory
//   phase B  lanes = accumulators: the 10 unique entries of H and 4 of b are summed in DOUBLE in the
//            reference's pixel order (a structurally singular H makes the result order-sensitive,
//            SURVEY.md F3); the float cost chain runs beside them
//   solve    every lane runs the 4x4 LLT redundantly (no broadcast needed)
// Shared memory per warp: NP * 56 + 128 bytes.
// =================================================================================================
__device__ __forceinline__ const unsigned char *pagk_level_ptr(const unsigned char *images, const PagkGeom &g,
                                                               int pair, int which, int level) {
  return images + (size_t)(pair * 2 + which) * g.slot_bytes + g.lv[level].offset;
}

__global__ void pagk_lk_kernel(const unsigned char *__restrict__ images, PagkGeom g,
                               const PagkPairConst *__restrict__ pcs, const float2 *__restrict__ keys_un,
                               PagkOutPtrs out, PagkMode mode, int max_keys, int warps_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int pair = blockIdx.y;
  const int i = blockIdx.x * warps_per_cta + wib;
  if (i >= pcs[pair].n_keys) return;
  const size_t o = (size_t)pair * max_keys + i;
  const int h = mode.half, P = 2 * h + 1, NP = P * P;
  const size_t per_warp = (size_t)NP * 56 + 128;
  unsigned char *base = smem_raw + per_warp * wib;
  double *sJ = reinterpret_cast<double *>(base);                  // [NP][5]: Ix, Iy, c, 1, -e
  double *sRed = reinterpret_cast<double *>(base + (size_t)NP * 40);  // [16]
  float *sE2 = reinterpret_cast<float *>(base + (size_t)NP * 40 + 128);
  float *sWx = sE2 + NP, *sWy = sWx + NP, *sT = sWy + NP;

  const float2 pt1 = keys_un[o];
  float2 pt2 = mode.gyro_init ? out.pt_predict_un[o] : pt1;
  const unsigned char gst = out.gyro_status[o];
  int n_iter = 0;
  bool succ = false;
  float lastCost = 0.f;
  if (gst) {
    const float4 A = out.affine[o];
    // accumulator role of this lane: acc += J[ia] * J[ib]; column 4 holds -e, so that the b lanes
    // compute b_r += J_r * (-e), the same value as the reference's b += -J * e (negation is exact)
    const int ia_tab[14] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3, 0, 1, 2, 3};
    const int ib_tab[14] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3, 4, 4, 4, 4};
    const int role = lane < 14 ? lane : 0;
    const int ia = ia_tab[role], ib = ib_tab[role];
    for (int p = lane; p < NP; p += 32) {
      const int y = p / P - h, x = p % P - h;
      float wx = (float)x, wy = (float)y;
      if (mode.affine) {
        wx = A.x * (float)x + A.y * (float)y;
        wy = A.z * (float)x + A.w * (float)y;
      }
      sWx[p] = wx; sWy[p] = wy;
    }
    for (int level = mode.levels - 1; level >= 0; --level) {
      const unsigned char *I1 = pagk_level_ptr(images, g, pair, 0, level);
      const unsigned char *I2 = pagk_level_ptr(images, g, pair, 1, level);
      const int cols = g.lv[level].cols, rows = g.lv[level].rows, pitch = g.lv[level].pitch;
      const float scale = 1.0f / (float)(1 << level);
      const float ptx = pt1.x * scale, pty = pt1.y * scale;
      float nx, ny;
      if (level == mode.levels - 1) { nx = pt2.x * scale; ny = pt2.y * scale; }
      else { nx = pt2.x * 2.0f; ny = pt2.y * 2.0f; }
      float dx = nx - ptx, dy = ny - pty;
      float dg = 0.f, db = 0.f, cost = 0.f;
      lastCost = 0.f;
      succ = true;
      const float cval = -pagk_sample(I1, pitch, cols, rows, ptx, pty);
      for (int p = lane; p < NP; p += 32) {
        const int y = p / P - h, x = p % P - h;
        sT[p] = pagk_sample(I1, pitch, cols, rows, ptx + (float)x, pty + (float)y);
        sJ[p * 5 + 2] = (double)cval;
        sJ[p * 5 + 3] = 1.0;
      }
      for (int iter = 0; iter < mode.iterations; ++iter) {
        ++n_iter;
        const float bx = ptx + dx, by = pty + dy;
        const float gain = 1.0f + dg;
        for (int p = lane; p < NP; p += 32) {
          const float sx = bx + sWx[p], sy = by + sWy[p];
          const float e = (pagk_sample(I2, pitch, cols, rows, sx, sy) + db) - gain * sT[p];
          const float gx = pagk_sample(I2, pitch, cols, rows, sx + 1.0f, sy) - pagk_sample(I2, pitch, cols, rows, sx - 1.0f, sy);
          const float gy = pagk_sample(I2, pitch, cols, rows, sx, sy + 1.0f) - pagk_sample(I2, pitch, cols, rows, sx, sy - 1.0f);
          sJ[p * 5 + 0] = 0.5 * (double)gx;  // (float)(0.5 * (double)diff) is an exact halving
          sJ[p * 5 + 1] = 0.5 * (double)gy;
          sJ[p * 5 + 4] = -(double)e;
          sE2[p] = e * e;
        }
        __syncwarp();
        double acc = 0.0;
        cost = 0.f;
#pragma unroll 4
        for (int p = 0; p < NP; ++p) {
          acc = fma(sJ[p * 5 + ia], sJ[p * 5 + ib], acc);
          cost = cost + sE2[p];
        }
        if (lane < 14) sRed[lane] = acc;
        __syncwarp();
        double h00 = sRed[0], h10 = sRed[1], h20 = sRed[2], h30 = sRed[3], h11 = sRed[4], h21 = sRed[5], h31 = sRed[6],
               h22 = sRed[7], h32 = sRed[8], h33 = sRed[9], b0 = sRed[10], b1 = sRed[11], b2 = sRed[12], b3 = sRed[13];
        __syncwarp();
        if (mode.regular) {  // reference src/patch_match.cpp:302-314
          const double d = (double)sqrtf(dx * dx + dy * dy);
          const float li = mode.lambda * mode.inv_log_max_dist;
          const double ad1 = (double)mode.alpha * d + 1.0;
          const double e_pen = (double)li * log(ad1);
          const double jx = ((double)(li * mode.alpha) / ad1) * ((double)dx / d);
          const double jy = ((double)(li * mode.alpha) / ad1) * ((double)dy / d);
          h00 += jx * jx; h10 += jy * jx; h11 += jy * jy;
          h20 += 0.0 * jx; h21 += 0.0 * jy; h30 += 0.0 * jx; h31 += 0.0 * jy;
          b0 += jx * e_pen; b1 += jy * e_pen; b2 += 0.0 * e_pen; b3 += 0.0 * e_pen;
          cost = (float)((double)cost + e_pen * e_pen);
        }
        double u0, u1, u2, u3;
        pagk_llt_solve4(h00, h10, h11, h20, h21, h22, h30, h31, h32, h33, b0, b1, b2, b3, u0, u1, u2, u3);
        if (isnan(u0)) { succ = false; break; }
        if (iter > 0 && cost > lastCost) break;
        dx = (float)((double)dx + u0);
        dy = (float)((double)dy + u1);
        if (mode.illum) { dg = (float)((double)dg + u2); db = (float)((double)db + u3); }
        lastCost = cost;
        succ = true;
        const double nrm = sqrt((u0 * u0 + u2 * u2) + (u1 * u1 + u3 * u3));
        if (nrm < 1e-2) break;
      }
      pt2.x = ptx + dx; pt2.y = pty + dy;
    }
  }
  if (lane == 0) {
    out.pm_un[o] = pt2;
    out.pm_status[o] = (gst && succ) ? 1 : 0;
    out.pix_err[o] = gst ? sqrt((double)lastCost * mode.win_size_inv) : 0.0;
    out.ncc[o] = gst ? 1.0f : 0.0f;
    out.iters[o] = n_iter;
  }
}

// =================================================================================================
// PatchMatch::NCC (reference src/patch_match.cpp:433-469), run when bCalculateNCC_ is set (off by default,
// include/patch_match.h:49).  The reference evaluates it after every level with the level-0 images and keeps the
// last value (:356-366), i.e. the one taken at the final level-0 position.  Zero-mean NCC over the P x P samples in
// x-outer / y-inner order with sequential float sums: one thread per feature walks them in that order (the second
// pass recomputes the samples instead of storing them; the values are identical).
// =================================================================================================
__global__ void __launch_bounds__(128) pagk_ncc_kernel(const unsigned char *__restrict__ images, PagkGeom g,
                                                     const PagkPairConst *__restrict__ pcs,
                                                     const float2 *__restrict__ keys_un, PagkOutPtrs out, PagkMode mode,
                                                     int max_keys) {
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pcs[pair].n_keys) return;
  const size_t o = (size_t)pair * max_keys + i;
  if (!out.gyro_status[o]) return;  // skipped at :173, mvNcc stays 0
  const unsigned char *I1 = pagk_level_ptr(images, g, pair, 0, 0), *I2 = pagk_level_ptr(images, g, pair, 1, 0);
  const int cols = g.lv[0].cols, rows = g.lv[0].rows, pitch = g.lv[0].pitch, h = mode.half;
  const float2 pr = keys_un[o], pc = out.pm_un[o];
  const float4 A = out.affine[o];
  float mean_ref = 0.f, mean_cur = 0.f;
  for (int x = -h; x <= h; ++x)
    for (int y = -h; y <= h; ++y) {
      float wx = (float)x, wy = (float)y;
      if (mode.affine) { wx = A.x * (float)x + A.y * (float)y; wy = A.z * (float)x + A.w * (float)y; }
      mean_ref += pagk_sample(I1, pitch, cols, rows, pr.x + (float)x, pr.y + (float)y);
      mean_cur += pagk_sample(I2, pitch, cols, rows, pc.x + wx, pc.y + wy);
    }
  const float n = (float)((2 * h + 1) * (2 * h + 1));
  mean_ref /= n;
  mean_cur /= n;
  float num = 0.f, d1 = 0.f, d2 = 0.f;
  for (int x = -h; x <= h; ++x)
    for (int y = -h; y <= h; ++y) {
      float wx = (float)x, wy = (float)y;
      if (mode.affine) { wx = A.x * (float)x + A.y * (float)y; wy = A.z * (float)x + A.w * (float)y; }
      const float a = pagk_sample(I1, pitch, cols, rows, pr.x + (float)x, pr.y + (float)y) - mean_ref;
      const float b = pagk_sample(I2, pitch, cols, rows, pc.x + wx, pc.y + wy) - mean_cur;
      num += a * b;
      d1 += a * a;
      d2 += b * b;
    }
  out.ncc[o] = (float)((double)num / sqrt((double)(d1 * d2) + 1e-10));
}

// =================================================================================================
// K4: DistortPoints + SetMatcher + the threshold filter.  One CTA per frame pair.  The mean pixel
// error is a double sum in feature-index order (src/gyro_aided_tracker.cpp:294-305): thread 0 adds
// chunk after chunk from shared memory so that the rounding sequence is the reference's.
// =================================================================================================
// Pairs with many features (BASELINE configs C and E: 8192 and 32768) as three launches, because one CTA per pair walking its
// features in strides leaves the machine to a handful of CTAs (config E, four pairs: 0.65 ms): the two elementwise phases over a
// grid of features, and between them the only serial part -- the mean pixel error, an ordered double sum per pair -- by
// one thread that adds from shared memory while the other warps of its CTA fetch the next chunk.
#define EPIW_THREADS 256
__global__ void __launch_bounds__(EPIW_THREADS) pagk_epilogue_points_kernel(const PagkPairConst *__restrict__ pcs, PagkOutPtrs out, int max_keys) {
  const int pair = blockIdx.y, i = blockIdx.x * EPIW_THREADS + threadIdx.x;
  const PagkPairConst &c = pcs[pair];
  if (i >= c.n_keys) return;
  const size_t o = (size_t)pair * max_keys + i;
  // distort + drift distance (src/patch_match.cpp:378-387, 409-416)
  const float2 p = out.pm_un[o];
  out.pm[o] = (c.k1 == 0.0f) ? p : pagk_distort(c, p);
  const float2 base = out.pt_predict_un[o];  // still the gyro prediction here (mvPtPredictUn, :384)
  const float ddx = base.x - p.x, ddy = base.y - p.y;
  out.dist[o] = (double)sqrtf(ddx * ddx + ddy * ddy);
}

#define EPIS_THREADS 1024
#define EPIS_CHUNK 2048
__global__ void __launch_bounds__(EPIS_THREADS) pagk_epilogue_mean_kernel(const PagkPairConst *__restrict__ pcs, PagkOutPtrs out, int max_keys,
                                                                      PagkPairResult *__restrict__ res) {
  __shared__ double s_err[2][EPIS_CHUNK];
  const int pair = blockIdx.x, t = threadIdx.x;
  const int N = pcs[pair].n_keys;
  const size_t o0 = (size_t)pair * max_keys;
  // a feature that is not ok contributes +0.0, which leaves a sum >= +0 unchanged
  auto fetch = [&](int c0, int b, int first, int stride) {
    int n = 0;
    for (int i = first; i < EPIS_CHUNK; i += stride) {
      const bool ok = c0 + i < N && out.pm_status[o0 + c0 + i] != 0;
      s_err[b][i] = ok ? out.pix_err[o0 + c0 + i] : 0.0;
      n += ok ? 1 : 0;
    }
    return n;
  };
  int my_ok = fetch(0, 0, t, EPIS_THREADS);
  double sum = 0.0;
  for (int c0 = 0, b = 0; c0 < N; c0 += EPIS_CHUNK, b ^= 1) {
    __syncthreads();  // chunk c0 is in buffer b; the other buffer has been summed
    if (t >= 32) {    // warps 1..31 fetch the next chunk
      if (c0 + EPIS_CHUNK < N) my_ok += fetch(c0 + EPIS_CHUNK, b ^ 1, t - 32, EPIS_THREADS - 32);
    } else if (t == 0) {
      // one thread, index order: the rounding sequence is the reference's (src/gyro_aided_tracker.cpp:294-305)
      const int m = min(EPIS_CHUNK, N - c0);
      const double *e = s_err[b];
#pragma unroll 16
      for (int i = 0; i < m; ++i) sum += e[i];
    }
  }
  __shared__ int s_cnt;
  if (t == 0) s_cnt = 0;
  __syncthreads();
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) my_ok += __shfl_xor_sync(0xffffffffu, my_ok, d);
  if ((t & 31) == 0) atomicAdd(&s_cnt, my_ok);
  __syncthreads();
  if (t == 0) {
    const int cnt = s_cnt;
    res[pair].avg_pixel_error = sum / (double)cnt;  // cnt == 0 -> NaN -> threshold falls back to half (:305-308)
    res[pair].cnt_pm_ok = cnt;
    res[pair].n_predict = 0;
    res[pair].n_iterations = 0;  // the filter kernel adds to both
  }
}

__global__ void __launch_bounds__(EPIW_THREADS) pagk_epilogue_filter_kernel(const PagkPairConst *__restrict__ pcs, PagkOutPtrs out, PagkMode mode,
                                                                        int max_keys, PagkPairResult *__restrict__ res, int do_filter) {
  const int pair = blockIdx.y, i = blockIdx.x * EPIW_THREADS + threadIdx.x;
  const int N = pcs[pair].n_keys;
  const size_t o = (size_t)pair * max_keys + i;
  const double avg = res[pair].avg_pixel_error, hp = (double)mode.half;
  const double thPix = (4.0 * avg > hp) ? 4.0 * avg : hp, thDist = hp * 4.0;
  int kc = 0;
  unsigned long long its = 0ull;
  if (i < N) {
    its = (unsigned long long)out.iters[o];
    if (do_filter) {
      const bool keep = out.pm_status[o] && out.pix_err[o] < thPix && out.dist[o] < thDist;
      if (keep) {
        out.pt_predict[o] = out.pm[o];
        out.pt_predict_un[o] = out.pm_un[o];
      }
      out.status[o] = keep ? 1 : 0;
      kc = keep ? 1 : 0;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kc += __shfl_xor_sync(0xffffffffu, kc, d);
    its += __shfl_xor_sync(0xffffffffu, its, d);
  }
  if ((threadIdx.x & 31) == 0) {
    if (kc) atomicAdd(&res[pair].n_predict, kc);
    if (its) atomicAdd(reinterpret_cast<unsigned long long *>(&res[pair].n_iterations), its);
  }
}

// The same for pairs of at most EPI1_THREADS features (the usual case: BASELINE config B has 1024): one feature per
// thread, every input read once and kept in registers across the three phases, so the kernel is one global round trip,
// the ordered sum, and the stores -- the general kernel above walks its features in four dependent rounds per phase.
#define EPI1_THREADS 1024

__global__ void __launch_bounds__(EPI1_THREADS) pagk_epilogue1_kernel(const PagkPairConst *__restrict__ pcs, PagkOutPtrs out,
                                                                   PagkMode mode, int max_keys,
                                                                   PagkPairResult *__restrict__ res, int do_filter) {
  __shared__ double s_err[EPI1_THREADS];
  __shared__ double s_th[2];
  __shared__ int s_cnt;
  __shared__ unsigned long long s_it;
  const int pair = blockIdx.x, t = threadIdx.x;
  const PagkPairConst &c = pcs[pair];
  const int N = c.n_keys;
  const size_t o = (size_t)pair * max_keys + t;
  const bool mine = t < N;
  if (t == 0) { s_cnt = 0; s_it = 0ull; }
  float2 p = make_float2(0.f, 0.f), pd = p;
  double err = 0.0, dist = 0.0;
  unsigned char ok = 0;
  int it = 0;
  if (mine) {
    p = out.pm_un[o];
    const float2 base = out.pt_predict_un[o];  // still the gyro prediction here (mvPtPredictUn, src/patch_match.cpp:384)
    err = out.pix_err[o];
    ok = out.pm_status[o];
    it = out.iters[o];
    // distort + drift distance (src/patch_match.cpp:378-387, 409-416)
    pd = (c.k1 == 0.0f) ? p : pagk_distort(c, p);
    const float ddx = base.x - p.x, ddy = base.y - p.y;
    dist = (double)sqrtf(ddx * ddx + ddy * ddy);
    out.pm[o] = pd;
    out.dist[o] = dist;
  }
  // a feature that is not ok contributes +0.0, which leaves a sum >= +0 unchanged
  s_err[t] = (mine && ok) ? err : 0.0;
  const int n_ok = __syncthreads_count(mine && ok);
  if (t == 0) {
    // one thread, index order: the rounding sequence is the reference's (src/gyro_aided_tracker.cpp:294-305)
    double sum = 0.0;
#pragma unroll 8
    for (int i = 0; i < N; ++i) sum += s_err[i];
    const double avg = sum / (double)n_ok;  // n_ok == 0 -> NaN -> threshold falls back to half (:305-308)
    const double hp = (double)mode.half;
    s_th[0] = (4.0 * avg > hp) ? 4.0 * avg : hp;
    s_th[1] = hp * 4.0;
    res[pair].avg_pixel_error = avg;
    res[pair].cnt_pm_ok = n_ok;
  }
  __syncthreads();
  bool keep = false;
  if (mine && do_filter) {
    keep = ok && err < s_th[0] && dist < s_th[1];
    if (keep) {
      out.pt_predict[o] = pd;
      out.pt_predict_un[o] = p;
    }
    out.status[o] = keep ? 1 : 0;
  }
  // per-warp partial sums, then one shared atomic per warp
  int kc = keep ? 1 : 0;
  unsigned long long its = (unsigned long long)it;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    kc += __shfl_xor_sync(0xffffffffu, kc, d);
    its += __shfl_xor_sync(0xffffffffu, its, d);
  }
  if ((t & 31) == 0) { atomicAdd(&s_cnt, kc); atomicAdd(&s_it, its); }
  __syncthreads();
  if (t == 0) { res[pair].n_predict = s_cnt; res[pair].n_iterations = (long long)s_it; }
}

// counts the gyro-predicted features of a pair (eType 1: TrackFeatures returns GyroPredictFeatures())
__global__ void __launch_bounds__(256) pagk_count_status_kernel(const PagkPairConst *__restrict__ pcs, PagkOutPtrs out,
                                                              int max_keys, PagkPairResult *__restrict__ res) {
  __shared__ int s_cnt;
  const int pair = blockIdx.x;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int c = 0;
  for (int i = threadIdx.x; i < pcs[pair].n_keys; i += blockDim.x) c += out.status[(size_t)pair * max_keys + i] ? 1 : 0;
  atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) { res[pair].n_predict = s_cnt; res[pair].n_iterations = 0; res[pair].cnt_pm_ok = 0; res[pair].avg_pixel_error = 0.0; }
}

// =================================================================================================
// GeometryValidation without its two RANSAC estimators (reference src/gyro_aided_tracker.cpp:429-508): CheckHomography
// (:589-678, symmetric transfer error) and CheckFundamental (:680-768, point-to-epipolar-line distance) on the status-1
// correspondences, the model choice RH = SH / (SH + SF) > 0.45 and the outlier marking.  One CTA per pair.  The chi-square
// terms are lane-parallel; each score is a float accumulated in correspondence order, two terms per correspondence, by one
// thread (thread 0: homography, thread 32: fundamental matrix), chunk after chunk from shared memory.
// While the scores are being summed status[i] carries the two inlier bits: 1 | inH << 1 | inF << 2.
// =================================================================================================
#define GEO_THREADS 1024

__global__ void __launch_bounds__(GEO_THREADS) pagk_geometry_kernel(const PagkGeoModel *__restrict__ models,
                                                                  const float2 *__restrict__ keys_un,
                                                                  const float2 *__restrict__ pred_un,
                                                                  unsigned char *__restrict__ status, int max_keys,
                                                                  PagkGeoResult *__restrict__ res) {
  __shared__ float s_chi[4][GEO_THREADS];
  __shared__ unsigned char s_st[GEO_THREADS];
  __shared__ float s_score[2];
  __shared__ int s_use;
  const int pair = blockIdx.x, t = threadIdx.x;
  const PagkGeoModel &M = models[pair];
  const int N = M.n_keys;
  const size_t o0 = (size_t)pair * max_keys;
  int cand = 0;
  for (int c0 = 0; c0 < N; c0 += GEO_THREADS) cand += __syncthreads_count(c0 + t < N && status[o0 + c0 + t] != 0);
  if (cand <= 8) {  // if (vPts1.size() > 8) ... : nothing is validated (:446)
    if (t == 0) { PagkGeoResult r; r.score_H = 0.f; r.score_F = 0.f; r.used_H = 0; r.n_candidates = cand; r.n_inlier = 0; r.pad = 0; res[pair] = r; }
    return;
  }
  const double h11 = M.H21[0], h12 = M.H21[1], h13 = M.H21[2], h21 = M.H21[3], h22 = M.H21[4], h23 = M.H21[5], h31 = M.H21[6],
               h32 = M.H21[7], h33 = M.H21[8];
  const double h11inv = M.H12[0], h12inv = M.H12[1], h13inv = M.H12[2], h21inv = M.H12[3], h22inv = M.H12[4], h23inv = M.H12[5],
               h31inv = M.H12[6], h32inv = M.H12[7], h33inv = M.H12[8];
  const double f11 = M.F21[0], f12 = M.F21[1], f13 = M.F21[2], f21 = M.F21[3], f22 = M.F21[4], f23 = M.F21[5], f31 = M.F21[6],
               f32 = M.F21[7], f33 = M.F21[8];
  const float thH = 5.99, thF = 3.84, thScore = 5.99;
  const float invSigmaSquare = (float)(1.0 / (double)(M.sigma * M.sigma));
  float sH = 0.f, sF = 0.f;
  for (int c0 = 0; c0 < N; c0 += GEO_THREADS) {
    const int i = c0 + t;
    unsigned char st = 0;
    if (i < N && status[o0 + i]) {
      const float2 p1 = keys_un[o0 + i], p2 = pred_un[o0 + i];
      const float u1 = p1.x, v1 = p1.y, u2 = p2.x, v2 = p2.y;
      // CheckHomography (:631-664)
      const float w1in2inv = (float)(1.0 / (h31 * (double)u1 + h32 * (double)v1 + h33));
      const float u1in2 = (float)((h11 * (double)u1 + h12 * (double)v1 + h13) * (double)w1in2inv);
      const float v1in2 = (float)((h21 * (double)u1 + h22 * (double)v1 + h23) * (double)w1in2inv);
      const float chi2H = ((u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2)) * invSigmaSquare;
      const float w2in1inv = (float)(1.0 / (h31inv * (double)u2 + h32inv * (double)v2 + h33inv));
      const float u2in1 = (float)((h11inv * (double)u2 + h12inv * (double)v2 + h13inv) * (double)w2in1inv);
      const float v2in1 = (float)((h21inv * (double)u2 + h22inv * (double)v2 + h23inv) * (double)w2in1inv);
      const float chi1H = ((u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1)) * invSigmaSquare;
      // CheckFundamental (:721-754)
      const float a2 = (float)(f11 * (double)u1 + f12 * (double)v1 + f13);
      const float b2 = (float)(f21 * (double)u1 + f22 * (double)v1 + f23);
      const float c2 = (float)(f31 * (double)u1 + f32 * (double)v1 + f33);
      const float num2 = a2 * u2 + b2 * v2 + c2;
      const float chi2F = (num2 * num2 / (a2 * a2 + b2 * b2)) * invSigmaSquare;
      const float a1 = (float)((double)u2 * f11 + (double)v2 * f21 + f31);
      const float b1 = (float)((double)u2 * f12 + (double)v2 * f22 + f32);
      const float c1 = (float)((double)u2 * f13 + (double)v2 * f23 + f33);
      const float num1 = a1 * u1 + b1 * v1 + c1;
      const float chi1F = (num1 * num1 / (a1 * a1 + b1 * b1)) * invSigmaSquare;
      const bool inH = !(chi2H > thH) && !(chi1H > thH), inF = !(chi2F > thF) && !(chi1F > thF);
      st = (unsigned char)(1 | (inH ? 2 : 0) | (inF ? 4 : 0));
      s_chi[0][t] = chi2H; s_chi[1][t] = chi1H; s_chi[2][t] = chi2F; s_chi[3][t] = chi1F;
      status[o0 + i] = st;
    }
    s_st[t] = st;
    __syncthreads();
    const int m = min(GEO_THREADS, N - c0);
    if (t == 0) {
      for (int j = 0; j < m; ++j)
        if (s_st[j]) {
          const float x2 = s_chi[0][j], x1 = s_chi[1][j];
          if (!(x2 > thH)) sH += thH - x2;
          if (!(x1 > thH)) sH += thH - x1;
        }
    } else if (t == 32) {
      for (int j = 0; j < m; ++j)
        if (s_st[j]) {
          const float x2 = s_chi[2][j], x1 = s_chi[3][j];
          if (!(x2 > thF)) sF += thScore - x2;
          if (!(x1 > thF)) sF += thScore - x1;
        }
    }
    __syncthreads();
  }
  if (t == 0) s_score[0] = sH;
  if (t == 32) s_score[1] = sF;
  __syncthreads();
  if (t == 0) {
    const float RH = s_score[0] / (s_score[1] + s_score[0]);  // :466
    s_use = ((double)RH > 0.45) ? 1 : 0;
  }
  __syncthreads();
  const int useH = s_use;
  int inl = 0;
  for (int c0 = 0; c0 < N; c0 += GEO_THREADS) {
    const int i = c0 + t;
    bool keep = false;
    if (i < N) {
      const unsigned char st = status[o0 + i];
      keep = (st & 1) && (useH ? (st & 2) : (st & 4));
      status[o0 + i] = keep ? 1 : 0;
    }
    inl += __syncthreads_count(keep);
  }
  if (t == 0) {
    PagkGeoResult r;
    r.score_H = s_score[0]; r.score_F = s_score[1]; r.used_H = useH; r.n_candidates = cand; r.n_inlier = inl; r.pad = 0;
    res[pair] = r;
  }
}

// =================================================================================================
// Frame::SetPredictKeyPointsAndMask (reference src/frame.cpp:115-153): the surviving predictions of a pair become the
// reference keypoints of the next one.  One CTA per pair; survivors are compacted in index order (warp ballots + a scan of
// the warp totals), each survivor zeroes its 14 x 14 square of the occupancy mask (the mask is set to ones beforehand;
// overlapping squares all write zero).
// =================================================================================================
#define CARRY_THREADS 1024

__global__ void __launch_bounds__(CARRY_THREADS) pagk_carry_kernel(const PagkCarryConst *__restrict__ ccs,
                                                                const float2 *__restrict__ pt_predict,
                                                                const float2 *__restrict__ pt_predict_un,
                                                                const unsigned char *__restrict__ status,
                                                                const float2 *__restrict__ normal_last, int max_keys,
                                                                float2 *__restrict__ keys, float2 *__restrict__ keys_un,
                                                                float2 *__restrict__ keys_normal, int *__restrict__ index_in_last,
                                                                float2 *__restrict__ flow_last, int *__restrict__ n_out,
                                                                unsigned char *__restrict__ mask, unsigned long long mask_stride) {
  __shared__ int s_warp[CARRY_THREADS / 32];
  __shared__ int s_base;
  const int pair = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const PagkCarryConst c = ccs[pair];
  const size_t o0 = (size_t)pair * max_keys;
  unsigned char *m = mask ? mask + (size_t)pair * mask_stride : nullptr;
  const int half_path_size = 7;
  if (t == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < c.n_keys; c0 += CARRY_THREADS) {
    const int i = c0 + t;
    const bool alive = i < c.n_keys && status[o0 + i] != 0;
    const unsigned ballot = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    if (alive) {
      const int k = before + __popc(ballot & ((1u << lane) - 1u));
      const float2 pd = pt_predict[o0 + i], pu = pt_predict_un[o0 + i], nl = normal_last[o0 + i];
      const float2 pn = make_float2((pu.x - c.cx) * c.fx_inv, (pu.y - c.cy) * c.fy_inv);
      keys[o0 + k] = pd; keys_un[o0 + k] = pu; keys_normal[o0 + k] = pn;
      index_in_last[o0 + k] = i;
      // dv = dpt_normal / (mTimeStamp - mpLastFrame->mTimeStamp): Point2f / double = float(double(x) / dt)
      flow_last[o0 + k] = make_float2((float)((double)(pn.x - nl.x) / c.dt), (float)((double)(pn.y - nl.y) / c.dt));
      if (m) {
        const int x = min(max(0, (int)pu.x - half_path_size), c.width - 2 * half_path_size);
        const int y = min(max(0, (int)pu.y - half_path_size), c.height - 2 * half_path_size);
        for (int r = 0; r < 2 * half_path_size; ++r) {
          unsigned char *row = m + (size_t)(y + r) * c.width + x;
#pragma unroll
          for (int q = 0; q < 2 * half_path_size; ++q) row[q] = 0;
        }
      }
    }
    __syncthreads();
    if (t == 0) { int tot = 0; for (int w = 0; w < CARRY_THREADS / 32; ++w) tot += s_warp[w]; s_base += tot; }
    __syncthreads();
  }
  if (t == 0) n_out[pair] = s_base;
}

// =================================================================================================
// cv::FAST(image, keypoints, threshold, nonmaxSuppression, TYPE_9_16) -- OpenCV modules/features2d/src/fast.cpp.
// A pixel is a corner when nine contiguous pixels of the 16-pixel circle are all darker than v - t or all brighter than
// v + t; its score is cornerScore<16> (the largest threshold for which it still is one); with non-maximum suppression a
// corner is kept when its score is strictly greater than the scores of its eight neighbours.  Keypoints come out in
// row-major order.  Four small kernels: scores (one thread per pixel), keep flags + per-row counts (one warp per row),
// scan of the row counts (one block), ordered emit (one warp per row).  score image: (1 << 8 | score) for a corner, else 0.
// =================================================================================================
__constant__ int c_fast_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_fast_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

__global__ void __launch_bounds__(256) pagk_fast_score_kernel(const unsigned char *__restrict__ img, int cols, int rows, int threshold,
                                                            int nonmax, unsigned short *__restrict__ score) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= cols || y >= rows) return;
  unsigned short out = 0;
  if (x >= 3 && x < cols - 3 && y >= 3 && y < rows - 3) {
    const unsigned char *ptr = img + (size_t)y * cols + x;
    const int v = ptr[0];
    int d[25];
    unsigned dark = 0, bright = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int p = ptr[c_fast_dx[k] + c_fast_dy[k] * cols];
      d[k] = v - p;
      dark |= (p < v - threshold ? 1u : 0u) << k;
      bright |= (p > v + threshold ? 1u : 0u) << k;
    }
#pragma unroll
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    // nine contiguous set bits on the circle: AND of the mask with its rotations by 1..8
    auto arc9 = [](unsigned m) {
      m |= m << 16;
      unsigned a = m;
#pragma unroll
      for (int r = 1; r <= 8; ++r) a &= (m >> r);
      return (a & 0xffffu) != 0u;
    };
    if (arc9(dark) || arc9(bright)) {
      int sc = 0;
      if (nonmax) {  // cornerScore<16>
        int a0 = threshold;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          int a = min(d[k + 1], d[k + 2]);
          a = min(a, d[k + 3]);
          if (a <= a0) continue;
          a = min(a, d[k + 4]); a = min(a, d[k + 5]); a = min(a, d[k + 6]); a = min(a, d[k + 7]); a = min(a, d[k + 8]);
          a0 = max(a0, min(a, d[k]));
          a0 = max(a0, min(a, d[k + 9]));
        }
        int b0 = -a0;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
          int b = max(d[k + 1], d[k + 2]);
          b = max(b, d[k + 3]); b = max(b, d[k + 4]); b = max(b, d[k + 5]);
          if (b >= b0) continue;
          b = max(b, d[k + 6]); b = max(b, d[k + 7]); b = max(b, d[k + 8]);
          b0 = min(b0, max(b, d[k]));
          b0 = min(b0, max(b, d[k + 9]));
        }
        sc = (-b0 - 1) & 0xff;  // curr[j] = (uchar)cornerScore(...)
      }
      out = (unsigned short)(0x100 | sc);
    }
  }
  score[(size_t)y * cols + x] = out;
}

__device__ __forceinline__ bool pagk_fast_keep(const unsigned short *__restrict__ score, const unsigned char *__restrict__ mask, int cols,
                                               int rows, int x, int y, int nonmax) {
  if (x < 3 || x >= cols - 3 || y < 3 || y >= rows - 3) return false;
  const unsigned short *s = score + (size_t)y * cols + x;
  if (!(s[0] & 0x100)) return false;
  if (nonmax) {
    const int s0 = s[0] & 0xff;
    const bool mx = s0 > (s[1] & 0xff) && s0 > (s[-1] & 0xff) && s0 > (s[-cols - 1] & 0xff) && s0 > (s[-cols] & 0xff) &&
                    s0 > (s[-cols + 1] & 0xff) && s0 > (s[cols - 1] & 0xff) && s0 > (s[cols] & 0xff) && s0 > (s[cols + 1] & 0xff);
    if (!mx) return false;
  }
  return !mask || mask[(size_t)y * cols + x] != 0;
}

// one warp per row: keep flags and the row's count
__global__ void __launch_bounds__(256) pagk_fast_rows_kernel(const unsigned short *__restrict__ score, const unsigned char *__restrict__ mask,
                                                           int cols, int rows, int nonmax, unsigned char *__restrict__ keep,
                                                           int *__restrict__ row_count) {
  const int y = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y >= rows) return;
  int cnt = 0;
  for (int x0 = 0; x0 < cols; x0 += 32) {
    const int x = x0 + lane;
    const bool k = x < cols && pagk_fast_keep(score, mask, cols, rows, x, y, nonmax);
    if (x < cols) keep[(size_t)y * cols + x] = k ? 1 : 0;
    cnt += __popc(__ballot_sync(0xffffffffu, k));
  }
  if (lane == 0) row_count[y] = cnt;
}

// exclusive scan of the row counts by one block; row_offset[rows] = total
__global__ void __launch_bounds__(1024) pagk_fast_scan_kernel(const int *__restrict__ row_count, int rows, int *__restrict__ row_offset) {
  __shared__ int s_part[1024];
  const int t = threadIdx.x, per = (rows + 1023) / 1024, b = t * per;
  int sum = 0;
  for (int i = b; i < min(b + per, rows); ++i) sum += row_count[i];
  s_part[t] = sum;
  __syncthreads();
  if (t == 0) { int run = 0; for (int i = 0; i < 1024; ++i) { const int v = s_part[i]; s_part[i] = run; run += v; } row_offset[rows] = run; }
  __syncthreads();
  int run = s_part[t];
  for (int i = b; i < min(b + per, rows); ++i) { row_offset[i] = run; run += row_count[i]; }
}

// one warp per row: keypoints of the row at row_offset[y] + rank within the row
__global__ void __launch_bounds__(256) pagk_fast_emit_kernel(const unsigned short *__restrict__ score, const unsigned char *__restrict__ keep,
                                                           const int *__restrict__ row_offset, int cols, int rows, int max_out,
                                                           float2 *__restrict__ xy, float *__restrict__ response) {
  const int y = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y >= rows) return;
  int base = row_offset[y];
  if (row_offset[y + 1] == base) return;
  for (int x0 = 0; x0 < cols; x0 += 32) {
    const int x = x0 + lane;
    const bool k = x < cols && keep[(size_t)y * cols + x] != 0;
    const unsigned b = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int pos = base + __popc(b & ((1u << lane) - 1u));
      if (pos < max_out) { xy[pos] = make_float2((float)x, (float)y); response[pos] = (float)(score[(size_t)y * cols + x] & 0xff); }
    }
    base += __popc(b);
  }
}

// =================================================================================================
// The detection half of ORBextractor::ComputeKeyPointsOctTree for one level (reference src/ORBextractor.cc:789-852) plus the
// mask filter of DetectFeatures (:1200-1203).  cv::FAST on a cell's sub-image sees the same 16-pixel circles as on the whole
// image, a corner's score does not depend on the threshold it was found with, and it is a corner at threshold T exactly when
// its score is >= T; so one score image at min_th serves both thresholds of every cell.  What is per cell is the non-maximum
// suppression: neighbours outside the cell's interior (its sub-image minus FAST's 3-pixel border) or below the threshold
// count as 0.  The interiors of the cells tile the image without overlap.  One warp per cell.
// =================================================================================================
__device__ __forceinline__ void pagk_cell_box(const PagkCellGrid &g, int ci, int cj, int &x0, int &y0, int &x1, int &y1, bool &live) {
  const float iniY = (float)(g.min_y + ci * g.h_cell), iniX = (float)(g.min_x + cj * g.w_cell);
  float maxY = iniY + (float)g.h_cell + 6.0f, maxX = iniX + (float)g.w_cell + 6.0f;
  live = !(iniY >= (float)(g.max_y - 3)) && !(iniX >= (float)(g.max_x - 6));
  if (maxY > (float)g.max_y) maxY = (float)g.max_y;
  if (maxX > (float)g.max_x) maxX = (float)g.max_x;
  // interior of the sub-image [ini, max): FAST skips a 3-pixel border
  x0 = (int)iniX + 3; y0 = (int)iniY + 3; x1 = (int)maxX - 3; y1 = (int)maxY - 3;
}

__device__ __forceinline__ int pagk_cell_score(const unsigned short *__restrict__ score, int cols, int x, int y, int x0, int y0, int x1,
                                               int y1, int th) {
  if (x < x0 || x >= x1 || y < y0 || y >= y1) return 0;
  const unsigned short s = score[(size_t)y * cols + x];
  return ((s & 0x100) && (int)(s & 0xff) >= th) ? (int)(s & 0xff) : 0;
}

__device__ __forceinline__ bool pagk_cell_keep(const unsigned short *__restrict__ score, int cols, int x, int y, int x0, int y0, int x1,
                                               int y1, int th) {
  const unsigned short s = score[(size_t)y * cols + x];
  if (!(s & 0x100) || (int)(s & 0xff) < th) return false;
  const int s0 = s & 0xff;
  return s0 > pagk_cell_score(score, cols, x + 1, y, x0, y0, x1, y1, th) && s0 > pagk_cell_score(score, cols, x - 1, y, x0, y0, x1, y1, th) &&
         s0 > pagk_cell_score(score, cols, x - 1, y - 1, x0, y0, x1, y1, th) && s0 > pagk_cell_score(score, cols, x, y - 1, x0, y0, x1, y1, th) &&
         s0 > pagk_cell_score(score, cols, x + 1, y - 1, x0, y0, x1, y1, th) && s0 > pagk_cell_score(score, cols, x - 1, y + 1, x0, y0, x1, y1, th) &&
         s0 > pagk_cell_score(score, cols, x, y + 1, x0, y0, x1, y1, th) && s0 > pagk_cell_score(score, cols, x + 1, y + 1, x0, y0, x1, y1, th);
}

__global__ void __launch_bounds__(256) pagk_orb_cells_kernel(const unsigned short *__restrict__ score, const unsigned char *__restrict__ mask,
                                                           int cols, PagkCellGrid g, int ini_th, int min_th,
                                                           unsigned char *__restrict__ keep, int *__restrict__ cell_count) {
  const int cell = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (cell >= g.n_cols * g.n_rows) return;
  const int ci = cell / g.n_cols, cj = cell - ci * g.n_cols;
  int x0, y0, x1, y1;
  bool live;
  pagk_cell_box(g, ci, cj, x0, y0, x1, y1, live);
  const int w = x1 - x0, hgt = y1 - y0, npx = (live && w > 0 && hgt > 0) ? w * hgt : 0;
  // does ini_th leave anything in this cell?
  int n_ini = 0;
  for (int p0 = 0; p0 < npx; p0 += 32) {
    const int p = p0 + lane;
    bool k = false;
    if (p < npx) { const int y = y0 + p / w, x = x0 + p % w; k = pagk_cell_keep(score, cols, x, y, x0, y0, x1, y1, ini_th); }
    n_ini += __popc(__ballot_sync(0xffffffffu, k));
  }
  const int th = n_ini > 0 ? ini_th : min_th;
  int cnt = 0;
  for (int p0 = 0; p0 < npx; p0 += 32) {
    const int p = p0 + lane;
    bool k = false;
    if (p < npx) {
      const int y = y0 + p / w, x = x0 + p % w;
      k = pagk_cell_keep(score, cols, x, y, x0, y0, x1, y1, th) && (!mask || mask[(size_t)y * cols + x] != 0);
      keep[(size_t)y * cols + x] = k ? 1 : 0;
    }
    cnt += __popc(__ballot_sync(0xffffffffu, k));
  }
  if (lane == 0) cell_count[cell] = cnt;
}

__global__ void __launch_bounds__(256) pagk_orb_emit_kernel(const unsigned short *__restrict__ score, const unsigned char *__restrict__ keep,
                                                          const int *__restrict__ cell_offset, int cols, PagkCellGrid g, int max_out,
                                                          float2 *__restrict__ xy, float *__restrict__ response) {
  const int cell = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (cell >= g.n_cols * g.n_rows) return;
  int base = cell_offset[cell];
  if (cell_offset[cell + 1] == base) return;
  const int ci = cell / g.n_cols, cj = cell - ci * g.n_cols;
  int x0, y0, x1, y1;
  bool live;
  pagk_cell_box(g, ci, cj, x0, y0, x1, y1, live);
  const int w = x1 - x0, npx = w * (y1 - y0);
  for (int p0 = 0; p0 < npx; p0 += 32) {
    const int p = p0 + lane;
    bool k = false;
    int x = 0, y = 0;
    if (p < npx) { y = y0 + p / w; x = x0 + p % w; k = keep[(size_t)y * cols + x] != 0; }
    const unsigned b = __ballot_sync(0xffffffffu, k);
    if (k) {
      const int pos = base + __popc(b & ((1u << lane) - 1u));
      if (pos < max_out) { xy[pos] = make_float2((float)x, (float)y); response[pos] = (float)(score[(size_t)y * cols + x] & 0xff); }
    }
    base += __popc(b);
  }
}

// =================================================================================================
// cv::remap(src, dst, map_x, map_y, INTER_LINEAR) for CV_8UC1 and CV_32FC1 maps with the default constant 0 border
// (OpenCV modules/imgproc/src/imgwarp.cpp, remapBilinear): coordinates rounded to 1/32 pixel (cvRound = round half to even),
// integer part saturated to short, 15-bit weights (32 - fy)(32 - fx) * 32 ..., (sum + 2^14) >> 15.  One thread per output pixel.
// =================================================================================================
__global__ void __launch_bounds__(256) pagk_remap_kernel(const unsigned char *__restrict__ src, int cols, int rows,
                                                       const float *__restrict__ map_x, const float *__restrict__ map_y, int dcols,
                                                       int drows, unsigned char *__restrict__ dst) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= dcols || y >= drows) return;
  const size_t o = (size_t)y * dcols + x;
  const int sx = __float2int_rn(map_x[o] * 32.0f), sy = __float2int_rn(map_y[o] * 32.0f);
  const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767), fx = sx & 31, fy = sy & 31;
  const bool x0 = ix >= 0 && ix < cols, x1 = ix + 1 >= 0 && ix + 1 < cols, y0 = iy >= 0 && iy < rows, y1 = iy + 1 >= 0 && iy + 1 < rows;
  const unsigned char *p = src + (long long)iy * cols + ix;
  const int p00 = (x0 && y0) ? p[0] : 0, p01 = (x1 && y0) ? p[1] : 0, p10 = (x0 && y1) ? p[cols] : 0, p11 = (x1 && y1) ? p[cols + 1] : 0;
  const int v = (p00 * ((32 - fy) * (32 - fx) * 32) + p01 * ((32 - fy) * fx * 32) + p10 * (fy * (32 - fx) * 32) + p11 * (fy * fx * 32) + (1 << 14)) >> 15;
  dst[o] = (unsigned char)min(max(v, 0), 255);
}

__device__ __forceinline__ unsigned int pagk_remap_px(const unsigned char *__restrict__ src, int cols, int rows, float mx, float my) {
  const int sx = __float2int_rn(mx * 32.0f), sy = __float2int_rn(my * 32.0f);
  const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767), fx = sx & 31, fy = sy & 31;
  const bool x0 = ix >= 0 && ix < cols, x1 = ix + 1 >= 0 && ix + 1 < cols, y0 = iy >= 0 && iy < rows, y1 = iy + 1 >= 0 && iy + 1 < rows;
  const unsigned char *p = src + (long long)iy * cols + ix;
  const int p00 = (x0 && y0) ? __ldg(p) : 0, p01 = (x1 && y0) ? __ldg(p + 1) : 0, p10 = (x0 && y1) ? __ldg(p + cols) : 0,
            p11 = (x1 && y1) ? __ldg(p + cols + 1) : 0;
  const int v = (p00 * ((32 - fy) * (32 - fx) * 32) + p01 * ((32 - fy) * fx * 32) + p10 * (fy * (32 - fx) * 32) + p11 * (fy * fx * 32) + (1 << 14)) >> 15;
  return (unsigned int)min(max(v, 0), 255);
}

// The maps are the same for every image of a batch, so a thread prepares its four output pixels once -- source offset,
// 1/32-pixel fractions and which of the four taps lie inside the source -- and then walks REMAP_ZB images with them: per
// image and pixel four byte loads, the 15-bit weighted sum and a shift (the map loads, the rounding and the border tests
// are paid once per REMAP_ZB images; 16-byte map loads, one 4-byte store per image).
#define REMAP_ZB 8
struct PagkRemapTap { int off; unsigned int fxy; };  // fxy: fx | fy << 8 | inside mask (p00 p01 p10 p11) << 16
__device__ __forceinline__ PagkRemapTap pagk_remap_prepare(int cols, int rows, float mx, float my) {
  const int sx = __float2int_rn(mx * 32.0f), sy = __float2int_rn(my * 32.0f);
  const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767), fx = sx & 31, fy = sy & 31;
  const bool x0 = ix >= 0 && ix < cols, x1 = ix + 1 >= 0 && ix + 1 < cols, y0 = iy >= 0 && iy < rows, y1 = iy + 1 >= 0 && iy + 1 < rows;
  PagkRemapTap t;
  t.off = iy * cols + ix;
  t.fxy = (unsigned int)fx | ((unsigned int)fy << 8) | ((unsigned int)(x0 && y0) << 16) | ((unsigned int)(x1 && y0) << 17) |
          ((unsigned int)(x0 && y1) << 18) | ((unsigned int)(x1 && y1) << 19);
  return t;
}
__device__ __forceinline__ unsigned int pagk_remap_apply(const unsigned char *__restrict__ src, int cols, const PagkRemapTap t) {
  const int fx = (int)(t.fxy & 31u), fy = (int)((t.fxy >> 8) & 31u);
  const unsigned char *p = src + t.off;
  const int p00 = (t.fxy & (1u << 16)) ? __ldg(p) : 0, p01 = (t.fxy & (1u << 17)) ? __ldg(p + 1) : 0;
  const int p10 = (t.fxy & (1u << 18)) ? __ldg(p + cols) : 0, p11 = (t.fxy & (1u << 19)) ? __ldg(p + cols + 1) : 0;
  const int v = (p00 * ((32 - fy) * (32 - fx) * 32) + p01 * ((32 - fy) * fx * 32) + p10 * (fy * (32 - fx) * 32) + p11 * (fy * fx * 32) + (1 << 14)) >> 15;
  return (unsigned int)min(max(v, 0), 255);
}

__global__ void __launch_bounds__(256, 4) pagk_remap_slots_kernel(const unsigned char *__restrict__ raw, unsigned char *__restrict__ images,
                                                             PagkGeom g, const float *__restrict__ map_x, const float *__restrict__ map_y,
                                                             int n_images, int z_stride, int z_offset) {
  const int cols = g.lv[0].cols, rows = g.lv[0].rows, pitch = g.lv[0].pitch;  // the raw images and the maps are dense
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (y >= rows) return;
  const int z0 = blockIdx.z * REMAP_ZB, z1 = min(z0 + REMAP_ZB, n_images);
  const size_t img_bytes = (size_t)cols * rows;
  if ((cols & 3) == 0) {
    const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    if (x >= cols) return;
    const size_t o = (size_t)y * cols + x;
    const float4 mx = *reinterpret_cast<const float4 *>(map_x + o), my = *reinterpret_cast<const float4 *>(map_y + o);
    const PagkRemapTap t0 = pagk_remap_prepare(cols, rows, mx.x, my.x), t1 = pagk_remap_prepare(cols, rows, mx.y, my.y);
    const PagkRemapTap t2 = pagk_remap_prepare(cols, rows, mx.z, my.z), t3 = pagk_remap_prepare(cols, rows, mx.w, my.w);
#pragma unroll 1
    for (int z = z0; z < z1; ++z) {
      const unsigned char *src = raw + (size_t)z * img_bytes;
      unsigned char *dst = images + (size_t)(z * z_stride + z_offset) * g.slot_bytes + g.lv[0].offset;
      const unsigned int v = pagk_remap_apply(src, cols, t0) | (pagk_remap_apply(src, cols, t1) << 8) |
                             (pagk_remap_apply(src, cols, t2) << 16) | (pagk_remap_apply(src, cols, t3) << 24);
      *reinterpret_cast<unsigned int *>(dst + (size_t)y * pitch + x) = v;
    }
  } else {
    for (int k = 0; k < 4; ++k) {
      const int x = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4 + k;
      if (x >= cols) return;
      const size_t o = (size_t)y * cols + x;
      const PagkRemapTap t = pagk_remap_prepare(cols, rows, map_x[o], map_y[o]);
      for (int z = z0; z < z1; ++z)
        images[(size_t)(z * z_stride + z_offset) * g.slot_bytes + g.lv[0].offset + (size_t)y * pitch + x] =
            (unsigned char)pagk_remap_apply(raw + (size_t)z * img_bytes, cols, t);
    }
  }
}

// =================================================================================================
// The two robust estimators of GeometryValidation (reference src/gyro_aided_tracker.cpp:597, :691; pagk_ransac.h says what
// is and is not shared with OpenCV).  One CTA per (pair, model): model 0 the homography, model 1 the fundamental matrix.
//   1  the status-1 correspondences in index order -> a compact index list (vPts1 / vPts2, :432-440)
//   2  kHypotheses hypotheses, one thread each (four per thread): minimal sample from the counter-based generator, model,
//      inlier count over the whole list; the best one (most inliers, then the lowest hypothesis number) by a block reduction
//   3  refit on the inliers of the best hypothesis: every entry of the 9 x 9 normal matrix is owned by one thread that walks
//      the list in order (deterministic sums), thread 0 takes the null vector (Jacobi) and finishes the model
//   4  homography only: five Gauss-Newton steps on the forward reprojection error of those inliers, same ownership scheme
// A pair with at most eight candidates is left alone (the reference does not validate it, :446).
// =================================================================================================
#define RANSAC_THREADS 256

__device__ __forceinline__ void pagk_ransac_point(const float2 *__restrict__ keys_un, const float2 *__restrict__ pred_un, size_t o,
                                                  double &x, double &y, double &u, double &v) {
  const float2 a = keys_un[o], b = pred_un[o];
  x = (double)a.x; y = (double)a.y; u = (double)b.x; v = (double)b.y;
}

__global__ void __launch_bounds__(RANSAC_THREADS) pagk_ransac_kernel(const float2 *__restrict__ keys_un, const float2 *__restrict__ pred_un,
                                                                   const unsigned char *__restrict__ status, int max_keys, unsigned int seed,
                                                                   int *__restrict__ scratch_idx, unsigned char *__restrict__ scratch_in,
                                                                   PagkGeoModel *__restrict__ models, const unsigned char *__restrict__ estimate) {
  using namespace pagk_ransac;
  const int pair = blockIdx.x, model = blockIdx.y, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (!estimate[pair]) return;
  PagkGeoModel &G = models[pair];
  const int N = G.n_keys;
  const size_t o0 = (size_t)pair * max_keys;
  int *list = scratch_idx + ((size_t)pair * 2 + model) * max_keys;
  unsigned char *inl = scratch_in + ((size_t)pair * 2 + model) * max_keys;
  __shared__ int s_warp[RANSAC_THREADS / 32];
  __shared__ int s_base, s_M, s_best_cnt, s_best_hyp, s_nin;
  __shared__ double s_model[9], s_mat[81], s_vec[16], s_sim[6];
  __shared__ int s_cnt[RANSAC_THREADS / 32], s_hyp[RANSAC_THREADS / 32];
  // ---- 1: compact list, order preserved
  if (t == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < N; c0 += RANSAC_THREADS) {
    const int i = c0 + t;
    const bool f = i < N && status[o0 + i] != 0;
    const unsigned int bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (f) list[off + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (t == 0) { int tot = 0; for (int w = 0; w < RANSAC_THREADS / 32; ++w) tot += s_warp[w]; s_base += tot; }
    __syncthreads();
  }
  const int M = s_base;
  double *out = model == 0 ? G.H21 : G.F21;
  if (M <= 8) {  // nothing is validated: a neutral model
    if (t < 9) out[t] = (model == 0 && (t == 0 || t == 4 || t == 8)) ? 1.0 : 0.0;
    if (model == 0 && t < 9) G.H12[t] = (t == 0 || t == 4 || t == 8) ? 1.0 : 0.0;
    return;
  }
  __threadfence_block();
  __syncthreads();
  // ---- 2: hypotheses
  int best_cnt = -1, best_hyp = 0x7fffffff;
  for (int hyp = t; hyp < kHypotheses; hyp += RANSAC_THREADS) {
    double Mdl[9];
    bool ok;
    if (model == 0) {
      int id[4];
      sample<4>(seed, (unsigned int)pair, (unsigned int)hyp, M, id);
      double x[4], y[4], u[4], v[4];
      for (int k = 0; k < 4; ++k) pagk_ransac_point(keys_un, pred_un, o0 + list[id[k]], x[k], y[k], u[k], v[k]);
      ok = h_from_4(x, y, u, v, Mdl);
    } else {
      int id[8];
      sample<8>(seed, (unsigned int)pair, (unsigned int)hyp + kHypotheses, M, id);
      double x[8], y[8], u[8], v[8];
      for (int k = 0; k < 8; ++k) pagk_ransac_point(keys_un, pred_un, o0 + list[id[k]], x[k], y[k], u[k], v[k]);
      ok = f_from_8(x, y, u, v, Mdl);
    }
    if (!ok) continue;
    int cnt = 0;
    for (int k = 0; k < M; ++k) {
      double x, y, u, v;
      pagk_ransac_point(keys_un, pred_un, o0 + list[k], x, y, u, v);
      const double e = model == 0 ? h_error(Mdl, x, y, u, v) : f_error(Mdl, x, y, u, v);
      cnt += (e <= kThreshold2) ? 1 : 0;
    }
    if (cnt > best_cnt) { best_cnt = cnt; best_hyp = hyp; }  // hypotheses of a thread come in increasing order
  }
  for (int d = 16; d >= 1; d >>= 1) {
    const int oc = __shfl_xor_sync(0xffffffffu, best_cnt, d), oh = __shfl_xor_sync(0xffffffffu, best_hyp, d);
    if (oc > best_cnt || (oc == best_cnt && oh < best_hyp)) { best_cnt = oc; best_hyp = oh; }
  }
  if (lane == 0) { s_cnt[warp] = best_cnt; s_hyp[warp] = best_hyp; }
  __syncthreads();
  if (t == 0) {
    int bc = s_cnt[0], bh = s_hyp[0];
    for (int w = 1; w < RANSAC_THREADS / 32; ++w)
      if (s_cnt[w] > bc || (s_cnt[w] == bc && s_hyp[w] < bh)) { bc = s_cnt[w]; bh = s_hyp[w]; }
    s_best_cnt = bc; s_best_hyp = bh;
    // recompute the winning model (cheaper than keeping four models per thread)
    double Mdl[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (bc >= 0) {
      if (model == 0) {
        int id[4];
        sample<4>(seed, (unsigned int)pair, (unsigned int)bh, M, id);
        double x[4], y[4], u[4], v[4];
        for (int k = 0; k < 4; ++k) pagk_ransac_point(keys_un, pred_un, o0 + list[id[k]], x[k], y[k], u[k], v[k]);
        h_from_4(x, y, u, v, Mdl);
      } else {
        int id[8];
        sample<8>(seed, (unsigned int)pair, (unsigned int)bh + kHypotheses, M, id);
        double x[8], y[8], u[8], v[8];
        for (int k = 0; k < 8; ++k) pagk_ransac_point(keys_un, pred_un, o0 + list[id[k]], x[k], y[k], u[k], v[k]);
        f_from_8(x, y, u, v, Mdl);
      }
    } else if (model == 1) {
      for (int i = 0; i < 9; ++i) Mdl[i] = 0.0;
    }
    for (int i = 0; i < 9; ++i) s_model[i] = Mdl[i];
  }
  __syncthreads();
  // ---- 3: inliers of the best hypothesis, refit on them
  const int need = model == 0 ? 4 : 8;
  if (s_best_cnt >= need) {
    for (int k = t; k < M; k += RANSAC_THREADS) {
      double x, y, u, v;
      pagk_ransac_point(keys_un, pred_un, o0 + list[k], x, y, u, v);
      const double e = model == 0 ? h_error(s_model, x, y, u, v) : f_error(s_model, x, y, u, v);
      inl[k] = (e <= kThreshold2) ? 1 : 0;
    }
    __threadfence_block();
    __syncthreads();
    // Hartley similarities of the inliers: four ordered sums, then two
    if (t < 4) {
      double sum = 0.0;
      int n = 0;
      for (int k = 0; k < M; ++k)
        if (inl[k]) {
          double x, y, u, v;
          pagk_ransac_point(keys_un, pred_un, o0 + list[k], x, y, u, v);
          sum += t == 0 ? x : t == 1 ? y : t == 2 ? u : v;
          ++n;
        }
      s_sim[t] = sum / n;
      if (t == 0) s_nin = n;
    }
    __syncthreads();
    if (t < 2) {
      double sum = 0.0;
      const double cx = s_sim[2 * t], cy = s_sim[2 * t + 1];
      for (int k = 0; k < M; ++k)
        if (inl[k]) {
          double x, y, u, v;
          pagk_ransac_point(keys_un, pred_un, o0 + list[k], x, y, u, v);
          const double px = t == 0 ? x : u, py = t == 0 ? y : v;
          sum += sqrt((px - cx) * (px - cx) + (py - cy) * (py - cy));
        }
      const double d = sum / s_nin;
      s_sim[4 + t] = d > 1e-12 ? 1.4142135623730951 / d : 1.0;
    }
    __syncthreads();
    const Sim t1 = {s_sim[4], s_sim[0], s_sim[1]}, t2 = {s_sim[5], s_sim[2], s_sim[3]};
    if (t < 81) {
      const int r = t / 9, c = t - 9 * r;
      double acc = 0.0;
      if (c >= r) {  // the upper triangle; mirrored below
        for (int k = 0; k < M; ++k)
          if (inl[k]) {
            double x, y, u, v;
            pagk_ransac_point(keys_un, pred_un, o0 + list[k], x, y, u, v);
            const double xn = t1.s * (x - t1.cx), yn = t1.s * (y - t1.cy), un = t2.s * (u - t2.cx), vn = t2.s * (v - t2.cy);
            if (model == 0) {
              double r0[9], r1[9];
              h_rows(xn, yn, un, vn, r0, r1);
              acc += r0[r] * r0[c] + r1[r] * r1[c];
            } else {
              double rr[9];
              f_row(xn, yn, un, vn, rr);
              acc += rr[r] * rr[c];
            }
          }
      }
      s_mat[t] = acc;
    }
    __syncthreads();
    if (t < 81) { const int r = t / 9, c = t - 9 * r; if (c < r) s_mat[t] = s_mat[c * 9 + r]; }
    __syncthreads();
    if (t == 0) {
      double A[81], V[81], vec[9], Mdl[9];
      for (int i = 0; i < 81; ++i) A[i] = s_mat[i];
      const int k = jacobi_smallest<9>(A, V);
      for (int i = 0; i < 9; ++i) vec[i] = V[i * 9 + k];
      const bool ok = model == 0 ? h_denormalise(vec, t1, t2, Mdl) : f_finish(vec, t1, t2, Mdl);
      if (ok) for (int i = 0; i < 9; ++i) s_model[i] = Mdl[i];
    }
    __syncthreads();
    // ---- 4: Gauss-Newton polish of the homography on the same inliers
    if (model == 0) {
      for (int it = 0; it < 5; ++it) {
        if (t < 72) {  // 64 entries of J^T J, 8 of J^T r
          const int r = t < 64 ? t / 8 : t - 64, c = t < 64 ? t - 8 * (t / 8) : -1;
          double acc = 0.0;
          if (c < 0 || c >= r) {
            for (int k = 0; k < M; ++k)
              if (inl[k]) {
                double x, y, u, v, ju[8], jv[8], ru, rv;
                pagk_ransac_point(keys_un, pred_un, o0 + list[k], x, y, u, v);
                h_jacobian(s_model, x, y, u, v, ju, jv, &ru, &rv);
                acc += c < 0 ? (ju[r] * ru + jv[r] * rv) : (ju[r] * ju[c] + jv[r] * jv[c]);
              }
          }
          if (t < 64) s_mat[t] = acc; else s_vec[t - 64] = acc;
        }
        __syncthreads();
        if (t < 64) { const int r = t / 8, c = t - 8 * r; if (c < r) s_mat[t] = s_mat[c * 8 + r]; }
        __syncthreads();
        if (t == 0) {
          double A[64], b[8], d[8];
          for (int i = 0; i < 64; ++i) A[i] = s_mat[i];
          for (int i = 0; i < 8; ++i) b[i] = s_vec[i];
          if (solve8(A, b, d)) for (int i = 0; i < 8; ++i) s_model[i] -= d[i];
        }
        __syncthreads();
      }
    }
  }
  if (t < 9) out[t] = s_model[t];
  if (model == 0 && t == 0) inv3(s_model, G.H12);  // cv::Mat H12 = H21.inv(), src/gyro_aided_tracker.cpp:597
}

// =================================================================================================
// launch wrappers (host)
// =================================================================================================
int pagk_pyramid_fused_max_level() {  // a level's tile has to be at least four pixels wide
  int l = 1, tw = PYR_TW / 2, th = PYR_TH / 2;
  while ((tw >> 1) >= 4 && (th >> 1) > 0) { tw >>= 1; th >>= 1; ++l; }
  return l;
}

int pagk_launch_pyramids(unsigned char *images, const PagkGeom &g, int n_images, int z_stride, cudaStream_t st,
                         long long *launches) {
  // levels 1..n_fused form the exact-2x chain and come out of the fused kernel
  int n_fused = 0;
  for (int l = 1; l < g.levels; ++l) {
    if (g.lv[l - 1].cols == 2 * g.lv[l].cols && g.lv[l - 1].rows == 2 * g.lv[l].rows && l <= pagk_pyramid_fused_max_level()) n_fused = l;
    else break;
  }
  dim3 grid((g.lv[0].cols + PYR_TW - 1) / PYR_TW, (g.lv[0].rows + PYR_TH - 1) / PYR_TH, n_images);
  if ((g.lv[0].cols & 15) == 0) pagk_pyramid_fused_kernel<true><<<grid, 256, 0, st>>>(images, g, n_fused, z_stride);
  else pagk_pyramid_fused_kernel<false><<<grid, 256, 0, st>>>(images, g, n_fused, z_stride);
  ++*launches;
  for (int l = n_fused + 1; l < g.levels; ++l) {
    dim3 gg((g.lv[l].cols + 255) / 256, g.lv[l].rows, n_images);
    pagk_pyramid_general_kernel<<<gg, 256, 0, st>>>(images, g, l, z_stride);
    ++*launches;
  }
  return (int)cudaGetLastError();
}

int pagk_launch_predict(const PagkPairConst *pcs, const float2 *keys_un, const float2 *keys, const PagkOutPtrs &out,
                        const PagkMode &mode, int max_keys, int n_max, int n_pairs, int width, int height,
                        const float *ntab, unsigned long long ntab_stride, cudaStream_t st, long long *launches) {
  if (n_max <= 0 || n_pairs <= 0) return 0;
  dim3 grid((n_max + 127) / 128, n_pairs);
  pagk_gyro_predict_kernel<<<grid, 128, 0, st>>>(pcs, keys_un, keys, out, mode, max_keys, width, height, ntab, ntab_stride);
  ++*launches;
  return (int)cudaGetLastError();
}

// per device, from pagk_create after cudaSetDevice (a function attribute belongs to the device it was set on)
int pagk_configure_kernels() {
  return (int)cudaFuncSetAttribute(pagk_lk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

size_t pagk_lk_smem_per_warp(int half) {
  const size_t NP = (size_t)(2 * half + 1) * (2 * half + 1);
  return NP * 56 + 128;
}

int pagk_launch_lk(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                   const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs, cudaStream_t st,
                   long long *launches) {
  if (n_max <= 0 || n_pairs <= 0) return 0;
  const size_t pw = pagk_lk_smem_per_warp(mode.half);
  int wpc = 8;
  while (wpc > 1 && pw * wpc > 200 * 1024) wpc >>= 1;
  const size_t smem = pw * wpc;
  if (smem > 227 * 1024) return (int)cudaErrorInvalidValue;
  dim3 grid((n_max + wpc - 1) / wpc, n_pairs);
  pagk_lk_kernel<<<grid, wpc * 32, smem, st>>>(images, g, pcs, keys_un, out, mode, max_keys, wpc);
  ++*launches;
  return (int)cudaGetLastError();
}

int pagk_launch_ncc(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                    const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs, cudaStream_t st,
                    long long *launches) {
  if (n_max <= 0 || n_pairs <= 0) return 0;
  dim3 grid((n_max + 127) / 128, n_pairs);
  pagk_ncc_kernel<<<grid, 128, 0, st>>>(images, g, pcs, keys_un, out, mode, max_keys);
  ++*launches;
  return (int)cudaGetLastError();
}

int pagk_launch_epilogue(const PagkPairConst *pcs, const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max,
                         int n_pairs, PagkPairResult *res, int do_filter, cudaStream_t st, long long *launches) {
  if (n_pairs <= 0) return 0;
  if (n_max <= EPI1_THREADS) {
    pagk_epilogue1_kernel<<<n_pairs, EPI1_THREADS, 0, st>>>(pcs, out, mode, max_keys, res, do_filter);
    ++*launches;
  } else {
    const dim3 grid((n_max + EPIW_THREADS - 1) / EPIW_THREADS, n_pairs);
    pagk_epilogue_points_kernel<<<grid, EPIW_THREADS, 0, st>>>(pcs, out, max_keys);
    pagk_epilogue_mean_kernel<<<n_pairs, EPIS_THREADS, 0, st>>>(pcs, out, max_keys, res);
    pagk_epilogue_filter_kernel<<<grid, EPIW_THREADS, 0, st>>>(pcs, out, mode, max_keys, res, do_filter);
    *launches += 3;
  }
  return (int)cudaGetLastError();
}

int pagk_launch_geometry(const PagkGeoModel *models, const float2 *keys_un, const float2 *pred_un, unsigned char *status,
                         int max_keys, int n_pairs, PagkGeoResult *res, cudaStream_t st, long long *launches) {
  if (n_pairs <= 0) return 0;
  pagk_geometry_kernel<<<n_pairs, GEO_THREADS, 0, st>>>(models, keys_un, pred_un, status, max_keys, res);
  ++*launches;
  return (int)cudaGetLastError();
}

int pagk_launch_ransac(const float2 *keys_un, const float2 *pred_un, const unsigned char *status, int max_keys, int n_pairs,
                       unsigned int seed, int *scratch_idx, unsigned char *scratch_in, PagkGeoModel *models,
                       const unsigned char *estimate, cudaStream_t st, long long *launches) {
  if (n_pairs <= 0) return 0;
  pagk_ransac_kernel<<<dim3(n_pairs, 2), RANSAC_THREADS, 0, st>>>(keys_un, pred_un, status, max_keys, seed, scratch_idx, scratch_in, models,
                                                                 estimate);
  ++*launches;
  return (int)cudaGetLastError();
}

int pagk_launch_carry(const PagkCarryConst *cc, const float2 *pt_predict, const float2 *pt_predict_un, const unsigned char *status,
                      const float2 *normal_last, int max_keys, int n_pairs, float2 *keys, float2 *keys_un, float2 *keys_normal,
                      int *index_in_last, float2 *flow_last, int *n_out, unsigned char *mask, unsigned long long mask_stride,
                      cudaStream_t st, long long *launches) {
  if (n_pairs <= 0) return 0;
  pagk_carry_kernel<<<n_pairs, CARRY_THREADS, 0, st>>>(cc, pt_predict, pt_predict_un, status, normal_last, max_keys, keys, keys_un,
                                                      keys_normal, index_in_last, flow_last, n_out, mask, mask_stride);
  ++*launches;
  return (int)cudaGetLastError();
}

int pagk_launch_fast(const unsigned char *img, int cols, int rows, int threshold, int nonmax, const unsigned char *mask,
                     unsigned short *score, unsigned char *keep, int *row_count, int *row_offset, int max_out, float2 *xy,
                     float *response, cudaStream_t st, long long *launches) {
  if (cols <= 0 || rows <= 0) return 0;
  threshold = threshold < 0 ? 0 : threshold > 255 ? 255 : threshold;
  dim3 grid((cols + 31) / 32, (rows + 7) / 8);
  pagk_fast_score_kernel<<<grid, 256, 0, st>>>(img, cols, rows, threshold, nonmax, score);
  pagk_fast_rows_kernel<<<(rows + 7) / 8, 256, 0, st>>>(score, mask, cols, rows, nonmax, keep, row_count);
  pagk_fast_scan_kernel<<<1, 1024, 0, st>>>(row_count, rows, row_offset);
  pagk_fast_emit_kernel<<<(rows + 7) / 8, 256, 0, st>>>(score, keep, row_offset, cols, rows, max_out, xy, response);
  *launches += 4;
  return (int)cudaGetLastError();
}

int pagk_launch_orb_cells(const unsigned char *img, int cols, int rows, int ini_th, int min_th, const unsigned char *mask,
                          unsigned short *score, unsigned char *keep, int *cell_count, int *cell_offset, PagkCellGrid grid,
                          int max_out, float2 *xy, float *response, cudaStream_t st, long long *launches) {
  const int n_cells = grid.n_cols * grid.n_rows;
  if (cols <= 0 || rows <= 0 || n_cells <= 0) return 0;
  const int t = min_th < 0 ? 0 : min_th > 255 ? 255 : min_th;
  dim3 g((cols + 31) / 32, (rows + 7) / 8);
  pagk_fast_score_kernel<<<g, 256, 0, st>>>(img, cols, rows, t, 1, score);
  pagk_orb_cells_kernel<<<(n_cells + 7) / 8, 256, 0, st>>>(score, mask, cols, grid, ini_th, min_th, keep, cell_count);
  pagk_fast_scan_kernel<<<1, 1024, 0, st>>>(cell_count, n_cells, cell_offset);
  pagk_orb_emit_kernel<<<(n_cells + 7) / 8, 256, 0, st>>>(score, keep, cell_offset, cols, grid, max_out, xy, response);
  *launches += 4;
  return (int)cudaGetLastError();
}

int pagk_launch_remap(const unsigned char *src, int cols, int rows, const float *map_x, const float *map_y, int dcols, int drows,
                      unsigned char *dst, cudaStream_t st, long long *launches) {
  if (dcols <= 0 || drows <= 0) return 0;
  dim3 g((dcols + 31) / 32, (drows + 7) / 8);
  pagk_remap_kernel<<<g, 256, 0, st>>>(src, cols, rows, map_x, map_y, dcols, drows, dst);
  ++*launches;
  return (int)cudaGetLastError();
}

int pagk_launch_remap_slots(const unsigned char *raw, unsigned char *images, const PagkGeom &g, const float *map_x, const float *map_y,
                            int n_images, int z_stride, int z_offset, cudaStream_t st, long long *launches) {
  if (n_images <= 0) return 0;
  dim3 grid((g.lv[0].cols + 127) / 128, (g.lv[0].rows + 7) / 8, (n_images + REMAP_ZB - 1) / REMAP_ZB);
  pagk_remap_slots_kernel<<<grid, 256, 0, st>>>(raw, images, g, map_x, map_y, n_images, z_stride, z_offset);
  ++*launches;
  return (int)cudaGetLastError();
}

int pagk_launch_count_status(const PagkPairConst *pcs, const PagkOutPtrs &out, int max_keys, int n_pairs,
                             PagkPairResult *res, cudaStream_t st, long long *launches) {
  if (n_pairs <= 0) return 0;
  pagk_count_status_kernel<<<n_pairs, 256, 0, st>>>(pcs, out, max_keys, res);
  ++*launches;
  return (int)cudaGetLastError();
}
