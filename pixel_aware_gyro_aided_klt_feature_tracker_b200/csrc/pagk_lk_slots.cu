// pagk_lk_slots.cu -- K3, the production patch-alignment kernel (11x11 patches).
//
// Reference: PatchMatch::OpticalFlowMultiLevel + OpticalFlowConsideringIlluminationChange_onePixel,
// src/patch_match.cpp:79-142 and :167-367 (forward-additive Gauss-Newton on (dx, dy, dg, db)).
//
// Why this shape.  One Gauss-Newton pass has two halves with opposite parallelism:
//   A  121 independent pixels: five bilinear samples of the current image each (FP32)
//   B  14 sums over those pixels that MUST be taken in the reference's pixel order in double
//      (H is structurally singular, SURVEY.md F3: any other order changes the solve), then a 4x4
//      LLT and the convergence logic -- serial per feature, independent across features.
// A persistent CTA therefore owns SLOTS = 32 features ("slots") at once:
//   phase A  each of the 8 warps takes 4 slots; its 32 lanes are the pixels of one slot at a time.
//            The current-image window of a slot is staged once per level in shared memory as
//            float (no u8 conversions or 64-bit addressing in the loop); the template value T of
//            each pixel is kept beside the pixel's (Ix, Iy, e) record, so one float4 per pixel
//            is the whole A->B hand-over.
//   phase B  one warp, lane = slot: 12 dependent DFMA chains per lane walk the slot's 121 records
//            in order (LDS.128 per step, conflict-free slot stride), then every lane solves its own
//            4x4 system, applies the update and advances its slot's state machine (next iteration,
//            next level, or fetch the next feature from a global counter).
// Slots refill independently, so lanes never wait for the slowest feature of a group.
//
// Bit-exactness of the window path.  The reference samples at (sx, sy), (sx+-1, sy), (sx, sy+-1) with
// sx+-1 rounded in float.  When no clamp fires: sx-1 is always exact (finer binade), so that sample
// shares floor-1 and the fractional weights of sx; sx+1 may round when it crosses a binade, so its
// weights are recomputed from fl(sx+1) and only the index floor(sx)+1 is assumed -- the one case
// where that fails (fl(sx+1) lands on an integer) is detected and the slot falls back to the
// per-sample global path for that pass, as do patches that touch the image border or do not fit the
// window.  Horizontal interpolations a*d0 + xx*d1 are shared between samples only where the
// reference would compute the identical expression.
#include "pagk_device.cuh"
#include "pagk_kernels.h"

namespace {

constexpr int HALF = 5;
constexpr int P = 2 * HALF + 1;
constexpr int NP = P * P;                 // 121
constexpr int PASSES = (NP + 31) / 32;    // 4
constexpr int SLOTS = 32;
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int SLOTS_PER_WARP = SLOTS / WARPS;
constexpr int WIN_W = 21;                 // window width == row stride (21: rows of a pass land in distinct banks)
constexpr int WIN_H = 18;
constexpr int REC_STRIDE = NP;            // float4 records per slot; 484 words = 4 mod 32 -> LDS.128 conflict-free over slots

struct SlotShared {
  float4 rec[SLOTS][REC_STRIDE];          // (Ix, Iy, e, T) per pixel
  float win[SLOTS][WIN_W * WIN_H];
  // slot state, written by the B lane (and the window origin / cval by the A warp)
  int feat[SLOTS];                        // pair * max_keys + i, or -1
  int pair[SLOTS];
  int level[SLOTS];
  int needs_setup[SLOTS];
  int win_x0[SLOTS], win_y0[SLOTS], win_valid[SLOTS];
  float ptx[SLOTS], pty[SLOTS], dx[SLOTS], dy[SLOTS], dg[SLOTS], db[SLOTS], cval[SLOTS];
  float a00[SLOTS], a01[SLOTS], a10[SLOTS], a11[SLOTS];
  float wxmin[SLOTS], wxmax[SLOTS], wymin[SLOTS], wymax[SLOTS];
  int any_active;
};

__device__ __forceinline__ const unsigned char *level_ptr(const unsigned char *images, const PagkGeom &g, int pair,
                                                          int which, int level) {
  return images + (size_t)(pair * 2 + which) * g.slot_bytes + g.lv[level].offset;
}

// ---------------------------------------------------------------------------------------------
// phase A for one slot (whole warp)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void phase_a_slot(SlotShared &S, int s, int lane, const unsigned char *__restrict__ images,
                                             const PagkGeom &g, const PagkMode &mode) {
  const int level = S.level[s], pair = S.pair[s];
  const unsigned char *I1 = level_ptr(images, g, pair, 0, level);
  const unsigned char *I2 = level_ptr(images, g, pair, 1, level);
  const int cols = g.lv[level].cols, rows = g.lv[level].rows;
  const float ptx = S.ptx[s], pty = S.pty[s];
  float4 *rec = S.rec[s];
  float *win = S.win[s];

  if (S.needs_setup[s]) {  // new level: template values and c = -I1(pt)
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int p = lane + 32 * k;
      if (p < NP) {
        const int y = p / P - HALF, x = p % P - HALF;
        rec[p].w = pagk_sample(I1, cols, rows, ptx + (float)x, pty + (float)y);
      }
    }
    const float c = -pagk_sample(I1, cols, rows, ptx, pty);
    __syncwarp();
    if (lane == 0) { S.cval[s] = c; S.needs_setup[s] = 0; S.win_valid[s] = 0; }
    __syncwarp();
  }

  const float bx = ptx + S.dx[s], by = pty + S.dy[s];
  const float db = S.db[s], gain = 1.0f + S.dg[s];
  const float a00 = S.a00[s], a01 = S.a01[s], a10 = S.a10[s], a11 = S.a11[s];
  // extreme sample coordinates over the patch (monotone in the corner offsets)
  const float x2min = (bx + S.wxmin[s]) - 1.0f, x1max = (bx + S.wxmax[s]) + 1.0f;
  const float y2min = (by + S.wymin[s]) - 1.0f, y1max = (by + S.wymax[s]) + 1.0f;
  bool use_window = (x2min >= 0.0f) && (x1max < (float)cols) && (y2min >= 0.0f) && (y1max < (float)rows);
  int wx0 = 0, wy0 = 0;
  if (use_window) {
    const int ixlo = (int)floorf(x2min), ixhi = (int)floorf(x1max) + 1;
    const int iylo = (int)floorf(y2min), iyhi = (int)floorf(y1max) + 1;
    const int needw = ixhi - ixlo + 1, needh = iyhi - iylo + 1;
    if (needw > WIN_W || needh > WIN_H) {
      use_window = false;
    } else {
      wx0 = S.win_x0[s]; wy0 = S.win_y0[s];
      const bool ok = S.win_valid[s] && ixlo >= wx0 && ixhi <= wx0 + WIN_W - 1 && iylo >= wy0 && iyhi <= wy0 + WIN_H - 1;
      if (!ok) {  // (re)stage, centred on the needed box
        wx0 = ixlo - (WIN_W - needw) / 2;
        wy0 = iylo - (WIN_H - needh) / 2;
        __syncwarp();
        for (int idx = lane; idx < WIN_W * WIN_H; idx += 32) {
          const int j = idx % WIN_W, i = idx / WIN_W;
          const int gx = wx0 + j, gy = wy0 + i;
          float v = 0.0f;
          if (gx >= 0 && gx <= cols && gy >= 0 && gy <= rows) v = (float)__ldg(I2 + (size_t)gy * cols + gx);
          win[idx] = v;
        }
        if (lane == 0) { S.win_x0[s] = wx0; S.win_y0[s] = wy0; S.win_valid[s] = 1; }
        __syncwarp();
      }
    }
  }

  bool bad = false;
  if (use_window) {
#pragma unroll
    for (int k = 0; k < PASSES; ++k) {
      const int p = lane + 32 * k;
      if (p < NP) {
        const float xf = (float)(p % P - HALF), yf = (float)(p / P - HALF);
        float wx = xf, wy = yf;
        if (mode.affine) { wx = a00 * xf + a01 * yf; wy = a10 * xf + a11 * yf; }
        const float sx = bx + wx, sy = by + wy;
        const float fx = floorf(sx), fy = floorf(sy);
        const float xx = sx - fx, yy = sy - fy;
        const float a = 1.0f - xx, b = 1.0f - yy;
        const float X1 = sx + 1.0f, Y1 = sy + 1.0f;
        const float fX1 = floorf(X1), fY1 = floorf(Y1);
        const float xx1 = X1 - fX1, yy1 = Y1 - fY1;
        const float a1 = 1.0f - xx1, b1 = 1.0f - yy1;
        bad |= (fX1 != fx + 1.0f) | (fY1 != fy + 1.0f);
        const float *w = win + ((int)fy - wy0) * WIN_W + ((int)fx - wx0);
        const float m0 = w[-WIN_W], m1 = w[-WIN_W + 1];
        const float c_1 = w[-1], c0 = w[0], c1 = w[1], c2 = w[2];
        const float d_1 = w[WIN_W - 1], d0 = w[WIN_W], d1 = w[WIN_W + 1], d2 = w[WIN_W + 2];
        const float n0 = w[2 * WIN_W], n1 = w[2 * WIN_W + 1];
        const float Hm = a * m0 + xx * m1;
        const float H0 = a * c0 + xx * c1, H0p = a1 * c1 + xx1 * c2, H0m = a * c_1 + xx * c0;
        const float H1 = a * d0 + xx * d1, H1p = a1 * d1 + xx1 * d2, H1m = a * d_1 + xx * d0;
        const float H2 = a * n0 + xx * n1;
        const float v0 = b * H0 + yy * H1;
        const float vx1 = b * H0p + yy * H1p, vx2 = b * H0m + yy * H1m;
        const float vy1 = b1 * H1 + yy1 * H2, vy2 = b * Hm + yy * H0;
        const float T = rec[p].w;
        const float e = (v0 + db) - gain * T;
        rec[p] = make_float4(0.5f * (vx1 - vx2), 0.5f * (vy1 - vy2), e, T);
      }
    }
    bad = __any_sync(0xffffffffu, bad);
  }
  if (!use_window || bad) {  // per-sample path straight from the level (L1/L2), any coordinates
#pragma unroll 1
    for (int k = 0; k < PASSES; ++k) {
      const int p = lane + 32 * k;
      if (p < NP) {
        const float xf = (float)(p % P - HALF), yf = (float)(p / P - HALF);
        float wx = xf, wy = yf;
        if (mode.affine) { wx = a00 * xf + a01 * yf; wy = a10 * xf + a11 * yf; }
        const float sx = bx + wx, sy = by + wy;
        const float T = rec[p].w;
        const float e = (pagk_sample(I2, cols, rows, sx, sy) + db) - gain * T;
        const float gx = pagk_sample(I2, cols, rows, sx + 1.0f, sy) - pagk_sample(I2, cols, rows, sx - 1.0f, sy);
        const float gy = pagk_sample(I2, cols, rows, sx, sy + 1.0f) - pagk_sample(I2, cols, rows, sx, sy - 1.0f);
        rec[p] = make_float4(0.5f * gx, 0.5f * gy, e, T);
      }
    }
  }
}

// per-lane (= per-slot) state of the B warp that is not needed by phase A
struct LaneState {
  float pt1x, pt1y;      // mvKeysRefUn[i].pt
  float cost, lastCost;
  int iter;              // iteration index inside the level
  int n_iter;            // passes over all levels (output)
  bool succ;
};

}  // namespace

__global__ void __launch_bounds__(THREADS, 2)
pagk_lk_slots_kernel(const unsigned char *__restrict__ images, PagkGeom g, const PagkPairConst *__restrict__ pcs,
                     const float2 *__restrict__ keys_un, PagkOutPtrs out, PagkMode mode, int max_keys, int n_max,
                     int n_pairs, int *__restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlotShared &S = *reinterpret_cast<SlotShared *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b_warp = blockIdx.x % WARPS;  // spread the FP64-heavy warp of co-resident CTAs over the SM sub-partitions
  const int total_work = n_pairs * n_max;
  const int top = mode.levels - 1;

  LaneState L;
  L.pt1x = L.pt1y = 0.f; L.cost = L.lastCost = 0.f; L.iter = 0; L.n_iter = 0; L.succ = true;
  bool exhausted = false;
  if (warp == b_warp) { S.feat[lane] = -1; S.needs_setup[lane] = 0; S.win_valid[lane] = 0; }

  // fetch features for every empty slot of this CTA (B warp only, lane = slot)
  auto refill = [&]() {
    while (true) {
      const bool want = (S.feat[lane] < 0) && !exhausted;
      const unsigned m = __ballot_sync(0xffffffffu, want);
      if (m == 0u) break;
      int base = 0;
      if (lane == 0) base = atomicAdd(work_counter, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (want) {
        const int wi = base + __popc(m & ((1u << lane) - 1u));
        if (wi >= total_work) {
          exhausted = true;
        } else {
          const int pair = wi / n_max, i = wi % n_max;
          if (i < pcs[pair].n_keys) {
            const size_t o = (size_t)pair * max_keys + i;
            const float2 pt1 = keys_un[o];
            const float2 pt2 = mode.gyro_init ? out.pt_predict_un[o] : pt1;
            if (!out.gyro_status[o]) {  // skipped by the reference (src/patch_match.cpp:173): defaults only
              out.pm_un[o] = pt2; out.pm_status[o] = 0; out.pix_err[o] = 0.0; out.ncc[o] = 0.f; out.iters[o] = 0;
            } else {
              const float4 A = out.affine[o];
              const float scale = 1.0f / (float)(1 << top);
              const float ptx = pt1.x * scale, pty = pt1.y * scale;
              S.feat[lane] = (int)o; S.pair[lane] = pair; S.level[lane] = top; S.needs_setup[lane] = 1;
              S.ptx[lane] = ptx; S.pty[lane] = pty;
              S.dx[lane] = pt2.x * scale - ptx; S.dy[lane] = pt2.y * scale - pty;
              S.dg[lane] = 0.f; S.db[lane] = 0.f;
              S.a00[lane] = A.x; S.a01[lane] = A.y; S.a10[lane] = A.z; S.a11[lane] = A.w;
              const float hf = (float)HALF;
              float x0 = -hf, x1 = hf, y0 = -hf, y1 = hf;
              if (mode.affine) {  // the warp offsets at the four patch corners, exactly as phase A computes them
                const float c0x = A.x * -hf + A.y * -hf, c1x = A.x * hf + A.y * -hf, c2x = A.x * -hf + A.y * hf, c3x = A.x * hf + A.y * hf;
                const float c0y = A.z * -hf + A.w * -hf, c1y = A.z * hf + A.w * -hf, c2y = A.z * -hf + A.w * hf, c3y = A.z * hf + A.w * hf;
                x0 = fminf(fminf(c0x, c1x), fminf(c2x, c3x)); x1 = fmaxf(fmaxf(c0x, c1x), fmaxf(c2x, c3x));
                y0 = fminf(fminf(c0y, c1y), fminf(c2y, c3y)); y1 = fmaxf(fmaxf(c0y, c1y), fmaxf(c2y, c3y));
              }
              S.wxmin[lane] = x0; S.wxmax[lane] = x1; S.wymin[lane] = y0; S.wymax[lane] = y1;
              L.pt1x = pt1.x; L.pt1y = pt1.y; L.cost = 0.f; L.lastCost = 0.f; L.iter = 0; L.n_iter = 0; L.succ = true;
            }
          }
        }
      }
    }
  };

  if (warp == b_warp) {
    refill();
    const unsigned act = __ballot_sync(0xffffffffu, S.feat[lane] >= 0);
    if (lane == 0) S.any_active = (act != 0u);
  }
  __syncthreads();

  while (S.any_active) {
    // ------------------------------ phase A: all warps, SLOTS_PER_WARP slots each ------------------------------
#pragma unroll 1
    for (int q = 0; q < SLOTS_PER_WARP; ++q) {
      const int s = warp + q * WARPS;
      if (S.feat[s] >= 0) phase_a_slot(S, s, lane, images, g, mode);
    }
    __syncthreads();
    // ------------------------------ phase B: one warp, lane = slot ------------------------------
    if (warp == b_warp) {
      const int s = lane;
      const bool active = S.feat[s] >= 0;
      double h00 = 0, h10 = 0, h11 = 0, h20 = 0, h21 = 0, h22 = 0, h30 = 0, h31 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
      float cost = 0.f;
      const double c = (double)S.cval[s];
      const float4 *rec = S.rec[s];
#pragma unroll 2
      for (int p = 0; p < NP; ++p) {
        const float4 r = rec[p];
        const double ix = (double)r.x, iy = (double)r.y, me = -(double)r.z;
        h00 = fma(ix, ix, h00); h10 = fma(iy, ix, h10); h11 = fma(iy, iy, h11);
        h20 = fma(c, ix, h20); h21 = fma(c, iy, h21); h22 = fma(c, c, h22);
        h30 = h30 + ix; h31 = h31 + iy;
        b0 = fma(ix, me, b0); b1 = fma(iy, me, b1); b2 = fma(c, me, b2); b3 = b3 + me;
        cost = cost + r.z * r.z;
      }
      if (active) {
        double h32 = c * (double)NP, h33 = (double)NP;  // sum of c and of 1 over the patch: exact in double
        float dx = S.dx[s], dy = S.dy[s], dg = S.dg[s], db = S.db[s];
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
        ++L.n_iter;
        bool level_done = false;
        if (isnan(u0)) {
          L.succ = false; level_done = true;
        } else if (L.iter > 0 && cost > L.lastCost) {
          level_done = true;
        } else {
          dx = (float)((double)dx + u0);
          dy = (float)((double)dy + u1);
          if (mode.illum) { dg = (float)((double)dg + u2); db = (float)((double)db + u3); }
          L.lastCost = cost;
          L.succ = true;
          ++L.iter;
          const double nrm = sqrt((u0 * u0 + u2 * u2) + (u1 * u1 + u3 * u3));
          if (nrm < 1e-2 || L.iter >= mode.iterations) level_done = true;
        }
        if (!level_done) {
          S.dx[s] = dx; S.dy[s] = dy; S.dg[s] = dg; S.db[s] = db;
        } else {
          const float p2x = S.ptx[s] + dx, p2y = S.pty[s] + dy;  // mvPtPyr2Un[i] = pt + (dx, dy)
          const int level = S.level[s];
          if (level == 0) {
            const size_t o = (size_t)S.feat[s];
            out.pm_un[o] = make_float2(p2x, p2y);
            out.pm_status[o] = L.succ ? 1 : 0;
            out.pix_err[o] = sqrt((double)L.lastCost * mode.win_size_inv);
            out.ncc[o] = 1.0f;
            out.iters[o] = L.n_iter;
            S.feat[s] = -1;
          } else {
            const int nl = level - 1;
            const float scale = 1.0f / (float)(1 << nl);
            const float ptx = L.pt1x * scale, pty = L.pt1y * scale;
            S.level[s] = nl; S.needs_setup[s] = 1;
            S.ptx[s] = ptx; S.pty[s] = pty;
            S.dx[s] = p2x * 2.0f - ptx; S.dy[s] = p2y * 2.0f - pty;
            S.dg[s] = 0.f; S.db[s] = 0.f;
            L.iter = 0; L.lastCost = 0.f; L.succ = true;
          }
        }
      }
      refill();
      const unsigned act = __ballot_sync(0xffffffffu, S.feat[lane] >= 0);
      if (lane == 0) S.any_active = (act != 0u);
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------
size_t pagk_lk_slots_smem() { return sizeof(SlotShared); }

bool pagk_lk_slots_supported(const PagkMode &mode) {
  return mode.half == HALF && mode.iterations >= 1 && !mode.calc_ncc;
}

int pagk_launch_lk_slots(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                         const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs,
                         int *work_counter, int n_sms, cudaStream_t st, long long *launches) {
  if (n_max <= 0 || n_pairs <= 0) return 0;
  const size_t smem = sizeof(SlotShared);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(pagk_lk_slots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(int), st);
  if (e != cudaSuccess) return (int)e;
  const long long total = (long long)n_max * n_pairs;
  long long ctas = 2LL * n_sms;  // persistent: two CTAs per SM
  const long long needed = (total + SLOTS - 1) / SLOTS;
  if (ctas > needed) ctas = needed;
  pagk_lk_slots_kernel<<<(unsigned)ctas, THREADS, smem, st>>>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs,
                                                              work_counter);
  ++*launches;
  return (int)cudaGetLastError();
}
