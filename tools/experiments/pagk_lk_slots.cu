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
// One persistent CTA per SM owns SLOTS = 64 features ("slots") in two groups of 32 and runs a
// two-stage software pipeline: in every stage the warps do phase A on one group while phase B runs
// on the other group, then the groups swap (one __syncthreads per stage).
//   phase A  a warp takes one slot at a time from a shared cursor; its 32 lanes are the pixels.
//            The current-image window of a slot is staged once per level in shared memory as
//            float (no u8 conversions or 64-bit addressing in the loop); the template value T of
//            each pixel is kept beside the pixel's (Ix, Iy, -e) record, so one float4 per pixel
//            is the whole A->B hand-over.
//   phase B  lane = slot of the group, full lane occupancy (the float->double conversions run on
//            the 16-lane conversion pipe, so they must not be replicated).  The solver warp walks the
//            slot's 121 records in order with six H chains, the producer warp with the four b chains,
//            the two plain sums and the float cost chain (LDS.128 per step, conflict-free slot stride,
//            register-prefetched four pixels ahead); the solver warp then solves the
//            4x4 system of every slot, applies the update and advances the slot's state machine
//            (next iteration, next level, or the next feature).  Warp 1 meanwhile tops up a queue
//            of prefetched feature descriptors from the global work counter, so fetching never
//            sits on the solver's critical path.  Both join phase A when they are done.
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
constexpr int SLOTS = 64;
constexpr int HSLOTS = SLOTS / 2;         // one group = 32 slots = the lanes of a B warp
constexpr int WARPS = 16;
constexpr int THREADS = WARPS * 32;
constexpr int WIN_W = 21;                 // window width == row stride (21: rows of a pass land in distinct banks)
constexpr int WIN_H = 17;
constexpr int REC_STRIDE = NP;            // float4 records per slot; 484 words = 4 mod 32 -> LDS.128 conflict-free over slots
constexpr int QCAP = 64;
#ifndef PAGK_PARK
#define PAGK_PARK 2                       // phase-A warps parked per B sub-partition while its B warp works (0..3)
#endif                  // prefetched feature descriptors

struct SlotShared {
  float4 rec[SLOTS][REC_STRIDE];          // (Ix, Iy, -e, T) per pixel
  float win[SLOTS][WIN_W * WIN_H];
  double exch[HSLOTS][4];                 // b0..b3 of the producer warp, handed to the solver
  float exch_cost[HSLOTS];
  // slot state, written by the solver lane (and the window origin / cval by the A warp)
  int feat[SLOTS];                        // pair * max_keys + i, or -1
  int pair[SLOTS];
  int level[SLOTS];
  int needs_setup[SLOTS];
  int rec_valid[SLOTS];
  int win_x0[SLOTS], win_y0[SLOTS], win_valid[SLOTS];
  float ptx[SLOTS], pty[SLOTS], dx[SLOTS], dy[SLOTS], dg[SLOTS], db[SLOTS], cval[SLOTS];
  float a00[SLOTS], a01[SLOTS], a10[SLOTS], a11[SLOTS];
  float wxmin[SLOTS], wxmax[SLOTS], wymin[SLOTS], wymax[SLOTS];
  // solver-only state (kept here because a B lane serves one slot of each group)
  float pt1x[SLOTS], pt1y[SLOTS], lastCost[SLOTS];
  int iter[SLOTS], n_iter[SLOTS], succ[SLOTS];
  // queue of prefetched trackable features (ring; producer = warp W_PROD, consumer = warp W_SOLVE)
  int q_feat[QCAP], q_pair[QCAP];
  float q_pt1x[QCAP], q_pt1y[QCAP], q_pt2x[QCAP], q_pt2y[QCAP], q_a00[QCAP], q_a01[QCAP], q_a10[QCAP], q_a11[QCAP];
  int q_head, q_tail;                     // monotonically increasing; index = value % QCAP
  int lv_cols[PAGK_MAX_LEVELS], lv_rows[PAGK_MAX_LEVELS];
  unsigned int lv_off[PAGK_MAX_LEVELS];
  int a_cursor[2];                        // double-buffered by stage parity
  int cons_more[2], prod_more[2];         // "keep going" votes of the consumer and the producer
};

// floor(x) as float and int for 0 <= x < 2^22 on the FP32/ALU pipes: every FRND / F2I / I2F / F2F costs about
// 8 issue cycles of the 16-lane conversion pipe per warp (tools/ubench.cu), and the B warps need that
// pipe for their float -> double conversions.  x + 2^23 rounds x to the nearest integer (ulp is 1 there),
// subtracting 2^23 back is exact, and one compare fixes the round-up case.  Exact for every input in range.
__device__ __forceinline__ float floor_nonneg(float x, int &i) {
  const float t = x + 8388608.0f;
  float r = t - 8388608.0f;
  i = __float_as_int(t) - 0x4B000000;
  if (r > x) { r -= 1.0f; i -= 1; }
  return r;
}
__device__ __forceinline__ float floor_nonneg(float x) {
  const float t = x + 8388608.0f;
  float r = t - 8388608.0f;
  if (r > x) r -= 1.0f;
  return r;
}
// exact u8 -> float without the conversion pipe: 0x4B000000 | b is the float 2^23 + b
__device__ __forceinline__ float u8_to_float(unsigned char b) {
  return __uint_as_float(0x4B000000u | (unsigned int)b) - 8388608.0f;
}

// Stage the WIN_W x WIN_H window with origin (x0, y0) of a level into shared memory as float; lane j < WIN_W
// owns column j.  Coordinates are clamped into [0, cols] x [0, rows] (the level plus its wrap column and
// guard row), so every load is in bounds and unconditional: the 17 loads of a lane are issued back to back.
// Elements whose true position lies outside that range hold an arbitrary in-bounds pixel; callers never
// read them.
__device__ __forceinline__ void stage_window(float *__restrict__ win, const unsigned char *__restrict__ img, int cols,
                                             int rows, int x0, int y0, int lane) {
  if (lane < WIN_W) {
    const int gx = min(max(x0 + lane, 0), cols);
    unsigned char v[WIN_H];
#pragma unroll
    for (int i = 0; i < WIN_H; ++i) {
      const int gy = min(max(y0 + i, 0), rows);
      v[i] = __ldg(img + gy * cols + gx);
    }
#pragma unroll
    for (int i = 0; i < WIN_H; ++i) win[i * WIN_W + lane] = u8_to_float(v[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// phase A for one slot (whole warp).  xf/yf: this lane's pixel offsets of the four passes.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void phase_a_slot(SlotShared &S, int s, int lane, const float (&xf)[PASSES],
                                             const float (&yf)[PASSES], const unsigned char *__restrict__ images,
                                             unsigned long long slot_bytes, const PagkMode &mode) {
  const int level = S.level[s], pair = S.pair[s];
  const int cols = S.lv_cols[level], rows = S.lv_rows[level];
  const unsigned char *I1 = images + (size_t)(pair * 2) * slot_bytes + S.lv_off[level];
  const unsigned char *I2 = I1 + slot_bytes;
  const float ptx = S.ptx[s], pty = S.pty[s];
  float4 *rec = S.rec[s];
  float *win = S.win[s];
  const float hf = (float)HALF;

  if (S.needs_setup[s]) {  // new level: template values T = I1(pt + (x, y)) and c = -I1(pt)
    const float txlo = ptx + (-hf), txhi = ptx + hf, tylo = pty + (-hf), tyhi = pty + hf;
    float c;
    if (txlo >= 0.0f && txhi < (float)cols && tylo >= 0.0f && tyhi < (float)rows) {
      // no clamp fires anywhere in the template: sample from a staged window of the reference level
      const int x0 = (int)txlo, y0 = (int)tylo;
      __syncwarp();  // taps reach floor(hi) + 1 <= x0 + 12: inside the 21 x 17 window
      stage_window(win, I1, cols, rows, x0, y0, lane);
      __syncwarp();
#pragma unroll
      for (int k = 0; k < PASSES; ++k) {
        const int p = lane + 32 * k;
        if (p < NP) {
          const float cx = ptx + xf[k], cy = pty + yf[k];
          int ix, iy;
          const float fx = floor_nonneg(cx, ix), fy = floor_nonneg(cy, iy);
          const float xx = cx - fx, yy = cy - fy, a = 1.0f - xx, b = 1.0f - yy;
          const float *q = win + (iy - y0) * WIN_W + (ix - x0);
          rec[p].w = b * (a * q[0] + xx * q[1]) + yy * (a * q[WIN_W] + xx * q[WIN_W + 1]);
        }
      }
      {
        int ix, iy;
        const float fx = floor_nonneg(ptx, ix), fy = floor_nonneg(pty, iy);
        const float xx = ptx - fx, yy = pty - fy, a = 1.0f - xx, b = 1.0f - yy;
        const float *q = win + (iy - y0) * WIN_W + (ix - x0);
        c = -(b * (a * q[0] + xx * q[1]) + yy * (a * q[WIN_W] + xx * q[WIN_W + 1]));
      }
    } else {
#pragma unroll 1
      for (int k = 0; k < PASSES; ++k) {
        const int p = lane + 32 * k;
        if (p < NP) rec[p].w = pagk_sample_call(I1, cols, rows, ptx + xf[k], pty + yf[k]);
      }
      c = -pagk_sample_call(I1, cols, rows, ptx, pty);
    }
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
    const int ixlo = (int)x2min, ixhi = (int)x1max + 1;  // non-negative: truncation == floor
    const int iylo = (int)y2min, iyhi = (int)y1max + 1;
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
        stage_window(win, I2, cols, rows, wx0, wy0, lane);
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
        float wx = xf[k], wy = yf[k];
        if (mode.affine) { wx = a00 * xf[k] + a01 * yf[k]; wy = a10 * xf[k] + a11 * yf[k]; }
        const float sx = bx + wx, sy = by + wy;
        int ix, iy;
        const float fx = floor_nonneg(sx, ix), fy = floor_nonneg(sy, iy);
        const float xx = sx - fx, yy = sy - fy;
        const float a = 1.0f - xx, b = 1.0f - yy;
        const float X1 = sx + 1.0f, Y1 = sy + 1.0f;
        const float fX1 = floor_nonneg(X1), fY1 = floor_nonneg(Y1);
        const float xx1 = X1 - fX1, yy1 = Y1 - fY1;
        const float a1 = 1.0f - xx1, b1 = 1.0f - yy1;
        bad |= (fX1 != fx + 1.0f) | (fY1 != fy + 1.0f);
        const float *w = win + (iy - wy0) * WIN_W + (ix - wx0);
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
        rec[p] = make_float4(0.5f * (vx1 - vx2), 0.5f * (vy1 - vy2), -e, T);
      }
    }
    bad = __any_sync(0xffffffffu, bad);
  }
  if (!use_window || bad) {  // per-sample path straight from the level (L1/L2), any coordinates
#pragma unroll 1
    for (int k = 0; k < PASSES; ++k) {
      const int p = lane + 32 * k;
      if (p < NP) {
        float wx = xf[k], wy = yf[k];
        if (mode.affine) { wx = a00 * xf[k] + a01 * yf[k]; wy = a10 * xf[k] + a11 * yf[k]; }
        const float sx = bx + wx, sy = by + wy;
        const float T = rec[p].w;
        const float e = (pagk_sample_call(I2, cols, rows, sx, sy) + db) - gain * T;
        const float gx = pagk_sample_call(I2, cols, rows, sx + 1.0f, sy) - pagk_sample_call(I2, cols, rows, sx - 1.0f, sy);
        const float gy = pagk_sample_call(I2, cols, rows, sx, sy + 1.0f) - pagk_sample_call(I2, cols, rows, sx, sy - 1.0f);
        rec[p] = make_float4(0.5f * gx, 0.5f * gy, -e, T);
      }
    }
  }
  if (lane == 0) S.rec_valid[s] = 1;
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_barrier_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace

__global__ void __launch_bounds__(THREADS, 1)
pagk_lk_slots_kernel(const unsigned char *__restrict__ images, PagkGeom g, const PagkPairConst *__restrict__ pcs,
                     const float2 *__restrict__ keys_un, PagkOutPtrs out, PagkMode mode, int max_keys, int n_max,
                     int n_pairs, int *__restrict__ work_counter, long long *__restrict__ dbg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SlotShared &S = *reinterpret_cast<SlotShared *>(smem_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Roles: the SM's warp arbiter favours the highest warp id of a sub-partition (B300_MICROARCH.md), so the
  // two latency-critical B warps take the two highest ids (they also sit on different sub-partitions).
  constexpr int W_SOLVE = WARPS - 1, W_PROD = WARPS - 2;
  const int total_work = n_pairs * n_max;
  const int top = mode.levels - 1;
  const unsigned long long slot_bytes = g.slot_bytes;

  // optional timeline of CTA 0 (PAGK_LK_TIMELINE): dbg[stage * 16 + slot] = clock64() at named points
#define PAGK_TL(stage, slot) do { if (dbg && blockIdx.x == 0 && lane == 0 && (stage) < 400) dbg[(stage) * 16 + (slot)] = clock64(); } while (0)
  const long long t_kernel0 = clock64();
  float xf[PASSES], yf[PASSES];
#pragma unroll
  for (int k = 0; k < PASSES; ++k) {
    const int p = lane + 32 * k;
    xf[k] = (float)(p % P - HALF);
    yf[k] = (float)(p / P - HALF);
  }

  bool exhausted = false;  // producer warp only
  if (warp == W_SOLVE) {
    for (int s = lane; s < SLOTS; s += 32) { S.feat[s] = -1; S.needs_setup[s] = 0; S.win_valid[s] = 0; S.rec_valid[s] = 0; }
    if (lane < PAGK_MAX_LEVELS) { S.lv_cols[lane] = g.lv[lane].cols; S.lv_rows[lane] = g.lv[lane].rows; S.lv_off[lane] = g.lv[lane].offset; }
    if (lane == 0) { S.a_cursor[0] = 0; S.a_cursor[1] = 0; S.q_head = 0; S.q_tail = 0; }
  }
  __syncthreads();

  // producer (warp W_PROD): top the queue up to QCAP entries counted from `head`; features the reference
  // skips (gyro status 0, src/patch_match.cpp:173) get their default outputs here and are not queued
  auto produce = [&](int head, int tail) -> int {
    int t = tail;
    while (!exhausted && (t - head) + 32 <= QCAP) {
      int base = 0;
      if (lane == 0) base = atomicAdd(work_counter, 32);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base >= total_work) { exhausted = true; break; }
      const int wi = base + lane;
      bool ok = false;
      int pair = 0;
      size_t o = 0;
      float2 pt1 = make_float2(0.f, 0.f), pt2 = pt1;
      float4 A = make_float4(1.f, 0.f, 0.f, 1.f);
      if (wi < total_work) {
        pair = wi / n_max;
        const int i = wi % n_max;
        if (i < pcs[pair].n_keys) {
          o = (size_t)pair * max_keys + i;
          pt1 = keys_un[o];
          pt2 = mode.gyro_init ? out.pt_predict_un[o] : pt1;
          if (!out.gyro_status[o]) {
            out.pm_un[o] = pt2; out.pm_status[o] = 0; out.pix_err[o] = 0.0; out.ncc[o] = 0.f; out.iters[o] = 0;
          } else {
            A = out.affine[o];
            ok = true;
          }
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int q = (t + __popc(m & ((1u << lane) - 1u))) % QCAP;
        S.q_feat[q] = (int)o; S.q_pair[q] = pair;
        S.q_pt1x[q] = pt1.x; S.q_pt1y[q] = pt1.y; S.q_pt2x[q] = pt2.x; S.q_pt2y[q] = pt2.y;
        S.q_a00[q] = A.x; S.q_a01[q] = A.y; S.q_a10[q] = A.z; S.q_a11[q] = A.w;
      }
      t += __popc(m);
    }
    return t;
  };

  // consumer (warp W_SOLVE): give every empty slot of group `grp` a queued feature; returns the new head
  auto consume = [&](int grp, int head, int tail) -> int {
    const int s = grp * HSLOTS + lane;
    const bool want = S.feat[s] < 0;
    const unsigned m = __ballot_sync(0xffffffffu, want);
    const int rank = __popc(m & ((1u << lane) - 1u));
    if (want && head + rank < tail) {
      const int q = (head + rank) % QCAP;
      const float pt1x = S.q_pt1x[q], pt1y = S.q_pt1y[q], pt2x = S.q_pt2x[q], pt2y = S.q_pt2y[q];
      const float a00 = S.q_a00[q], a01 = S.q_a01[q], a10 = S.q_a10[q], a11 = S.q_a11[q];
      const float scale = 1.0f / (float)(1 << top);
      const float ptx = pt1x * scale, pty = pt1y * scale;
      S.feat[s] = S.q_feat[q]; S.pair[s] = S.q_pair[q]; S.level[s] = top; S.needs_setup[s] = 1; S.rec_valid[s] = 0;
      S.ptx[s] = ptx; S.pty[s] = pty;
      S.dx[s] = pt2x * scale - ptx; S.dy[s] = pt2y * scale - pty;
      S.dg[s] = 0.f; S.db[s] = 0.f;
      S.a00[s] = a00; S.a01[s] = a01; S.a10[s] = a10; S.a11[s] = a11;
      const float hf = (float)HALF;
      float x0 = -hf, x1 = hf, y0 = -hf, y1 = hf;
      if (mode.affine) {  // the warp offsets at the four patch corners, exactly as phase A computes them
        const float c0x = a00 * -hf + a01 * -hf, c1x = a00 * hf + a01 * -hf, c2x = a00 * -hf + a01 * hf, c3x = a00 * hf + a01 * hf;
        const float c0y = a10 * -hf + a11 * -hf, c1y = a10 * hf + a11 * -hf, c2y = a10 * -hf + a11 * hf, c3y = a10 * hf + a11 * hf;
        x0 = fminf(fminf(c0x, c1x), fminf(c2x, c3x)); x1 = fmaxf(fmaxf(c0x, c1x), fmaxf(c2x, c3x));
        y0 = fminf(fminf(c0y, c1y), fminf(c2y, c3y)); y1 = fmaxf(fmaxf(c0y, c1y), fmaxf(c2y, c3y));
      }
      S.wxmin[s] = x0; S.wxmax[s] = x1; S.wymin[s] = y0; S.wymax[s] = y1;
      S.pt1x[s] = pt1x; S.pt1y[s] = pt1y; S.lastCost[s] = 0.f; S.iter[s] = 0; S.n_iter[s] = 0; S.succ[s] = 1;
    }
    const int taken = min(__popc(m), tail - head);
    return head + taken;
  };

  // No separate prologue: the first stages find empty slots and an empty queue; the producer fills the queue
  // in stage 0, the solver hands features to group 0 in stage 1 and to group 1 in stage 2.
  if (threadIdx.x == 0) { S.cons_more[0] = 1; S.prod_more[0] = 1; }
  __syncthreads();

  // stage t: phase A on group (t & 1), phase B on the other group.  Cursors and votes are
  // double-buffered by stage parity so that one barrier per stage is enough.
  for (int t = 0; S.cons_more[t & 1] | S.prod_more[t & 1]; ++t) {
    const int hA = t & 1, hB = hA ^ 1;
    if (threadIdx.x == 0) S.a_cursor[hB] = 0;  // the cursor of the next stage; nobody reads it during this one
    if (warp == W_SOLVE) PAGK_TL(t, 0);
    // Sub-partition reservation.  The B warps run long dependent FP64 chains; sharing a scheduler with three
    // phase-A warps that always have independent FP32 work ready slows them about 2.5x (measured), and they
    // are the critical path of the stage.  So the phase-A warps of the producer's sub-partition (warp % 4 == 2)
    // and of the solver's (warp % 4 == 3) park on a named barrier -- costing no issue slots -- until that B
    // warp has finished its part of the stage; phase A meanwhile runs on the other two sub-partitions.
    if (warp < W_PROD && (warp & 3) >= 2 && (warp >> 2) < PAGK_PARK) named_barrier_sync(warp & 3, 32 * (PAGK_PARK + 1));
    if (warp >= W_PROD) {
      // ------------------------------ phase B: lane = slot of group hB ------------------------------
      const int s = hB * HSLOTS + lane;
      const int q_head = S.q_head, q_tail = S.q_tail;  // stable since the barrier
      const bool active = S.feat[s] >= 0 && S.rec_valid[s];
      const bool any = __any_sync(0xffffffffu, active);
      const double c = (double)S.cval[s];
      const float4 *rec = S.rec[s];
      if (warp == W_PROD) {
        // ---- b chains (b += -J * e; the record holds -e) and the float cost chain, then prefetch ----
        double b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        float cost = 0.f;
        if (any) {
          // software pipeline: the records of the next four pixels are loaded while the current four are
          // converted and accumulated (left to itself the compiler puts each LDS right before its use and
          // every step then pays the full load + conversion latency, tools/ubench_acc.cu).  Reading up to
          // three records past the slot's 121 stays inside the shared struct; those values are not used.
          float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
#pragma unroll 1
          for (int p = 0; p < NP - 1; p += 4) {
            const float4 n0 = rec[p + 4], n1 = rec[p + 5], n2 = rec[p + 6], n3 = rec[p + 7];
            const double x0 = (double)r0.x, y0 = (double)r0.y, m0 = (double)r0.z;
            const double x1 = (double)r1.x, y1 = (double)r1.y, m1 = (double)r1.z;
            const double x2 = (double)r2.x, y2 = (double)r2.y, m2 = (double)r2.z;
            const double x3 = (double)r3.x, y3 = (double)r3.y, m3 = (double)r3.z;
            b0 = fma(x0, m0, b0); b1 = fma(y0, m0, b1); b2 = fma(c, m0, b2); b3 = b3 + m0;
            cost = cost + r0.z * r0.z;
            b0 = fma(x1, m1, b0); b1 = fma(y1, m1, b1); b2 = fma(c, m1, b2); b3 = b3 + m1;
            cost = cost + r1.z * r1.z;
            b0 = fma(x2, m2, b0); b1 = fma(y2, m2, b1); b2 = fma(c, m2, b2); b3 = b3 + m2;
            cost = cost + r2.z * r2.z;
            b0 = fma(x3, m3, b0); b1 = fma(y3, m3, b1); b2 = fma(c, m3, b2); b3 = b3 + m3;
            cost = cost + r3.z * r3.z;
            r0 = n0; r1 = n1; r2 = n2; r3 = n3;
          }
          {  // pixel 120 (NP = 4 * 30 + 1)
            const double x0 = (double)r0.x, y0 = (double)r0.y, m0 = (double)r0.z;
            b0 = fma(x0, m0, b0); b1 = fma(y0, m0, b1); b2 = fma(c, m0, b2); b3 = b3 + m0;
            cost = cost + r0.z * r0.z;
          }
        }
        S.exch[lane][0] = b0; S.exch[lane][1] = b1; S.exch[lane][2] = b2; S.exch[lane][3] = b3;
        S.exch_cost[lane] = cost;
        PAGK_TL(t, 6);
        named_barrier_sync(1, 64);
        const int nt = produce(q_head, q_tail);
        if (PAGK_PARK > 0) named_barrier_arrive(2, 32 * (PAGK_PARK + 1));  // releases the parked warps of sub-partition 2
        PAGK_TL(t, 7);
        if (lane == 0) { S.q_tail = nt; S.prod_more[hB] = (nt > q_tail) || !exhausted; }
      } else {
        // ---- H chains, then the solve and the slot state machine ----
        double h00 = 0, h10 = 0, h11 = 0, h20 = 0, h21 = 0, h22 = 0, h30 = 0, h31 = 0;
        if (any) {  // software-pipelined like the producer's loop
          float4 r0 = rec[0], r1 = rec[1], r2 = rec[2], r3 = rec[3];
#pragma unroll 1
          for (int p = 0; p < NP - 1; p += 4) {
            const float4 n0 = rec[p + 4], n1 = rec[p + 5], n2 = rec[p + 6], n3 = rec[p + 7];
            const double x0 = (double)r0.x, y0 = (double)r0.y, x1 = (double)r1.x, y1 = (double)r1.y;
            const double x2 = (double)r2.x, y2 = (double)r2.y, x3 = (double)r3.x, y3 = (double)r3.y;
            h00 = fma(x0, x0, h00); h10 = fma(y0, x0, h10); h11 = fma(y0, y0, h11); h20 = fma(c, x0, h20); h21 = fma(c, y0, h21); h22 = fma(c, c, h22); h30 = h30 + x0; h31 = h31 + y0;
            h00 = fma(x1, x1, h00); h10 = fma(y1, x1, h10); h11 = fma(y1, y1, h11); h20 = fma(c, x1, h20); h21 = fma(c, y1, h21); h22 = fma(c, c, h22); h30 = h30 + x1; h31 = h31 + y1;
            h00 = fma(x2, x2, h00); h10 = fma(y2, x2, h10); h11 = fma(y2, y2, h11); h20 = fma(c, x2, h20); h21 = fma(c, y2, h21); h22 = fma(c, c, h22); h30 = h30 + x2; h31 = h31 + y2;
            h00 = fma(x3, x3, h00); h10 = fma(y3, x3, h10); h11 = fma(y3, y3, h11); h20 = fma(c, x3, h20); h21 = fma(c, y3, h21); h22 = fma(c, c, h22); h30 = h30 + x3; h31 = h31 + y3;
            r0 = n0; r1 = n1; r2 = n2; r3 = n3;
          }
          {
            const double x0 = (double)r0.x, y0 = (double)r0.y;
            h00 = fma(x0, x0, h00); h10 = fma(y0, x0, h10); h11 = fma(y0, y0, h11); h20 = fma(c, x0, h20); h21 = fma(c, y0, h21); h22 = fma(c, c, h22); h30 = h30 + x0; h31 = h31 + y0;
          }
        }
        // The factorisation needs only H, which this warp owns: do it before the hand-over so that it overlaps
        // the producer's (conversion-bound, slower) loop; only the substitutions wait for b.
        double h32 = c * (double)NP, h33 = (double)NP;  // sum of c and of 1 over the patch: exact in double
        float dx = 0.f, dy = 0.f, dg = 0.f, db = 0.f, lastCost = 0.f;
        int iter = 0, succ = 1;
        double jx = 0.0, jy = 0.0, e_pen = 0.0;
        if (active) {
          dx = S.dx[s]; dy = S.dy[s]; dg = S.dg[s]; db = S.db[s];
          lastCost = S.lastCost[s]; iter = S.iter[s]; succ = S.succ[s];
          if (mode.regular) {  // reference src/patch_match.cpp:302-314
            const double d = (double)sqrtf(dx * dx + dy * dy);
            const float li = mode.lambda * mode.inv_log_max_dist;
            const double ad1 = (double)mode.alpha * d + 1.0;
            e_pen = (double)li * log(ad1);
            jx = ((double)(li * mode.alpha) / ad1) * ((double)dx / d);
            jy = ((double)(li * mode.alpha) / ad1) * ((double)dy / d);
            h00 += jx * jx; h10 += jy * jx; h11 += jy * jy;
            h20 += 0.0 * jx; h21 += 0.0 * jy; h30 += 0.0 * jx; h31 += 0.0 * jy;
          }
          pagk_llt_factor4(h00, h10, h11, h20, h21, h22, h30, h31, h32, h33);
        }
        PAGK_TL(t, 1);
        named_barrier_sync(1, 64);
        PAGK_TL(t, 2);
        if (active) {
          double b0 = S.exch[lane][0], b1 = S.exch[lane][1], b2 = S.exch[lane][2], b3 = S.exch[lane][3];
          float cost = S.exch_cost[lane];
          if (mode.regular) {
            b0 += jx * e_pen; b1 += jy * e_pen; b2 += 0.0 * e_pen; b3 += 0.0 * e_pen;
            cost = (float)((double)cost + e_pen * e_pen);
          }
          double u0, u1, u2, u3;
          PAGK_TL(t, 12);
          pagk_llt_subst4(h00, h10, h11, h20, h21, h22, h30, h31, h32, h33, b0, b1, b2, b3, u0, u1, u2, u3);
          PAGK_TL(t, 13);
          const int n_iter = S.n_iter[s] + 1;
          S.n_iter[s] = n_iter;
          bool level_done = false;
          if (isnan(u0)) {
            succ = 0; level_done = true;
          } else if (iter > 0 && cost > lastCost) {
            level_done = true;
          } else {
            dx = (float)((double)dx + u0);
            dy = (float)((double)dy + u1);
            if (mode.illum) { dg = (float)((double)dg + u2); db = (float)((double)db + u3); }
            lastCost = cost;
            succ = 1;
            ++iter;
            const double nrm = sqrt((u0 * u0 + u2 * u2) + (u1 * u1 + u3 * u3));
            if (nrm < 1e-2 || iter >= mode.iterations) level_done = true;
          }
          S.rec_valid[s] = 0;
          PAGK_TL(t, 14);
          if (!level_done) {
            S.dx[s] = dx; S.dy[s] = dy; S.dg[s] = dg; S.db[s] = db;
            S.lastCost[s] = lastCost; S.iter[s] = iter; S.succ[s] = succ;
          } else {
            const float p2x = S.ptx[s] + dx, p2y = S.pty[s] + dy;  // mvPtPyr2Un[i] = pt + (dx, dy)
            const int level = S.level[s];
            if (level == 0) {
              const size_t o = (size_t)S.feat[s];
              out.pm_un[o] = make_float2(p2x, p2y);
              out.pm_status[o] = succ ? 1 : 0;
              out.pix_err[o] = sqrt((double)lastCost * mode.win_size_inv);
              out.ncc[o] = 1.0f;
              out.iters[o] = n_iter;
              S.feat[s] = -1;
            } else {
              const int nl = level - 1;
              const float scale = 1.0f / (float)(1 << nl);
              const float ptx = S.pt1x[s] * scale, pty = S.pt1y[s] * scale;
              S.level[s] = nl; S.needs_setup[s] = 1;
              S.ptx[s] = ptx; S.pty[s] = pty;
              S.dx[s] = p2x * 2.0f - ptx; S.dy[s] = p2y * 2.0f - pty;
              S.dg[s] = 0.f; S.db[s] = 0.f;
              S.iter[s] = 0; S.lastCost[s] = 0.f; S.succ[s] = 1;
            }
          }
        }
        __syncwarp();
        PAGK_TL(t, 3);
        const int nh = consume(hB, q_head, q_tail);
        if (lane == 0) S.q_head = nh;
        const unsigned act = __ballot_sync(0xffffffffu, S.feat[lane] >= 0 || S.feat[HSLOTS + lane] >= 0);
        if (lane == 0) S.cons_more[hB] = (act != 0u) || (q_tail - nh > 0);  // read at the top of stage t + 1
        if (PAGK_PARK > 0) named_barrier_arrive(3, 32 * (PAGK_PARK + 1));  // releases the parked warps of sub-partition 3
        PAGK_TL(t, 4);
      }
    }
    // ------------------------------ phase A on the other group (every warp, when free) ------------------------------
    while (true) {
      int k = 0;
      if (lane == 0) k = atomicAdd(&S.a_cursor[hA], 1);
      k = __shfl_sync(0xffffffffu, k, 0);
      if (k >= HSLOTS) break;
      const int s = hA * HSLOTS + k;
      if (S.feat[s] >= 0) phase_a_slot(S, s, lane, xf, yf, images, slot_bytes, mode);
    }
    if (warp == W_SOLVE) PAGK_TL(t, 5);
    if (warp == W_PROD) PAGK_TL(t, 8);
    if (warp == 2) PAGK_TL(t, 9);
    if (warp == 0) PAGK_TL(t, 10);
    __syncthreads();
    if (warp == W_SOLVE) PAGK_TL(t, 11);
    if (dbg && threadIdx.x == 0) { dbg[6400 + blockIdx.x * 2] = clock64() - t_kernel0; dbg[6400 + blockIdx.x * 2 + 1] = t + 1; }
  }
}

// -------------------------------------------------------------------------------------------------
size_t pagk_lk_slots_smem() { return sizeof(SlotShared); }

bool pagk_lk_slots_supported(const PagkMode &mode) {
  return mode.half == HALF && mode.iterations >= 1;
}

int pagk_launch_lk_slots(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                         const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs,
                         int *work_counter, int n_sms, cudaStream_t st, long long *launches, long long *dbg) {
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
  long long ctas = n_sms;  // persistent: one CTA per SM
  const long long needed = (total + SLOTS - 1) / SLOTS;
  if (ctas > needed) ctas = needed;
  pagk_lk_slots_kernel<<<(unsigned)ctas, THREADS, smem, st>>>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs,
                                                              work_counter, dbg);
  ++*launches;
  return (int)cudaGetLastError();
}
