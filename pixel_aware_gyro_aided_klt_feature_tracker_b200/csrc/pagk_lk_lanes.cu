// pagk_lk_lanes.cu -- K3, the production patch-alignment kernels: a template pass (K3a, fed by TMA tile loads) and the
// alignment kernel proper (K3b, one LANE per feature, windows by asynchronous copies).
//
// Reference: PatchMatch::OpticalFlowMultiLevel + OpticalFlowConsideringIlluminationChange_onePixel,
// src/patch_match.cpp:79-142 and :167-367 (forward-additive Gauss-Newton on (dx, dy, dg, db)).
//
// Why this shape.  The 4x4 normal matrix of the reference is structurally singular, so the
// 14 sums of one Gauss-Newton pass must be taken in double, in the reference's pixel order: a serial chain
// of P*P steps per feature.  Independent features are the only parallelism that keeps that order, so a
// lane owns a feature ("slot") for a whole pyramid level -- all its iterations -- and walks the patch two pixels
// at a time: five bilinear samples per pixel of the current image (packed FP32, no FMA contraction), residual and
// gradient, then straight into the lane's eleven FP64 accumulators and the float cost.  Nothing is handed over between
// threads inside a level, there is no block barrier and no role: every warp of the CTA is an independent worker with
// 32 slots, and a lane refills itself from the global work counter when its level is finished.
//
// What a level needs before its first pass, and where it comes from:
//   T[p] = I1(pt + (x, y)), the P*P template values, c = -I1(pt) and h22 = sum of c*c.  They depend on the
//        reference keypoint and the level only -- not on the tracking result of the level above -- so K3a computes
//        them for every (feature, level) of the batch up front, pixel-parallel, into one record per item in global
//        memory: an item's (P+2) x (P+2) tap block arrives by ONE TMA tile load (cp.async.bulk.tensor -> UTMALDG.3D,
//        completion on the warp's mbarrier).  K3b streams T from the record during the pass (LDG.128, the next
//        vectors prefetched) and takes the record's 16-byte tail (last T value, c, h22) with one vector load.
//   the WIN_W x WIN_H u8 window of the current image around the lane's sample box: 4-byte asynchronous copies
//        (cp.async -> LDGSTS, each lane copies its own window) straight from the pyramid level, which is why a
//        level's rows are aligned (PagkLevelGeom::pitch).
// All copies of a round -- every window that starts a level or moved -- are in flight together and the warp waits for
// them once; no tap is staged through registers and no lane-serial setup code is left in the alignment kernel (it was
// 31 % of a warp's time and a third of the kernel's code in round 1).
//
// Shared memory holds the windows (u8), laid out ONE BANK PER LANE: word w of lane l's window sits at
// (w * 32 + l) * 4 inside the warp's block, so whatever the lanes' sample offsets are, a tap load is one wavefront.
// (Slot-contiguous windows cost 2.65 wavefronts per load at production offsets: the pass was bound by the
// shared-memory pipe as soon as its arithmetic got cheaper, tools/ubench_pass2.cu.)  Behind the windows: the buffers
// of the pixel-parallel pass.
//
// The pass arithmetic (argued bit-exact above the loop; tests/test_sass.py counts its fused operations):
//   * two pixels per step, every FP32 operation packed across the pair (FADD2 / FMUL2 / FFMA2, sm_100): 34.5 packed
//     operations per pixel instead of 69 scalar ones;
//   * no u8 -> float conversion at all: a tap enters the arithmetic as the raw byte read as a SUBNORMAL float
//     (b * 2^-149), and the sample coordinates carry a factor 2^100, so that every product weight * tap is the
//     reference's product times 2^-49 exactly; the factor leaves the sums after the loop;
//   * software-pipelined by hand: a pair's coordinates, weights and tap loads ahead of the previous pair's sums.
//
// Per round a warp does, with warp-uniform control flow:
//   refill   idle lanes take the next items from the global counter
//   issue    lanes that start a level: tail of the template record; lanes whose sample box left the window (or that
//            start a level): window copy; one wait for everything
//   pass     lane = slot: P*P pixels, samples + ordered sums            <- the hot loop
//   pixel-parallel pass
//            lanes whose samples may clamp at the border, whose box does not fit the window, that hit the one
//            rounding case the shared-weight sampling does not cover, and every live lane of a nearly empty warp:
//            lanes = pixels (each sample with PatchMatch::GetPixelValue semantics where needed), then lanes =
//            accumulators, two slots at a time
//   solve    lane = slot: 4x4 LLT in Eigen's operation order, update, exits, next level / next item
//
#include <cstdlib>
#include "pagk_device.cuh"
#include "pagk_kernels.h"

namespace {

// a warp with at most this many live lanes runs them through the pixel-parallel pass (lanes = pixels, then lanes =
// accumulators: about 5 k cycles per slot, two slots at a time) instead of a lockstep pass (20 k to 28 k cycles whatever
// the number of live lanes).  Measured on config B: 2 beats 5, 8 and 12.
#ifndef PAGK_LANES_SPARSE
#define PAGK_LANES_SPARSE 3
#endif
// warps per SM for 11 x 11 patches, chosen per launch: twelve (three per scheduler, 170 registers per thread: the low-ILP
// phases of one warp -- claim, window copies, solve -- are covered by the passes of two others) are 9 % faster per pair on
// batches of 128 pairs and more; on config B's 64 pairs a launch is only 4.6 items per lane deep for 56 k lanes, a
// fifth of them idle in the tail or waiting for a level hand-over, and eight warps (6.9 items per lane) win by 12 %.
#ifndef PAGK_LANES_WARPS5
#define PAGK_LANES_WARPS5 8
#endif
#ifndef PAGK_LANES_WARPS5_BIG
#define PAGK_LANES_WARPS5_BIG 12
#endif
// features per launch from which the larger number of warps is used
#ifndef PAGK_LANES_BIG_FEATURES
#define PAGK_LANES_BIG_FEATURES 110000
#endif
// template vectors (four pixels = two packed pairs) per trip of the pass loop: the independent chains a lone warp has
// in flight
#ifndef PAGK_LANES_TRIP
#define PAGK_LANES_TRIP 2
#endif
// warps per CTA: the warps of an SM as several small CTAs leave the SM one by one when the work runs out, and the
// CTAs of the next launch (another stream) move in
// measured and left off (config B, 8 warps): the first template vectors loaded before the window copies (0.708 against
// 0.698 ms), the pair offsets from shared memory instead of the constant bank (0.784 ms)
#ifndef PAGK_LANES_TPRE
#define PAGK_LANES_TPRE 0
#endif
#ifndef PAGK_LANES_LDSPAIRS
#define PAGK_LANES_LDSPAIRS 0
#endif
// the pass software-pipelined by hand: a pair's front half (coordinates, weights, taps) ahead of the previous pair's back half
#ifndef PAGK_LANES_PIPE
#define PAGK_LANES_PIPE 1
#endif
// register cap of the alignment kernel as CTAs per SM in its launch bounds (0: what the warps per SM imply)
#ifndef PAGK_LANES_REGCTAS
#define PAGK_LANES_REGCTAS 0
#endif
#ifndef PAGK_LANES_CTA_WARPS
#define PAGK_LANES_CTA_WARPS 4
#endif

template <int HALF, int WSM = (HALF <= 5 ? PAGK_LANES_WARPS5 : 7)>
struct LanesCfg {
  static constexpr int P = 2 * HALF + 1;
  static constexpr int NP = P * P;
  // window of the current level: WIN_W x WIN_H bytes, copied as 4-byte words from a 4-byte aligned origin: the
  // origin is a multiple of ALIGN columns and the usable width for any box is WIN_W - (ALIGN - 1).
  static constexpr int ALIGN = 4;
  static constexpr int WIN_W = HALF > 5 ? 32 : 24;
  static constexpr int WIN_H = P + 6;
  static constexpr int WPR = WIN_W / 4;  // words per window row
  static constexpr int WIN_WORDS = WPR * WIN_H;
  // one bank per lane: byte i of lane l's window sits at (i >> 2) * 128 + (i & 3) + 4 * l of the warp's block
  static constexpr int ROW_BYTES = WPR * 128;  // from a byte to the byte below it
  // a template record in global memory: T[0 .. NP-2] (T_BULK bytes, a multiple of 16: the pass reads it with 16-byte
  // loads), then T[NP-1], c = -T[NP/2], and h22 = the ordered double sum of c*c over the patch
  static constexpr int T_BULK = (NP * 4) & ~15;
  static constexpr int REC_BYTES = T_BULK + 16;
  static_assert(NP * 4 - T_BULK == 4, "exactly the last template value sits in the record's tail");
  // Shared memory holds the windows (the template is streamed from its record) and the buffers of the pixel-parallel
  // pass: 12.8 + 3.9 KB per warp for 11 x 11, 27 + 1.7 KB for 21 x 21 (seven warps: 201 KB)
  static constexpr int SLOTS = 32;
  static constexpr int WARPS_SM = WSM;
  static constexpr int WARPS = WARPS_SM % PAGK_LANES_CTA_WARPS == 0 ? PAGK_LANES_CTA_WARPS : WARPS_SM;  // per CTA
  static constexpr int CTAS_SM = WARPS_SM / WARPS;
  // per-warp scratch of the pixel-parallel pass: COOP_BUFS buffers (one slot each).  A buffer = COOP_CH records of
  // 32 bytes (Ix, Iy, -e as doubles, e * e as float: converted by the lanes that sample, so that the serial
  // accumulation loop has no conversion in it) -- the whole 11 x 11 patch, a chunk of a 21 x 21 one --, the two constants
  // c and 1 (doubles), the twelve sums and the cost the accumulator lanes hand back, and a row-major copy of the slot's
  // window.  Indices in floats.
  static constexpr int COOP_CH = HALF <= 5 ? 128 : 64;
  static constexpr int COOP_BUFS = (HALF <= 5 && WSM <= 8) ? 2 : 1;
  static constexpr int COOP_CONST = COOP_CH * 8;                // c, 1
  static constexpr int COOP_RES = COOP_CONST + 4;               // twelve doubles, then the cost
  static constexpr int COOP_LIN = COOP_RES + 28;                // the row-major window copy
  static constexpr int BUF_FLOATS = (COOP_LIN + WIN_WORDS + 3) & ~3;
  static constexpr int SCRATCH_FLOATS = COOP_BUFS * BUF_FLOATS;
  static constexpr int WARP_BYTES = WIN_WORDS * 128 + SCRATCH_FLOATS * 4;
  static constexpr int SMEM_BYTES = WARPS * WARP_BYTES;
  static_assert(WIN_W % 4 == 0 && WIN_W - (ALIGN - 1) >= P + 4, "window narrower than a sample box");
  // static shared memory of the kernel: level geometry and the table of pixel-pair offsets
  static constexpr int STATIC_BYTES = 256 + (PAGK_LANES_LDSPAIRS ? ((NP + 1) / 2) * 16 : 0);
  static_assert(SMEM_BYTES + STATIC_BYTES <= 227 * 1024, "a CTA's windows exceed the shared memory of a block");
  static_assert(CTAS_SM * (SMEM_BYTES + STATIC_BYTES + 1024) <= 228 * 1024, "windows do not fit the SM");
};

// floor(x) as float for 0 <= x < 2^22 without the conversion pipe: x + 2^23 rounds to an integer,
// subtracting 2^23 back is exact, one compare fixes the round-up case.
__device__ __forceinline__ float floor_nn(float x, int &i) {
  const float t = x + 8388608.0f;
  float r = t - 8388608.0f;
  i = __float_as_int(t) - 0x4B000000;
  if (r > x) { r -= 1.0f; i -= 1; }
  return r;
}
// developer aid (-DPAGK_LANES_CHECK): trap on a window index outside the staged window (compute-sanitizer stand-in)
#ifdef PAGK_LANES_CHECK
#define CHECK_IDX(i, lo, hi) do { if ((i) < (lo) || (i) > (hi)) __trap(); } while (0)
#else
#define CHECK_IDX(i, lo, hi)
#endif
// exact u8 -> float on the ALU + FP32 pipes
__device__ __forceinline__ float u8f(unsigned int b) { return __uint_as_float(0x4B000000u | b) - 8388608.0f; }

// ---- asynchronous copies (PTX cp.async; SASS: LDGSTS) ----
__device__ __forceinline__ unsigned int smem_u32(const void *p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(unsigned int dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// ---- TMA tile loads (cp.async.bulk.tensor; SASS: UTMALDG) and the mbarrier they complete on (SYNCS) ----
__device__ __forceinline__ void mbar_init(unsigned int bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned int bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned int bar, unsigned int parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// box of a 3-D tensor (x, y: a pyramid level with its padding; z: the image slot) -> shared memory, rows dense.
// Coordinates may lie outside the tensor: those elements arrive as zeros (and are never sampled).
__device__ __forceinline__ void tma_load_3d(unsigned int dst, const CUtensorMap *map, int x, int y, int z, unsigned int bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}

// PatchMatch::GetPixelValue (reference src/patch_match.cpp:391-406) with the four taps taken from a row-major copy
// `lin` (WIN_W bytes per row) of a staged window with origin (wx0, wy0): the same clamps, the same expression tree.
template <int WIN_W, int WIN_H>
__device__ __forceinline__ float window_sample(const unsigned char *__restrict__ lin, int wx0, int wy0, float fcols,
                                               float fcm1, float frows, float frm1, float x, float y) {
  if (x < 0.f) x = 0.f;
  if (y < 0.f) y = 0.f;
  if (x >= fcols) x = fcm1;
  if (y >= frows) y = frm1;
  int ix, iy;
  const float fx = floor_nn(x, ix), fy = floor_nn(y, iy);
  const float xx = x - fx, yy = y - fy, wa = 1.0f - xx, wb = 1.0f - yy;
  const int i = (iy - wy0) * WIN_W + (ix - wx0);
  CHECK_IDX(i, 0, WIN_W * WIN_H - WIN_W - 2);
  const unsigned char *q = lin + i;
  return wb * (wa * u8f(q[0]) + xx * u8f(q[1])) + yy * (wa * u8f(q[WIN_W]) + xx * u8f(q[WIN_W + 1]));
}

// ---- packed FP32 (sm_100: add / mul / fma .f32x2 -> FADD2 / FMUL2 / FFMA2, IEEE round-to-nearest per element).
// ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even under -fmad=false (measured, CUDA 12.9), so a sum with a
// product among its operands is written fma2(product, ONE, other) with ONE = (1, 1) from a kernel argument:
// rn(a * 1 + b) == rn(a + b), and a product that feeds the MULTIPLICAND of an FMA cannot be contracted into it.
// tests/test_sass.py counts the FFMA2 of the built kernel against the number written here.
struct f2 { float x, y; };
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return add2(a, f2{-b.x, -b.y}); }
__device__ __forceinline__ f2 addrd2(f2 a, f2 b) {  // rounded towards minus infinity
  f2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rm.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}

// developer aid (-DPAGK_LANES_PROF): per-warp cycles of every phase, rounds and active lane-rounds, wall-clock marks
// (start, queue exhausted, end) and what the warp had done when the queue ran out -> prof[warp_global][16]
#ifdef PAGK_LANES_PROF
__device__ __forceinline__ long long prof_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PROF_DECL long long pf[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long pt = clock64(); pf[8] = prof_ns();
#define PROF(i) do { const long long now_ = clock64(); pf[i] += now_ - pt; pt = now_; } while (0)
#define PROF_ADD(i, v) pf[i] += (v)
#define PROF_SET(i, v) pf[i] = (v)
#define PROF_FLUSH() do { pf[10] = prof_ns(); if (prof && lane == 0) for (int i_ = 0; i_ < 16; ++i_) prof[(size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * 16 + i_] = pf[i_]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_ADD(i, v)
#define PROF_SET(i, v)
#define PROF_FLUSH()
#endif

// patch offsets of the pixel pair (p, q) = (2k, 2k + 1) in the reference's order (y outer, x inner,
// src/patch_match.cpp:233-246) as (x_p, x_q, y_p, y_q), read with a warp-uniform index in the pass; the last pair of
// an odd patch repeats its pixel
__constant__ float4 c_pair5[61];
__constant__ float4 c_pair10[221];

struct Sums {
  double h00, h10, h11, h20, h21, h30, h31, b0, b1, b2, b3;
  float cost;
};

}  // namespace

// =================================================================================================
// K3a: the template records.  One warp per 32 consecutive (pair, feature, level) items.  Lane j decodes item j (one
// coalesced load of the keypoints) and fetches the item's raw (P+2) x (P+2) tap block of the reference image with ONE
// TMA tile load (UTMALDG: a box of the level's 3-D tensor at an arbitrary byte position, completion on the warp's
// mbarrier); then for each item lanes = pixels compute T (the reference's GetPixelValue on the reference image,
// src/patch_match.cpp:253, :263) from the staged taps and store it.  Lane j finally runs item j's ordered sum of c*c
// (the h22 entry of the normal matrix, the same for every iteration of the level: src/patch_match.cpp:296 with
// J[2] = c) and writes the record's tail.
// =================================================================================================
template <int HALF>
struct TmplCfg {
  static constexpr int P = 2 * HALF + 1;
  // floor(fl(pt + x)) is floor(pt) + x, or one more when the sum is rounded up across a binade: taps span P + 2 columns
  // and rows from floor(pt) - HALF.  A TMA box starts at a 16-byte aligned byte of the row (a tile load whose innermost
  // coordinate is not a multiple of 16 bytes faults: measured, tools/probes/tma_probe2.cu), so the box is the aligned
  // ROW_BYTES around those columns x ROWS.
  static constexpr int ROWS = P + 2;
  // (32 bytes would do for 11 x 11; with 48 the rows that one tap load of the pixel-pair lanes touches fall into distinct
  // shared-memory banks: 65 -> 61 us)
  static constexpr int ROW_BYTES = HALF <= 5 ? 48 : (15 + P + 2 + 15) & ~15;
  static constexpr int BOX_BYTES = ROWS * ROW_BYTES;
  static constexpr int ITEM_BYTES = (BOX_BYTES + 127) & ~127;  // TMA destinations are 128-byte aligned
  // items per group; two groups in flight: 10 / 9 KB per warp (groups of four with twice the CTAs per SM: 70 us against 61)
  static constexpr int G = HALF <= 5 ? 8 : 4;
  static constexpr int WARPS = 4;
};

template <int HALF>
__global__ void __launch_bounds__(TmplCfg<HALF>::WARPS * 32) pagk_lk_template_kernel(const unsigned char *__restrict__ images, PagkGeom g,
                                                             const __grid_constant__ PagkTmaLevels maps,
                                                             const PagkPairConst *__restrict__ pcs,
                                                             const float2 *__restrict__ keys_un, int levels, int max_keys,
                                                             int n_max, int n_pairs, unsigned char *__restrict__ tmpl, float unit) {
  using C = LanesCfg<HALF>;
  using TC = TmplCfg<HALF>;
  constexpr int P = C::P, NP = C::NP, TK = (NP + 31) / 32, G = TC::G, RB = TC::ROW_BYTES;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(128) unsigned char s_raw[TC::WARPS][2][G][TC::ITEM_BYTES];
  __shared__ __align__(8) unsigned long long s_bar[TC::WARPS][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long wg = (long long)blockIdx.x * TC::WARPS + warp;
  const long long total = (long long)n_pairs * n_max * levels, base = wg * 32;
  if (base >= total) return;
  const unsigned int mbar = smem_u32(&s_bar[warp][0]);  // the barrier of buffer b is mbar + 8 * b
  if (lane == 0) {
    mbar_init(mbar, 1); mbar_init(mbar + 8u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const float hf = (float)HALF;
  float tpx[TK], tpy[TK];
  int toff[TK];  // tap offset of pixel p inside the staged block when nothing is rounded: py * RB + px
#pragma unroll
  for (int k = 0; k < TK; ++k) {
    const int p = lane + 32 * k, py = p / P, px = p - py * P;
    tpx[k] = (float)(px - HALF); tpy[k] = (float)(py - HALF);
    toff[k] = py * RB + px;
  }
  // pixel pairs (2k, 2k + 1), k = lane + 32 * kk, of the shared-weight path: tap offsets of both pixels
  constexpr int NPAIR = (NP + 1) / 2, TK2 = (NPAIR + 31) / 32;
  int toffp[TK2], toffq[TK2];
#pragma unroll
  for (int kk = 0; kk < TK2; ++kk) {
    const int p = 2 * (lane + 32 * kk), q = p + 1 < NP ? p + 1 : p;
    toffp[kk] = (p / P) * RB + p % P;
    toffq[kk] = (q / P) * RB + q % P;
  }
  // ---- lane j: item j
  bool valid = false, tin = false, uni = false;
  float ptx = 0.f, pty = 0.f;
  int cols = 1, rows = 1, pitch = 4, wx0 = 0, wy0 = 0, lv = 0, slot = 0;
  const unsigned char *img1 = images;
  unsigned char *rec = nullptr;
  {
    const long long idx = base + lane;
    const int per_pair = n_max * levels;
    if (idx < total) {
      const int pr = (int)(idx / per_pair), rem = (int)(idx - (long long)pr * per_pair);
      const int i = rem / levels;
      lv = rem - i * levels;
      if (i < pcs[pr].n_keys) {
        valid = true;
        slot = pr * 2;
        const size_t o = (size_t)pr * max_keys + i;
        const float2 p1 = keys_un[o];
        const float scale = 1.0f / (float)(1 << lv);
        ptx = p1.x * scale; pty = p1.y * scale;  // pt = mvKeysRefUn[i].pt * mvScales[level] (:177)
        cols = g.lv[lv].cols; rows = g.lv[lv].rows; pitch = g.lv[lv].pitch;
        img1 = images + (size_t)slot * g.slot_bytes + g.lv[lv].offset;
        rec = tmpl + ((size_t)lv * ((size_t)n_pairs * max_keys) + o) * C::REC_BYTES;  // level-major: a level's records lie together
        const float txlo = ptx + (-hf), txhi = ptx + hf, tylo = pty + (-hf), tyhi = pty + hf;
        // no clamp of GetPixelValue fires anywhere in the template, and no tap lies in column `cols`: in a continuous
        // level that column is the next row's first byte, which a tile of the (x, y, slot) tensor does not see
        tin = txlo >= 0.0f && txhi < (float)(cols - 1) && tylo >= 0.0f && tyhi < (float)rows;
        if (tin) {
          wx0 = (int)txlo; wy0 = (int)tylo;  // pt - HALF is exact: floor(pt) - HALF (the box starts at wx0 & ~15)
          // pt + HALF exact (the difference below is exact by Sterbenz: pt >= HALF) <=> every pt + x, |x| <= HALF, is
          // exact: all pixels then share the fractions of pt and pixel (px, py) has its taps at (wx0 + px, wy0 + py)
          uni = (txhi - ptx == hf) && (tyhi - pty == hf);
        }
      }
    }
  }
  float myc = 0.f, mylast = 0.f;
  // ---- the tap blocks of a group of G items: every owner lane issues its own tile load.  Two groups are in flight: the
  // loads of group n + 1 are issued before the pixels of group n are computed.
  const unsigned int tin_mask = __ballot_sync(FULL, valid && tin), uni_mask = __ballot_sync(FULL, valid && tin && uni);
  auto issue = [&](const int grp) {
    const int g0 = grp * G, b = grp & 1;
    const unsigned int mm = (tin_mask >> g0) & ((1u << G) - 1u);
    if (mm) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the buffer was read through the generic proxy
      if (lane == 0) mbar_expect_tx(mbar + 8u * b, (unsigned int)TC::BOX_BYTES * (unsigned int)__popc(mm));
      __syncwarp();
      if (lane >= g0 && lane < g0 + G && valid && tin)
        tma_load_3d(smem_u32(&s_raw[warp][b][lane - g0][0]), &maps.lv[lv], wx0 & ~15, wy0, slot, mbar + 8u * b);
    }
  };
  issue(0);
  unsigned int phases = 0;  // bit b: parity of the phase buffer b's next wait is for
#pragma unroll 1
  for (int grp = 0; grp < 32 / G; ++grp) {
    const int g0 = grp * G, buf = grp & 1;
    if (grp + 1 < 32 / G) issue(grp + 1);
    // a buffer's barrier advances one phase per group that loaded anything into it
    if ((tin_mask >> g0) & ((1u << G) - 1u)) {
      mbar_wait(mbar + 8u * buf, (phases >> buf) & 1u);
      phases ^= 1u << buf;
    }
    __syncwarp();
    // ---- items whose pixels all have the weights of pt (the usual case), TWO items at a time so that a warp has two
    // independent chains of loads and arithmetic in flight.  Lanes = pixel PAIRS, the arithmetic packed across the pair, the
    // taps taken as raw bytes read as subnormal floats (b * 2^-149) against weights scaled by 2^100: the same scaling
    // argument as in the pass of the alignment kernel (a weight is 0 or at least 2^-21 here: pt >= HALF), every intermediate
    // is the reference's times a power of two, and the last product with 2^-51 is exact.
    {
      struct Item { f2 WA, XX, WB, YY; const unsigned char *raw; float *T; };
      auto prep = [&](const int j, const int jj) {
        const float sptx = __shfl_sync(FULL, ptx, j), spty = __shfl_sync(FULL, pty, j);
        const int sx0 = __shfl_sync(FULL, wx0, j);
        Item it;
        it.T = reinterpret_cast<float *>(__shfl_sync(FULL, (unsigned long long)rec, j));
        it.raw = &s_raw[warp][buf][jj][0] + (sx0 & 15);  // the tap of column x, row y sits at (y - wy0) * RB + (x - (wx0 & ~15))
        // floor by a round-down add of 2^23 (exact for 0 <= x < 2^22)
        const float xx = sptx - (__fadd_rd(sptx, 8388608.0f) - 8388608.0f), yy = spty - (__fadd_rd(spty, 8388608.0f) - 8388608.0f);
        const float wa = 1.0f - xx, wb = 1.0f - yy;
        constexpr float SC = 1.2676506002282294e30f;  // 2^100
        it.WA = f2{wa * SC, wa * SC}; it.XX = f2{xx * SC, xx * SC}; it.WB = f2{wb * SC, wb * SC}; it.YY = f2{yy * SC, yy * SC};
        return it;
      };
      const f2 ONE = {unit, unit}, UN = {4.440892098500626e-16f, 4.440892098500626e-16f};  // 1, 2^-51
      auto pair_of = [&](const Item &it, const int kk) {
        const unsigned char *a = it.raw + toffp[kk], *b = it.raw + toffq[kk];
#define PAGK_TAP2(o) f2{__uint_as_float((unsigned int)a[o]), __uint_as_float((unsigned int)b[o])}
        const f2 t00 = PAGK_TAP2(0), t01 = PAGK_TAP2(1), t10 = PAGK_TAP2(RB), t11 = PAGK_TAP2(RB + 1);
#undef PAGK_TAP2
        const f2 top = fma2(mul2(it.WA, t00), ONE, mul2(it.XX, t01)), bot = fma2(mul2(it.WA, t10), ONE, mul2(it.XX, t11));
        return mul2(fma2(mul2(it.WB, top), ONE, mul2(it.YY, bot)), UN);
      };
      auto put = [&](const Item &it, const int k, const f2 v, float &tcv, float &tlv) {
        const int p = 2 * k;
        if (p + 1 < NP - 1) *reinterpret_cast<float2 *>(it.T + p) = make_float2(v.x, v.y);
        else if (p < NP - 1) it.T[p] = v.x;
        if (k == (NP / 2) / 2) tcv = (NP / 2) % 2 ? v.y : v.x;
        if (k == (NP - 1) / 2) tlv = v.x;
      };
      unsigned int m2 = (uni_mask >> g0) & ((1u << G) - 1u);
      while (m2) {
        const int ja = __ffs(m2) - 1;
        m2 &= m2 - 1;
        const bool two = m2 != 0u;
        const int jb = two ? __ffs(m2) - 1 : ja;
        m2 &= m2 - 1;
        const Item A = prep(g0 + ja, ja), B = prep(g0 + jb, jb);
        float tca = 0.f, tla = 0.f, tcb = 0.f, tlb = 0.f;
#pragma unroll
        for (int kk = 0; kk < TK2; ++kk) {
          const int k = lane + 32 * kk;
          if (k < NPAIR) {
            const f2 va = pair_of(A, kk), vb = pair_of(B, kk);
            put(A, k, va, tca, tla);
            if (two) put(B, k, vb, tcb, tlb);
          }
        }
        const float ca = __shfl_sync(FULL, tca, ((NP / 2) / 2) % 32), la = __shfl_sync(FULL, tla, ((NP - 1) / 2) % 32);
        const float cb = __shfl_sync(FULL, tcb, ((NP / 2) / 2) % 32), lb = __shfl_sync(FULL, tlb, ((NP - 1) / 2) % 32);
        if (lane == g0 + ja) { myc = -ca; mylast = la; }
        if (two && lane == g0 + jb) { myc = -cb; mylast = lb; }
      }
    }
    // ---- the other items, one at a time
#pragma unroll 1
    for (int jj = 0; jj < G; ++jj) {
      const int j = g0 + jj;
      if (!__shfl_sync(FULL, (int)(valid && !(tin && uni)), j)) continue;
      const int kind = __shfl_sync(FULL, (int)tin + (int)uni, j);  // 0: clamps may fire, 1: per-pixel weights
      const float sptx = __shfl_sync(FULL, ptx, j), spty = __shfl_sync(FULL, pty, j);
      float *T = reinterpret_cast<float *>(__shfl_sync(FULL, (unsigned long long)rec, j));
      // the tap of column x, row y sits at (y - wy0) * RB + (x - (wx0 & ~15))
      const int sx0 = __shfl_sync(FULL, wx0, j), sy0 = __shfl_sync(FULL, wy0, j);
      const unsigned char *raw = &s_raw[warp][buf][jj][0] + (sx0 & 15);
      float tv[TK];
      if (kind == 1) {
#pragma unroll
        for (int k = 0; k < TK; ++k) {
          const float cx = sptx + tpx[k], cy = spty + tpy[k];
          const float tx = __fadd_rd(cx, 8388608.0f), ty = __fadd_rd(cy, 8388608.0f);
          const float xx = cx - (tx - 8388608.0f), yy = cy - (ty - 8388608.0f), wa = 1.0f - xx, wb = 1.0f - yy;
          const int ix = __float_as_int(tx) - 0x4B000000, iy = __float_as_int(ty) - 0x4B000000;
          tv[k] = 0.f;
          if (lane + 32 * k < NP) {
            CHECK_IDX((iy - sy0) * RB + (ix - sx0) + (sx0 & 15), 0, TC::BOX_BYTES - RB - 2);
            const unsigned char *q = raw + (iy - sy0) * RB + (ix - sx0);
            tv[k] = wb * (wa * u8f(q[0]) + xx * u8f(q[1])) + yy * (wa * u8f(q[RB]) + xx * u8f(q[RB + 1]));
          }
        }
      } else {
        const int scols = __shfl_sync(FULL, cols, j), srows = __shfl_sync(FULL, rows, j), sp = __shfl_sync(FULL, pitch, j);
        const unsigned char *simg = reinterpret_cast<const unsigned char *>(__shfl_sync(FULL, (unsigned long long)img1, j));
#pragma unroll
        for (int k = 0; k < TK; ++k) {
          tv[k] = 0.f;
          if (lane + 32 * k < NP) tv[k] = pagk_sample_call(simg, sp, scols, srows, sptx + tpx[k], spty + tpy[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < TK; ++k) {
        const int p = lane + 32 * k;
        if (p < NP - 1) T[p] = tv[k];
      }
      // de_dg = -I1(pt) (src/patch_match.cpp:263) is minus the template value of the centre pixel: pt + (0, 0)
      const float tc = __shfl_sync(FULL, tv[(NP / 2) / 32], (NP / 2) % 32);
      const float tl = __shfl_sync(FULL, tv[(NP - 1) / 32], (NP - 1) % 32);
      if (lane == j) { myc = -tc; mylast = tl; }
    }
    __syncwarp();
  }
  if (valid) {
    const double c = (double)myc;
    double h22 = 0.0;
#pragma unroll 11
    for (int p = 0; p < NP; ++p) h22 = fma(c, c, h22);  // c*c is exact in double: DFMA rounds like mul + add
    float4 tail;
    tail.x = mylast; tail.y = myc;
    tail.z = __int_as_float(__double2loint(h22)); tail.w = __int_as_float(__double2hiint(h22));
    *reinterpret_cast<float4 *>(rec + C::T_BULK) = tail;
  }
}

// =================================================================================================
// K3b: the alignment kernel
// =================================================================================================
// The hand-over of a feature from one level to the next (level-granular work items): a 32-byte record per feature of
// three 8-byte words, each carrying the tag epoch * 8 + levels finished in its upper half:
//   (tag, pm.x) (tag, pm.y) (tag, passes so far)
// An aligned 8-byte access is single-copy atomic, so a reader that finds the expected tag in all three words has the
// three values of that publication whatever order the words were written or read in: no acquire load in front of the
// data, and the record is fetched together with everything else a claim needs (one global round trip).
__device__ __forceinline__ void handover_store(unsigned long long *rec, int tag, float x, float y, int n) {
  const unsigned long long t = (unsigned long long)(unsigned int)tag << 32;
  const unsigned long long q0 = t | (unsigned long long)__float_as_uint(x), q1 = t | (unsigned long long)__float_as_uint(y);
  const unsigned long long q2 = t | (unsigned long long)(unsigned int)n;
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(rec), "l"(q0), "l"(q1) : "memory");
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(rec + 2), "l"(q2) : "memory");
}
__device__ __forceinline__ void handover_load(const unsigned long long *rec, unsigned long long &q0, unsigned long long &q1, unsigned long long &q2) {
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(q0), "=l"(q1) : "l"(rec) : "memory");
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(q2) : "l"(rec + 2) : "memory");
}
__device__ __forceinline__ bool handover_valid(int tag, unsigned long long q0, unsigned long long q1, unsigned long long q2) {
  return (unsigned int)(q0 >> 32) == (unsigned int)tag && (unsigned int)(q1 >> 32) == (unsigned int)tag && (unsigned int)(q2 >> 32) == (unsigned int)tag;
}

template <int HALF, bool AFFINE, int WSM>
__global__ void __launch_bounds__(LanesCfg<HALF, WSM>::WARPS * 32, PAGK_LANES_REGCTAS && HALF <= 5 ? PAGK_LANES_REGCTAS : LanesCfg<HALF, WSM>::CTAS_SM)
pagk_lk_lanes_kernel(const unsigned char *__restrict__ images, PagkGeom g, const PagkPairConst *__restrict__ pcs,
                     const float2 *__restrict__ keys_un, PagkOutPtrs out, PagkMode mode, int max_keys, int n_max,
                     int n_pairs, int *__restrict__ work_counter, int *__restrict__ next_counter, int lane_cap,
                     int split, int *progress, int epoch_base, const unsigned char *__restrict__ tmpl, float unit,
                     long long *__restrict__ prof) {
  using C = LanesCfg<HALF, WSM>;
  constexpr int P = C::P, NP = C::NP, WIN_W = C::WIN_W, WIN_H = C::WIN_H;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wwin = smem_raw + (size_t)warp * C::WARP_BYTES;                          // [WIN_WORDS][32 lanes] words
  float *scratch = reinterpret_cast<float *>(wwin + C::WIN_WORDS * 128);                  // COOP_BUFS buffers of the pixel-parallel pass
  const unsigned char *mywin = wwin + lane * 4;                                            // byte 0 of this lane's window
  const int total_work = n_pairs * n_max;
  const int top = mode.levels - 1;
  unsigned long long *handover = reinterpret_cast<unsigned long long *>(progress);
  // template records are level-major ([level][pair * max_keys + feature]): the lanes work on one level at a time (the
  // queue is level-major too), so the records in use lie together and stay in L2 between a level's passes
  const size_t lv_stride = (size_t)n_pairs * (size_t)max_keys;
  // Work items.  split == 0: an item is a feature (all levels in one lane).  split != 0: an item is one LEVEL of a
  // feature, queued level-major (every feature's coarsest level first).  All a level hands to the next one is
  // mvPtPyr2Un[i] (src/patch_match.cpp:348; dg, db, cost restart per level), so the hand-over is that point and the
  // running pass count in the feature's tagged hand-over record (above; tag = epoch_base + levels finished).  The tail of
  // a launch -- lanes idling while the last items finish -- is then one level long instead of one feature life long.
  const int total_items = split ? total_work * mode.levels : total_work;
  const unsigned long long slot_bytes = g.slot_bytes;
  const float hf = (float)HALF;
  constexpr unsigned FULL = 0xffffffffu;
  // level geometry, indexed by each lane's own level
  __shared__ int s_cols[PAGK_MAX_LEVELS], s_rows[PAGK_MAX_LEVELS], s_pitch[PAGK_MAX_LEVELS];
  __shared__ unsigned int s_off[PAGK_MAX_LEVELS];
  // -DPAGK_LANES_LDSPAIRS: patch offsets of the pixel pairs from shared memory (one broadcast LDS.128 per pair)
#if PAGK_LANES_LDSPAIRS
  __shared__ float4 s_pairs[(NP + 1) / 2];
  for (int k = threadIdx.x; k < (NP + 1) / 2; k += blockDim.x) s_pairs[k] = (HALF == 5 ? c_pair5 : c_pair10)[k];
#endif
  // the two work counters of a handle alternate between launches: this launch zeroes the one the next launch uses
  if (blockIdx.x == 0 && threadIdx.x == 0) *next_counter = 0;
  if (threadIdx.x < PAGK_MAX_LEVELS) {
    s_cols[threadIdx.x] = g.lv[threadIdx.x].cols; s_rows[threadIdx.x] = g.lv[threadIdx.x].rows;
    s_pitch[threadIdx.x] = g.lv[threadIdx.x].pitch; s_off[threadIdx.x] = g.lv[threadIdx.x].offset;
  }
  __syncthreads();

  // accumulator role of this lane in the pixel-parallel pass: acc += A * B with A in {Ix, Iy, c, 1} and B in
  // {Ix, Iy, -e, c}; roles 0..11 = h00 h10 h11 h20 h21 h22 h30 h31 b0 b1 b2 b3.  An operand is a float in the lane's
  // buffer: a record field (stride 3) or one of the two constants behind the records (stride 0).  With two buffers
  // lanes 0..15 serve the first one, lanes 16..31 the second.
  const int cbuf = C::COOP_BUFS == 2 ? lane >> 4 : 0;
  const int role16 = C::COOP_BUFS == 2 ? lane & 15 : lane;
  const int role = role16 < 12 ? role16 : 0;
  const int selA = (int)((0x431044333110ull >> (4 * role)) & 0xfull);  // 0 Ix, 1 Iy, 3 c, 4 one
  const int selB = (int)((0x222210310100ull >> (4 * role)) & 0xfull);  // 0 Ix, 1 Iy, 2 -e, 3 c
  // offsets and strides in doubles: a record is four doubles wide
  const int offA = selA < 3 ? selA : C::COOP_CONST / 2 + (selA - 3);
  const int offB = selB < 3 ? selB : C::COOP_CONST / 2 + (selB - 3);
  // (opaque to the compiler: as plain registers the strides cost one add per operand in the serial walk; known to be 0 or 4
  // they become a predicated pair of instructions per operand -- 3.5 % of a single pair's patch alignment)
  int strideA = selA < 3 ? 4 : 0, strideB = selB < 3 ? 4 : 0;
  asm("" : "+r"(strideA), "+r"(strideB));

  // ---- slot state (registers of the owning lane) ----
  int feat = -1, pair = 0, level = 0, iter = 0, n_iter = 0, succ = 1;
  float pt1x = 0.f, pt1y = 0.f, ptx = 0.f, pty = 0.f, dx = 0.f, dy = 0.f, dg = 0.f, db = 0.f, lastCost = 0.f, cval = 0.f;
  const float4 *Tg = reinterpret_cast<const float4 *>(tmpl);  // the level's template record (any valid record when idle)
  float tlast = 0.f;   // T[NP - 1]
  double h22v = 0.0;   // the level's sum of c*c
  float a00 = 1.f, a01 = 0.f, a10 = 0.f, a11 = 1.f;
  float wxmin = -hf, wxmax = hf, wymin = -hf, wymax = hf;
  int win_x0 = 0, win_y0 = 0;
  bool needs_setup = false, win_valid = false;
  bool waiting = false;    // split mode: the item is claimed, the level above it is not finished yet
  int item_lo = 0;         // last level of the claimed item
  bool exhausted = false;  // warp-uniform
  PROF_DECL

  while (true) {
    // ------------------------------------------------------------------ refill
    while (!exhausted) {
      const bool want = feat < 0 && lane < lane_cap;
      const unsigned m = __ballot_sync(FULL, want);
      if (m == 0u) break;
      const int cnt = __popc(m);
      int base = 0;
      if (lane == 0) base = atomicAdd(work_counter, cnt);
      base = __shfl_sync(FULL, base, 0);
      if (base + cnt >= total_items) { exhausted = true; PROF_SET(9, prof_ns()); PROF_SET(11, pf[6]); PROF_SET(12, pf[7]); }
      if (want) {
        const int wi = base + __popc(m & ((1u << lane) - 1u));
        if (wi < total_items) {
          int idx = wi, lv = top, lo = 0;
          if (split) { const int k = wi / total_work; idx = wi - k * total_work; lv = top - k; lo = lv; }
          const int pr = idx / n_max, i = idx - pr * n_max;
          // everything the claim needs, loaded in one go (o is inside the arrays for any i < n_max <= max_keys): one global
          // round trip after the counter's instead of a chain of dependent ones
          const size_t o = (size_t)pr * max_keys + i;
          const int nk = pcs[pr].n_keys;
          const unsigned char st = out.gyro_status[o];
          const float2 p1 = keys_un[o];
          const float4 A = out.affine[o];
          float2 pp = p1;
          unsigned long long q0 = 0ull, q1 = 0ull, q2 = 0ull;
          if (lv == top) {
            if (mode.gyro_init) pp = out.pt_predict_un[o];
          } else {
            handover_load(handover + 4 * o, q0, q1, q2);
          }
          // the tail of the level's template record as well: its address depends on the item only
          const unsigned char *rec = tmpl + ((size_t)lv * lv_stride + o) * C::REC_BYTES;
          const float4 tl = __ldg(reinterpret_cast<const float4 *>(rec + C::T_BULK));
          if (i < nk) {
            if (!st) {  // the reference skips these (src/patch_match.cpp:173): default outputs
              if (lv == top) {
                out.pm_un[o] = pp;
                out.pm_status[o] = 0; out.pix_err[o] = 0.0; out.ncc[o] = 0.f; out.iters[o] = 0;
              }
            } else {
              const float scale = 1.0f / (float)(1 << lv);
              feat = (int)o; pair = pr; level = lv; item_lo = lo; needs_setup = false; win_valid = false;
              Tg = reinterpret_cast<const float4 *>(rec);
              tlast = tl.x; cval = tl.y;
              h22v = __hiloint2double(__float_as_int(tl.w), __float_as_int(tl.z));
              pt1x = p1.x; pt1y = p1.y;
              ptx = p1.x * scale; pty = p1.y * scale;
              if (lv == top) {
                dx = pp.x * scale - ptx; dy = pp.y * scale - pty;
                n_iter = 0; waiting = false;
              } else if (handover_valid(epoch_base + (top - lv), q0, q1, q2)) {  // the level above is already published (the usual case)
                n_iter = (int)(unsigned int)q2;
                dx = __uint_as_float((unsigned int)q0) * 2.0f - ptx; dy = __uint_as_float((unsigned int)q1) * 2.0f - pty;  // nextPt = mvPtPyr2Un[i] * 1.0f / mPyramidScale (:182)
                waiting = false;
              } else {
                waiting = true;  // polled once per round below
              }
              dg = 0.f; db = 0.f; lastCost = 0.f; iter = 0; succ = 1;
              a00 = A.x; a01 = A.y; a10 = A.z; a11 = A.w;
              wxmin = -hf; wxmax = hf; wymin = -hf; wymax = hf;
              if (AFFINE) {  // warp offsets at the four patch corners, exactly as the pass computes them
                const float c0x = a00 * -hf + a01 * -hf, c1x = a00 * hf + a01 * -hf, c2x = a00 * -hf + a01 * hf, c3x = a00 * hf + a01 * hf;
                const float c0y = a10 * -hf + a11 * -hf, c1y = a10 * hf + a11 * -hf, c2y = a10 * -hf + a11 * hf, c3y = a10 * hf + a11 * hf;
                wxmin = fminf(fminf(c0x, c1x), fminf(c2x, c3x)); wxmax = fmaxf(fmaxf(c0x, c1x), fmaxf(c2x, c3x));
                wymin = fminf(fminf(c0y, c1y), fminf(c2y, c3y)); wymax = fmaxf(fmaxf(c0y, c1y), fmaxf(c2y, c3y));
              }
            }
          }
        }
      }
    }
    // ------------------------------------------------------------------ a level that starts: its template record
    // (the pass streams T from it; c, h22 and the last T value come from the record's tail).  The template does not depend
    // on the level above, so a lane that still waits for its hand-over does this as well.
    if (feat >= 0 && needs_setup) {
      const unsigned char *rec = tmpl + ((size_t)level * lv_stride + (size_t)feat) * C::REC_BYTES;
      Tg = reinterpret_cast<const float4 *>(rec);
      const float4 tl = __ldg(reinterpret_cast<const float4 *>(rec + C::T_BULK));
      tlast = tl.x; cval = tl.y;
      h22v = __hiloint2double(__float_as_int(tl.w), __float_as_int(tl.z));
      needs_setup = false; win_valid = false;
    }
    if (split) {
      // a claimed level starts once the level above it has published its result.  The owner of that level is a
      // running lane of this launch (items are claimed in queue order), so polling once per round cannot deadlock.
      if (waiting) {
        unsigned long long q0, q1, q2;
        handover_load(handover + 4 * (size_t)feat, q0, q1, q2);
        if (handover_valid(epoch_base + (top - level), q0, q1, q2)) {
          n_iter = (int)(unsigned int)q2;
          dx = __uint_as_float((unsigned int)q0) * 2.0f - ptx; dy = __uint_as_float((unsigned int)q1) * 2.0f - pty;  // nextPt = mvPtPyr2Un[i] * 1.0f / mPyramidScale (:182)
          waiting = false;
        }
      }
    }
    const bool active = feat >= 0 && !waiting;
    const unsigned m_active = __ballot_sync(FULL, active);
    if (m_active == 0u) {
      if (__ballot_sync(FULL, waiting) != 0u) { __nanosleep(200); continue; }  // nothing to run this round: poll again
      break;
    }
    PROF(0); PROF_ADD(6, 1); PROF_ADD(7, __popc(m_active)); PROF_ADD(2, __popc(__ballot_sync(FULL, waiting)));

    const int cols = s_cols[level], rows = s_rows[level], pitch = s_pitch[level];
    const unsigned char *I2w = images + (size_t)(pair * 2 + 1) * slot_bytes + s_off[level];  // the plane the windows come from
    const float fcols = (float)cols, frows = (float)rows, fcm1 = (float)(cols - 1), frm1 = (float)(rows - 1);

    // ------------------------------------------------------------------ sample box of this pass (lane = slot)
    // Extreme sample coordinates over the patch: the warp offsets are monotone in x and in y, so their extremes
    // sit at the corners; the +-1 of the gradient samples and every rounding are monotone too.  `inside`: no
    // clamp of GetPixelValue can fire for any sample.  The box of the CLAMPED coordinates is what a window has to
    // hold (a level keeps its wrap column and its guard row), so border slots are windowable as well.
    const float bx = ptx + dx, by = pty + dy;
    bool windowable = false, fast = false, restage = false;
    int nx0 = 0, ny0 = 0;
    {
      const float x2min = (bx + wxmin) - 1.0f, x1max = (bx + wxmax) + 1.0f;
      const float y2min = (by + wymin) - 1.0f, y1max = (by + wymax) + 1.0f;
      const bool inside = (x2min >= 0.0f) && (x1max < fcols) && (y2min >= 0.0f) && (y1max < frows);
      const bool sane = fabsf(x2min) < 1.0e6f && fabsf(x1max) < 1.0e6f && fabsf(y2min) < 1.0e6f && fabsf(y1max) < 1.0e6f;  // false for NaN
      if (active && sane) {
        const int ixlo = (int)fminf(fmaxf(x2min, 0.0f), fcm1), ixhi = (int)fminf(fmaxf(x1max, 0.0f), fcm1) + 1;
        const int iylo = (int)fminf(fmaxf(y2min, 0.0f), frm1), iyhi = (int)fminf(fmaxf(y1max, 0.0f), frm1) + 1;
        const int needw = ixhi - ixlo + 1, needh = iyhi - iylo + 1;
        // the window's origin is a multiple of ALIGN columns: any box up to WIN_W - (ALIGN - 1) wide fits
        if (needw <= WIN_W - (C::ALIGN - 1) && needh <= WIN_H) {
          windowable = true;
          fast = inside;
          const bool ok = win_valid && ixlo >= win_x0 && ixhi <= win_x0 + WIN_W - 1 && iylo >= win_y0 && iyhi <= win_y0 + WIN_H - 1;
          if (!ok) {  // (re)stage, centred on the needed box: 0 <= x0 <= ixlo, x0 + WIN_W - 1 >= ixhi, the same in y;
                      // the rows of a level are followed by at least 32 - rows allocated rows and 64 bytes
            restage = true;
            nx0 = max((ixlo - (WIN_W - (C::ALIGN - 1) - needw) / 2) & ~(C::ALIGN - 1), 0);
            ny0 = max(min(iylo - (WIN_H - needh) / 2, rows + 1 - WIN_H), 0);
          }
        }
      }
    }

    // the first template vectors of the pass: in flight across the window copies below
#if PAGK_LANES_TPRE
    float4 t4[PAGK_LANES_TRIP];
#pragma unroll
    for (int u = 0; u < PAGK_LANES_TRIP; ++u) t4[u] = __ldg(Tg + u);
#endif

    // ------------------------------------------------------------------ window of the current level: LDGSTS, lane = slot
    // Every lane that (re)stages copies its own window word by word: the destinations of one instruction are the
    // lanes' own banks (no conflict), the sources WIN_H short row segments per lane.
    if (__ballot_sync(FULL, restage) != 0u) {
      if (restage) {
        const unsigned char *src = I2w + (size_t)ny0 * (size_t)pitch + (size_t)nx0;
        const unsigned int dst = smem_u32(mywin);
#pragma unroll 1
        for (int r = 0; r < WIN_H; ++r) {
#pragma unroll
          for (int w = 0; w < C::WPR; ++w) cp_async4(dst + (unsigned int)((r * C::WPR + w) * 128), src + 4 * w);
          src += pitch;
        }
        win_x0 = nx0; win_y0 = ny0; win_valid = true;
      }
    }
    // one wait for every copy of the round
    cp_async_wait_all();
    __syncwarp();
    PROF(1);

    Sums S;
    S.h00 = S.h10 = S.h11 = S.h20 = S.h21 = S.h30 = S.h31 = S.b0 = S.b1 = S.b2 = S.b3 = 0.0;
    S.cost = 0.f;
    const double c = (double)cval;
    const float gain = 1.0f + dg;
    const bool sparse = __popc(m_active) <= PAGK_LANES_SPARSE;
    bool coop = active && (sparse || !fast);
    bool lockbad = false;  // the lockstep pass met the rounding case its shared weights exclude

    // ------------------------------------------------------------------ the pass: lane = slot
    // Bit-exactness of the window path.  The reference samples at (sx, sy), (sx+-1, sy), (sx, sy+-1), each
    // through GetPixelValue (src/patch_match.cpp:391-406).  With no clamp firing (`fast`):
    //   sx - 1 is exact (same or finer binade), so that sample has floor(sx) - 1 and the weights of sx;
    //   X1 = fl(sx + 1) lies in [fx + 1, fx + 2]; X1 - (fx + 1) is exact (Sterbenz) and equals the reference's
    //   X1 - floor(X1) unless X1 == fx + 2, where it is exactly 1 -- flagged `bad`, the slot then redoes the
    //   pass through the pixel-parallel pass, which evaluates every sample on its own.
    // Horizontal interpolations are shared between samples only where the reference would evaluate the
    // identical expression.
    // Scaling.  The coordinate chain (pbx, the affine matrix, 1, 2^23) carries the factor SC = 2^100.  Every operation
    // on it is an addition, a subtraction or a product with a small integer, 0 <= coordinates < 2^22, so every
    // intermediate of the scaled chain is 2^100 times the reference's (scaling by a power of two commutes with
    // rounding while nothing overflows or goes subnormal: the largest value is 2^123 + 2^122, the smallest non-zero
    // one a weight of 2^-23 * 2^100).  A tap is the byte b read as the float b * 2^-149 (subnormal operands are full
    // speed).  A horizontal interpolation is then rn(w * b) * 2^-49 + ..., exactly the reference's times 2^-49 (smallest
    // non-zero value 2^-72), a sample -- one more scaled weight -- the reference's times 2^51 (smallest non-zero
    // product 2^-95, differences are multiples of 2^-118: all normal).  The residual takes the sample through
    // fma(v, 2^-51, db) = rn(v0 + db) (the product is exact), the gradient differences keep the factor into the double
    // sums, where it is a power of two again and leaves after the loop together with the 0.5 of the central difference.
    if (!sparse) {
      constexpr float SC = 1.2676506002282294e30f;  // 2^100
      constexpr float BIGS = 8388608.0f * SC;       // 2^123: a coordinate added to it (rounding down) loses its fraction
      constexpr float UNS = 4.440892098500626e-16f; // 2^-51
      float badv = 0.0f;
      // lanes without a fast slot walk a harmless patch in the middle of their window
      const float pbx = (fast ? bx : (float)(HALF + 5) + 0.5f) * SC, pby = (fast ? by : (float)(HALF + 3) + 0.5f) * SC;
      const f2 PBX = {pbx, pbx}, PBY = {pby, pby};
      const float q00 = (fast ? a00 : 1.0f) * SC, q01 = (fast ? a01 : 0.0f) * SC, q10 = (fast ? a10 : 0.0f) * SC, q11 = (fast ? a11 : 1.0f) * SC;
      const f2 Q00 = {q00, q00}, Q01 = {q01, q01}, Q10 = {q10, q10}, Q11 = {q11, q11};
      const f2 ONE = {unit, unit}, BIG = {BIGS, BIGS}, NBIG = {-BIGS, -BIGS}, P1 = {SC, SC}, UN = {UNS, UNS};
      const f2 DB = {db, db}, GAIN = {gain, gain};
      // window byte index of the tap left of (floor(sx), floor(sy)) from the mantissas of ty = 2^123 + floor(sy) * 2^100
      // and tx: bits(ty) * WIN_W + bits(tx) - kk, kk = (WIN_W + 1) * bits(2^123) + origin + 1 (mod 2^32)
      const unsigned int kk = (unsigned int)(WIN_W + 1) * 0x7D000000u + (unsigned int)(fast ? win_y0 * WIN_W + win_x0 : 0) + 1u;
#if PAGK_LANES_LDSPAIRS
      const float4 *pix = s_pairs;
#else
      const float4 *pix = HALF == 5 ? c_pair5 : c_pair10;
#endif
      // the four taps of a window row start at byte i1 of the lane's window; tap j sits at
      // A0 + j + 124 * ((s + j) >> 2), A0 = (i1 >> 2) * 128 + (i1 & 3), s = i1 & 3 (a word of the lane is 128 bytes from the
      // next one); rows are ROW_BYTES apart
      auto bases = [&](const unsigned int i1, const unsigned char *&b0, const unsigned char *&b1, const unsigned char *&b2, const unsigned char *&b3) {
        CHECK_IDX((int)i1, WIN_W, WIN_W * WIN_H - 2 * WIN_W - 4);
        const unsigned int s = i1 & 3u;
        b0 = mywin + ((i1 >> 2) * 124u + i1);
        b1 = b0 + ((s + 1u) & 4u) * 31u; b2 = b0 + ((s + 2u) & 4u) * 31u; b3 = b0 + ((s + 3u) & 4u) * 31u;
      };
      // A pixel pair p = 2 * k2 (.x), p + 1 (.y) in two halves.  front: sample coordinates, weights, tap loads (a pair's
      // 20 weights and 24 taps); back: the interpolations, residual, gradient and the ordered sums (`both` false: only .x
      // enters the sums -- the last pixel of an odd patch, whose pair repeats it).  The loop issues the front half of
      // pair k + 1 before the back half of pair k, so that a warp alone on its scheduler has a pair's loads and address
      // arithmetic in flight under the previous pair's arithmetic (tools/ubench_pass2.cu V5: 10 % fewer cycles per pass
      // for a lone warp, 9 % with three warps per scheduler).
      struct Front { f2 WA, XX, WB, YY, WA1, XX1, WB1, YY1, m0, m1, c_1, c0, c1, c2, d_1, d0, d1, d2, n0, n1; };
      auto front = [&](const int k2, Front &F) {
        const float4 xy = pix[k2];
        const f2 XF = {xy.x, xy.y}, YF = {xy.z, xy.w};
        f2 SX, SY;
        if (AFFINE) {
          const f2 WX = fma2(mul2(Q00, XF), ONE, mul2(Q01, YF)), WY = fma2(mul2(Q10, XF), ONE, mul2(Q11, YF));
          SX = add2(PBX, WX); SY = add2(PBY, WY);
        } else {  // the scaled offset x * 2^100 is exact: the fused form rounds once, like the sum it stands for
          SX = fma2(XF, P1, PBX); SY = fma2(YF, P1, PBY);
        }
        // floor for 0 <= x < 2^22: x + 2^23 rounded DOWN is 2^23 + floor(x) exactly; taking 2^23 off is exact
        const f2 TX = addrd2(SX, BIG), TY = addrd2(SY, BIG);
        const f2 FX = add2(TX, NBIG), FY = add2(TY, NBIG);
        F.XX = sub2(SX, FX); F.YY = sub2(SY, FY);
        F.WA = sub2(P1, F.XX); F.WB = sub2(P1, F.YY);
        const f2 X1 = add2(SX, P1), Y1 = add2(SY, P1);
        F.XX1 = sub2(X1, add2(FX, P1)); F.YY1 = sub2(Y1, add2(FY, P1));
        F.WA1 = sub2(P1, F.XX1); F.WB1 = sub2(P1, F.YY1);
        badv = fmaxf(badv, fmaxf(fmaxf(F.XX1.x, F.YY1.x), fmaxf(F.XX1.y, F.YY1.y)));
        const unsigned int ip = (unsigned int)__float_as_int(TY.x) * (unsigned int)WIN_W + (unsigned int)__float_as_int(TX.x) - kk;
        const unsigned int iq = (unsigned int)__float_as_int(TY.y) * (unsigned int)WIN_W + (unsigned int)__float_as_int(TX.y) - kk;
        const unsigned char *p0, *p1, *p2, *p3, *q0, *q1, *q2, *q3;
        bases(ip, p0, p1, p2, p3); bases(iq, q0, q1, q2, q3);
        constexpr int R = C::ROW_BYTES;
#define PAGK_TAP(pb, qb, o) f2{__uint_as_float((unsigned int)pb[o]), __uint_as_float((unsigned int)qb[o])}
        F.m0 = PAGK_TAP(p1, q1, 1 - R); F.m1 = PAGK_TAP(p2, q2, 2 - R);
        F.c_1 = PAGK_TAP(p0, q0, 0); F.c0 = PAGK_TAP(p1, q1, 1); F.c1 = PAGK_TAP(p2, q2, 2); F.c2 = PAGK_TAP(p3, q3, 3);
        F.d_1 = PAGK_TAP(p0, q0, R); F.d0 = PAGK_TAP(p1, q1, 1 + R); F.d1 = PAGK_TAP(p2, q2, 2 + R); F.d2 = PAGK_TAP(p3, q3, 3 + R);
        F.n0 = PAGK_TAP(p1, q1, 1 + 2 * R); F.n1 = PAGK_TAP(p2, q2, 2 + 2 * R);
#undef PAGK_TAP
      };
      auto back = [&](const Front &F, const f2 tv, const bool both) {
        // rn(rn(w * a) + rn(u * b)): two products and a sum that ptxas cannot contract
#define PAGK_LERP(w, a, u, b) fma2(mul2(w, a), ONE, mul2(u, b))
        const f2 Hm = PAGK_LERP(F.WA, F.m0, F.XX, F.m1);
        const f2 H0 = PAGK_LERP(F.WA, F.c0, F.XX, F.c1), H0p = PAGK_LERP(F.WA1, F.c1, F.XX1, F.c2), H0m = PAGK_LERP(F.WA, F.c_1, F.XX, F.c0);
        const f2 H1 = PAGK_LERP(F.WA, F.d0, F.XX, F.d1), H1p = PAGK_LERP(F.WA1, F.d1, F.XX1, F.d2), H1m = PAGK_LERP(F.WA, F.d_1, F.XX, F.d0);
        const f2 H2 = PAGK_LERP(F.WA, F.n0, F.XX, F.n1);
        const f2 v0 = PAGK_LERP(F.WB, H0, F.YY, H1);
        const f2 vx1 = PAGK_LERP(F.WB, H0p, F.YY, H1p), vx2 = PAGK_LERP(F.WB, H0m, F.YY, H1m);
        const f2 vy1 = PAGK_LERP(F.WB1, H1, F.YY1, H2), vy2 = PAGK_LERP(F.WB, Hm, F.YY, H0);
#undef PAGK_LERP
        // -e = gain * T - (v0 + db) (the reference's e = (v0 + db) - gain * T negated: exact)
        const f2 U = fma2(v0, UN, DB);
        const f2 MF = fma2(mul2(GAIN, tv), ONE, f2{-U.x, -U.y});
        // 2^52 times the gradient: Ix = 0.5 * (vx1 - vx2) is an exact halving, and every sum below that contains Ix or Iy is
        // the reference's sum times a power of two at every step, so the factors are applied once, after the loop
        const f2 GX = sub2(vx1, vx2), GY = sub2(vy1, vy2);
        const f2 M2 = mul2(MF, MF);
        // J = (Ix, Iy, c, 1) as double; b += -J * e; H += J * J^T; cost += e * e (float), reference :264-299.
        // Each product of two float-valued doubles is exact, so DFMA rounds like the separate mul + add.
        // (H[2][2] = sum of c*c does not depend on the samples: it comes with the template record.)
        {
          const double x = (double)GX.x, y = (double)GY.x, mm = (double)MF.x;
          S.h00 = fma(x, x, S.h00); S.h10 = fma(y, x, S.h10); S.h11 = fma(y, y, S.h11);
          S.h20 = fma(c, x, S.h20); S.h21 = fma(c, y, S.h21);
          S.h30 = S.h30 + x; S.h31 = S.h31 + y;
          S.b0 = fma(x, mm, S.b0); S.b1 = fma(y, mm, S.b1); S.b2 = fma(c, mm, S.b2); S.b3 = S.b3 + mm;
          S.cost = S.cost + M2.x;
        }
        if (both) {
          const double x = (double)GX.y, y = (double)GY.y, mm = (double)MF.y;
          S.h00 = fma(x, x, S.h00); S.h10 = fma(y, x, S.h10); S.h11 = fma(y, y, S.h11);
          S.h20 = fma(c, x, S.h20); S.h21 = fma(c, y, S.h21);
          S.h30 = S.h30 + x; S.h31 = S.h31 + y;
          S.b0 = fma(x, mm, S.b0); S.b1 = fma(y, mm, S.b1); S.b2 = fma(c, mm, S.b2); S.b3 = S.b3 + mm;
          S.cost = S.cost + M2.y;
        }
      };
      // one flat walk over the P * P pixels (row-major, the reference's order), two pairs per template vector (L1 / L2; the
      // next trip's vectors are loaded before this trip's pixels are computed); the last pixel's template value is a register
      static_assert((NP - 1) % (4 * PAGK_LANES_TRIP) == 0, "vector loads of the template per trip");
      constexpr int NV = (NP - 1) / 4;  // vectors of four template values
#if !PAGK_LANES_TPRE
      float4 t4[PAGK_LANES_TRIP];
#pragma unroll
      for (int u = 0; u < PAGK_LANES_TRIP; ++u) t4[u] = __ldg(Tg + u);
#endif
#if PAGK_LANES_PIPE
      Front F0, F1;
      front(0, F0);
#pragma unroll 1
      for (int j = 0; j < NV; j += PAGK_LANES_TRIP) {
        float4 nx[PAGK_LANES_TRIP];
#pragma unroll
        for (int u = 0; u < PAGK_LANES_TRIP; ++u) nx[u] = __ldg(Tg + min(j + PAGK_LANES_TRIP + u, NV - 1));
#pragma unroll
        for (int u = 0; u < PAGK_LANES_TRIP; ++u) {
          front(2 * (j + u) + 1, F1);
          back(F0, f2{t4[u].x, t4[u].y}, true);
          front(2 * (j + u) + 2, F0);
          back(F1, f2{t4[u].z, t4[u].w}, true);
        }
#pragma unroll
        for (int u = 0; u < PAGK_LANES_TRIP; ++u) t4[u] = nx[u];
      }
      back(F0, f2{tlast, tlast}, false);
#else
      auto pixels = [&](const int k2, const f2 tv, const bool both) { Front F; front(k2, F); back(F, tv, both); };
#pragma unroll 1
      for (int j = 0; j < NV; j += PAGK_LANES_TRIP) {
        float4 nx[PAGK_LANES_TRIP];
#pragma unroll
        for (int u = 0; u < PAGK_LANES_TRIP; ++u) nx[u] = __ldg(Tg + min(j + PAGK_LANES_TRIP + u, NV - 1));
#pragma unroll
        for (int u = 0; u < PAGK_LANES_TRIP; ++u) {
          pixels(2 * (j + u), f2{t4[u].x, t4[u].y}, true); pixels(2 * (j + u) + 1, f2{t4[u].z, t4[u].w}, true);
        }
#pragma unroll
        for (int u = 0; u < PAGK_LANES_TRIP; ++u) t4[u] = nx[u];
      }
      pixels((NP - 1) / 2, f2{tlast, tlast}, false);
#endif
      constexpr double G1 = 2.220446049250313e-16;  // 2^-52: the samples' 2^51 and the central difference's 2
      S.h00 *= G1 * G1; S.h10 *= G1 * G1; S.h11 *= G1 * G1;
      S.h20 *= G1; S.h21 *= G1; S.h30 *= G1; S.h31 *= G1; S.b0 *= G1; S.b1 *= G1;
      lockbad = active && fast && (badv >= SC);
      coop |= lockbad;
    }
    PROF(3);

    // ------------------------------------------------------------------ the pixel-parallel pass, COOP_BUFS slots at a time
    // For slots whose samples may clamp at the image border, slots that hit the rounding case above, slots whose
    // box does not fit the window (sampled straight from the level) and every live slot of a sparse warp (the tail of a
    // launch, small batches: a pass then takes about a microsecond per slot instead of ten for the lockstep pass).
    // Lanes = pixels: samples -> records (Ix, Iy, -e) in the slot's buffer; then lanes = accumulators: the twelve double
    // sums and the float cost walk the records in pixel order (lanes 0..15 the first buffer, 16..31 the second).
    // The window of a slot lives in ONE bank (its lane's), which lanes = pixels would hit 32 ways, so the slot's window
    // is first copied row-major into the buffer.  A slot without clamps (`fast`) uses the shared interpolations of the
    // lockstep pass (same argument); a pixel that meets the excluded rounding case evaluates its five samples one by
    // one like every pixel of the other slots (PatchMatch::GetPixelValue semantics).
    {
      unsigned m = __ballot_sync(FULL, coop);
      PROF_ADD(13, __popc(m)); PROF_ADD(14, sparse ? 1 : 0); PROF_ADD(15, __popc(__ballot_sync(FULL, restage)));
      constexpr int NB = C::COOP_BUFS, CH = C::COOP_CH;
      while (m) {
        int sl[NB];
        float sbx[NB], sby[NB], s00[NB], s01[NB], s10[NB], s11[NB], sdb[NB], sgain[NB], stl[NB];
        int scols[NB], srows[NB], sp[NB], swx0[NB], swy0[NB], skind[NB];
        const unsigned char *img2[NB];
        const float *Tp[NB];
        __syncwarp();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          sl[b] = m ? __ffs(m) - 1 : -1;
          m &= m - 1;
          const int s = sl[b] < 0 ? 0 : sl[b];
          sbx[b] = __shfl_sync(FULL, bx, s); sby[b] = __shfl_sync(FULL, by, s);
          s00[b] = __shfl_sync(FULL, a00, s); s01[b] = __shfl_sync(FULL, a01, s); s10[b] = __shfl_sync(FULL, a10, s); s11[b] = __shfl_sync(FULL, a11, s);
          sdb[b] = __shfl_sync(FULL, db, s); sgain[b] = __shfl_sync(FULL, gain, s); stl[b] = __shfl_sync(FULL, tlast, s);
          const float scv = __shfl_sync(FULL, cval, s);
          scols[b] = __shfl_sync(FULL, cols, s); srows[b] = __shfl_sync(FULL, rows, s); sp[b] = __shfl_sync(FULL, pitch, s);
          swx0[b] = __shfl_sync(FULL, win_x0, s); swy0[b] = __shfl_sync(FULL, win_y0, s);
          // 0: no clamp can fire (shared interpolations), 1: per-sample from the window, 2: per-sample from the level
          skind[b] = __shfl_sync(FULL, windowable ? (fast && !lockbad ? 0 : 1) : 2, s);
          const int spair = __shfl_sync(FULL, pair, s), slevel = __shfl_sync(FULL, level, s);
          img2[b] = images + (size_t)(spair * 2 + 1) * slot_bytes + s_off[slevel];  // the u8 level
          Tp[b] = reinterpret_cast<const float *>(__shfl_sync(FULL, (unsigned long long)Tg, s));
          if (sl[b] >= 0) {
            float *buf = scratch + b * C::BUF_FLOATS;
            if (lane == 0) { reinterpret_cast<double *>(buf + C::COOP_CONST)[0] = (double)scv; reinterpret_cast<double *>(buf + C::COOP_CONST)[1] = 1.0; }
            if (skind[b] != 2) {  // the slot's window, row-major
              const unsigned int *src = reinterpret_cast<const unsigned int *>(wwin) + s;
              unsigned int *lin = reinterpret_cast<unsigned int *>(buf + C::COOP_LIN);
#pragma unroll
              for (int k = 0; k < (C::WIN_WORDS + 31) / 32; ++k) {
                const int w = lane + 32 * k;
                if (w < C::WIN_WORDS) lin[w] = src[w * 32];
              }
            }
          }
        }
        double acc = 0.0;
        float cacc = 0.f;
#pragma unroll 1
        for (int p0 = 0; p0 < NP; p0 += CH) {
          __syncwarp();
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            if (sl[b] < 0) continue;
            float *buf = scratch + b * C::BUF_FLOATS;
            const unsigned char *lin = reinterpret_cast<const unsigned char *>(buf + C::COOP_LIN);
            const float gc = (float)scols[b], gr = (float)srows[b], gc1 = (float)(scols[b] - 1), gr1 = (float)(srows[b] - 1);
            const int kind = skind[b];
            float tv[CH / 32];
#pragma unroll
            for (int k = 0; k < CH / 32; ++k) {
              const int p = p0 + lane + 32 * k;
              tv[k] = p < NP - 1 ? __ldg(Tp[b] + p) : stl[b];
            }
            // (two pixels per trip; the template value of pixel k picked from registers: indexing tv[] with the trip
            // counter would put it on the stack, and the store would wait for the loads before the first pixel starts)
            static_assert(CH / 32 == 2 || CH / 32 == 4, "template values of a chunk: two or four per lane");
#pragma unroll 1
            for (int kk = 0; kk < CH / 32; kk += 2)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int k = kk + j;
              const float tvk = (CH / 32 == 4 && kk != 0) ? tv[2 + j] : tv[j];
              const int q = lane + 32 * k, p = p0 + q;
              if (p >= NP) continue;
              const int py = p / P, px = p - py * P;
              const float xf = (float)(px - HALF), yf = (float)(py - HALF);
              float wx = xf, wy = yf;
              if (AFFINE) { wx = s00[b] * xf + s01[b] * yf; wy = s10[b] * xf + s11[b] * yf; }
              const float sx = sbx[b] + wx, sy = sby[b] + wy;
              float v0, vx1, vx2, vy1, vy2;
              bool each = kind != 0;
              if (kind == 0) {
                // floor for 0 <= x < 2^22: x + 2^23 rounded DOWN is 2^23 + floor(x) exactly; taking 2^23 off is exact
                const float tx = __fadd_rd(sx, 8388608.0f), ty = __fadd_rd(sy, 8388608.0f);
                const float fx = tx - 8388608.0f, fy = ty - 8388608.0f;
                const float xx = sx - fx, yy = sy - fy;
                const float wa = 1.0f - xx, wb = 1.0f - yy;
                const float X1 = sx + 1.0f, Y1 = sy + 1.0f;
                const float xx1 = X1 - (fx + 1.0f), yy1 = Y1 - (fy + 1.0f);
                const float wa1 = 1.0f - xx1, wb1 = 1.0f - yy1;
                each = fmaxf(xx1, yy1) >= 1.0f;
                const int widx = (__float_as_int(ty) - 0x4B000000 - swy0[b]) * WIN_W + (__float_as_int(tx) - 0x4B000000 - swx0[b]);
                CHECK_IDX(widx, WIN_W + 1, WIN_W * WIN_H - 2 * WIN_W - 3);
                const unsigned char *w = lin + widx;
                const float m0 = u8f(w[-WIN_W]), m1 = u8f(w[-WIN_W + 1]);
                const float c_1 = u8f(w[-1]), c0 = u8f(w[0]), c1 = u8f(w[1]), c2 = u8f(w[2]);
                const float d_1 = u8f(w[WIN_W - 1]), d0 = u8f(w[WIN_W]), d1 = u8f(w[WIN_W + 1]), d2 = u8f(w[WIN_W + 2]);
                const float n0 = u8f(w[2 * WIN_W]), n1 = u8f(w[2 * WIN_W + 1]);
                const float Hm = wa * m0 + xx * m1;
                const float H0 = wa * c0 + xx * c1, H0p = wa1 * c1 + xx1 * c2, H0m = wa * c_1 + xx * c0;
                const float H1 = wa * d0 + xx * d1, H1p = wa1 * d1 + xx1 * d2, H1m = wa * d_1 + xx * d0;
                const float H2 = wa * n0 + xx * n1;
                v0 = wb * H0 + yy * H1;
                vx1 = wb * H0p + yy * H1p; vx2 = wb * H0m + yy * H1m;
                vy1 = wb1 * H1 + yy1 * H2; vy2 = wb * Hm + yy * H0;
              }
              if (each) {
                if (kind != 2) {
                  v0 = window_sample<WIN_W, WIN_H>(lin, swx0[b], swy0[b], gc, gc1, gr, gr1, sx, sy);
                  vx1 = window_sample<WIN_W, WIN_H>(lin, swx0[b], swy0[b], gc, gc1, gr, gr1, sx + 1.0f, sy);
                  vx2 = window_sample<WIN_W, WIN_H>(lin, swx0[b], swy0[b], gc, gc1, gr, gr1, sx - 1.0f, sy);
                  vy1 = window_sample<WIN_W, WIN_H>(lin, swx0[b], swy0[b], gc, gc1, gr, gr1, sx, sy + 1.0f);
                  vy2 = window_sample<WIN_W, WIN_H>(lin, swx0[b], swy0[b], gc, gc1, gr, gr1, sx, sy - 1.0f);
                } else {
                  v0 = pagk_sample_call(img2[b], sp[b], scols[b], srows[b], sx, sy);
                  vx1 = pagk_sample_call(img2[b], sp[b], scols[b], srows[b], sx + 1.0f, sy);
                  vx2 = pagk_sample_call(img2[b], sp[b], scols[b], srows[b], sx - 1.0f, sy);
                  vy1 = pagk_sample_call(img2[b], sp[b], scols[b], srows[b], sx, sy + 1.0f);
                  vy2 = pagk_sample_call(img2[b], sp[b], scols[b], srows[b], sx, sy - 1.0f);
                }
              }
              const float e = (v0 + sdb[b]) - sgain[b] * tvk, me = -e;
              double2 *rec = reinterpret_cast<double2 *>(buf + 8 * q);
              rec[0] = make_double2((double)(0.5f * (vx1 - vx2)), (double)(0.5f * (vy1 - vy2)));
              rec[1] = make_double2((double)me, __hiloint2double(0, __float_as_int(me * me)));
            }
          }
          __syncwarp();
          const int n = min(CH, NP - p0);
          const double *base = reinterpret_cast<const double *>(scratch + cbuf * C::BUF_FLOATS);
          const double *pa = base + offA, *pb = base + offB;
          const float *pm = reinterpret_cast<const float *>(base) + 6;  // e * e: the low word of a record's fourth double
#pragma unroll 8
          for (int q = 0; q < n; ++q) {
            const double da = *pa, db2 = *pb;
            const float m2 = *pm;
            pa += strideA; pb += strideB; pm += 8;
            acc = fma(da, db2, acc);
            cacc = cacc + m2;
          }
        }
        // the sums back to the slots' lanes through the buffers
        {
          float *base = scratch + cbuf * C::BUF_FLOATS;
          double *res = reinterpret_cast<double *>(base + C::COOP_RES);
          if (role16 < 12) res[role16] = acc;
          if (role16 == 0) base[C::COOP_RES + 24] = cacc;
        }
        __syncwarp();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (lane == sl[b]) {
            const float *base = scratch + b * C::BUF_FLOATS;
            const double *res = reinterpret_cast<const double *>(base + C::COOP_RES);
            S.h00 = res[0]; S.h10 = res[1]; S.h11 = res[2]; S.h20 = res[3]; S.h21 = res[4]; S.h30 = res[6]; S.h31 = res[7];
            S.b0 = res[8]; S.b1 = res[9]; S.b2 = res[10]; S.b3 = res[11]; S.cost = base[C::COOP_RES + 24];
          }
        }
      }
      __syncwarp();
    }
    PROF(4);

    // ------------------------------------------------------------------ solve, update, exits (lane = slot)
    if (active) {
      double h00 = S.h00, h10 = S.h10, h11 = S.h11, h20 = S.h20, h21 = S.h21, h22 = h22v, h30 = S.h30, h31 = S.h31;
      double h32 = c * (double)NP, h33 = (double)NP;  // sum of c and of 1 over the patch: exact in double
      double b0 = S.b0, b1 = S.b1, b2 = S.b2, b3 = S.b3;
      float cost = S.cost;
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
      ++n_iter;
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
      if (level_done) {
        const float p2x = ptx + dx, p2y = pty + dy;  // mvPtPyr2Un[i] = pt + (dx, dy)
        if (level == 0) {
          const size_t o = (size_t)feat;
          out.pm_un[o] = make_float2(p2x, p2y);
          out.pm_status[o] = succ ? 1 : 0;
          out.pix_err[o] = sqrt((double)lastCost * mode.win_size_inv);
          out.ncc[o] = 1.0f;
          out.iters[o] = n_iter;
          feat = -1;
        } else if (level == item_lo) {  // split mode: publish this level's result, the lane is free
          handover_store(handover + 4 * (size_t)feat, epoch_base + (top - level + 1), p2x, p2y, n_iter);
          feat = -1;
        } else {
          --level;
          const float scale = 1.0f / (float)(1 << level);
          ptx = pt1x * scale; pty = pt1y * scale;
          dx = p2x * 2.0f - ptx; dy = p2y * 2.0f - pty;
          dg = 0.f; db = 0.f; iter = 0; lastCost = 0.f; succ = 1;
          needs_setup = true; win_valid = false;
        }
      }
    }
    __syncwarp();
    PROF(5);
  }
  PROF_FLUSH();
}

// -------------------------------------------------------------------------------------------------
bool pagk_lk_lanes_supported(const PagkMode &mode) {
  return (mode.half == 5 || mode.half == 10) && mode.iterations >= 1;
}

// the box of the tensor maps the template kernel loads its tap blocks with
void pagk_lk_lanes_tma_box(int half, int *box_w, int *box_h) {
  *box_w = half == 5 ? TmplCfg<5>::ROW_BYTES : TmplCfg<10>::ROW_BYTES;
  *box_h = half == 5 ? TmplCfg<5>::ROWS : TmplCfg<10>::ROWS;
}

size_t pagk_lk_lanes_record_bytes(int half) {
  return half == 5 ? (size_t)LanesCfg<5>::REC_BYTES : half == 10 ? (size_t)LanesCfg<10>::REC_BYTES : 0;
}

template <int HALF, bool AFFINE, int WSM>
static cudaError_t configure_lanes() {
  return cudaFuncSetAttribute(pagk_lk_lanes_kernel<HALF, AFFINE, WSM>, cudaFuncAttributeMaxDynamicSharedMemorySize, LanesCfg<HALF, WSM>::SMEM_BYTES);
}

template <int HALF>
static cudaError_t upload_pairs(const float4 *sym) {
  constexpr int P = 2 * HALF + 1, NP = P * P, N2 = (NP + 1) / 2;
  float4 h[N2];
  for (int k = 0; k < N2; ++k) {
    const int p = 2 * k, q = p + 1 < NP ? p + 1 : p;
    h[k] = make_float4((float)(p % P - HALF), (float)(q % P - HALF), (float)(p / P - HALF), (float)(q / P - HALF));
  }
  return cudaMemcpyToSymbol(*reinterpret_cast<const float4(*)[N2]>(sym), h, sizeof(h));
}

// per device, from pagk_create after cudaSetDevice (function attributes and __constant__ data belong to the device)
int pagk_lk_lanes_configure() {
  cudaError_t e = upload_pairs<5>(c_pair5);
  if (e == cudaSuccess) e = upload_pairs<10>(c_pair10);
  if (e == cudaSuccess) e = configure_lanes<5, true, PAGK_LANES_WARPS5>();
  if (e == cudaSuccess) e = configure_lanes<5, false, PAGK_LANES_WARPS5>();
  if (e == cudaSuccess) e = configure_lanes<5, true, PAGK_LANES_WARPS5_BIG>();
  if (e == cudaSuccess) e = configure_lanes<5, false, PAGK_LANES_WARPS5_BIG>();
  if (e == cudaSuccess) e = configure_lanes<10, true, 7>();
  if (e == cudaSuccess) e = configure_lanes<10, false, 7>();
  return (int)e;
}

template <int HALF, bool AFFINE, int WSM>
static int launch_lanes_w(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                          const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs,
                          int *work_counter, int *next_counter, int *progress, int epoch, int n_sms, unsigned char *tmpl,
                          cudaStream_t st, long long *prof, int share) {
  using C = LanesCfg<HALF, WSM>;
  const long long total = (long long)n_max * n_pairs;
  long long ctas = (long long)n_sms * C::CTAS_SM;  // persistent: the SMs are filled once
  // pagk_set_device_share(h, n): the caller keeps the launches of n handles in flight on this device, so this one takes
  // 1/n of an SM's CTA slots and the other handles' launches run BESIDE it instead of behind it.  Two batches interleaved on
  // every SM are twice the independent work per lane: the tail of a launch (lanes idling while the last long items finish,
  // 20 % of a launch that has the device to itself at 64 pairs) is half as large a share, and it lies under the body of
  // the neighbour.  Config B over three rotating handles: 0.678 -> 0.645 ms per step.  (A launch that turns out to be alone
  // runs on 1/n of the device: the setting is the caller's statement about its own pipeline, not a heuristic.)
  // (only where the CTA slots divide evenly: the 12-warp shape of large batches has three, and a batch that large has the
  // independent work to fill the device by itself)
  const int eff_share = share > C::CTAS_SM ? C::CTAS_SM : share;  // more handles than CTA slots: one slot each
  if (eff_share > 1 && total >= 16384 && C::CTAS_SM % eff_share == 0) ctas = (long long)n_sms * (C::CTAS_SM / eff_share);
  // A small batch is spread over all warps (lane_cap features per warp at a time) instead of filling a few: a warp
  // with a handful of live lanes runs them through the pixel-parallel pass, several times faster per iteration
  // than a lockstep pass, which is what the latency of a single frame pair is made of.
  const long long warps = ctas * C::WARPS;
  int lane_cap = (int)((total + warps - 1) / warps);
  if (lane_cap > C::SLOTS) lane_cap = C::SLOTS;
  if (lane_cap < 1) lane_cap = 1;
  // Level-granular work items when the batch is more than one wave of lanes (the tail of the launch is what they
  // shorten); a batch that fits the lanes keeps a feature in its lane.  PAGK_LK_SPLIT=0|1 forces either.
  static const int forced = [] { const char *e = getenv("PAGK_LK_SPLIT"); return e ? atoi(e) : -1; }();
  int split = (progress != nullptr && mode.levels > 1 && total > warps * C::SLOTS) ? 1 : 0;
  if (forced >= 0 && progress != nullptr && mode.levels > 1) split = forced ? 1 : 0;
  pagk_lk_lanes_kernel<HALF, AFFINE, WSM><<<(unsigned)ctas, C::WARPS * 32, C::SMEM_BYTES, st>>>(
      images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs, work_counter, next_counter, lane_cap, split, progress, epoch * 8,
      tmpl, 1.0f, prof);
  return (int)cudaGetLastError();
}

template <int HALF, bool AFFINE>
static int launch_lanes(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                        const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs,
                        int *work_counter, int *next_counter, int *progress, int epoch, int n_sms, unsigned char *tmpl,
                        const PagkTmaLevels *tmaps, cudaStream_t st, long long *prof, int share) {
  const long long total = (long long)n_max * n_pairs;
  {  // K3a: 32 items per warp
    constexpr int W = TmplCfg<HALF>::WARPS;
    const long long items = total * mode.levels, groups = (items + 31) / 32;
    pagk_lk_template_kernel<HALF><<<(unsigned)((groups + W - 1) / W), W * 32, 0, st>>>(images, g, *tmaps, pcs, keys_un, mode.levels, max_keys,
                                                                                      n_max, n_pairs, tmpl, 1.0f);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  if constexpr (HALF <= 5) {
    // PAGK_LK_WARPS=<n> forces the small (n < 12) or the large number of warps per SM
    static const int forced_w = [] { const char *e = getenv("PAGK_LK_WARPS"); return e ? atoi(e) : 0; }();
    // (a pipeline of three or more handles, pagk_set_device_share: the 12-warp shape, one of its three CTA slots per launch --
    // three batches side by side are the independent work that shape needs: 0.611 ms per config-B step over four handles
    // against 0.635 with two 8-warp launches side by side)
    const bool big = forced_w ? forced_w >= PAGK_LANES_WARPS5_BIG : (total >= PAGK_LANES_BIG_FEATURES || (share >= 3 && total >= 16384));
    if (big)
      return launch_lanes_w<HALF, AFFINE, PAGK_LANES_WARPS5_BIG>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs, work_counter,
                                                                  next_counter, progress, epoch, n_sms, tmpl, st, prof, share);
    return launch_lanes_w<HALF, AFFINE, PAGK_LANES_WARPS5>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs, work_counter,
                                                            next_counter, progress, epoch, n_sms, tmpl, st, prof, share);
  } else
    return launch_lanes_w<HALF, AFFINE, 7>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs, work_counter, next_counter, progress, epoch,
                                         n_sms, tmpl, st, prof, share);
}

int pagk_launch_lk_lanes(const unsigned char *images, const PagkGeom &g, const PagkPairConst *pcs, const float2 *keys_un,
                         const PagkOutPtrs &out, const PagkMode &mode, int max_keys, int n_max, int n_pairs,
                         int *work_counters, int parity, int *progress, int epoch, int n_sms, unsigned char *tmpl,
                         const PagkTmaLevels *tmaps, cudaStream_t st, long long *launches, long long *prof, int share) {
  if (n_max <= 0 || n_pairs <= 0) return 0;
  // work_counters[0..1]: both zero when the handle is created; launch n uses [n & 1] and zeroes the other one
  int *work_counter = work_counters + (parity & 1), *next_counter = work_counters + ((parity + 1) & 1);
  int rc;
  if (mode.half == 5) {
    rc = mode.affine ? launch_lanes<5, true>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs, work_counter, next_counter, progress, epoch, n_sms, tmpl, tmaps, st, prof, share)
                     : launch_lanes<5, false>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs, work_counter, next_counter, progress, epoch, n_sms, tmpl, tmaps, st, prof, share);
  } else {
    rc = mode.affine ? launch_lanes<10, true>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs, work_counter, next_counter, progress, epoch, n_sms, tmpl, tmaps, st, prof, share)
                     : launch_lanes<10, false>(images, g, pcs, keys_un, out, mode, max_keys, n_max, n_pairs, work_counter, next_counter, progress, epoch, n_sms, tmpl, tmaps, st, prof, share);
  }
  *launches += 2;
  return rc;
}
