// ubench_pass2.cu -- the lockstep pass of K3b in isolation: instruction-count diets measured before they go into the
// production kernel.  Every lane owns a u8 window (production layout: 24 x 17, odd word stride) and walks the 121
// pixels of an affine-warped patch; template values stream from global memory (LDG.128).
//   V0  the production pixel (scalar FP32, u8 -> float by LOP3 + FADD)
//   V1  scalar, the conversion folded into an exact FMA:  w * b == fma(w, 2^23 + b, -(w * 2^23))
//   V2  two pixels per step on the packed FP32 pipe (FADD2 / FMUL2 / FFMA2, sm_100), conversion folded as in V1
//   V3  V2 with no conversion at all: the bytes are read as subnormal floats and the coordinates carry 2^100
//   V4  V3 on windows laid out one bank per lane (no bank conflicts whatever the lanes' offsets)
// The variants must agree bit for bit (the sums are printed as a checksum).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o ubench_pass2 ubench_pass2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int HALF = 5, P = 11, NP = 121, WIN_W = 24, WIN_H = 17, WIN_WORDS = WIN_W * WIN_H / 4, WIN_STRIDE = (WIN_WORDS | 1) * 4;
__constant__ float2 c_pix[NP];
__constant__ float4 c_pix2[(NP + 1) / 2];  // (xf_p, xf_q, yf_p, yf_q) of pixels p = 2k, q = 2k + 1

struct Sums {
  double h00, h10, h11, h20, h21, h30, h31, b0, b1, b2, b3;
  float cost;
};

__device__ __forceinline__ float u8f(unsigned int b) { return __uint_as_float(0x4B000000u | b) - 8388608.0f; }
__device__ __forceinline__ float u8b(unsigned int b) { return __uint_as_float(0x4B000000u | b); }

// ---- packed FP32 (sm_100: FADD2 / FMUL2 / FFMA2).  ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even under
// -fmad=false, so a sum whose operand is a product is written fma(a, ONE, b) with ONE = (1, 1) from a kernel
// argument: rn(a * 1 + b) == rn(a + b), and a product feeding the multiplicand of an FMA cannot be contracted.
struct f2 { float x, y; };
#define PK2(a, b) "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};"
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return add2(a, f2{-b.x, -b.y}); }
__device__ __forceinline__ f2 addrd2(f2 a, f2 b) {
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

#define TP(pb, qb, o) f2{__uint_as_float((unsigned int)pb[o]), __uint_as_float((unsigned int)qb[o])}
#define VV(wt, a, wu, b) fma2(mul2(wt, a), ONE, mul2(wu, b))
#ifndef UB_MINCTAS
#define UB_MINCTAS 3
#endif
template <int V>
__global__ void __launch_bounds__(128, UB_MINCTAS) k(const float4 *__restrict__ tmpl, float *out, long long *cycles, int rounds, float one_arg) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // V0..V3: a slot's window is WIN_W * WIN_H consecutive bytes (production layout).  V4: a lane owns a bank -- word w of
  // lane l's window sits at (w * 32 + l) * 4 inside the warp's block, so a lane's load never meets another lane's bank
  unsigned char *win = V >= 4 ? smem + (size_t)warp * (WIN_WORDS * 128) + lane * 4 : smem + (size_t)(warp * 32 + lane) * WIN_STRIDE;
  unsigned int rng = 1234567u + threadIdx.x * 7919u + (blockIdx.x % 148) * 104729u;
  auto rnd = [&]() { rng = rng * 1664525u + 1013904223u; return (rng >> 8) * (1.0f / 16777216.0f); };
  for (int i = 0; i < WIN_W * WIN_H; ++i) win[V >= 4 ? (i >> 2) * 128 + (i & 3) : i] = (unsigned char)(int)(rnd() * 255.0f);
  const int win_x0 = 100 + 4 * (int)(rnd() * 75.0f), win_y0 = 80 + (int)(rnd() * 200.0f);
  float bx = (float)win_x0 + 7.5f + 7.0f * rnd(), by = (float)win_y0 + 7.3f + 1.4f * rnd();  // anywhere the box fits
  const float a00 = 1.0f + 0.02f * (rnd() - 0.5f), a01 = 0.04f * (rnd() - 0.5f), a10 = 0.04f * (rnd() - 0.5f), a11 = 1.0f + 0.02f * (rnd() - 0.5f);
  const float db = rnd(), gain = 1.0f + 0.01f * rnd();
  const double c = -(double)(rnd() * 255.0f);
  const float4 *Tg = tmpl + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) % 4096) * 31;
  const float tlast = 17.25f;
  __syncthreads();
  Sums S;
  S.h00 = S.h10 = S.h11 = S.h20 = S.h21 = S.h30 = S.h31 = S.b0 = S.b1 = S.b2 = S.b3 = 0.0;
  S.cost = 0.f;
  float badv = 0.f;
  const f2 ONE = {one_arg, one_arg};
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const float pbx = bx, pby = by, q00 = a00, q01 = a01, q10 = a10, q11 = a11;
    const unsigned int kk = (unsigned int)(WIN_W + 1) * 0x4B000000u + (unsigned int)(win_y0 * WIN_W + win_x0);
    if (V == 0 || V == 1) {
      auto pixel = [&](const int p, const float tval) {
        const float2 xy = c_pix[p];
        const float xf = xy.x, yf = xy.y;
        const float wx = q00 * xf + q01 * yf, wy = q10 * xf + q11 * yf;
        const float sx = pbx + wx, sy = pby + wy;
        const float tx = __fadd_rd(sx, 8388608.0f), ty = __fadd_rd(sy, 8388608.0f);
        const float fx = tx - 8388608.0f, fy = ty - 8388608.0f;
        const float xx = sx - fx, yy = sy - fy;
        const float wa = 1.0f - xx, wb = 1.0f - yy;
        const float X1 = sx + 1.0f, Y1 = sy + 1.0f;
        const float xx1 = X1 - (fx + 1.0f), yy1 = Y1 - (fy + 1.0f);
        const float wa1 = 1.0f - xx1, wb1 = 1.0f - yy1;
        badv = fmaxf(badv, fmaxf(xx1, yy1));
        const int widx = (int)((unsigned int)__float_as_int(ty) * (unsigned int)WIN_W + (unsigned int)__float_as_int(tx) - kk);
        const unsigned char *w = win + widx;
        float Hm, H0, H0p, H0m, H1, H1p, H1m, H2;
        if (V == 0) {
          const float m0 = u8f(w[-WIN_W]), m1 = u8f(w[-WIN_W + 1]);
          const float c_1 = u8f(w[-1]), c0 = u8f(w[0]), c1 = u8f(w[1]), c2 = u8f(w[2]);
          const float d_1 = u8f(w[WIN_W - 1]), d0 = u8f(w[WIN_W]), d1 = u8f(w[WIN_W + 1]), d2 = u8f(w[WIN_W + 2]);
          const float n0 = u8f(w[2 * WIN_W]), n1 = u8f(w[2 * WIN_W + 1]);
          Hm = wa * m0 + xx * m1;
          H0 = wa * c0 + xx * c1; H0p = wa1 * c1 + xx1 * c2; H0m = wa * c_1 + xx * c0;
          H1 = wa * d0 + xx * d1; H1p = wa1 * d1 + xx1 * d2; H1m = wa * d_1 + xx * d0;
          H2 = wa * n0 + xx * n1;
        } else {
          const float m0 = u8b(w[-WIN_W]), m1 = u8b(w[-WIN_W + 1]);
          const float c_1 = u8b(w[-1]), c0 = u8b(w[0]), c1 = u8b(w[1]), c2 = u8b(w[2]);
          const float d_1 = u8b(w[WIN_W - 1]), d0 = u8b(w[WIN_W]), d1 = u8b(w[WIN_W + 1]), d2 = u8b(w[WIN_W + 2]);
          const float n0 = u8b(w[2 * WIN_W]), n1 = u8b(w[2 * WIN_W + 1]);
          const float na = wa * -8388608.0f, nx = xx * -8388608.0f, na1 = wa1 * -8388608.0f, nx1 = xx1 * -8388608.0f;
#define PR(wt, nw, tap) __fmaf_rn(wt, tap, nw)
          Hm = PR(wa, na, m0) + PR(xx, nx, m1);
          H0 = PR(wa, na, c0) + PR(xx, nx, c1); H0p = PR(wa1, na1, c1) + PR(xx1, nx1, c2); H0m = PR(wa, na, c_1) + PR(xx, nx, c0);
          H1 = PR(wa, na, d0) + PR(xx, nx, d1); H1p = PR(wa1, na1, d1) + PR(xx1, nx1, d2); H1m = PR(wa, na, d_1) + PR(xx, nx, d0);
          H2 = PR(wa, na, n0) + PR(xx, nx, n1);
        }
        const float v0 = wb * H0 + yy * H1;
        const float vx1 = wb * H0p + yy * H1p, vx2 = wb * H0m + yy * H1m;
        const float vy1 = wb1 * H1 + yy1 * H2, vy2 = wb * Hm + yy * H0;
        const float e = (v0 + db) - gain * tval;
        const float gxf = vx1 - vx2, gyf = vy1 - vy2, mf = -e;
        const double x = (double)gxf, y = (double)gyf, mm = (double)mf;
        S.h00 = fma(x, x, S.h00); S.h10 = fma(y, x, S.h10); S.h11 = fma(y, y, S.h11);
        S.h20 = fma(c, x, S.h20); S.h21 = fma(c, y, S.h21);
        S.h30 = S.h30 + x; S.h31 = S.h31 + y;
        S.b0 = fma(x, mm, S.b0); S.b1 = fma(y, mm, S.b1); S.b2 = fma(c, mm, S.b2); S.b3 = S.b3 + mm;
        S.cost = S.cost + mf * mf;
      };
      float4 t4 = __ldg(Tg);
#pragma unroll 1
      for (int j = 0; j < (NP - 1) / 4; ++j) {
        const float4 nx = __ldg(Tg + min(j + 1, (NP - 1) / 4 - 1));
        pixel(4 * j, t4.x); pixel(4 * j + 1, t4.y); pixel(4 * j + 2, t4.z); pixel(4 * j + 3, t4.w);
        t4 = nx;
      }
      pixel(NP - 1, tlast);
    } else if (V == 2) {
      // ---- two pixels per step, every FP32 operation packed across the pair (.x = pixel p, .y = pixel q = p + 1)
      const f2 PBX = {pbx, pbx}, PBY = {pby, pby}, Q00 = {q00, q00}, Q01 = {q01, q01}, Q10 = {q10, q10}, Q11 = {q11, q11};
      const f2 BIG = {8388608.0f, 8388608.0f}, NBIG = {-8388608.0f, -8388608.0f}, P1 = {1.0f, 1.0f};
      const f2 DB = {db, db}, GAIN = {gain, gain};
      const unsigned char *wbase = win - kk;  // byte address arithmetic modulo 2^32 on the low word is what the index needs
      auto pair = [&](const int k2, const f2 tv, const bool both) {
        const float4 xy = c_pix2[k2];
        const f2 XF = {xy.x, xy.y}, YF = {xy.z, xy.w};
        const f2 WX = fma2(mul2(Q00, XF), ONE, mul2(Q01, YF)), WY = fma2(mul2(Q10, XF), ONE, mul2(Q11, YF));
        const f2 SX = add2(PBX, WX), SY = add2(PBY, WY);
        const f2 TX = addrd2(SX, BIG), TY = addrd2(SY, BIG);
        const f2 FX = add2(TX, NBIG), FY = add2(TY, NBIG);
        const f2 XX = sub2(SX, FX), YY = sub2(SY, FY);
        const f2 WA = sub2(P1, XX), WB = sub2(P1, YY);
        const f2 X1 = add2(SX, P1), Y1 = add2(SY, P1);
        const f2 XX1 = sub2(X1, add2(FX, P1)), YY1 = sub2(Y1, add2(FY, P1));
        const f2 WA1 = sub2(P1, XX1), WB1 = sub2(P1, YY1);
        badv = fmaxf(badv, fmaxf(XX1.x, YY1.x));
        badv = fmaxf(badv, fmaxf(XX1.y, YY1.y));
        const unsigned int ip = (unsigned int)__float_as_int(TY.x) * (unsigned int)WIN_W + (unsigned int)__float_as_int(TX.x) - kk;
        const unsigned int iq = (unsigned int)__float_as_int(TY.y) * (unsigned int)WIN_W + (unsigned int)__float_as_int(TX.y) - kk;
        const unsigned char *wp = win + (int)ip, *wq = win + (int)iq;
        (void)wbase;
#define TAP(o) f2{u8b(wp[o]), u8b(wq[o])}
        const f2 m0 = TAP(-WIN_W), m1 = TAP(-WIN_W + 1);
        const f2 c_1 = TAP(-1), c0 = TAP(0), c1 = TAP(1), c2 = TAP(2);
        const f2 d_1 = TAP(WIN_W - 1), d0 = TAP(WIN_W), d1 = TAP(WIN_W + 1), d2 = TAP(WIN_W + 2);
        const f2 n0 = TAP(2 * WIN_W), n1 = TAP(2 * WIN_W + 1);
        const f2 NA = mul2(WA, NBIG), NX = mul2(XX, NBIG), NA1 = mul2(WA1, NBIG), NX1 = mul2(XX1, NBIG);
#define HH(wl, nl, tl, wr, nr, tr) add2(fma2(wl, tl, nl), fma2(wr, tr, nr))
        const f2 Hm = HH(WA, NA, m0, XX, NX, m1);
        const f2 H0 = HH(WA, NA, c0, XX, NX, c1), H0p = HH(WA1, NA1, c1, XX1, NX1, c2), H0m = HH(WA, NA, c_1, XX, NX, c0);
        const f2 H1 = HH(WA, NA, d0, XX, NX, d1), H1p = HH(WA1, NA1, d1, XX1, NX1, d2), H1m = HH(WA, NA, d_1, XX, NX, d0);
        const f2 H2 = HH(WA, NA, n0, XX, NX, n1);
#define VV(wt, a, wu, b) fma2(mul2(wt, a), ONE, mul2(wu, b))
        const f2 v0 = VV(WB, H0, YY, H1);
        const f2 vx1 = VV(WB, H0p, YY, H1p), vx2 = VV(WB, H0m, YY, H1m);
        const f2 vy1 = VV(WB1, H1, YY1, H2), vy2 = VV(WB, Hm, YY, H0);
        // mf = -e = gain * T - (v0 + db)
        const f2 U = add2(v0, DB);
        const f2 MF = fma2(mul2(GAIN, tv), ONE, f2{-U.x, -U.y});
        const f2 GX = sub2(vx1, vx2), GY = sub2(vy1, vy2);
        const f2 M2 = mul2(MF, MF);
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
      float4 t4 = __ldg(Tg);
#pragma unroll 1
      for (int j = 0; j < (NP - 1) / 4; ++j) {
        const float4 nx = __ldg(Tg + min(j + 1, (NP - 1) / 4 - 1));
        pair(2 * j, f2{t4.x, t4.y}, true); pair(2 * j + 1, f2{t4.z, t4.w}, true);
        t4 = nx;
      }
      pair((NP - 1) / 2, f2{tlast, tlast}, false);
    } else if (V == 5) {
      // ---- V5: V4 software-pipelined by hand: the coordinates, weights and tap loads of pair k + 1 are issued before the
      // interpolations and sums of pair k (a pair's front half and the previous pair's back half share a loop body)
      constexpr float SC = 1.2676506002282294e30f, BIGS = 8388608.0f * SC, UNS = 4.440892098500626e-16f;
      const f2 PBX = {pbx * SC, pbx * SC}, PBY = {pby * SC, pby * SC};
      const f2 Q00 = {q00 * SC, q00 * SC}, Q01 = {q01 * SC, q01 * SC}, Q10 = {q10 * SC, q10 * SC}, Q11 = {q11 * SC, q11 * SC};
      const f2 BIG = {BIGS, BIGS}, NBIG = {-BIGS, -BIGS}, P1 = {SC, SC}, UN = {UNS, UNS};
      const f2 DB = {db, db}, GAIN = {gain, gain};
      const unsigned int kks = (unsigned int)(WIN_W + 1) * 0x7D000000u + (unsigned int)(win_y0 * WIN_W + win_x0);
      struct Front { f2 WA, XX, WB, YY, WA1, XX1, WB1, YY1, m0, m1, c_1, c0, c1, c2, d_1, d0, d1, d2, n0, n1; };
      auto front = [&](const int k2, Front &F) {
        const float4 xy = c_pix2[k2];
        const f2 XF = {xy.x, xy.y}, YF = {xy.z, xy.w};
        const f2 WX = fma2(mul2(Q00, XF), ONE, mul2(Q01, YF)), WY = fma2(mul2(Q10, XF), ONE, mul2(Q11, YF));
        const f2 SX = add2(PBX, WX), SY = add2(PBY, WY);
        const f2 TX = addrd2(SX, BIG), TY = addrd2(SY, BIG);
        const f2 FX = add2(TX, NBIG), FY = add2(TY, NBIG);
        F.XX = sub2(SX, FX); F.YY = sub2(SY, FY);
        F.WA = sub2(P1, F.XX); F.WB = sub2(P1, F.YY);
        const f2 X1 = add2(SX, P1), Y1 = add2(SY, P1);
        F.XX1 = sub2(X1, add2(FX, P1)); F.YY1 = sub2(Y1, add2(FY, P1));
        F.WA1 = sub2(P1, F.XX1); F.WB1 = sub2(P1, F.YY1);
        badv = fmaxf(badv, fmaxf(F.XX1.x, F.YY1.x));
        badv = fmaxf(badv, fmaxf(F.XX1.y, F.YY1.y));
        const unsigned int ip = (unsigned int)__float_as_int(TY.x) * (unsigned int)WIN_W + (unsigned int)__float_as_int(TX.x) - kks;
        const unsigned int iq = (unsigned int)__float_as_int(TY.y) * (unsigned int)WIN_W + (unsigned int)__float_as_int(TX.y) - kks;
        auto bases = [&](unsigned int i, const unsigned char *&b0, const unsigned char *&b1, const unsigned char *&b2, const unsigned char *&b3) {
          const unsigned int i1 = i - 1u, s = i1 & 3u;
          b0 = win + (i1 >> 2) * 124u + i1;
          b1 = b0 + ((s + 1u) & 4u) * 31u; b2 = b0 + ((s + 2u) & 4u) * 31u; b3 = b0 + ((s + 3u) & 4u) * 31u;
        };
        const unsigned char *p0, *p1, *p2, *p3, *q0, *q1, *q2, *q3;
        bases(ip, p0, p1, p2, p3); bases(iq, q0, q1, q2, q3);
        constexpr int R = WIN_W / 4 * 128;
        F.m0 = TP(p1, q1, 1 - R); F.m1 = TP(p2, q2, 2 - R);
        F.c_1 = TP(p0, q0, 0); F.c0 = TP(p1, q1, 1); F.c1 = TP(p2, q2, 2); F.c2 = TP(p3, q3, 3);
        F.d_1 = TP(p0, q0, R); F.d0 = TP(p1, q1, 1 + R); F.d1 = TP(p2, q2, 2 + R); F.d2 = TP(p3, q3, 3 + R);
        F.n0 = TP(p1, q1, 1 + 2 * R); F.n1 = TP(p2, q2, 2 + 2 * R);
      };
      auto back = [&](const Front &F, const f2 tv, const bool both) {
        const f2 Hm = VV(F.WA, F.m0, F.XX, F.m1);
        const f2 H0 = VV(F.WA, F.c0, F.XX, F.c1), H0p = VV(F.WA1, F.c1, F.XX1, F.c2), H0m = VV(F.WA, F.c_1, F.XX, F.c0);
        const f2 H1 = VV(F.WA, F.d0, F.XX, F.d1), H1p = VV(F.WA1, F.d1, F.XX1, F.d2), H1m = VV(F.WA, F.d_1, F.XX, F.d0);
        const f2 H2 = VV(F.WA, F.n0, F.XX, F.n1);
        const f2 v0 = VV(F.WB, H0, F.YY, H1);
        const f2 vx1 = VV(F.WB, H0p, F.YY, H1p), vx2 = VV(F.WB, H0m, F.YY, H1m);
        const f2 vy1 = VV(F.WB1, H1, F.YY1, H2), vy2 = VV(F.WB, Hm, F.YY, H0);
        const f2 U = fma2(v0, UN, DB);
        const f2 MF = fma2(mul2(GAIN, tv), ONE, f2{-U.x, -U.y});
        const f2 GX = sub2(vx1, vx2), GY = sub2(vy1, vy2);
        const f2 M2 = mul2(MF, MF);
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
      float4 t4 = __ldg(Tg);
      Front F0, F1;
      front(0, F0);
#pragma unroll 1
      for (int j = 0; j < (NP - 1) / 4; ++j) {
        const float4 nx = __ldg(Tg + min(j + 1, (NP - 1) / 4 - 1));
        front(2 * j + 1, F1);
        back(F0, f2{t4.x, t4.y}, true);
        front(2 * j + 2, F0);
        back(F1, f2{t4.z, t4.w}, true);
        t4 = nx;
      }
      back(F0, f2{tlast, tlast}, false);
    } else {
      // ---- V3 / V4: as V2, and the taps enter the arithmetic as the raw bytes read as (subnormal) floats, b * 2^-149: the
      // coordinate chain is scaled by 2^100 (every operation on it commutes with a power-of-two scale: nothing leaves the
      // normal range), so a horizontal interpolation is the reference's times 2^-49 and a sample the reference's times 2^51
      constexpr float SC = 1.2676506002282294e30f;             // 2^100
      constexpr float BIGS = 8388608.0f * SC;                  // 2^123
      constexpr float UNS = 4.440892098500626e-16f;            // 2^-51
      const f2 PBX = {pbx * SC, pbx * SC}, PBY = {pby * SC, pby * SC};
      const f2 Q00 = {q00 * SC, q00 * SC}, Q01 = {q01 * SC, q01 * SC}, Q10 = {q10 * SC, q10 * SC}, Q11 = {q11 * SC, q11 * SC};
      const f2 BIG = {BIGS, BIGS}, NBIG = {-BIGS, -BIGS}, P1 = {SC, SC}, UN = {UNS, UNS};
      const f2 DB = {db, db}, GAIN = {gain, gain};
      const unsigned int kks = (unsigned int)(WIN_W + 1) * 0x7D000000u + (unsigned int)(win_y0 * WIN_W + win_x0);
      auto pair = [&](const int k2, const f2 tv, const bool both) {
        const float4 xy = c_pix2[k2];
        const f2 XF = {xy.x, xy.y}, YF = {xy.z, xy.w};
        const f2 WX = fma2(mul2(Q00, XF), ONE, mul2(Q01, YF)), WY = fma2(mul2(Q10, XF), ONE, mul2(Q11, YF));
        const f2 SX = add2(PBX, WX), SY = add2(PBY, WY);
        const f2 TX = addrd2(SX, BIG), TY = addrd2(SY, BIG);
        const f2 FX = add2(TX, NBIG), FY = add2(TY, NBIG);
        const f2 XX = sub2(SX, FX), YY = sub2(SY, FY);
        const f2 WA = sub2(P1, XX), WB = sub2(P1, YY);
        const f2 X1 = add2(SX, P1), Y1 = add2(SY, P1);
        const f2 XX1 = sub2(X1, add2(FX, P1)), YY1 = sub2(Y1, add2(FY, P1));
        const f2 WA1 = sub2(P1, XX1), WB1 = sub2(P1, YY1);
        badv = fmaxf(badv, fmaxf(XX1.x, YY1.x));
        badv = fmaxf(badv, fmaxf(XX1.y, YY1.y));
        const unsigned int ip = (unsigned int)__float_as_int(TY.x) * (unsigned int)WIN_W + (unsigned int)__float_as_int(TX.x) - kks;
        const unsigned int iq = (unsigned int)__float_as_int(TY.y) * (unsigned int)WIN_W + (unsigned int)__float_as_int(TX.y) - kks;
        f2 m0, m1, c_1, c0, c1, c2, d_1, d0, d1, d2, n0, n1;
        if (V == 3) {
          const unsigned char *wp = win + (int)ip, *wq = win + (int)iq;
#undef TAP
#define TAP(o) f2{__uint_as_float((unsigned int)wp[o]), __uint_as_float((unsigned int)wq[o])}
          m0 = TAP(-WIN_W); m1 = TAP(-WIN_W + 1);
          c_1 = TAP(-1); c0 = TAP(0); c1 = TAP(1); c2 = TAP(2);
          d_1 = TAP(WIN_W - 1); d0 = TAP(WIN_W); d1 = TAP(WIN_W + 1); d2 = TAP(WIN_W + 2);
          n0 = TAP(2 * WIN_W); n1 = TAP(2 * WIN_W + 1);
        } else {
          // byte i of the lane's window sits at (i >> 2) * 128 + (i & 3): the four taps of a row start at i1 = i - 1, tap j
          // at A0 + j + 124 * ((s + j) >> 2) with s = i1 & 3; rows are WIN_W / 4 words = 768 bytes apart
          auto bases = [&](unsigned int i, const unsigned char *&b0, const unsigned char *&b1, const unsigned char *&b2, const unsigned char *&b3) {
            const unsigned int i1 = i - 1u, s = i1 & 3u;
            b0 = win + (i1 >> 2) * 124u + i1;
            b1 = b0 + ((s + 1u) & 4u) * 31u; b2 = b0 + ((s + 2u) & 4u) * 31u; b3 = b0 + ((s + 3u) & 4u) * 31u;
          };
          const unsigned char *p0, *p1, *p2, *p3, *q0, *q1, *q2, *q3;
          bases(ip, p0, p1, p2, p3); bases(iq, q0, q1, q2, q3);
          constexpr int R = WIN_W / 4 * 128;
          m0 = TP(p1, q1, 1 - R); m1 = TP(p2, q2, 2 - R);
          c_1 = TP(p0, q0, 0); c0 = TP(p1, q1, 1); c1 = TP(p2, q2, 2); c2 = TP(p3, q3, 3);
          d_1 = TP(p0, q0, R); d0 = TP(p1, q1, 1 + R); d1 = TP(p2, q2, 2 + R); d2 = TP(p3, q3, 3 + R);
          n0 = TP(p1, q1, 1 + 2 * R); n1 = TP(p2, q2, 2 + 2 * R);
        }
        const f2 Hm = VV(WA, m0, XX, m1);
        const f2 H0 = VV(WA, c0, XX, c1), H0p = VV(WA1, c1, XX1, c2), H0m = VV(WA, c_1, XX, c0);
        const f2 H1 = VV(WA, d0, XX, d1), H1p = VV(WA1, d1, XX1, d2), H1m = VV(WA, d_1, XX, d0);
        const f2 H2 = VV(WA, n0, XX, n1);
        const f2 v0 = VV(WB, H0, YY, H1);
        const f2 vx1 = VV(WB, H0p, YY, H1p), vx2 = VV(WB, H0m, YY, H1m);
        const f2 vy1 = VV(WB1, H1, YY1, H2), vy2 = VV(WB, Hm, YY, H0);
        // mf = -e = gain * T - (v0 + db), v0 = v0s * 2^-51 (exact), so v0 + db is one FMA
        const f2 U = fma2(v0, UN, DB);
        const f2 MF = fma2(mul2(GAIN, tv), ONE, f2{-U.x, -U.y});
        const f2 GX = sub2(vx1, vx2), GY = sub2(vy1, vy2);
        const f2 M2 = mul2(MF, MF);
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
      float4 t4 = __ldg(Tg);
#pragma unroll 1
      for (int j = 0; j < (NP - 1) / 4; ++j) {
        const float4 nx = __ldg(Tg + min(j + 1, (NP - 1) / 4 - 1));
        pair(2 * j, f2{t4.x, t4.y}, true); pair(2 * j + 1, f2{t4.z, t4.w}, true);
        t4 = nx;
      }
      pair((NP - 1) / 2, f2{tlast, tlast}, false);
    }
    bx += 0.001f; by -= 0.001f;
  }
  const long long t1 = clock64();
  if (V >= 3) {  // gradients carry 2^51: the sums 2^102 / 2^51
    const double s1 = 4.440892098500626e-16, s2 = s1 * s1;
    S.h00 *= s2; S.h10 *= s2; S.h11 *= s2; S.h20 *= s1; S.h21 *= s1; S.h30 *= s1; S.h31 *= s1; S.b0 *= s1; S.b1 *= s1;
    badv = badv >= 1.2676506002282294e30f ? 1.0f : badv * 7.888609052210118e-31f;
  }
  S.h00 *= 0.25; S.h10 *= 0.25; S.h11 *= 0.25;
  out[blockIdx.x * blockDim.x + threadIdx.x] =
      (float)(S.h00 + S.h10 + S.h11 + S.h20 + S.h21 + S.h30 + S.h31 + S.b0 + S.b1 + S.b2 + S.b3) + S.cost + badv;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int V>
void run(const char *name, int ctas_sm, const float4 *tmpl, double *checksum) {
  float *out; long long *cyc, h;
  const int threads = 128, grid = 148 * ctas_sm;
  cudaMalloc(&out, (size_t)grid * threads * 4); cudaMalloc(&cyc, 8);
  const int rounds = 40;
  const size_t smem = V >= 4 ? (size_t)(threads / 32) * WIN_WORDS * 128 : (size_t)threads * WIN_STRIDE;
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V><<<grid, threads, smem>>>(tmpl, out, cyc, rounds, 1.0f);
  cudaEventRecord(e0);
  k<V><<<grid, threads, smem>>>(tmpl, out, cyc, rounds, 1.0f);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  static float host[148 * 4 * 128];
  cudaMemcpy(host, out, (size_t)148 * threads * 4, cudaMemcpyDeviceToHost);  // the first 148 CTAs: the same data in every configuration
  double cs = 0; for (int i = 0; i < 148 * threads; ++i) cs += (double)host[i] * (1 + i % 7);
  *checksum = cs;
  const double cyc_round = (double)h / rounds;
  const int warps = 4 * ctas_sm;
  printf("%-4s warps/SM %2d: %8.0f cycles/pass/warp = %6.1f cyc/pixel/warp, %6.1f cycles per pixel-round per scheduler pair... SM rate %.4f slot-passes/cycle (%.3f ms) checksum %.17g %s\n",
         name, warps, cyc_round, cyc_round / NP, cyc_round / NP, warps * 32 / cyc_round, ms, cs, err == cudaSuccess ? "" : cudaGetErrorString(err));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  float2 h[NP]; float4 h2[(NP + 1) / 2];
  for (int p = 0; p < NP; ++p) h[p] = make_float2((float)(p % P - HALF), (float)(p / P - HALF));
  for (int k2 = 0; k2 < (NP + 1) / 2; ++k2) {
    const int p = 2 * k2, q = p + 1 < NP ? p + 1 : p;
    h2[k2] = make_float4(h[p].x, h[q].x, h[p].y, h[q].y);
  }
  cudaMemcpyToSymbol(c_pix, h, sizeof(h)); cudaMemcpyToSymbol(c_pix2, h2, sizeof(h2));
  float4 *tmpl; cudaMalloc(&tmpl, (size_t)4096 * 31 * 16);
  {
    static float ht[4096 * 31 * 4];
    unsigned int rng = 99u;
    for (size_t i = 0; i < sizeof(ht) / 4; ++i) { rng = rng * 1664525u + 1013904223u; ht[i] = (rng >> 8) * (255.0f / 16777216.0f); }
    cudaMemcpy(tmpl, ht, sizeof(ht), cudaMemcpyHostToDevice);
  }
  double c0, c4, c5;
  for (int ctas = 1; ctas <= UB_MINCTAS; ++ctas) {
    run<0>("V0", ctas, tmpl, &c0); run<4>("V4", ctas, tmpl, &c4); run<5>("V5", ctas, tmpl, &c5);
    printf("  bit-equal: V4 %s, V5 %s\n", c0 == c4 ? "yes" : "NO", c0 == c5 ? "yes" : "NO");
  }
  return 0;
}
