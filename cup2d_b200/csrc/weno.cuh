// WENO5 line core shared by the advect kernels (uniform tiles: advect.cu; multi-level blocks: amr_fast.cu).
//
// Upwind WENO5 differences of both velocity components along one line of 8 cells (window of 14 values per component),
// written in first/second differences with shared fluxes, smoothness indicators scaled by 4 and ONE reciprocal per
// flux — see the header of advect.cu for the algebra (reference: weno5_plus/minus/derivative, main.cpp:162-208).
#pragma once
#include "common.cuh"

namespace cup2d {

// WENO constants live in constant memory so that they are DFMA constant-bank operands; as literals the
// compiler materialises each with a pair of UMOVs (36 UMOV per cell in profiles/r01c).
enum { K_G = 0, K_EPS, K_D1, K_D2, K_D3,            // 13/3, 4e-6, den weights .1 .6 .3
       K_PA1, K_PB1, K_PA2, K_PB2, K_PA3, K_PB3,    // plus-flux phi coefficients (gamma folded in)
       K_MA1, K_MB1, K_MA2, K_MB2, K_MA3, K_MB3, K_N };
static __constant__ double cW[K_N];
static const double hW[K_N] = {13.0 / 3.0, 4e-6, 0.1, 0.6, 0.3,
                               0.1 * 5.0 / 6.0, -0.1 / 3.0, 0.6 / 6.0, 0.6 / 3.0, 0.3 * 2.0 / 3.0, -0.3 / 6.0,
                               -0.3 * 2.0 / 3.0, 0.3 / 6.0, -0.6 / 3.0, -0.6 / 6.0, -0.1 * 5.0 / 6.0, 0.1 / 3.0};

struct LineState {
  double dm2, dm1, d0, dp1; // D[w-2..w+1]
  double Gm1, G0, Gp1;      // 13/3 D2^2 + 4e-6 at w-1, w, w+1
  double qlast;             // q[w+2]
  double rP1, rP2, rM1;     // ratioP(w-1), ratioP(w-2), ratioM(w-1)
};

__device__ __forceinline__ double Gfun(double D2) {
  return fma(cW[K_G] * D2, D2, cW[K_EPS]);
}
__device__ __forceinline__ void line_init(LineState &s, const double *q, int es) {
  double q0 = q[0], q1 = q[es], q2 = q[2 * es], q3 = q[3 * es], q4 = q[4 * es];
  s.dm2 = q1 - q0; // w = 2: D[0]
  s.dm1 = q2 - q1; // D[1]
  s.d0 = q3 - q2;  // D[2]
  s.dp1 = q4 - q3; // D[3]
  s.Gm1 = Gfun(s.dm1 - s.dm2);
  s.G0 = Gfun(s.d0 - s.dm1);
  s.Gp1 = Gfun(s.dp1 - s.d0);
  s.qlast = q4;
  s.rP1 = s.rP2 = s.rM1 = 0.0;
}
__device__ __forceinline__ void line_betas(const LineState &s, double &s1, double &s2, double &s3) {
  const double e1 = fma(3.0, s.dm1, -s.dm2);
  const double e2 = s.dm1 + s.d0;
  const double e3 = fma(3.0, s.d0, -s.dp1);
  const double B1 = fma(e1, e1, s.Gm1);
  const double B2 = fma(e2, e2, s.G0);
  const double B3 = fma(e3, e3, s.Gp1);
  const double q1 = B2 * B3, q2 = B1 * B3, q3 = B1 * B2;
  s1 = q1 * q1;
  s2 = q2 * q2;
  s3 = q3 * q3;
}
// Two arithmetic variants prepared for measurement, OFF by default (the validated kernels are built without them):
//   CUP2D_WENO_CUBIC_RCP   one cubic step r0 (1 + e + e^2) instead of two Newton steps: 3 FP64 operations instead of 4 per
//                          reciprocal, error (seed error)^3 ~ 2^-63
//   CUP2D_WENO_LAZY_BETAS  no smoothness indicators at a window position where neither flux family is needed (the first
//                          position when the flow there is not positive, the last when it is)
#ifndef CUP2D_WENO_CUBIC_RCP
#define CUP2D_WENO_CUBIC_RCP 0
#endif
#ifndef CUP2D_WENO_LAZY_BETAS
#define CUP2D_WENO_LAZY_BETAS 0
#endif

// 1/x for x > 0, normal: MUFU.RCP64H seed (~2^-20) + two Newton steps = full double accuracy.  (One step
// leaves 1.3e-13 relative error for 3 % of the kernel time, profiles/r01h; not worth it.)
__device__ __forceinline__ double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#if CUP2D_WENO_CUBIC_RCP
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
#else
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
  }
  return r;
#endif
}
// upwind-from-the-left flux ratio at face w+1/2 (weno5_plus, main.cpp:162-181; gammas .1,.6,.3)
__device__ __forceinline__ double ratio_plus(const LineState &s, double s1, double s2, double s3) {
  const double den = fma(cW[K_D1], s1, fma(cW[K_D3], s3, cW[K_D2] * s2));
  const double p1 = fma(cW[K_PA1], s.dm1, cW[K_PB1] * s.dm2);
  const double p2 = fma(cW[K_PA2], s.dm1, cW[K_PB2] * s.d0);
  const double p3 = fma(cW[K_PA3], s.d0, cW[K_PB3] * s.dp1);
  const double num = fma(s1, p1, fma(s3, p3, s2 * p2));
  return num * rcp_pos(den);
}
// upwind-from-the-right flux ratio at face w-1/2 (weno5_minus, main.cpp:182-201; gammas .3,.6,.1)
__device__ __forceinline__ double ratio_minus(const LineState &s, double s1, double s2, double s3) {
  const double den = fma(cW[K_D3], s1, fma(cW[K_D1], s3, cW[K_D2] * s2));
  const double p1 = fma(cW[K_MA1], s.dm1, cW[K_MB1] * s.dm2);
  const double p2 = fma(cW[K_MA2], s.dm1, cW[K_MB2] * s.d0);
  const double p3 = fma(cW[K_MA3], s.d0, cW[K_MB3] * s.dp1);
  const double num = fma(s1, p1, fma(s3, p3, s2 * p2));
  return num * rcp_pos(den);
}
__device__ __forceinline__ void line_advance(LineState &s, double qn, double rP, double rM) {
  s.rP2 = s.rP1;
  s.rP1 = rP;
  s.rM1 = rM;
  s.dm2 = s.dm1;
  s.dm1 = s.d0;
  s.d0 = s.dp1;
  s.dp1 = qn - s.qlast;
  s.qlast = qn;
  s.Gm1 = s.G0;
  s.G0 = s.Gp1;
  s.Gp1 = Gfun(s.dp1 - s.d0);
}

// Upwind WENO5 differences of both components along one line of 8 cells (window of 14 values per
// component, element stride es).  qa = advecting component (sign + multiplier), qb = the other one.
// emit(c, Ua, Ub, da, db, D2a, D2b) is called once per cell c = 0..7 with the cell values, the undivided
// differences (reference `derivative`, main.cpp:202-208) and the second differences (diffusion term).
template <class Emit>
__device__ __forceinline__ void weno_line(const double *__restrict__ qa, const double *__restrict__ qb,
                                          const int es, Emit emit) {
  LineState A, B;
  line_init(A, qa, es);
  line_init(B, qb, es);
  // sign of the advecting velocity at window indices 2..12 (bit k <-> index k), one pass, no FP64 pipe
  unsigned pos = 0;
#pragma unroll
  for (int k = 2; k <= 12; k++) pos |= is_pos(qa[k * es]) ? (1u << k) : 0u;
  double Ubm1 = qb[2 * es]; // qb at window index w-1 (cell value of the other component)
  double Uam1 = qa[2 * es];
#pragma unroll
  for (int w = 2; w <= 11; ++w) {
    const bool vc = (unsigned)(w - 3) < 8u, vn = (unsigned)(w - 2) < 8u, vp = (unsigned)(w - 4) < 8u;
    const unsigned pw = pos >> (w - 1); // bit 0: cell w-1, bit 1: cell w, bit 2: cell w+1
    const bool posp = pw & 1u;
    // flux families needed at this window position (masks are compile-time after unrolling)
    const bool needP = (pw & ((vc ? 2u : 0u) | (vn ? 4u : 0u))) != 0u;
    const bool needM = (~pw & ((vc ? 2u : 0u) | (vp ? 1u : 0u))) != 0u;
    double a1 = 0, a2 = 0, a3 = 0, b1 = 0, b2 = 0, b3 = 0;
    if (!CUP2D_WENO_LAZY_BETAS || needP || needM) {
      line_betas(A, a1, a2, a3);
      line_betas(B, b1, b2, b3);
    }
    double rPa = 0, rPb = 0, rMa = 0, rMb = 0;
    if (needP) {
      rPa = ratio_plus(A, a1, a2, a3);
      rPb = ratio_plus(B, b1, b2, b3);
    }
    if (needM) {
      rMa = ratio_minus(A, a1, a2, a3);
      rMb = ratio_minus(B, b1, b2, b3);
    }
    if (vp) { // finalize cell c = w-4 (window index w-1)
      double da, db;
      if (posp) {
        da = A.dm2 + (A.rP1 - A.rP2);
        db = B.dm2 + (B.rP1 - B.rP2);
      } else {
        da = A.dm1 + (rMa - A.rM1);
        db = B.dm1 + (rMb - B.rM1);
      }
      emit(w - 4, Uam1, Ubm1, da, db, A.dm1 - A.dm2, B.dm1 - B.dm2);
    }
    if (w < 11) {
      Uam1 = qa[w * es];
      Ubm1 = qb[w * es];
      const double qna = qa[(w + 3) * es], qnb = qb[(w + 3) * es];
      line_advance(A, qna, rPa, rMa);
      line_advance(B, qnb, rPb, rMb);
    }
  }
}

} // namespace cup2d
