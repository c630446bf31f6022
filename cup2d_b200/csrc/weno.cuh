// WENO5 line core shared by the advect kernels (uniform tiles: advect.cu; multi-level blocks: amr_fast.cu).
//
// Upwind WENO5 differences of both velocity components along one line of 8 cells (window of 14 values per component)
// (reference: weno5_plus/minus/derivative, main.cpp:162-208), re-derived so that the FP64 pipe — the unit that bounds the
// advect kernels — sees as few instructions as the algebra allows.  With D[k] = q[k+1]-q[k], E(k) = D[k]-D[k-1] (second
// difference at k), every smoothness indicator scaled by 4 (the weights are ratios, so a common factor cancels):
//   B1 = 13/3 E(w-1)^2 + (3 D[w-1] - D[w-2])^2 + 4e-6,   B2 = 13/3 E(w)^2 + (D[w-1]+D[w])^2 + 4e-6,
//   B3 = 13/3 E(w+1)^2 + (3 D[w] - D[w+1])^2 + 4e-6                      ( = 4 (beta_k + 1e-6) )
//   s1 = (B2 B3)^2, s2 = (B1 B3)^2, s3 = (B1 B2)^2        ( the reference's alpha_k = gamma_k/(beta_k+eps)^2 times (B1 B2 B3)^2 )
// The three candidate fluxes differ from the central one by THIRD differences only:
//   phi_1 - phi_2 = -(E(w) - E(w-1))/3 = -T1/3,      phi_3 - phi_2 = -(E(w+1) - E(w))/6 = -T3/6
// so the left-biased flux at face w+1/2 is  q[w] + ratioP(w)  with (gammas .1 .6 .3, numerator and denominator times 10)
//   3 ratioP = (D[w-1]/2 + D[w]) - (s1 T1 + 1.5 s3 T3) / (s1 + 6 s2 + 3 s3)
// and, mirrored, the right-biased flux at face w-1/2 is  q[w] + ratioM(w)  with
//   3 ratioM = -(D[w]/2 + D[w-1]) + (s3 T3 + 1.5 s1 T1) / (s3 + 6 s2 + 3 s1).
// Everything downstream works with 3 x (undivided upwind difference); the callers fold the 1/3 into their advection
// factor.  Per flux: 2 + 2 + 1 + 1 products/sums, one reciprocal (MUFU seed + one cubic step), one fused add: 10 FP64
// instructions (round 1's sum-of-candidates form: 17), on top of 17 per window position for differences and indicators.
// Same real-number result as the reference; rounding differs at the 1e-16 relative level (tests bound it at 1e-12).
#pragma once
#include "common.cuh"

namespace cup2d {

// 13/3 and 4e-6 need all 64 bits: they live in constant memory so that they are DFMA constant-bank operands (as literals
// the compiler materialises each with a pair of UMOVs); 3, 6, 1.5, .5 fit the 32-bit FP64 immediate of DFMA/DMUL.
enum { K_G = 0, K_EPS, K_N };
static __constant__ double cW[K_N];
static const double hW[K_N] = {13.0 / 3.0, 4e-6};

struct LineState {
  double dm2, dm1, d0, dp1; // D[w-2..w+1]
  double Em1, E0, Ep1;      // second differences at w-1, w, w+1
  double Gm1, G0, Gp1;      // 13/3 E^2 + 4e-6 at w-1, w, w+1
  double T1;                // E(w) - E(w-1)
  double qlast;             // q[w+2]
  double rP1, rP2, rM1;     // 3 ratioP(w-1), 3 ratioP(w-2), 3 ratioM(w-1)
};

__device__ __forceinline__ double Gfun(double E) {
  return fma(cW[K_G] * E, E, cW[K_EPS]);
}
__device__ __forceinline__ void line_init(LineState &s, double q0, double q1, double q2, double q3, double q4) {
  s.dm2 = q1 - q0; // w = 2: D[0]
  s.dm1 = q2 - q1; // D[1]
  s.d0 = q3 - q2;  // D[2]
  s.dp1 = q4 - q3; // D[3]
  s.Em1 = s.dm1 - s.dm2;
  s.E0 = s.d0 - s.dm1;
  s.Ep1 = s.dp1 - s.d0;
  s.Gm1 = Gfun(s.Em1);
  s.G0 = Gfun(s.E0);
  s.Gp1 = Gfun(s.Ep1);
  s.T1 = s.E0 - s.Em1;
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
// 1/x for x > 0, normal: MUFU.RCP64H seed (relative error ~2^-20) + one cubic step r0 (1 + e + e^2), e = 1 - x r0:
// error e^3 ~ 2^-60, below the rounding of the three operations themselves.  (Two Newton steps: one FP64 instruction more.)
__device__ __forceinline__ double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double e = fma(-x, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
// 3 x (left-biased flux at face w+1/2 minus q[w])  (weno5_plus, main.cpp:162-181);  u1 = s1 T1, u3 = s3 T3
__device__ __forceinline__ double ratio_plus(const LineState &s, double s1, double s2, double s3, double u1, double u3) {
  const double den = fma(6.0, s2, fma(3.0, s3, s1));
  const double num = fma(1.5, u3, u1);
  const double lin = fma(0.5, s.dm1, s.d0);
  return fma(-num, rcp_pos(den), lin);
}
// 3 x (right-biased flux at face w-1/2 minus q[w])  (weno5_minus, main.cpp:182-201)
__device__ __forceinline__ double ratio_minus(const LineState &s, double s1, double s2, double s3, double u1, double u3) {
  const double den = fma(6.0, s2, fma(3.0, s1, s3));
  const double num = fma(1.5, u1, u3);
  const double lin = fma(0.5, s.d0, s.dm1);
  return fma(num, rcp_pos(den), -lin);
}
__device__ __forceinline__ void line_advance(LineState &s, double qn, double T3, double rP, double rM) {
  s.rP2 = s.rP1;
  s.rP1 = rP;
  s.rM1 = rM;
  s.dm2 = s.dm1;
  s.dm1 = s.d0;
  s.d0 = s.dp1;
  s.dp1 = qn - s.qlast;
  s.qlast = qn;
  s.Em1 = s.E0;
  s.E0 = s.Ep1;
  s.Ep1 = s.dp1 - s.d0;
  s.Gm1 = s.G0;
  s.G0 = s.Gp1;
  s.Gp1 = Gfun(s.Ep1);
  s.T1 = T3;
}

// ---- the line cores -----------------------------------------------------------------------------------------------------
// ld(k, x, y): both components at window index k (0..13; cell c is index c+3).  emit(c, dx3, dy3, Ex, Ey) is called once
// per cell c = 0..7 with 3 x the undivided upwind differences (reference `derivative`, main.cpp:202-208) of the two
// components and their second differences (diffusion term).  The two components are treated alike: the line direction and
// which of them advects enter only through ld and the sign bits, so one copy of this code can serve both passes.
//
// weno_line_upwind: all eight cells are advected from the left (index 0 side).  Straight-line code — nine left-biased
// fluxes, no branch, nothing evaluated that is not used — which the scheduler can interleave across window positions.
// A line whose eight cells are all advected from the RIGHT is the same computation on the mirrored line: the caller hands
// in ld(k) = q[13-k], writes cell 7-c where emit says c, and negates the differences (flux^-(c+1/2) of q is
// flux^+(c'-1/2) of the mirrored line, c' = 7-c, so D^-(c) = -D'^+(c'); second differences are symmetric).
template <class Ld, class Emit>
__device__ __forceinline__ void weno_line_upwind(Ld ld, Emit emit) {
  LineState A, B;
  {
    double x0, x1, x2, x3, x4, y0, y1, y2, y3, y4;
    ld(0, x0, y0);
    ld(1, x1, y1);
    ld(2, x2, y2);
    ld(3, x3, y3);
    ld(4, x4, y4);
    line_init(A, x0, x1, x2, x3, x4);
    line_init(B, y0, y1, y2, y3, y4);
  }
#pragma unroll
  for (int w = 2; w <= 10; ++w) {
    if (w >= 4) // finalize cell c = w-4 (window index w-1): flux(w-1) - flux(w-2) + D[w-2]
      emit(w - 4, fma(3.0, A.dm2, A.rP1 - A.rP2), fma(3.0, B.dm2, B.rP1 - B.rP2), A.Em1, B.Em1);
    const double T3a = A.Ep1 - A.E0, T3b = B.Ep1 - B.E0;
    double a1, a2, a3, b1, b2, b3;
    line_betas(A, a1, a2, a3);
    line_betas(B, b1, b2, b3);
    const double rPa = ratio_plus(A, a1, a2, a3, a1 * A.T1, a3 * T3a);
    const double rPb = ratio_plus(B, b1, b2, b3, b1 * B.T1, b3 * T3b);
    double qna = 0.0, qnb = 0.0;
    if (w < 10) ld(w + 3, qna, qnb); // index 13 only enters right-biased fluxes
    line_advance(A, qna, T3a, rPa, 0.0);
    line_advance(B, qnb, T3b, rPb, 0.0);
  }
  emit(7, fma(3.0, A.dm2, A.rP1 - A.rP2), fma(3.0, B.dm2, B.rP1 - B.rP2), A.Em1, B.Em1); // window position 11
}

// weno_line_core: any sign pattern.  pos: bit k set <-> the advecting component is positive at window index k; only the
// bits of the eight cells (k = 3..10) are looked at.  (+0 may be reported either way: the term it selects is multiplied
// by U = 0; the reference's `U > 0` sends it to the right-biased side.)
template <class Ld, class Emit>
__device__ __forceinline__ void weno_line_core(Ld ld, const unsigned pos, Emit emit) {
  LineState A, B;
  {
    double x0, x1, x2, x3, x4, y0, y1, y2, y3, y4;
    ld(0, x0, y0);
    ld(1, x1, y1);
    ld(2, x2, y2);
    ld(3, x3, y3);
    ld(4, x4, y4);
    line_init(A, x0, x1, x2, x3, x4);
    line_init(B, y0, y1, y2, y3, y4);
  }
#pragma unroll
  for (int w = 2; w <= 11; ++w) {
    const bool vc = (unsigned)(w - 3) < 8u, vn = (unsigned)(w - 2) < 8u, vp = (unsigned)(w - 4) < 8u;
    const unsigned pw = pos >> (w - 1); // bit 0: cell w-1, bit 1: cell w, bit 2: cell w+1
    const bool posp = pw & 1u;
    // flux families needed at this window position (masks are compile-time after unrolling)
    const bool needP = (pw & ((vc ? 2u : 0u) | (vn ? 4u : 0u))) != 0u;
    const bool needM = (~pw & ((vc ? 2u : 0u) | (vp ? 1u : 0u))) != 0u;
    const double T3a = A.Ep1 - A.E0, T3b = B.Ep1 - B.E0;
    double rPa = 0, rPb = 0, rMa = 0, rMb = 0;
    double a1, a2, a3, b1, b2, b3;
    line_betas(A, a1, a2, a3);
    line_betas(B, b1, b2, b3);
    const double ua1 = a1 * A.T1, ua3 = a3 * T3a, ub1 = b1 * B.T1, ub3 = b3 * T3b;
    if (needP) {
      rPa = ratio_plus(A, a1, a2, a3, ua1, ua3);
      rPb = ratio_plus(B, b1, b2, b3, ub1, ub3);
    }
    if (needM) {
      rMa = ratio_minus(A, a1, a2, a3, ua1, ua3);
      rMb = ratio_minus(B, b1, b2, b3, ub1, ub3);
    }
    if (vp) { // finalize cell c = w-4 (window index w-1)
      double da, db;
      if (posp) {
        da = fma(3.0, A.dm2, A.rP1 - A.rP2);
        db = fma(3.0, B.dm2, B.rP1 - B.rP2);
      } else {
        da = fma(3.0, A.dm1, rMa - A.rM1);
        db = fma(3.0, B.dm1, rMb - B.rM1);
      }
      emit(w - 4, da, db, A.Em1, B.Em1);
    }
    if (w < 11) {
      double qna, qnb;
      ld(w + 3, qna, qnb);
      line_advance(A, qna, T3a, rPa, rMa);
      line_advance(B, qnb, T3b, rPb, rMb);
    }
  }
}

// Separate component planes (amr_fast.cu): qa = advecting component, qb = the other one, element stride es.
// emit(c, Ua, Ub, da3, db3, D2a, D2b): cell values, 3 x the undivided upwind differences, second differences; c may be a
// run-time value (lines advected from the right are walked backwards).
template <class Emit>
__device__ __forceinline__ void weno_line(const double *__restrict__ qa, const double *__restrict__ qb,
                                          const int es, Emit emit) {
  unsigned pos = 0;
#pragma unroll
  for (int k = 3; k <= 10; k++) pos |= __double2hiint(qa[k * es]) >= 0 ? (1u << k) : 0u;
  if (pos == 0x7f8u || pos == 0u) {
    const bool rev = pos == 0u;
    const double *pa = rev ? qa + 13 * es : qa, *pb = rev ? qb + 13 * es : qb;
    const int ee = rev ? -es : es;
    const double sg = rev ? -1.0 : 1.0;
    weno_line_upwind([&](int k, double &x, double &y) { x = pa[k * ee]; y = pb[k * ee]; },
                     [&](int c, double da3, double db3, double Ea, double Eb) {
                       // U, da3 enter the callers only through the product U * d3: the mirror's sign goes to U's copy
                       emit(rev ? 7 - c : c, pa[(c + 3) * ee], pb[(c + 3) * ee], sg * da3, sg * db3, Ea, Eb);
                     });
  } else {
    weno_line_core([&](int k, double &x, double &y) { x = qa[k * es]; y = qb[k * es]; }, pos,
                   [&](int c, double da3, double db3, double Ea, double Eb) {
                     emit(c, qa[(c + 3) * es], qb[(c + 3) * es], da3, db3, Ea, Eb);
                   });
  }
}

} // namespace cup2d
