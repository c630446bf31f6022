// Fused advection-diffusion RK stage:  out = old + coef * K(in) / h^2
//
// K is the reference's KernelAdvectDiffuse (main.cpp:5441-5503): WENO5 upwind differences
// (main.cpp:162-208) of both velocity components + 5-point diffusion, undivided.  The Lab assembly
// (BlockLab::load, main.cpp:2270-2440) and the free-slip ghost fill (VectorLab::applyBCface,
// main.cpp:3131-3154) become the tile loader of this kernel; the RK update loops
// (main.cpp:6618-6626, 6634-6642) are fused into the store.
//
// One CTA = one tile of 4x4 blocks (32x32 cells), 128 threads, 4 CTAs per SM.  Data path:
//   HBM --cp.async.bulk (1-D TMA, one 1 KB copy per 8x8 block, mbarrier complete_tx)--> staging smem
//   staging (AoS, block layout) --repack + ghost synthesis--> padded SoA planes su/sv (38x38, +3 ring)
//   x pass (lanes = rows, thread = 8 cells of a row)    -> partial result planes Ru/Rv (alias staging)
//   y pass (lanes = columns, thread = 8 cells of a column) -> + old, 128-bit coalesced stores
//
// Arithmetic (all FP64; this kernel is bound by the FP64 pipe, not by HBM — see DESIGN.md).  Per line
// the WENO fluxes are shared between neighbouring cells and only the upwind family that some cell needs
// is evaluated.  In differences D[k] = q[k+1]-q[k], D2(k) = D[k]-D[k-1], with every smoothness indicator
// scaled by 4 (the weights are ratios, so a common factor cancels):
//   B1 = 13/3 D2(w-1)^2 + (3 D[w-1] - D[w-2])^2 + 4e-6,   B2 = 13/3 D2(w)^2 + (D[w-1]+D[w])^2 + 4e-6,
//   B3 = 13/3 D2(w+1)^2 + (3 D[w] - D[w+1])^2 + 4e-6                      ( = 4 (beta_k + 1e-6) )
//   flux(w) = q[w] + (sum_k s_k gamma_k phi_k)/(sum_k gamma_k s_k),   s_k = (B_j B_l)^2
// which is the reference's w_k = (gamma_k/(beta_k+eps)^2)/sum with numerator and denominator multiplied
// by (B1 B2 B3)^2: one division per flux instead of four.  Same real-number result; rounding differs
// from the reference's CPU evaluation at the 1e-16 relative level (tests bound it at 1e-12).
#include "sim.h"
#include <vector>

namespace cup2d {

constexpr int TC = 32;          // tile cells per side
constexpr int GH = 3;           // ghost width (stencil -3..+3, main.cpp:5442)
constexpr int TW = TC + 2 * GH; // 38
constexpr int SP = 39;          // plane row stride (odd: conflict-free for lanes along y)
constexpr int RP = 33;          // partial-result plane stride
constexpr int NT_ADV = 128;
constexpr int STG_BYTES = 24 * 1024 + 8 * 384; // 16 interior + 4 W + 4 E blocks, 4 S + 4 N 3-row strips
constexpr int OFF_RU = 0;                      // Ru/Rv alias the staging area (dead after the repack)
constexpr int OFF_RV = OFF_RU + TC * RP * 8;
static_assert(OFF_RV + TC * RP * 8 <= STG_BYTES, "R planes must fit in the staging area");
constexpr int OFF_SU = STG_BYTES;
constexpr int OFF_SV = OFF_SU + TW * SP * 8;
constexpr int OFF_BAR = OFF_SV + TW * SP * 8;
constexpr int OFF_SLOTS = OFF_BAR + 16;
constexpr int ADV_SMEM = OFF_SLOTS + TILE_SLOTS * 4; // 51.5 KB -> 4 CTAs/SM

// WENO constants live in constant memory so that they are DFMA constant-bank operands; as literals the
// compiler materialises each with a pair of UMOVs (36 UMOV per cell in profiles/r01c).
enum { K_G = 0, K_EPS, K_D1, K_D2, K_D3,            // 13/3, 4e-6, den weights .1 .6 .3
       K_PA1, K_PB1, K_PA2, K_PB2, K_PA3, K_PB3,    // plus-flux phi coefficients (gamma folded in)
       K_MA1, K_MB1, K_MA2, K_MB2, K_MA3, K_MB3, K_N };
__constant__ double cW[K_N];
static const double hW[K_N] = {13.0 / 3.0, 4e-6, 0.1, 0.6, 0.3,
                               0.1 * 5.0 / 6.0, -0.1 / 3.0, 0.6 / 6.0, 0.6 / 3.0, 0.3 * 2.0 / 3.0, -0.3 / 6.0,
                               -0.3 * 2.0 / 3.0, 0.3 / 6.0, -0.6 / 3.0, -0.6 / 6.0, -0.1 * 5.0 / 6.0, 0.1 / 3.0};

constexpr int ADV_LUT_N = TW * TW - 4 * GH * GH; // 1408 cells of the cross-shaped footprint = 11 * 128
static_assert(ADV_LUT_N % NT_ADV == 0, "table must divide evenly among the threads");
// staging slot (in double2 units) of tile-local cell (cx,cy), -3 <= cx,cy < 35, not a corner
__host__ __device__ __forceinline__ int adv_src_slot(int cx, int cy) {
  if ((unsigned)cx < (unsigned)TC && (unsigned)cy < (unsigned)TC)
    return ((cy >> 3) * 4 + (cx >> 3)) * 64 + (cy & 7) * 8 + (cx & 7);
  if (cx < 0) return (16 + (cy >> 3)) * 64 + (cy & 7) * 8 + (8 + cx);
  if (cx >= TC) return (20 + (cy >> 3)) * 64 + (cy & 7) * 8 + (cx - TC);
  if (cy < 0) return 24 * 64 + (cx >> 3) * 24 + (3 + cy) * 8 + (cx & 7);
  return 24 * 64 + 4 * 24 + (cx >> 3) * 24 + (cy - TC) * 8 + (cx & 7);
}

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
// 1/x for x > 0, normal: MUFU.RCP64H seed (~2^-20) + two Newton steps = full double accuracy.  (One step
// leaves 1.3e-13 relative error for 3 % of the kernel time, profiles/r01h; not worth it.)
__device__ __forceinline__ double rcp_pos(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
  }
  return r;
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
    double a1, a2, a3, b1, b2, b3;
    line_betas(A, a1, a2, a3);
    line_betas(B, b1, b2, b3);
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

// MODE 0: out = tot (raw K, undivided)   1: old == in (stage 1)   2: old is a separate field (stage 2)
// Both passes run through ONE copy of the fully unrolled line code (a 2-trip loop with run-time strides)
// instead of two specialised copies: halves the instruction footprint (I-cache), profiles/r01g.
template <int MODE>
__global__ void __launch_bounds__(NT_ADV, 4)
advect_stage_kernel(const double *__restrict__ in, const double *__restrict__ old,
                    double *__restrict__ out, const int *__restrict__ tiles,
                    const int *__restrict__ tile_org, const unsigned *__restrict__ lut, int nbx, int nby,
                    int nloc, double afac, double dfac, double ofac) {
  extern __shared__ __align__(128) unsigned char smem[];
  double2 *stg = reinterpret_cast<double2 *>(smem);
  double *su = reinterpret_cast<double *>(smem + OFF_SU);
  double *sv = reinterpret_cast<double *>(smem + OFF_SV);
  double *Ru = reinterpret_cast<double *>(smem + OFF_RU);
  double *Rv = reinterpret_cast<double *>(smem + OFF_RV);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  int *s_slots = reinterpret_cast<int *>(smem + OFF_SLOTS);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  if (tid < TILE_SLOTS) s_slots[tid] = tiles[tile * TILE_SLOTS + tid];
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  // ---- stage 0: TMA bulk loads, one per block / strip, issued by the lanes of warp 0 ----
  if (warp == 0) {
    const int slot = s_slots[lane];
    const uint32_t bytes = slot >= 0 ? (lane < 24 ? 1024u : 384u) : 0u;
    uint32_t tot = bytes;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) mbar_arrive_expect_tx(bar, tot);
    __syncwarp();
    if (slot >= 0) {
      const unsigned char *src = reinterpret_cast<const unsigned char *>(in + (size_t)slot * 128);
      if (lane >= 24 && lane < 28) src += 5 * 128; // S strip = rows 5..7 of the block below
      unsigned char *dst = smem + (lane < 24 ? lane * 1024 : 24 * 1024 + (lane - 24) * 384);
      tma_load_1d(dst, src, bytes, bar);
    }
    // (an L2 prefetch of the next wave's tile here costs 7 %: profiles/r01h_ab_test.jsonl)
  }
  // y-pass ownership (known now, so the `old` loads of stage 2 can be in flight during everything else)
  const int yx = lane, ys = warp;
  const int yb = ys * 4 + (yx >> 3);
  const int yslot = s_slots[yb];
  const bool store = yslot >= 0 && yslot < nloc;
  double2 oldv[8];
  if (MODE == 2) {
    const double2 *oldp = reinterpret_cast<const double2 *>(old) + (size_t)(store ? yslot : 0) * 64 + (yx & 7);
#pragma unroll
    for (int c = 0; c < 8; c++) oldv[c] = store ? oldp[c * 8] : make_double2(0.0, 0.0);
  }
  const int gx0 = tile_org[2 * tile] * CUP2D_BS, gy0 = tile_org[2 * tile + 1] * CUP2D_BS;
  const int NX = nbx * CUP2D_BS, NY = nby * CUP2D_BS;
  if (warp == 0) mbar_wait(bar, 0); // one warp polls the mbarrier; the others park at the CTA barrier
  __syncthreads();

  // ---- stage 1: repack AoS blocks -> padded SoA planes, synthesising wall ghosts ----
  // Tiles whose ghost ring lies inside the domain (all but the perimeter tiles) use a precomputed table
  // (source slot in the staging area, destination in the planes) for the 1408 cells of the cross-shaped
  // footprint: 11 table entries per thread.  Index arithmetic used to be 29 % of the kernel's instructions.
  const bool edge = gx0 < GH || gy0 < GH || gx0 + TC + GH > NX || gy0 + TC + GH > NY;
  if (!edge) {
#pragma unroll
    for (int k = 0; k < ADV_LUT_N / NT_ADV; k++) {
      const unsigned e = __ldg(lut + k * NT_ADV + tid);
      const double2 v = stg[e & 0xffffu];
      su[e >> 16] = v.x;
      sv[e >> 16] = v.y;
    }
  } else {
    // (VectorLab::applyBCface main.cpp:3131-3154: ghost = wall-adjacent cell, normal component negated)
    for (int idx = tid; idx < TW * TW; idx += NT_ADV) {
      const int ty = idx / TW, tx = idx - ty * TW;
      const int lx = tx - GH, ly = ty - GH;
      const bool xin = (unsigned)lx < (unsigned)TC, yin = (unsigned)ly < (unsigned)TC;
      if (!xin && !yin) continue; // corner ghosts are never read (cross-shaped stencil)
      int gx = gx0 + lx, gy = gy0 + ly;
      double sgu = 1.0, sgv = 1.0;
      if (gx < 0) { gx = 0; sgu = -1.0; } else if (gx >= NX) { gx = NX - 1; sgu = -1.0; }
      if (gy < 0) { gy = 0; sgv = -1.0; } else if (gy >= NY) { gy = NY - 1; sgv = -1.0; }
      const double2 v = stg[adv_src_slot(gx - gx0, gy - gy0)];
      su[ty * SP + tx] = sgu * v.x;
      sv[ty * SP + tx] = sgv * v.y;
    }
  }
  __syncthreads(); // staging is dead from here on: Ru/Rv reuse it

  double2 *outp = reinterpret_cast<double2 *>(out) + (size_t)(store ? yslot : 0) * 64 + (yx & 7);
  // x pass: lanes = rows, thread = 8 consecutive cells of one row (advecting component u)
  // y pass: lanes = columns, thread = 8 consecutive cells of one column = one block (advecting v)
  auto emit_x = [&](int c, double U, double, double du, double dv, double D2u, double D2v) {
    double *ru = Ru + lane * RP + 8 * warp, *rv = Rv + lane * RP + 8 * warp;
    const double aU = afac * U;
    ru[c] = fma(aU, du, dfac * D2u); // afac*u*dudx + dfac*(u_E + u_W - 2u)
    rv[c] = fma(aU, dv, dfac * D2v);
  };
  auto emit_y = [&](int c, double V, double Uc, double dv, double du, double D2v, double D2u) {
    const double *ru = Ru + (8 * ys) * RP + yx, *rv = Rv + (8 * ys) * RP + yx;
    const double aV = afac * V;
    const double tu = ru[c * RP] + fma(aV, du, dfac * D2u);
    const double tv = rv[c * RP] + fma(aV, dv, dfac * D2v);
    if (store) {
      double2 o;
      if (MODE == 0) {
        o.x = tu;
        o.y = tv;
      } else if (MODE == 1) { // old == in: the cell values are already in registers
        o.x = fma(ofac, tu, Uc);
        o.y = fma(ofac, tv, V);
      } else {
        o.x = fma(ofac, tu, oldv[c].x); // V = Vold + coef*tmpV/h^2, main.cpp:6618-6626
        o.y = fma(ofac, tv, oldv[c].y);
      }
      outp[c * 8] = o;
    }
  };
#pragma unroll 1
  for (int pass = 0; pass < 2; pass++) {
    const double *qa = pass == 0 ? su + (lane + GH) * SP + 8 * warp : sv + (8 * ys) * SP + (yx + GH);
    const double *qb = pass == 0 ? sv + (lane + GH) * SP + 8 * warp : su + (8 * ys) * SP + (yx + GH);
    const int es = pass == 0 ? 1 : SP;
    weno_line(qa, qb, es, [&](int c, double Ua, double Ub, double da, double db, double D2a, double D2b) {
      if (pass == 0) emit_x(c, Ua, Ub, da, db, D2a, D2b);
      else emit_y(c, Ua, Ub, da, db, D2a, D2b);
    });
    __syncthreads();
  }
}

typedef void (*adv_fn)(const double *, const double *, double *, const int *, const int *, const unsigned *, int,
                       int, int, double, double, double);

int launch_advect(cup2d_sim *s, const double *in, const double *old, double *out, double coef,
                  double dt, bool raw) {
  static bool constants_up = false;
  if (!constants_up) {
    CUP2D_CUDA(cudaMemcpyToSymbol(cW, hW, sizeof hW));
    constants_up = true;
  }
  if (!s->d_adv_lut) { // repack table of interior tiles: (destination in the planes) << 16 | source slot
    std::vector<unsigned> lut;
    for (int ty = 0; ty < TW; ty++)
      for (int tx = 0; tx < TW; tx++) {
        const int lx = tx - GH, ly = ty - GH;
        if (!((unsigned)lx < (unsigned)TC) && !((unsigned)ly < (unsigned)TC)) continue;
        lut.push_back((unsigned)(ty * SP + tx) << 16 | (unsigned)adv_src_slot(lx, ly));
      }
    CUP2D_CUDA(cudaMalloc(&s->d_adv_lut, lut.size() * sizeof(unsigned)));
    CUP2D_CUDA(cudaMemcpy(s->d_adv_lut, lut.data(), lut.size() * sizeof(unsigned), cudaMemcpyHostToDevice));
  }
  const int mode = raw ? 0 : (old == in ? 1 : 2);
  const adv_fn fn = mode == 0 ? advect_stage_kernel<0> : mode == 1 ? advect_stage_kernel<1> : advect_stage_kernel<2>;
  static bool configured[3] = {false, false, false};
  if (!configured[mode]) {
    CUP2D_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ADV_SMEM));
    configured[mode] = true;
  }
  const double afac = -dt * s->h; // main.cpp:5447
  const double dfac = s->nu * dt; // main.cpp:5446
  const double ofac = coef / (s->h * s->h);
  ProfScope prof(s, KC_ADVECT);
  fn<<<s->ntiles, NT_ADV, ADV_SMEM, s->stream>>>(in, old, out, s->d_tiles, s->d_tile_org, s->d_adv_lut, s->nbx,
                                                 s->nby, (int)s->nloc, afac, dfac, ofac);
  s->launches++;
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

} // namespace cup2d
