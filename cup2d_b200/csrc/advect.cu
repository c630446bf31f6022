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
#include "weno.cuh"
#include <vector>

namespace cup2d {

#ifndef CUP2D_ADV_WARP_ROWS
#define CUP2D_ADV_WARP_ROWS 0
#endif

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

// MODE 0: out = tot (raw K, undivided)   1: old == in (stage 1)   2: old is a separate field (stage 2)
// Both passes run through ONE copy of the fully unrolled line code (a 2-trip loop with run-time strides)
// instead of two specialised copies: halves the instruction footprint (I-cache), profiles/r01g.
// DEVFAC: the dt-dependent factors come from device memory (StepFactors, written by k_step_factors) instead of the
// by-value arguments, so that a time step captured in a CUDA graph needs no host-supplied dt.
template <int MODE, bool DEVFAC>
__global__ void __launch_bounds__(NT_ADV, 4)
advect_stage_kernel(const double *__restrict__ in, const double *__restrict__ old,
                    double *__restrict__ out, const int *__restrict__ tiles,
                    const int *__restrict__ tile_org, const unsigned *__restrict__ lut, int nbx, int nby,
                    int nloc, double afac_arg, double dfac_arg, double ofac, const StepFactors *__restrict__ sf) {
  const double afac = DEVFAC ? sf->afac : afac_arg, dfac = DEVFAC ? sf->dfac : dfac_arg;
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
  // x-pass ownership.  Default: lane = row, warp = 8-cell segment (every warp touches all 32 rows, so the y pass, whose warp
  // ys needs rows 8ys..8ys+7 complete, waits at a CTA barrier).  CUP2D_ADV_WARP_ROWS (prepared for measurement, off in the
  // validated build): warp = the 8 rows of ITS OWN block row, lane = (segment, row): the y pass of a warp then only reads what
  // the same warp wrote and the CTA barrier between the passes becomes a __syncwarp (stall_barrier was 20 % in
  // profiles/r01i_advect_ncu.md).  Bank-conflict-free either way (row stride 39 / 33 doubles).
#if CUP2D_ADV_WARP_ROWS
  const int xrow = 8 * warp + (lane & 7), xseg = lane >> 3;
#else
  const int xrow = lane, xseg = warp;
#endif
  auto emit_x = [&](int c, double U, double, double du, double dv, double D2u, double D2v) {
    double *ru = Ru + xrow * RP + 8 * xseg, *rv = Rv + xrow * RP + 8 * xseg;
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
    const double *qa = pass == 0 ? su + (xrow + GH) * SP + 8 * xseg : sv + (8 * ys) * SP + (yx + GH);
    const double *qb = pass == 0 ? sv + (xrow + GH) * SP + 8 * xseg : su + (8 * ys) * SP + (yx + GH);
    const int es = pass == 0 ? 1 : SP;
    weno_line(qa, qb, es, [&](int c, double Ua, double Ub, double da, double db, double D2a, double D2b) {
      if (pass == 0) emit_x(c, Ua, Ub, da, db, D2a, D2b);
      else emit_y(c, Ua, Ub, da, db, D2a, D2b);
    });
#if CUP2D_ADV_WARP_ROWS
    __syncwarp();
#else
    __syncthreads();
#endif
  }
}

typedef void (*adv_fn)(const double *, const double *, double *, const int *, const int *, const unsigned *, int,
                       int, int, double, double, double, const StepFactors *);

// dev: null (factors from dt) or the device-resident factors of the current step
int launch_advect(cup2d_sim *s, const double *in, const double *old, double *out, double coef,
                  double dt, bool raw, const StepFactors *dev) {
  static PerDeviceOnce constants;
  int rc = constants.run(s->device, []() -> int {
    CUP2D_CUDA(cudaMemcpyToSymbol(cW, hW, sizeof hW));
    return (int)CUP2D_OK;
  });
  if (rc) return rc;
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
  const adv_fn fn = dev ? (mode == 0 ? advect_stage_kernel<0, true> : mode == 1 ? advect_stage_kernel<1, true> : advect_stage_kernel<2, true>)
                        : (mode == 0 ? advect_stage_kernel<0, false> : mode == 1 ? advect_stage_kernel<1, false> : advect_stage_kernel<2, false>);
  static PerDeviceOnce configured[6];
  rc = configured[mode + (dev ? 3 : 0)].run(s->device, [fn]() -> int {
    CUP2D_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ADV_SMEM));
    return (int)CUP2D_OK;
  });
  if (rc) return rc;
  const double afac = -dt * s->h; // main.cpp:5447
  const double dfac = s->nu * dt; // main.cpp:5446
  const double ofac = coef / (s->h * s->h);
  ProfScope prof(s, KC_ADVECT);
  fn<<<s->ntiles, NT_ADV, ADV_SMEM, s->stream>>>(in, old, out, s->d_tiles, s->d_tile_org, s->d_adv_lut, s->nbx,
                                                 s->nby, (int)s->nloc, afac, dfac, ofac, dev);
  s->launches++;
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

} // namespace cup2d
