// Advect-diffuse right-hand side on multi-level meshes, per-block lab loader on the WENO line core.
//
// STATUS: like csrc/amr_ops.cu — compiled for sm_100a, its logic checked on the CPU through tests/host_emu/ (cooperative
// thread emulation), NOT YET RUN ON HARDWARE.  cup2d_amr_advect_diffuse_rhs_fast gives the same result as the table-gather
// baseline cup2d_amr_advect_diffuse_rhs.
//
// One warp owns four blocks.  Per block two padded planes (u, v: 14 rows x 15) in shared memory hold the lab:
//   interior      the block itself (1 KB, coalesced)
//   face ghosts   3 layers behind each of the 4 faces (the stencil of KernelAdvectDiffuse is cross-shaped: main.cpp:5441-5503
//                 reads +-3 along the axes only): same-level neighbour -> copy; domain wall -> the wall-adjacent cell with
//                 the normal component negated (main.cpp:3144-3154); coarser / finer neighbour -> left to the table pass
//   table pass    blocks that have a non-same-level neighbour get their ghost rows from the compact ghost-stencil table of
//                 the host plan (cup2d_amr_plan_ghosts): value = sum_e w[e] * vel[src[e]]
// then the x pass (lanes = 4 blocks x 8 rows) and the y pass (lanes = 4 blocks x 8 columns) of weno.cuh, exactly the
// arithmetic of the uniform-grid kernel (advect.cu) with the cell size of the block.  Warps never synchronise with each
// other.  Blocks with a coarse-fine face also store their face fluxes dfac*(inner - ghost) (main.cpp:5515-5569) for the
// correction kernel below, which is the coarse-face formulation of fillcases (see amr_ops.cu) reading those buffers.
#include "sim.h"
#include "weno.cuh"
#include <algorithm>
#include <vector>
#include "amr.h"

extern "C" {
int64_t cup2d_amr_plan_irregular(cup2d_amr_plan *p, int32_t *blocks_out);
int64_t cup2d_amr_plan_ghosts(cup2d_amr_plan *p, int which, int64_t *nrows, int64_t *rowptr, int32_t *dst, int32_t *src_block,
                              int32_t *src_cellcomp, double *weight);
int cup2d_amr_plan_neighbours(cup2d_amr_plan *p, int32_t *out);
}

namespace cup2d {

constexpr int AF_WARPS = 4;                // warps per CTA
constexpr int AF_NT = 32 * AF_WARPS;
constexpr int AF_BPW = 4;                  // blocks per warp
constexpr int AF_PS = 15;                  // plane row stride (odd: lanes along y are conflict-free)
constexpr int AF_PLANE = 14 * AF_PS;       // doubles per component plane
constexpr int AF_RS = 9;                   // partial-result row stride
constexpr int AF_BLK = 2 * AF_PLANE + 2 * 8 * AF_RS; // doubles of shared memory per block: su, sv, Ru, Rv
constexpr int AF_SMEM = AF_WARPS * AF_BPW * AF_BLK * 8; // 72 192 B per CTA -> 3 CTAs/SM

__global__ void __launch_bounds__(AF_NT)
amr_advect_fast_kernel(const double *__restrict__ vel, double *__restrict__ out, const int4 *__restrict__ nbr4,
                       const int *__restrict__ irr_of, const int64_t *__restrict__ grow, const int64_t *__restrict__ growptr,
                       const int *__restrict__ gdst, const int *__restrict__ gsb, const int *__restrict__ gsc,
                       const double *__restrict__ gw, const double *__restrict__ hb, double *__restrict__ faceflux,
                       int nb, double nu, double dt) {
  extern __shared__ __align__(16) double af_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane >> 3, r = lane & 7;
  const int b0 = (blockIdx.x * AF_WARPS + warp) * AF_BPW; // first block of this warp
  if (b0 >= nb) return;
  double *base = af_smem + (size_t)warp * AF_BPW * AF_BLK;
  const double2 *vel2 = reinterpret_cast<const double2 *>(vel);

  // ---- load: interior + same-level / wall face ghosts of the warp's (up to) four blocks ----
  for (int qq = 0; qq < AF_BPW; qq++) {
    const int k = b0 + qq;
    if (k >= nb) break;
    double *su = base + qq * AF_BLK, *sv = su + AF_PLANE;
    for (int i = lane; i < 64; i += 32) {
      const double2 v = vel2[(size_t)k * 64 + i];
      const int p = ((i >> 3) + 3) * AF_PS + (i & 7) + 3;
      su[p] = v.x;
      sv[p] = v.y;
    }
    const int4 nbk = nbr4[k];
    for (int g = lane; g < 96; g += 32) { // face f, position t along it, layer d = 1..3 behind it
      const int f = g / 24, t = g % 8, d = (g % 24) / 8 + 1;
      const int nbf = f == 0 ? nbk.x : f == 1 ? nbk.y : f == 2 ? nbk.z : nbk.w;
      if (nbf < -1) continue; // coarser / finer: the table pass writes this ghost
      int ix, iy, sx, sy; // ghost position (block-relative) and source cell
      if (f < 2) {
        ix = f == 0 ? -d : 7 + d, iy = t;
        sx = nbf >= 0 ? (f == 0 ? 8 - d : d - 1) : (f == 0 ? 0 : 7), sy = t;
      } else {
        ix = t, iy = f == 2 ? -d : 7 + d;
        sx = t, sy = nbf >= 0 ? (f == 2 ? 8 - d : d - 1) : (f == 2 ? 0 : 7);
      }
      double2 v = vel2[(size_t)(nbf >= 0 ? nbf : k) * 64 + sy * 8 + sx];
      if (nbf < 0) { // wall: normal component negated, tangential copied
        if (f < 2) v.x = -v.x; else v.y = -v.y;
      }
      const int p = (iy + 3) * AF_PS + ix + 3;
      su[p] = v.x;
      sv[p] = v.y;
    }
  }
  __syncwarp();
  // ---- table pass: ghost rows of the blocks with a coarser / finer neighbour ----
  for (int qq = 0; qq < AF_BPW; qq++) {
    const int k = b0 + qq;
    if (k >= nb) break;
    const int qi = irr_of[k];
    if (qi < 0) continue;
    double *su = base + qq * AF_BLK;
    for (int64_t row = grow[qi] + lane; row < grow[qi + 1]; row += 32) {
      double acc = 0.0;
      for (int64_t e = growptr[row]; e < growptr[row + 1]; e++) acc += gw[e] * vel[(size_t)gsb[e] * 128 + gsc[e]];
      const int cc = gdst[row] % (14 * 14 * 2); // (lab row * 14 + lab column) * 2 + component
      const int comp = cc & 1, lx = (cc >> 1) % 14, ly = (cc >> 1) / 14;
      su[comp * AF_PLANE + ly * AF_PS + lx] = acc;
    }
  }
  __syncwarp();
  const int k = b0 + q;
  const bool live = k < nb;
  double *su = base + q * AF_BLK, *sv = su + AF_PLANE, *Ru = sv + AF_PLANE, *Rv = Ru + 8 * AF_RS;
  const double h = live ? hb[k] : 1.0;
  const double dfac = nu * dt, afac = -dt * h; // main.cpp:5446-5447
  // ---- face fluxes of the blocks with a coarse-fine face: lane = (face q', position r) of block qq ----
  for (int qq = 0; qq < AF_BPW; qq++) {
    const int kk = b0 + qq;
    if (kk >= nb) break;
    const int qi = irr_of[kk];
    if (qi < 0) continue;
    const double *pu = base + qq * AF_BLK, *pv = pu + AF_PLANE;
    const int f = q, t = r;
    int ix, iy, gx, gy;
    if (f < 2) ix = f == 0 ? 0 : 7, iy = t, gx = f == 0 ? -1 : 8, gy = t;
    else ix = t, iy = f == 2 ? 0 : 7, gx = t, gy = f == 2 ? -1 : 8;
    const int pi = (iy + 3) * AF_PS + ix + 3, pg = (gy + 3) * AF_PS + gx + 3;
    double *ff = faceflux + ((size_t)qi * 32 + f * 8 + t) * 2;
    ff[0] = dfac * (pu[pi] - pu[pg]);
    ff[1] = dfac * (pv[pi] - pv[pg]);
  }
  // (lanes of a partial last warp keep running on their unused planes and simply do not store)
  // ---- x pass: lane = row r of block q ----
  weno_line(su + (r + 3) * AF_PS, sv + (r + 3) * AF_PS, 1,
            [&](int c, double U, double, double du, double dv, double D2u, double D2v) {
              const double aU = afac * U;
              Ru[r * AF_RS + c] = fma(aU, du, dfac * D2u);
              Rv[r * AF_RS + c] = fma(aU, dv, dfac * D2v);
            });
  __syncwarp();
  // ---- y pass: lane = column r of block q; advecting component is v ----
  double2 *outp = reinterpret_cast<double2 *>(out) + (size_t)(live ? k : 0) * 64 + r;
  weno_line(sv + (r + 3), su + (r + 3), AF_PS, [&](int c, double V, double, double dv, double du, double D2v, double D2u) {
    const double aV = afac * V;
    double2 o;
    o.x = Ru[c * AF_RS + r] + fma(aV, du, dfac * D2u);
    o.y = Rv[c * AF_RS + r] + fma(aV, dv, dfac * D2v);
    if (live) outp[c * 8] = o;
  });
}

// fillcases per coarse face from the stored face fluxes (see amr_fluxcorr_kernel in amr_ops.cu for the formulation)
__global__ void amr_fluxcorr_faces_kernel(const CoarseFace *__restrict__ cf, int ncf, const double *__restrict__ faceflux,
                                          const int *__restrict__ irr_of, double *__restrict__ result) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncf * 16) return;
  const int comp = i & 1, t = (i >> 1) & 7;
  const CoarseFace f = cf[i >> 4];
  auto flux = [&](int blk, int face, int pos) { return faceflux[(((size_t)irr_of[blk] * 4 + face) * 8 + pos) * 2 + comp]; };
  double acc = flux(f.coarse, f.face, t);
  const int fb = f.fine[t >> 2];
  if (fb >= 0) {
    const int t2 = 2 * (t & 3);
    acc += flux(fb, f.face ^ 1, t2) + flux(fb, f.face ^ 1, t2 + 1);
  }
  const int ix = f.face < 2 ? (f.face == 0 ? 0 : 7) : t, iy = f.face < 2 ? t : (f.face == 2 ? 0 : 7);
  double *dst = result + ((size_t)f.coarse * 64 + iy * 8 + ix) * 2 + comp;
  double v = *dst + acc;
  if (f.fine[0] >= 0 && f.fine[1] >= 0 && 2 * t + comp >= 9) v += acc; // second pass of fillcase1 (property 3)
  *dst = v;
}

static int fast_setup(cup2d_amr *a) {
  if (a->d_nbr4) return CUP2D_OK;
  const int64_t nb = a->nb;
  std::vector<int32_t> n8(8 * nb), n4(4 * nb), irr_of(nb, -1);
  if (cup2d_amr_plan_neighbours(a->plan, n8.data())) return CUP2D_EINVAL;
  for (int64_t k = 0; k < nb; k++) { // order of the 8: (-1,-1),(0,-1),(1,-1),(-1,0),(1,0),(-1,1),(0,1),(1,1)
    const int pick[4] = {3, 4, 1, 6};  // W, E, S, N
    for (int j = 0; j < 4; j++) n4[4 * k + j] = n8[8 * k + pick[j]] < -1 ? -2 : n8[8 * k + pick[j]];
  }
  const int64_t nirr = cup2d_amr_plan_irregular(a->plan, nullptr);
  std::vector<int32_t> irr(std::max<int64_t>(nirr, 1));
  cup2d_amr_plan_irregular(a->plan, irr.data());
  for (int64_t qi = 0; qi < nirr; qi++) irr_of[irr[qi]] = (int32_t)qi;
  int64_t nrows = 0;
  const int64_t nnz = cup2d_amr_plan_ghosts(a->plan, 0, &nrows, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (nnz < 0) return CUP2D_EINVAL;
  std::vector<int64_t> rp(nrows + 1), grow(nirr + 1, nrows);
  std::vector<int32_t> dst(std::max<int64_t>(nrows, 1)), sb(std::max<int64_t>(nnz, 1)), sc(std::max<int64_t>(nnz, 1));
  std::vector<double> w(std::max<int64_t>(nnz, 1));
  cup2d_amr_plan_ghosts(a->plan, 0, &nrows, rp.data(), dst.data(), sb.data(), sc.data(), w.data());
  for (int64_t row = nrows - 1; row >= 0; row--) grow[dst[row] / (14 * 14 * 2)] = row; // rows are grouped by block, in order
  for (int64_t qi = nirr - 1; qi >= 0; qi--) grow[qi] = std::min(grow[qi], grow[qi + 1]); // blocks without rows
  a->nirr = nirr;
  auto up = [](auto **d, const auto &h) -> cudaError_t {
    using T = typename std::remove_reference<decltype(h[0])>::type;
    cudaError_t e = cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  };
  CUP2D_CUDA(up(&a->d_nbr4, n4));
  CUP2D_CUDA(up(&a->d_irr_of, irr_of));
  CUP2D_CUDA(up(&a->d_grow, grow));
  CUP2D_CUDA(up(&a->d_growptr, rp));
  CUP2D_CUDA(up(&a->d_gdst, dst));
  CUP2D_CUDA(up(&a->d_gsb, sb));
  CUP2D_CUDA(up(&a->d_gsc, sc));
  CUP2D_CUDA(up(&a->d_gw, w));
  CUP2D_CUDA(cudaMalloc(&a->d_faceflux, std::max<int64_t>(nirr, 1) * 64 * sizeof(double)));
  CUP2D_CUDA(cudaMemcpyToSymbol(cW, hW, sizeof hW));
  CUP2D_CUDA(cudaFuncSetAttribute(amr_advect_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM));
  return CUP2D_OK;
}

} // namespace cup2d

using namespace cup2d;

extern "C" {

/* tmpV = KernelAdvectDiffuse(vel), flux-corrected (main.cpp:6611-6617): same result as cup2d_amr_advect_diffuse_rhs */
int cup2d_amr_advect_diffuse_rhs_fast(cup2d_amr *a, double dt) {
  if (!a) {
    set_error("null cup2d_amr handle");
    return CUP2D_EINVAL;
  }
  CUP2D_CUDA(cudaSetDevice(a->device));
  int rc = fast_setup(a);
  if (rc) return rc;
  const int per_cta = AF_WARPS * AF_BPW;
  const int grid = (int)((a->nb + per_cta - 1) / per_cta);
  amr_advect_fast_kernel<<<grid, AF_NT, AF_SMEM, a->stream>>>(
      a->f[CUP2D_VEL], a->f[CUP2D_TMPV], reinterpret_cast<const int4 *>(a->d_nbr4), a->d_irr_of, a->d_grow, a->d_growptr,
      a->d_gdst, a->d_gsb, a->d_gsc, a->d_gw, a->d_h, a->d_faceflux, (int)a->nb, a->nu, dt);
  CUP2D_CUDA(cudaGetLastError());
  for (int dir = 0; dir < 2; dir++) { // x faces, then y faces (fillcases order)
    const int n = a->ncf[dir] * 16;
    if (n == 0) continue;
    amr_fluxcorr_faces_kernel<<<(n + 127) / 128, 128, 0, a->stream>>>(a->d_cf[dir], a->ncf[dir], a->d_faceflux, a->d_irr_of,
                                                                    a->f[CUP2D_TMPV]);
  }
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

} // extern "C"
