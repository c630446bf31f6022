// dt control, Poisson right-hand side and pressure correction kernels.
//
// Thread mapping for the block-structured scalar kernels: one thread = one row of 8 cells of one
// 8x8 block (64 B of a scalar field / 128 B of a vector field, contiguous), 8 consecutive lanes = one
// block, one warp = 4 blocks = 2 KB (scalar) contiguous.  Neighbour rows/ghost cells come through the
// per-block neighbour table d_nbr[slot][W,E,S,N] (-1 = domain wall), which replaces the reference's
// BlockLab assembly (main.cpp:2270-2440) and ghost fill (main.cpp:3131-3154, 3210-3245).
#include "sim.h"

namespace cup2d {

constexpr int NT = 256;

__device__ __forceinline__ void load_row(const double *__restrict__ f, int slot, int y, double (&c)[8]) {
  const double4 *p = reinterpret_cast<const double4 *>(f + (size_t)slot * 64 + y * 8);
  double4 a = p[0], b = p[1];
  c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w;
  c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
}
__device__ __forceinline__ void store_row(double *__restrict__ f, int slot, int y, const double (&c)[8]) {
  double4 *p = reinterpret_cast<double4 *>(f + (size_t)slot * 64 + y * 8);
  p[0] = make_double4(c[0], c[1], c[2], c[3]);
  p[1] = make_double4(c[4], c[5], c[6], c[7]);
}
__device__ __forceinline__ void load_row2(const double *__restrict__ f, int slot, int y, double2 (&c)[8]) {
  const double4 *p = reinterpret_cast<const double4 *>(f + (size_t)slot * 128 + y * 16);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    double4 a = p[k];
    c[2 * k] = make_double2(a.x, a.y);
    c[2 * k + 1] = make_double2(a.z, a.w);
  }
}

// ------------------------------------------------------------------------------------------------
// umax = max |vel| over both components (main.cpp:6585-6591)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) umax_kernel(const double *__restrict__ vel, size_t n4,
                                                  double *partials, unsigned int *counter,
                                                  Comm comm, double *out) {
  double m = 0;
  const double4 *p = reinterpret_cast<const double4 *>(vel);
  for (size_t i = (size_t)blockIdx.x * NT + threadIdx.x; i < n4; i += (size_t)gridDim.x * NT) {
    double4 a = p[i];
    m = fmax(m, fmax(fmax(fabs(a.x), fabs(a.y)), fmax(fabs(a.z), fabs(a.w))));
  }
  double dummy[1] = {0};
  grid_reduce<1, NT>(dummy, m, partials, counter, comm, [=](const double *, double mx) { out[0] = mx; });
}

int launch_umax(cup2d_sim *s, double *umax_out) {
  const size_t n4 = (size_t)s->nloc * 128 / 4;
  int grid = s->num_sms * 4;
  {
  ProfScope prof(s, KC_UMAX);
  umax_kernel<<<grid, NT, 0, s->stream>>>(s->f[CUP2D_VEL], n4, s->d_partials, s->d_counter, s->comm, s->d_scal);
  }
  s->launches++;
  CUP2D_CUDA(cudaGetLastError());
  CUP2D_CUDA(cudaMemcpyAsync(s->h_scal, s->d_scal, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUP2D_CUDA(cudaStreamSynchronize(s->stream));
  *umax_out = s->h_scal[0];
  return CUP2D_OK;
}

// ------------------------------------------------------------------------------------------------
// undivided divergence of a vector field along one block row, free-slip ghosts
// (pressure_rhs main.cpp:6105-6139: (u_E - u_W + v_N) - v_S, left to right)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void div_row(const double *__restrict__ f, int slot, int y,
                                        const int4 nb, double (&d)[8]) {
  double2 c[8];
  load_row2(f, slot, y, c);
  const double2 *f2 = reinterpret_cast<const double2 *>(f);
  const double uW = nb.x >= 0 ? __ldg(f2 + (size_t)nb.x * 64 + y * 8 + 7).x : -c[0].x;
  const double uE = nb.y >= 0 ? __ldg(f2 + (size_t)nb.y * 64 + y * 8 + 0).x : -c[7].x;
  double2 up[8], dn[8];
  if (y < 7) load_row2(f, slot, y + 1, up);
  else if (nb.w >= 0) load_row2(f, nb.w, 0, up);
  else {
#pragma unroll
    for (int i = 0; i < 8; i++) up[i] = make_double2(c[i].x, -c[i].y);
  }
  if (y > 0) load_row2(f, slot, y - 1, dn);
  else if (nb.z >= 0) load_row2(f, nb.z, 7, dn);
  else {
#pragma unroll
    for (int i = 0; i < 8; i++) dn[i] = make_double2(c[i].x, -c[i].y);
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double e = i < 7 ? c[i + 1].x : uE;
    const double w = i > 0 ? c[i - 1].x : uW;
    d[i] = ((e - w) + up[i].y) - dn[i].y;
  }
}

// undivided 5-point Laplacian of a scalar along one block row, Neumann ghosts (ghost = adjacent cell)
// (pressure_rhs1 main.cpp:6209-6230: l1 + l2 + l3 + l4 - 4 l0 with l1=W, l2=E, l3=S, l4=N)
__device__ __forceinline__ void lap_row(const double *__restrict__ f, int slot, int y, const int4 nb,
                                        double (&l)[8]) {
  double c[8], up[8], dn[8];
  load_row(f, slot, y, c);
  const double gW = nb.x >= 0 ? __ldg(f + (size_t)nb.x * 64 + y * 8 + 7) : c[0];
  const double gE = nb.y >= 0 ? __ldg(f + (size_t)nb.y * 64 + y * 8 + 0) : c[7];
  if (y < 7) load_row(f, slot, y + 1, up);
  else if (nb.w >= 0) load_row(f, nb.w, 0, up);
  else {
#pragma unroll
    for (int i = 0; i < 8; i++) up[i] = c[i];
  }
  if (y > 0) load_row(f, slot, y - 1, dn);
  else if (nb.z >= 0) load_row(f, nb.z, 7, dn);
  else {
#pragma unroll
    for (int i = 0; i < 8; i++) dn[i] = c[i];
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double e = i < 7 ? c[i + 1] : gE;
    const double w = i > 0 ? c[i - 1] : gW;
    l[i] = (((w + e) + dn[i]) + up[i]) - 4.0 * c[i];
  }
}

// tmp = fac*(div vel) - fac*chi*(div udef) - lap(pold);  pres = 0     (main.cpp:7011-7027)
__global__ void __launch_bounds__(NT)
pressure_rhs_kernel(const double *__restrict__ vel, const double *__restrict__ udef,
                    const double *__restrict__ chi, const double *__restrict__ pold,
                    double *__restrict__ tmp, double *__restrict__ pres, const int4 *__restrict__ nbr,
                    int nrows, double fac) {
  for (int row = blockIdx.x * NT + threadIdx.x; row < nrows; row += gridDim.x * NT) {
    const int slot = row >> 3, y = row & 7;
    const int4 nb = __ldg(nbr + slot);
    double dv[8], du[8], lp[8], x[8], out[8];
    div_row(vel, slot, y, nb, dv);
    div_row(udef, slot, y, nb, du);
    load_row(chi, slot, y, x);
    lap_row(pold, slot, y, nb, lp);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = (fac * dv[i] - fac * x[i] * du[i]) - lp[i];
    store_row(tmp, slot, y, out);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = 0.0;
    store_row(pres, slot, y, out);
  }
}

int launch_pressure_rhs(cup2d_sim *s, double dt) {
  // pold <- pres is a pointer swap; the kernel then zeroes the new pres (main.cpp:7016-7021)
  swap_fields(s, CUP2D_PRES, CUP2D_POLD);
  if (s->nranks > 1) {
    int rc;
    if ((rc = halo_exchange_ptr(s, s->f[CUP2D_VEL], 2, CUP2D_VEL))) return rc;
    if ((rc = halo_exchange_ptr(s, s->f[CUP2D_TMPV], 2, CUP2D_TMPV))) return rc;
    if ((rc = halo_exchange_ptr(s, s->f[CUP2D_POLD], 1, CUP2D_POLD))) return rc;
  }
  const int nrows = (int)s->nloc * 8;
  const int grid = min((nrows + NT - 1) / NT, s->num_sms * 8);
  const double fac = 0.5 * s->h / dt; // main.cpp:6119
  ProfScope prof(s, KC_RHS);
  pressure_rhs_kernel<<<grid, NT, 0, s->stream>>>(s->f[CUP2D_VEL], s->f[CUP2D_TMPV], s->f[CUP2D_CHI],
                                                  s->f[CUP2D_POLD], s->f[CUP2D_TMP], s->f[CUP2D_PRES],
                                                  reinterpret_cast<const int4 *>(s->d_nbr), nrows, fac);
  s->launches++;
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

// pres = (x - avg) + pold ; vel += (-0.5 dt h) * grad(pres) / h^2        (main.cpp:7120-7187)
// x = Poisson solution (lives in a Krylov buffer), avg = its volume-weighted mean.  The reference's
// second mean (of the already mean-free field, main.cpp:7149-7166) is rounding noise and is dropped.
__global__ void __launch_bounds__(NT)
pressure_correct_kernel(const double *__restrict__ x, const double *__restrict__ pold,
                        double *__restrict__ pres, double *__restrict__ vel,
                        const int4 *__restrict__ nbr, int nrows, const double *__restrict__ xsum,
                        double inv_ncells, double pfac_ih2) {
  const double avg = xsum[0] * inv_ncells;
  for (int row = blockIdx.x * NT + threadIdx.x; row < nrows; row += gridDim.x * NT) {
    const int slot = row >> 3, y = row & 7;
    const int4 nb = __ldg(nbr + slot);
    double c[8], up[8], dn[8], t[8];
    // P = (x - avg) + pold on the row and its four neighbours
    load_row(x, slot, y, c);
    load_row(pold, slot, y, t);
#pragma unroll
    for (int i = 0; i < 8; i++) c[i] = (c[i] - avg) + t[i];
    double gW, gE;
    if (nb.x >= 0) gW = (__ldg(x + (size_t)nb.x * 64 + y * 8 + 7) - avg) + __ldg(pold + (size_t)nb.x * 64 + y * 8 + 7);
    else gW = c[0];
    if (nb.y >= 0) gE = (__ldg(x + (size_t)nb.y * 64 + y * 8) - avg) + __ldg(pold + (size_t)nb.y * 64 + y * 8);
    else gE = c[7];
    {
      int s2 = slot, y2 = y + 1;
      bool have = true;
      if (y == 7) { s2 = nb.w; y2 = 0; have = nb.w >= 0; }
      if (have) {
        load_row(x, s2, y2, up);
        load_row(pold, s2, y2, t);
#pragma unroll
        for (int i = 0; i < 8; i++) up[i] = (up[i] - avg) + t[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) up[i] = c[i];
      }
    }
    {
      int s2 = slot, y2 = y - 1;
      bool have = true;
      if (y == 0) { s2 = nb.z; y2 = 7; have = nb.z >= 0; }
      if (have) {
        load_row(x, s2, y2, dn);
        load_row(pold, s2, y2, t);
#pragma unroll
        for (int i = 0; i < 8; i++) dn[i] = (dn[i] - avg) + t[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) dn[i] = c[i];
      }
    }
    store_row(pres, slot, y, c);
    double4 *vp = reinterpret_cast<double4 *>(vel + (size_t)slot * 128 + y * 16);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double4 v = vp[k];
      const int i0 = 2 * k, i1 = 2 * k + 1;
      const double e0 = i0 < 7 ? c[i0 + 1] : gE, w0 = i0 > 0 ? c[i0 - 1] : gW;
      const double e1 = i1 < 7 ? c[i1 + 1] : gE, w1 = i1 > 0 ? c[i1 - 1] : gW;
      v.x = fma(pfac_ih2, e0 - w0, v.x);
      v.y = fma(pfac_ih2, up[i0] - dn[i0], v.y);
      v.z = fma(pfac_ih2, e1 - w1, v.z);
      v.w = fma(pfac_ih2, up[i1] - dn[i1], v.w);
      vp[k] = v;
    }
  }
}

int launch_pressure_correct(cup2d_sim *s, double dt) {
  const double *x = s->kx[s->h_state->opt];
  if (s->nranks > 1) {
    int rc;
    if ((rc = halo_exchange_ptr(s, s->kx[s->h_state->opt], 1, CUP2D_NFIELDS + 1 + s->h_state->opt))) return rc;
    if ((rc = halo_exchange_ptr(s, s->f[CUP2D_POLD], 1, CUP2D_POLD))) return rc;
  }
  const int nrows = (int)s->nloc * 8;
  const int grid = min((nrows + NT - 1) / NT, s->num_sms * 8);
  const double pfac = -0.5 * dt * s->h;          // main.cpp:6028
  const double ih2 = 1.0 / s->h / s->h;          // main.cpp:7182
  ProfScope prof(s, KC_CORRECT);
  pressure_correct_kernel<<<grid, NT, 0, s->stream>>>(
      x, s->f[CUP2D_POLD], s->f[CUP2D_PRES], s->f[CUP2D_VEL], reinterpret_cast<const int4 *>(s->d_nbr),
      nrows, &s->d_state->xsum, 1.0 / ((double)s->nglobal * 64.0), pfac * ih2);
  s->launches++;
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

} // namespace cup2d
