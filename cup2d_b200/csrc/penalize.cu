// Penalisation phase on device-resident fields (main.cpp:6643-6681, 6944-7002).
//
// A shape is what the reference keeps per shape in `obstacleBlocks` (main.cpp:3283-3286, 4245-4263): the list of
// blocks its body touches with, per block, its own chi X[8][8] and deformation velocity udef[8][8][2].  The host body
// model (out of scope) produces these every step; they are uploaded with cup2d_shape_set — a few hundred KB — and the
// velocity field never leaves the device:
//   cup2d_shape_integrals : the 7 sums {PM,PJ,PX,PY,UM,VM,AM} of main.cpp:6648-6679 (global across ranks); the 3x3
//                           solve for (u,v,omega) and the collision logic stay on the host
//   cup2d_penalize        : V = alpha V + (1-alpha)(u_s + omega x r + udef) where the shape owns the cell (6944-6979)
//   cup2d_udef_assemble   : tmpV = sum over shapes of udef where the shape's chi is not below the field's (6980-7002)
// One thread per cell of an obstacle block, 4 blocks per CTA.  The blend and the assembly use explicit
// round-to-nearest operations (no FMA contraction): they reproduce the reference bit for bit; the sums are reduced
// in a fixed order (deterministic, but not the reference's sequential order: 1e-13 relative).
#include "sim.h"

namespace cup2d {

constexpr int NT = 256;

struct ShapeView {
  const int *ids;
  const double *X, *udef;
  int nob;
};

// cell centre relative to the shape's centre of mass: p = origin + h (i + 0.5) - C   (main.cpp:6667-6670)
__device__ __forceinline__ void rel_pos(const int2 b, double h, int ix, int iy, double cx, double cy, double &px,
                                        double &py) {
  px = __dsub_rn(__dadd_rn(__dmul_rn((double)(b.x * CUP2D_BS), h), __dmul_rn(h, (double)ix + 0.5)), cx);
  py = __dsub_rn(__dadd_rn(__dmul_rn((double)(b.y * CUP2D_BS), h), __dmul_rn(h, (double)iy + 0.5)), cy);
}

// PART 0: {PM, PJ, PX, PY} (geometry)   PART 1: {UM, VM, AM} (momentum)
template <int PART>
__global__ void __launch_bounds__(NT)
shape_integrals_kernel(ShapeView sh, const double *__restrict__ vel, const int2 *__restrict__ ij, double h,
                       double lambdt, double cx, double cy, double *partials, unsigned int *counter, Comm comm,
                       double *out) {
  constexpr int NS = PART == 0 ? 4 : 3;
  double sums[NS];
#pragma unroll
  for (int k = 0; k < NS; k++) sums[k] = 0.0;
  const int cell = threadIdx.x & 63, ix = cell & 7, iy = cell >> 3;
  const double hsq = h * h;
  for (int k = blockIdx.x * (NT / 64) + (threadIdx.x >> 6); k < sh.nob; k += gridDim.x * (NT / 64)) {
    const double x = sh.X[(size_t)k * 64 + cell];
    if (x <= 0) continue;
    const int id = sh.ids[k];
    const double xl = x >= 0.5 ? lambdt : 0.0;
    const double F = hsq * xl / (1 + xl);
    double px, py;
    rel_pos(ij[id], h, ix, iy, cx, cy, px, py);
    if (PART == 0) {
      sums[0] += F;
      sums[1] += F * (px * px + py * py);
      sums[2] += F * px;
      sums[3] += F * py;
    } else {
      const double2 v = reinterpret_cast<const double2 *>(vel)[(size_t)id * 64 + cell];
      const double2 ud = reinterpret_cast<const double2 *>(sh.udef)[(size_t)k * 64 + cell];
      const double du = v.x - ud.x, dv = v.y - ud.y;
      sums[0] += F * du;
      sums[1] += F * dv;
      sums[2] += F * (px * dv - py * du);
    }
  }
  grid_reduce<NS, NT>(sums, 0.0, partials, counter, comm, [=](const double *t, double) {
    for (int k = 0; k < NS; k++) out[(PART == 0 ? 0 : 4) + k] = t[k];
  });
}

__global__ void __launch_bounds__(NT)
penalize_kernel(ShapeView sh, double *__restrict__ vel, const double *__restrict__ chi,
                const int2 *__restrict__ ij, double h, double inv1lam, double cx, double cy, double us, double vs,
                double omega) {
  const int cell = threadIdx.x & 63, ix = cell & 7, iy = cell >> 3;
  for (int k = blockIdx.x * (NT / 64) + (threadIdx.x >> 6); k < sh.nob; k += gridDim.x * (NT / 64)) {
    const double x = sh.X[(size_t)k * 64 + cell];
    const int id = sh.ids[k];
    if (chi[(size_t)id * 64 + cell] > x || x <= 0) continue;
    double px, py;
    rel_pos(ij[id], h, ix, iy, cx, cy, px, py);
    const double alpha = x > 0.5 ? inv1lam : 1.0, beta = __dsub_rn(1.0, alpha);
    const double2 ud = reinterpret_cast<const double2 *>(sh.udef)[(size_t)k * 64 + cell];
    const double US = __dadd_rn(__dsub_rn(us, __dmul_rn(omega, py)), ud.x);
    const double VS = __dadd_rn(__dadd_rn(vs, __dmul_rn(omega, px)), ud.y);
    double2 *vp = reinterpret_cast<double2 *>(vel) + (size_t)id * 64 + cell;
    double2 v = *vp;
    v.x = __dadd_rn(__dmul_rn(alpha, v.x), __dmul_rn(beta, US));
    v.y = __dadd_rn(__dmul_rn(alpha, v.y), __dmul_rn(beta, VS));
    *vp = v;
  }
}

__global__ void __launch_bounds__(NT)
udef_add_kernel(ShapeView sh, double *__restrict__ tmpv, const double *__restrict__ chi) {
  const int cell = threadIdx.x & 63;
  for (int k = blockIdx.x * (NT / 64) + (threadIdx.x >> 6); k < sh.nob; k += gridDim.x * (NT / 64)) {
    const int id = sh.ids[k];
    if (sh.X[(size_t)k * 64 + cell] < chi[(size_t)id * 64 + cell]) continue;
    const double2 ud = reinterpret_cast<const double2 *>(sh.udef)[(size_t)k * 64 + cell];
    double2 *tp = reinterpret_cast<double2 *>(tmpv) + (size_t)id * 64 + cell;
    double2 t = *tp;
    t.x += ud.x;
    t.y += ud.y;
    *tp = t;
  }
}

static ShapeView view_of(const cup2d_sim::Shape &sh) { return ShapeView{sh.d_ids, sh.d_X, sh.d_udef, sh.nob}; }
static int ob_grid(const cup2d_sim *s, int nob) {
  const int g = (nob + NT / 64 - 1) / (NT / 64);
  return g < 1 ? 1 : (g > s->num_sms * 8 ? s->num_sms * 8 : g);
}

int ensure_block_ij(cup2d_sim *s) {
  if (s->d_ij) return CUP2D_OK;
  CUP2D_CUDA(cudaMalloc(&s->d_ij, (size_t)s->nloc * 2 * sizeof(int)));
  CUP2D_CUDA(cudaMemcpy(s->d_ij, s->ij.data() + 2 * s->gbegin, (size_t)s->nloc * 2 * sizeof(int),
                        cudaMemcpyHostToDevice));
  return CUP2D_OK;
}

int shape_set(cup2d_sim *s, int shape, int nob, const int32_t *ids, const double *X, const double *udef) {
  if ((int)s->shapes.size() <= shape) s->shapes.resize(shape + 1);
  cup2d_sim::Shape &sh = s->shapes[shape];
  if (nob > sh.cap) {
    cudaFree(sh.d_ids); cudaFree(sh.d_X); cudaFree(sh.d_udef);
    sh.d_ids = nullptr; sh.d_X = sh.d_udef = nullptr;
    sh.cap = 0;
    const int cap = nob + nob / 4 + 16; // the body moves: leave room so that most steps do not reallocate
    CUP2D_CUDA(cudaMalloc(&sh.d_ids, (size_t)cap * sizeof(int)));
    CUP2D_CUDA(cudaMalloc(&sh.d_X, (size_t)cap * 64 * sizeof(double)));
    CUP2D_CUDA(cudaMalloc(&sh.d_udef, (size_t)cap * 128 * sizeof(double)));
    sh.cap = cap;
  }
  sh.nob = nob;
  if (nob > 0) {
    CUP2D_CUDA(cudaMemcpyAsync(sh.d_ids, ids, (size_t)nob * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CUP2D_CUDA(cudaMemcpyAsync(sh.d_X, X, (size_t)nob * 64 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    CUP2D_CUDA(cudaMemcpyAsync(sh.d_udef, udef, (size_t)nob * 128 * sizeof(double), cudaMemcpyHostToDevice, s->stream));
  }
  return CUP2D_OK;
}

void shapes_free(cup2d_sim *s) {
  for (auto &sh : s->shapes) {
    cudaFree(sh.d_ids); cudaFree(sh.d_X); cudaFree(sh.d_udef);
  }
  s->shapes.clear();
}

int shape_integrals(cup2d_sim *s, int shape, double lambda, double dt, double cx, double cy, double *out) {
  int rc = ensure_block_ij(s);
  if (rc) return rc;
  const ShapeView v = view_of(s->shapes[shape]);
  const int grid = ob_grid(s, v.nob);
  const int2 *ij = reinterpret_cast<const int2 *>(s->d_ij);
  // ranks that hold no block of this shape still take part in the cross-rank sum (grid of one idle CTA)
  shape_integrals_kernel<0><<<grid, NT, 0, s->stream>>>(v, s->f[CUP2D_VEL], ij, s->h, lambda * dt, cx, cy,
                                                       s->d_partials, s->d_counter, s->comm, s->d_scal);
  shape_integrals_kernel<1><<<grid, NT, 0, s->stream>>>(v, s->f[CUP2D_VEL], ij, s->h, lambda * dt, cx, cy,
                                                       s->d_partials, s->d_counter, s->comm, s->d_scal);
  s->launches += 2;
  CUP2D_CUDA(cudaGetLastError());
  CUP2D_CUDA(cudaMemcpyAsync(s->h_scal, s->d_scal, 7 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
  CUP2D_CUDA(cudaStreamSynchronize(s->stream));
  for (int k = 0; k < 7; k++) out[k] = s->h_scal[k];
  return CUP2D_OK;
}

int shape_penalize(cup2d_sim *s, int shape, double lambda, double dt, double cx, double cy, double us, double vs,
                   double omega) {
  int rc = ensure_block_ij(s);
  if (rc) return rc;
  const ShapeView v = view_of(s->shapes[shape]);
  if (v.nob == 0) return CUP2D_OK;
  penalize_kernel<<<ob_grid(s, v.nob), NT, 0, s->stream>>>(v, s->f[CUP2D_VEL], s->f[CUP2D_CHI],
                                                           reinterpret_cast<const int2 *>(s->d_ij), s->h,
                                                           1 / (1 + lambda * dt), cx, cy, us, vs, omega);
  s->launches++;
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

int udef_assemble(cup2d_sim *s) {
  CUP2D_CUDA(cudaMemsetAsync(s->f[CUP2D_TMPV], 0, (size_t)s->nloc * 128 * sizeof(double), s->stream));
  for (const auto &sh : s->shapes) {
    if (sh.nob == 0) continue;
    udef_add_kernel<<<ob_grid(s, sh.nob), NT, 0, s->stream>>>(view_of(sh), s->f[CUP2D_TMPV], s->f[CUP2D_CHI]);
    s->launches++;
  }
  CUP2D_CUDA(cudaGetLastError());
  return CUP2D_OK;
}

} // namespace cup2d
