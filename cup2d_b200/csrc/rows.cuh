// Warp-cooperative row access for the block-structured kernels.
//
// Compute mapping ("row layout"): one lane owns one row of 8 cells of one 8x8 block (8 lanes = one
// block, one warp = 4 consecutive blocks = 32 rows = one CHUNK of 2 KB (scalar) / 4 KB (vector) of
// CONTIGUOUS memory).  If every lane read its own row with 128-bit loads, each load instruction would
// touch 16-32 different 128-B lines and the kernels become L1/TEX-bound (measured: 78-98 % l1tex
// throughput, profiles/r01_summary.md).  So global memory is only ever touched in "chunk layout": lane
// l holds the 16-byte pieces j*32+l (j = 0..3) of the chunk, i.e. fully coalesced 128-bit accesses.
// Element-wise work (axpy, dots, norms) is done directly in chunk layout; only stencils and the block
// preconditioner need row layout, and the two layouts are converted through a per-warp shared-memory
// scratch.  The scratch is unpadded and XOR-swizzled: piece p of row r lives at 16-byte slot
// 4r + (p ^ ((r>>1)&3)), which is bank-conflict-free both for the coalesced side (8 lanes = 2 rows x 4
// pieces) and for the row side (8 lanes = 8 consecutive rows, same piece).
#pragma once
#include "common.cuh"

namespace cup2d {

constexpr int RS2 = 9;                     // vector row stride in double2 (8 used + 1 pad = 144 B)
constexpr int ROWS_SCRATCH = 32 * RS2 * 2; // doubles per warp for kernels that touch vector fields
constexpr int SCR1 = 288;                  // doubles per warp for scalar-only kernels (preconditioner: 4*72)

__device__ __forceinline__ int swz(int r, int p) { return r * 4 + (p ^ ((r >> 1) & 3)); }

// ---- chunk layout <-> global ------------------------------------------------------------------
__device__ __forceinline__ void chunk_ld(const double *__restrict__ f, int row0, int nvalid, int lane,
                                         double2 (&c)[4]) {
  const double2 *src = reinterpret_cast<const double2 *>(f) + (size_t)row0 * 4;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int i = j * 32 + lane;
    c[j] = (i >> 2) < nvalid ? src[i] : make_double2(0.0, 0.0);
  }
}
__device__ __forceinline__ void chunk_st(double *__restrict__ f, int row0, int nvalid, int lane,
                                         const double2 (&c)[4]) {
  double2 *dst = reinterpret_cast<double2 *>(f) + (size_t)row0 * 4;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int i = j * 32 + lane;
    if ((i >> 2) < nvalid) dst[i] = c[j];
  }
}
// ---- chunk layout <-> row layout through the scratch -----------------------------------------------
// after chunk_to_rows the whole chunk stays parked in the scratch (rows_peek1 may read other rows)
__device__ __forceinline__ void chunk_to_rows(double *sw, int lane, const double2 (&c)[4], double (&r)[8]) {
  double2 *s2 = reinterpret_cast<double2 *>(sw);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int i = j * 32 + lane;
    s2[swz(i >> 2, i & 3)] = c[j];
  }
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const double2 v = s2[swz(lane, p)];
    r[2 * p] = v.x;
    r[2 * p + 1] = v.y;
  }
}
__device__ __forceinline__ void rows_to_chunk(double *sw, int lane, const double (&r)[8], double2 (&c)[4]) {
  double2 *s2 = reinterpret_cast<double2 *>(sw);
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 4; p++) s2[swz(lane, p)] = make_double2(r[2 * p], r[2 * p + 1]);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int i = j * 32 + lane;
    c[j] = s2[swz(i >> 2, i & 3)];
  }
}
__device__ __forceinline__ void rows_peek1(const double *sw, int r, double (&c)[8]) {
  const double2 *s2 = reinterpret_cast<const double2 *>(sw);
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const double2 v = s2[swz(r, p)];
    c[2 * p] = v.x;
    c[2 * p + 1] = v.y;
  }
}
// convenience: rows [row0, row0+nvalid) of a scalar field -> lane l gets row row0+l (zeros beyond)
__device__ __forceinline__ void rows_load1(const double *__restrict__ f, int row0, int nvalid,
                                           double *sw, int lane, double (&c)[8]) {
  double2 t[4];
  chunk_ld(f, row0, nvalid, lane, t);
  chunk_to_rows(sw, lane, t, c);
}
__device__ __forceinline__ void rows_store1(double *__restrict__ f, int row0, int nvalid, double *sw,
                                            int lane, const double (&c)[8]) {
  double2 t[4];
  rows_to_chunk(sw, lane, c, t);
  chunk_st(f, row0, nvalid, lane, t);
}

// ---- vector field (u,v interleaved): lane l gets row row0+l as 8 double2 (padded scratch) -----------
// chunk layout of a vector field: lane l holds pieces j*32+l, j = 0..7 (4 KB per warp, coalesced)
__device__ __forceinline__ void chunk2_ld(const double *__restrict__ f, int row0, int nvalid, int lane,
                                          double2 (&c)[8]) {
  const double2 *src = reinterpret_cast<const double2 *>(f) + (size_t)row0 * 8;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = j * 32 + lane;
    c[j] = (i >> 3) < nvalid ? src[i] : make_double2(0.0, 0.0);
  }
}
__device__ __forceinline__ void chunk2_st(double *__restrict__ f, int row0, int nvalid, int lane,
                                          const double2 (&c)[8]) {
  double2 *dst = reinterpret_cast<double2 *>(f) + (size_t)row0 * 8;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = j * 32 + lane;
    if ((i >> 3) < nvalid) dst[i] = c[j];
  }
}
__device__ __forceinline__ void chunk2_to_rows(double *sw, int nvalid, int lane, const double2 (&t)[8],
                                               double2 (&c)[8]) {
  double2 *s2 = reinterpret_cast<double2 *>(sw);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = j * 32 + lane;
    s2[(i >> 3) * RS2 + (i & 7)] = t[j];
  }
  __syncwarp();
  if (lane < nvalid) {
#pragma unroll
    for (int p = 0; p < 8; p++) c[p] = s2[lane * RS2 + p];
  } else {
#pragma unroll
    for (int p = 0; p < 8; p++) c[p] = make_double2(0.0, 0.0);
  }
}
__device__ __forceinline__ void rows2_to_chunk(double *sw, int lane, const double2 (&c)[8], double2 (&t)[8]) {
  double2 *s2 = reinterpret_cast<double2 *>(sw);
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 8; p++) s2[lane * RS2 + p] = c[p];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = j * 32 + lane;
    t[j] = s2[(i >> 3) * RS2 + (i & 7)];
  }
}
__device__ __forceinline__ void rows_load2(const double *__restrict__ f, int row0, int nvalid,
                                           double *sw, int lane, double2 (&c)[8]) {
  const double2 *src = reinterpret_cast<const double2 *>(f) + (size_t)row0 * 8;
  double2 *s2 = reinterpret_cast<double2 *>(sw);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = j * 32 + lane;
    if ((i >> 3) < nvalid) s2[(i >> 3) * RS2 + (i & 7)] = src[i];
  }
  __syncwarp();
  if (lane < nvalid) {
#pragma unroll
    for (int p = 0; p < 8; p++) c[p] = s2[lane * RS2 + p];
  } else {
#pragma unroll
    for (int p = 0; p < 8; p++) c[p] = make_double2(0.0, 0.0);
  }
}
__device__ __forceinline__ void rows_peek2(const double *sw, int r, double2 (&c)[8]) {
  const double2 *s2 = reinterpret_cast<const double2 *>(sw);
#pragma unroll
  for (int p = 0; p < 8; p++) c[p] = s2[r * RS2 + p];
}
__device__ __forceinline__ void rows_store2(double *__restrict__ f, int row0, int nvalid, double *sw,
                                            int lane, const double2 (&c)[8]) {
  double2 *dst = reinterpret_cast<double2 *>(f) + (size_t)row0 * 8;
  double2 *s2 = reinterpret_cast<double2 *>(sw);
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 8; p++) s2[lane * RS2 + p] = c[p];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = j * 32 + lane;
    if ((i >> 3) < nvalid) dst[i] = s2[(i >> 3) * RS2 + (i & 7)];
  }
}
// plain per-lane global row access (used only by the few lanes that touch a NEIGHBOUR block's edge row)
__device__ __forceinline__ void grow_load1(const double *__restrict__ f, int slot, int y, double (&c)[8]) {
  const double2 *p = reinterpret_cast<const double2 *>(f + (size_t)slot * 64 + y * 8);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double2 v = p[k];
    c[2 * k] = v.x;
    c[2 * k + 1] = v.y;
  }
}
__device__ __forceinline__ void grow_load2(const double *__restrict__ f, int slot, int y, double2 (&c)[8]) {
  const double2 *p = reinterpret_cast<const double2 *>(f + (size_t)slot * 128 + y * 16);
#pragma unroll
  for (int k = 0; k < 8; k++) c[k] = p[k];
}

// Rows of the Poisson matrix that are NOT the same-level 5-point stencil (coarse-fine interpolation rows
// of main.cpp:5915-5997 pushed through cooPushBackRow).  They live in a small
// CSR side table and override the stencil result of the cells they belong to: blk[slot] = -1 (block is
// fully regular) or k, tab[k*64 + cell] = -1 or the CSR row index.
struct IrrView {
  const int *blk = nullptr;
  const int *tab = nullptr;
  const int *rowptr = nullptr;
  const int *col = nullptr;
  const double *val = nullptr;
};

// Neighbour rows that live in HALO slots (slot >= nloc) of a multi-rank context.  In the Krylov loop the producing
// kernel of a vector pushes its perimeter rows straight into the neighbours' halo slots over NVLink and raises a flag
// (poisson.cu: push_rows / push_finish); the consuming stencil kernel waits for the flags of the ranks it has halo
// blocks of — lazily, only in the lanes that touch a halo slot, so that interior rows never wait — and then reads the
// slots with coherent loads (the data arrived while this kernel was running).
struct NoGate {
  static constexpr bool on = false;
  int nloc = 0;
  __device__ __forceinline__ void wait() const {}
};
struct HaloGate {
  static constexpr bool on = true;
  int nloc;                  // slots >= nloc are halo slots
  unsigned mask;             // ranks this rank has halo blocks of
  unsigned long long target; // push epoch the flags must have reached
  Comm comm;
  __device__ __forceinline__ void wait() const {
    const unsigned long long *mine = comm.mb[comm.rank];
    for (int r = 0; r < comm.nranks; r++)
      if ((mask >> r) & 1u) wait_flag(mine + MB_PUSHED + r, target, comm, CW_PUSHED, r);
  }
};
template <class G> __device__ __forceinline__ double nb_ld1(const double *z, int slot, int off, const G &g) {
  const double *p = z + (size_t)slot * 64 + off;
  if (G::on && slot >= g.nloc) return ld_coherent(p);
  return *p;
}
template <class G> __device__ __forceinline__ void nb_row1(const double *z, int slot, int y, double (&c)[8], const G &g) {
  const double2 *p = reinterpret_cast<const double2 *>(z + (size_t)slot * 64 + y * 8);
  const bool halo = G::on && slot >= g.nloc;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double2 v = halo ? ld_coherent2(p + k) : p[k];
    c[2 * k] = v.x;
    c[2 * k + 1] = v.y;
  }
}

// Undivided 5-point Laplacian rows of a scalar field for the warp's 32 rows, ghost = the cell itself at
// a domain wall (Neumann rows of main.cpp:7100-7107 / ScalarLab::Neumann2D main.cpp:3210-3245).
// Returns the own row in c and the Laplacian in out.  Summation order S,W,E,N then -4C.
// rows_lap_c: the caller already holds the chunk of z (so that other global loads can be in flight too)
template <class G = NoGate>
__device__ __forceinline__ void rows_lap_c(const double2 (&cz)[4], const double *__restrict__ z, int row0,
                                           int nvalid, const int4 *__restrict__ nbr, double *sw, int lane,
                                           double (&c)[8], double (&out)[8], const IrrView irr = IrrView(),
                                           const G gate = G());
__device__ __forceinline__ void rows_lap(const double *__restrict__ z, int row0, int nvalid,
                                         const int4 *__restrict__ nbr, double *sw, int lane,
                                         double (&c)[8], double (&out)[8], const IrrView irr = IrrView()) {
  double2 cz[4];
  chunk_ld(z, row0, nvalid, lane, cz);
  rows_lap_c(cz, z, row0, nvalid, nbr, sw, lane, c, out, irr);
}
template <class G>
__device__ __forceinline__ void rows_lap_c(const double2 (&cz)[4], const double *__restrict__ z, int row0,
                                           int nvalid, const int4 *__restrict__ nbr, double *sw, int lane,
                                           double (&c)[8], double (&out)[8], const IrrView irr, const G gate) {
  chunk_to_rows(sw, lane, cz, c);
  const int row = row0 + lane, slot = row >> 3, y = row & 7;
  if (lane < nvalid) {
    const int4 nb = nbr[slot];
    if (G::on) { // pushed halo rows: wait (once, only here) for the ranks that own them
      if (nb.x >= gate.nloc || nb.y >= gate.nloc || (y == 7 && nb.w >= gate.nloc) || (y == 0 && nb.z >= gate.nloc)) gate.wait();
    }
    double up[8], dn[8];
    if (y < 7) rows_peek1(sw, lane + 1, up);
    else if (nb.w >= 0) nb_row1(z, nb.w, 0, up, gate);
    else {
#pragma unroll
      for (int i = 0; i < 8; i++) up[i] = c[i];
    }
    if (y > 0) rows_peek1(sw, lane - 1, dn);
    else if (nb.z >= 0) nb_row1(z, nb.z, 7, dn, gate);
    else {
#pragma unroll
      for (int i = 0; i < 8; i++) dn[i] = c[i];
    }
    const double gW = nb.x >= 0 ? nb_ld1(z, nb.x, y * 8 + 7, gate) : c[0];
    const double gE = nb.y >= 0 ? nb_ld1(z, nb.y, y * 8 + 0, gate) : c[7];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const double e = i < 7 ? c[i + 1] : gE;
      const double w = i > 0 ? c[i - 1] : gW;
      out[i] = (((dn[i] + w) + e) + up[i]) - 4.0 * c[i];
    }
    if (irr.blk) { // general rows override the stencil (rare: block faces at coarse-fine interfaces)
      const int k = irr.blk[slot];
      if (k >= 0) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int idx = irr.tab[k * 64 + y * 8 + i];
          if (idx >= 0) {
            double acc = 0.0;
            for (int j = irr.rowptr[idx]; j < irr.rowptr[idx + 1]; j++) {
              const int cj = irr.col[j];
              acc = fma(irr.val[j], (G::on && cj >= gate.nloc * 64) ? ld_coherent(z + cj) : z[cj], acc);
            }
            out[i] = acc;
          }
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = 0.0;
  }
}

} // namespace cup2d
