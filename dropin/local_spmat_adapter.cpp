// LocalSpMatDnVec on the B200 hot path: link THIS object (+ libcup2d_b200.so) instead of the reference's
// cuda.o and the unmodified main.cpp runs its pressure solves through cup2d_b200.
//
// It implements the class exactly as the reference declares it (cuda.h:26-79 is included from the
// reference tree at build time; nothing of it is copied here), so main.cpp's call sites
//   ctor main.cpp:6489 | reserve 7038 | cooPushBackVal 7075-7087 | cooPushBackRow 7109 | make 7113 |
//   solveWithUpdate 7115 | solveNoUpdate 7118 | get_x/get_b/get_h2 6002-6004, 7122
// bind unchanged.  What differs from cuda.cu:549-699 is what happens behind them:
//   * the pushed COO is not shipped to the device as a matrix; `make` splits it into the same-level
//     5-point stencil (a block neighbour table, applied matrix-free) plus the rows that are anything
//     else (coarse-fine interpolation rows on AMR grids, main.cpp:5915-5997) in a small CSR side table;
//   * solve* = upload b_, x_  ->  cup2d_poisson_solve (matrix-free BiCGSTAB, same stopping rule as
//     cuda.cu:403-548)  ->  download x_.
// Single rank (the image has no MPI; multi-rank integration goes through cup2d_create + the patched
// time loop of INTEGRATION.md §2, which also removes the per-solve PCIe round trip).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mpi.h>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>
#include "cuda.h"           // the reference's declaration (-I<reference tree>)
#include "cup2d_b200.h"     // the C ABI (-I<repo>/include)

class BiCGSTABSolver { // the reference forward-declares this name as the solver pimpl (cuda.h:25)
public:
  explicit BiCGSTABSolver(int blen) : blen_(blen) {}
  ~BiCGSTABSolver() { cup2d_destroy(sim_); }
  void set_matrix(std::vector<int32_t> nbr, std::vector<int32_t> irr_rows, std::vector<int32_t> irr_rowptr,
                  std::vector<int32_t> irr_col, std::vector<double> irr_val) {
    nbr_ = std::move(nbr);
    irr_rows_ = std::move(irr_rows);
    irr_rowptr_ = std::move(irr_rowptr);
    irr_col_ = std::move(irr_col);
    irr_val_ = std::move(irr_val);
    fprintf(stderr, "cup2d_b200 adapter: %zu blocks, %zu rows through the CSR side table (%.2f%%)\n", nbr_.size() / 4,
            irr_rows_.size(), 100.0 * irr_rows_.size() / (16.0 * nbr_.size()));
  }
  void rebuild() {
    cup2d_destroy(sim_);
    sim_ = nullptr;
    const char *dev = getenv("CUP2D_DEVICE");
    if (cup2d_poisson_create_general((int64_t)nbr_.size() / 4, nbr_.data(), (int64_t)irr_rows_.size(), irr_rows_.data(),
                                     irr_rowptr_.data(), irr_col_.data(), irr_val_.data(), dev ? atoi(dev) : 0, &sim_))
      throw std::runtime_error(std::string("cup2d_poisson_create_general: ") + cup2d_last_error());
  }
  // host-only self check (CUP2D_ADAPTER_CHECK=1): the split representation applied to a pseudo-random
  // vector must reproduce the pushed COO exactly; prints the result and exits (no GPU needed)
  void host_check(int m, const std::vector<int> &rp, const int *col, const double *val) const {
    std::vector<double> z(m), y0(m), y1(m);
    unsigned long long st = 88172645463325252ULL;
    for (int i = 0; i < m; i++) {
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      z[i] = (double)(st % 2000001ULL) / 1e6 - 1.0;
    }
    for (int r = 0; r < m; r++) {
      double s = 0;
      for (int k = rp[r]; k < rp[r + 1]; k++) s += val[k] * z[col[k]];
      y0[r] = s;
    }
    const int B = blen_, BS = CUP2D_BS;
    for (int r = 0; r < m; r++) {
      const int b = r / B, lr = r % B, x = lr % BS, y = lr / BS;
      const int32_t *nb = &nbr_[(size_t)b * 4];
      const double c = z[r];
      const double w = x > 0 ? z[r - 1] : (nb[0] >= 0 ? z[nb[0] * B + y * BS + BS - 1] : c);
      const double e = x < BS - 1 ? z[r + 1] : (nb[1] >= 0 ? z[nb[1] * B + y * BS] : c);
      const double s = y > 0 ? z[r - BS] : (nb[2] >= 0 ? z[nb[2] * B + (BS - 1) * BS + x] : c);
      const double n = y < BS - 1 ? z[r + BS] : (nb[3] >= 0 ? z[nb[3] * B + x] : c);
      y1[r] = (((s + w) + e) + n) - 4.0 * c;
    }
    for (size_t k = 0; k < irr_rows_.size(); k++) {
      double s = 0;
      for (int j = irr_rowptr_[k]; j < irr_rowptr_[k + 1]; j++) s += irr_val_[j] * z[irr_col_[j]];
      y1[irr_rows_[k]] = s;
    }
    double md = 0;
    for (int r = 0; r < m; r++) md = std::max(md, std::fabs(y0[r] - y1[r]));
    fprintf(stderr, "cup2d_b200 adapter check: rows %d, general rows %zu, max |COO z - (stencil+CSR) z| = %.3e\n", m,
            irr_rows_.size(), md);
    if (getenv("CUP2D_ADAPTER_CHECK")[0] == '2') exit(md < 1e-12 ? 0 : 3);
  }
  void solve(std::vector<double> &x, const std::vector<double> &b, double tol, double rtol, int restarts) {
    int iters = 0;
    double err = 0;
    if (!sim_) throw std::runtime_error("LocalSpMatDnVec: solveNoUpdate before the first solveWithUpdate");
    if (cup2d_field_upload(sim_, CUP2D_TMP, b.data()) || cup2d_field_upload(sim_, CUP2D_PRES, x.data()) ||
        cup2d_poisson_solve(sim_, tol, rtol, restarts, 1000 /* cuda.cu:438 */, &iters, &err) ||
        cup2d_field_download(sim_, CUP2D_PRES, x.data()))
      throw std::runtime_error(std::string("cup2d_b200: ") + cup2d_last_error());
    last_iters = iters;
    last_err = err;
  }
  int last_iters = 0;
  double last_err = 0;

private:
  int blen_;
  cup2d_sim *sim_ = nullptr;
  std::vector<int32_t> nbr_, irr_rows_, irr_rowptr_, irr_col_;
  std::vector<double> irr_val_;
};

// same diagnostics symbols the oracle harness reads (oracle/ref_harness.cpp)
int cup2d_ref_last_iters = 0;
double cup2d_ref_last_err = 0;
int cup2d_ref_force_iters = -1; // not honoured: the product runs the reference's 1000-iteration cap
int cup2d_ref_fixed_iters = 1000;

LocalSpMatDnVec::LocalSpMatDnVec(MPI_Comm m_comm, const int BLEN, const bool bMeanConstraint,
                                 const std::vector<double> &)
    : m_comm_(m_comm), BLEN_(BLEN) {
  MPI_Comm_rank(m_comm_, &rank_);
  MPI_Comm_size(m_comm_, &comm_size_);
  if (comm_size_ != 1 || bMeanConstraint || BLEN != CUP2D_BS * CUP2D_BS) {
    fprintf(stderr, "cup2d_b200 adapter: single rank, 8x8 blocks, bMeanConstraint=0 only\n");
    abort();
  }
  // P_inv (main.cpp:6451-6488) is not needed: the preconditioner is applied by fast diagonalisation
  solver_ = std::make_unique<BiCGSTABSolver>(BLEN);
}
LocalSpMatDnVec::~LocalSpMatDnVec() {}

void LocalSpMatDnVec::reserve(const int N) {
  m_ = N;
  bMeanRow_ = -1;
  loc_cooValA_.clear();
  loc_cooRowA_long_.clear();
  loc_cooColA_long_.clear();
  loc_cooValA_.reserve(6 * (size_t)N);
  loc_cooRowA_long_.reserve(6 * (size_t)N);
  loc_cooColA_long_.reserve(6 * (size_t)N);
  x_.resize(N);
  b_.resize(N);
  h2_.resize(N / BLEN_);
}
void LocalSpMatDnVec::cooPushBackVal(const double val, const long long row, const long long col) {
  loc_cooValA_.push_back(val);
  loc_cooRowA_long_.push_back(row);
  loc_cooColA_long_.push_back(col);
}
void LocalSpMatDnVec::cooPushBackRow(const SpRowInfo &row) {
  for (const auto &i : row.loc_colval_) {
    loc_cooValA_.push_back(i.second);
    loc_cooRowA_long_.push_back(row.idx_);
    loc_cooColA_long_.push_back(i.first);
  }
  if (!row.neirank_cols_.empty()) {
    fprintf(stderr, "cup2d_b200 adapter: off-rank columns at size 1\n");
    abort();
  }
}

// Split the pushed COO into "same-level 5-point stencil described by a block neighbour table" + "general
// rows in CSR".  Nothing about the reference's row construction is assumed: a face of a block is served
// by the stencil only if all 8 of its cells carry exactly one weight-1 entry to the facing cell of one and
// the same other block; a row is served by the stencil only if its entries equal the stencil row that the
// neighbour table implies.  Everything else (coarse-fine interpolation rows main.cpp:5915-5997, or anything
// a future assembly pushes) goes to the CSR side table and is applied verbatim by the solver.
void LocalSpMatDnVec::make(const std::vector<long long> &Nrows_xcumsum) {
  loc_nnz_ = (int)loc_cooValA_.size();
  bd_nnz_ = 0;
  halo_ = 0;
  const long long shift = -Nrows_xcumsum[rank_];
  const int B = BLEN_, nb = m_ / B, BS = CUP2D_BS;
  // row pointers (rows are pushed contiguously and in order, main.cpp:7051-7111)
  std::vector<int> rp(m_ + 1, 0);
  loc_cooColA_int_.resize(loc_nnz_);
  for (int k = 0; k < loc_nnz_; k++) {
    const long long r = loc_cooRowA_long_[k] + shift, c = loc_cooColA_long_[k] + shift;
    if (r < 0 || r >= m_ || c < 0 || c >= m_ || (k > 0 && loc_cooRowA_long_[k] < loc_cooRowA_long_[k - 1])) {
      fprintf(stderr, "cup2d_b200 adapter: COO row/column outside the local system or rows not in order\n");
      abort();
    }
    rp[r + 1]++;
    loc_cooColA_int_[k] = (int)c;
  }
  for (int r = 0; r < m_; r++) rp[r + 1] += rp[r];
  const int *col = loc_cooColA_int_.data();
  const double *val = loc_cooValA_.data();
  // facing cell of (block-local cell lr) across face d, as a block-local index; -1 if lr is not on face d
  auto facing = [&](int lr, int d) -> int {
    const int x = lr % BS, y = lr / BS;
    switch (d) {
    case 0: return x == 0 ? y * BS + (BS - 1) : -1;
    case 1: return x == BS - 1 ? y * BS : -1;
    case 2: return y == 0 ? (BS - 1) * BS + x : -1;
    default: return y == BS - 1 ? x : -1;
    }
  };
  // 1. faces
  std::vector<int32_t> nbr((size_t)nb * 4, -1);
  for (int b = 0; b < nb; b++)
    for (int d = 0; d < 4; d++) {
      int cand = -2; // -2 unset, -1 face has no stencil neighbour
      for (int lr = 0; lr < B && cand != -1; lr++) {
        const int f = facing(lr, d);
        if (f < 0) continue;
        const int r = b * B + lr;
        int found = -1;
        for (int k = rp[r]; k < rp[r + 1]; k++) {
          const int bc = col[k] / B;
          if (bc != b && col[k] % B == f && val[k] == 1.0) found = found == -1 ? bc : -3; // two candidates: ambiguous
        }
        if (found < 0) cand = -1;
        else if (cand == -2) cand = found;
        else if (cand != found) cand = -1;
      }
      nbr[(size_t)b * 4 + d] = cand < 0 ? -1 : cand;
    }
  // 2. rows: regular iff entries == stencil row implied by nbr
  std::vector<int32_t> irr_rows, irr_rowptr(1, 0), irr_col;
  std::vector<double> irr_val;
  for (int r = 0; r < m_; r++) {
    const int b = r / B, lr = r % B, x = lr % BS, y = lr / BS;
    int expect[4], ne = 0;
    if (x > 0) expect[ne++] = r - 1; else if (nbr[(size_t)b * 4 + 0] >= 0) expect[ne++] = nbr[(size_t)b * 4 + 0] * B + facing(lr, 0);
    if (x < BS - 1) expect[ne++] = r + 1; else if (nbr[(size_t)b * 4 + 1] >= 0) expect[ne++] = nbr[(size_t)b * 4 + 1] * B + facing(lr, 1);
    if (y > 0) expect[ne++] = r - BS; else if (nbr[(size_t)b * 4 + 2] >= 0) expect[ne++] = nbr[(size_t)b * 4 + 2] * B + facing(lr, 2);
    if (y < BS - 1) expect[ne++] = r + BS; else if (nbr[(size_t)b * 4 + 3] >= 0) expect[ne++] = nbr[(size_t)b * 4 + 3] * B + facing(lr, 3);
    bool regular = (rp[r + 1] - rp[r]) == ne + 1;
    int hits = 0;
    for (int k = rp[r]; k < rp[r + 1] && regular; k++) {
      if (col[k] == r) regular = val[k] == -(double)ne;
      else {
        bool ok = false;
        for (int e = 0; e < ne; e++) ok |= expect[e] == col[k];
        regular = ok && val[k] == 1.0;
        hits++;
      }
    }
    regular = regular && hits == ne;
    if (!regular) {
      irr_rows.push_back(r);
      for (int k = rp[r]; k < rp[r + 1]; k++) {
        irr_col.push_back(col[k]);
        irr_val.push_back(val[k]);
      }
      irr_rowptr.push_back((int32_t)irr_col.size());
    }
  }
  // a face whose cells are not all regular must not feed the stencil of the regular ones either way:
  // regular rows only ever read through nbr, so nothing else to do.
  solver_->set_matrix(std::move(nbr), std::move(irr_rows), std::move(irr_rowptr), std::move(irr_col), std::move(irr_val));
  if (getenv("CUP2D_ADAPTER_CHECK")) solver_->host_check(m_, rp, col, val);
}
void LocalSpMatDnVec::solveWithUpdate(const double max_error, const double max_rel_error,
                                      const int max_restarts) {
  solver_->rebuild();
  solveNoUpdate(max_error, max_rel_error, max_restarts);
}
void LocalSpMatDnVec::solveNoUpdate(const double max_error, const double max_rel_error,
                                    const int max_restarts) {
  solver_->solve(x_, b_, max_error, max_rel_error, max_restarts);
  cup2d_ref_last_iters = solver_->last_iters;
  cup2d_ref_last_err = solver_->last_err;
}
