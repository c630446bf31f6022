// TEST INFRASTRUCTURE (oracle) — not part of the product path.
//
// Operator-level harness around the UNMODIFIED reference translation unit.  It includes
// /root/reference/main.cpp verbatim (only `main` is renamed by a macro), lets the
// reference's own init code build a uniform 2^L x 2^L-block grid
// (-bpdx 1 -bpdy 1 -levelStart L -levelMax L+1 -Ctol 0), and takes control
// at the first MPI_Allreduce(MPI_MAX) of the time loop (main.cpp:6592) through the hook of
// oracle/shim/mpi.h.  From there it calls the reference's own operators
// (computeA<VectorLab>(KernelAdvectDiffuse(), ...) etc., main.cpp:6616, 7011, 7026, 7178)
// on seeded fields, or lets the reference time loop run and records every step.
//
// All files are raw little-endian doubles in GLOBAL row-major cell order
// (index = iy*N + ix, N = 8*2^L), so they do not depend on the block ordering.
//
//   ref_harness order L out.bin
//        out: int32 pairs (i,j) of every block in reference `infos` order (Hilbert id order)
//   ref_harness ops   L nu dt in.bin out.bin
//        in : u v p chi udef_u udef_v            (6 N^2 doubles)
//        out: K(vel)_u K(vel)_v  rhs  rhs1  gradp_u gradp_v   (6 N^2 doubles)
//             K = KernelAdvectDiffuse output tmpV (undivided), rhs = pressure_rhs(vel,udef,chi),
//             rhs1 = rhs - lap(p) (pressure_rhs1 with pold = p), gradp = pressureCorrectionKernel(p)
//   ref_harness vort  L in.bin out.bin
//        in : as ops;  out: KernelVorticity(vel) written to tmp (N^2 doubles) = adapt()'s tagging field
//   ref_harness tags  L rtol extra in.bin out.bin
//        adapt()'s first two lines (main.cpp:4659-4660) with sim.Rtol = rtol and levelMax = L+1+extra
//        (extra = 0: the grid is at the finest level -> GradChiOnTmp looks 4 cells around a block, else 2);
//        out: tmp (N^2 doubles) = vorticity with the 4 centre cells of body-adjacent blocks set to 2 Rtol
//   ref_harness dump  L time in.bin prefix
//        the reference's own dump() (main.cpp:3367-3467) of vel -> prefix.xyz.raw, prefix.attr.raw, prefix.xdmf2
//   ref_harness penal L nsteps kiter out.bin
//        two self-propelled fish on a uniform level-L grid (extent 1, no regridding); records the penalisation phase
//        of every step (main.cpp:6643-6681, 6944-7002) as a stream of records [tag, n, n doubles]:
//        tag 1 (per shape, at the 7-sum of main.cpp:6681): step, k, lambda, dt, Cx, Cy, Q[7], nob, ids[nob],
//              chi[nob*64], udef[nob*128]      (the shape's Obstacle blocks, main.cpp:3283-3286)
//        tag 2 (once per step, shape 0's hook): u, v before penalisation (2 N^2)
//        tag 3 (at main.cpp:7138, i.e. after penalisation + solve, before the correction): per shape (u, v, omega),
//              then u, v after penalisation, tmpV = summed udef (2 N^2), chi field (N^2)
//   ref_harness amrlab levelMax nsteps out.bin
//        the run.sh case for nsteps steps (so that the mesh is genuinely multi-level), then seeded fields on that mesh
//        and the reference's own ghost assembly + operators on them (round-2 groundwork);
//        record stream [tag, n, n doubles]:  10: nu, dt, h0, bpdx, bpdy, levelMax   11: (level,i,j) per block
//        12: vel blocks (128/block)  13: pres = pold blocks  14: chi blocks  15: udef blocks (128/block)
//        20: VectorLab of vel, stencil {-3,-3,4,4,tensorial} (14*14*2/block)   21: VectorLab of vel {-1,-1,2,2} (10*10*2)
//        22: ScalarLab of pres {-1,-1,2,2} (10*10)
//   ref_harness adump levelMax nsteps prefix
//        the run.sh case for nsteps steps, then the reference's dump() (main.cpp:3367-3467) of its velocity on that mesh:
//        prefix.xdmf2 / .xyz.raw / .attr.raw, and prefix.bin = 10: time, h0, bpdx, bpdy  11: mesh  12: vel blocks
//   ref_harness atags levelMax nsteps out.bin
//        the run.sh case for nsteps steps, then what adapt() looks at (main.cpp:4676-4678) on the fields as they are:
//        10: Rtol, Ctol, levelMax, h0, bpdx, bpdy   11: mesh   12: vel blocks   14: chi blocks
//        23: ScalarLab of chi, stencil {-4,-4,5,5,tensorial} (16*16/block: GradChiOnTmp's lab)
//        40: tmp after KernelVorticity   41: tmp after GradChiOnTmp (what the per-block L-inf is taken of)
//        30: tmpV after KernelAdvectDiffuse + flux correction   31: tmp after pressure_rhs (+fc)
//        32: tmp after pressure_rhs1 (+fc)   33: tmpV after pressureCorrectionKernel (+fc)
//   ref_harness fsteps L nsteps kiter out.bin
//        two fish on a uniform level-L grid (the configuration of `penal`), recorded at the dt reduction of every step
//        (so it also works for the patched drivers): per step dt, then u v p (1 + 3 N^2 doubles), then per shape
//        centerOfMass[2], u, v, omega
//   ref_harness asteps levelMax nsteps kiter out.bin
//        the run.sh case (2 fish, regridding) recorded at the dt reduction of every step, so that it also works for the
//        patched multi-level driver: per step double dt, double nblocks, then (level,i,j) as doubles, vel blocks, pres blocks
//   ref_harness steps L nu cfl nsteps kiter in.bin out.bin
//        in : as above (udef ignored: no shapes => udef = 0, chi = 0 ; p = initial pres)
//        out: per step: dt, then u v p  b x   (1 + 5 N^2 doubles), b/x = Poisson rhs / solution
//   ref_harness amr   levelMax nsteps kiter out.bin
//        the reference's own run.sh case (2 fish, AMR from level 5 to levelMax-1); per step:
//        int64 nrows, int64 nblocks, double dt, b[nrows], x[nrows], int32 (level,i,j)[nblocks]
//   ref_harness time  L reps kiter
//        prints one JSON line with per-operator CPU times (seconds, median of reps)
#define CUP2D_REF_HOOK_TU 1
#define main ref_main
#if defined(CUP2D_PATCHED_MAIN) && CUP2D_PATCHED_MAIN == 4
#include "main_amrresident.cpp" // oracle/_ref/: multi-level and device-resident with bodies (oracle/Makefile, ref_amrresident)
#elif defined(CUP2D_PATCHED_MAIN) && CUP2D_PATCHED_MAIN == 3
#include "main_amrloop.cpp" // oracle/_ref/: the multi-level form (oracle/Makefile, ref_amrloop)
#elif defined(CUP2D_PATCHED_MAIN) && CUP2D_PATCHED_MAIN == 2
#include "main_resident.cpp" // oracle/_ref/: the device-resident form with bodies (oracle/Makefile, ref_resident)
#elif defined(CUP2D_PATCHED_MAIN)
#include "main_patched.cpp" // oracle/_ref/: main.cpp with its hot path spliced onto cup2d_b200 (oracle/Makefile, ref_patched)
#else
#include "main.cpp" // resolved with -I/root/reference
#endif
#undef main
#include <chrono>

extern int cup2d_ref_last_iters;
extern double cup2d_ref_last_err;
extern int cup2d_ref_force_iters;
extern int cup2d_ref_fixed_iters;

namespace {
enum Mode { ORDER, OPS, STEPS, TIME, AMR, VORT, TAGS, DUMP, PENAL, AMRLAB, FSTEPS, ASTEPS, ATAGS, ADUMP } g_mode;
int g_sum7 = 0, g_sum2 = 0, g_step = 0;
int g_extra = 0;
double g_rtol = 0, g_time = 0;
int g_L, g_N, g_NY, g_bx = 1, g_by = 1, g_nsteps, g_reps, g_kiter;
double g_nu, g_dt, g_cfl;
std::string g_in, g_out;
int g_calls = 0;
std::vector<double> g_input;
FILE *g_fout = nullptr;

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void read_input(int nfields) {
  g_input.resize((size_t)nfields * g_N * g_NY);
  FILE *f = fopen(g_in.c_str(), "rb");
  if (!f || fread(g_input.data(), sizeof(double), g_input.size(), f) != g_input.size()) {
    fprintf(stderr, "ref_harness: cannot read %s\n", g_in.c_str());
    exit(2);
  }
  fclose(f);
}
// scatter a global row-major field into a reference grid (dim 1 or 2 components)
void scatter(Grid *g, int dim, const double *c0, const double *c1) {
  for (auto &info : g->infos) {
    const int bi = info.index[0], bj = info.index[1];
    for (int iy = 0; iy < _BS_; iy++)
      for (int ix = 0; ix < _BS_; ix++) {
        size_t gidx = (size_t)(bj * _BS_ + iy) * g_N + (bi * _BS_ + ix);
        info.block[dim * (_BS_ * iy + ix)] = c0[gidx];
        if (dim == 2) info.block[dim * (_BS_ * iy + ix) + 1] = c1[gidx];
      }
  }
}
void gather(Grid *g, int dim, double *c0, double *c1) {
  for (auto &info : g->infos) {
    const int bi = info.index[0], bj = info.index[1];
    for (int iy = 0; iy < _BS_; iy++)
      for (int ix = 0; ix < _BS_; ix++) {
        size_t gidx = (size_t)(bj * _BS_ + iy) * g_N + (bi * _BS_ + ix);
        c0[gidx] = info.block[dim * (_BS_ * iy + ix)];
        if (dim == 2) c1[gidx] = info.block[dim * (_BS_ * iy + ix) + 1];
      }
  }
}
void write_field(Grid *g, int dim) {
  std::vector<double> a((size_t)g_N * g_NY), b((size_t)g_N * g_NY);
  gather(g, dim, a.data(), b.data());
  fwrite(a.data(), sizeof(double), a.size(), g_fout);
  if (dim == 2) fwrite(b.data(), sizeof(double), b.size(), g_fout);
}
void write_vec(const std::vector<double> &x) {
  // solver vectors are block-major in `infos` order (main.cpp:5753-5771): back to global order
  std::vector<double> a((size_t)g_N * g_NY, 0.0);
  auto &infos = var.tmp->infos;
  if (x.size() < infos.size() * _BS_ * _BS_) { // the patched loop never fills the host solver's vectors: zeros
    fwrite(a.data(), sizeof(double), a.size(), g_fout);
    return;
  }
  for (size_t i = 0; i < infos.size(); i++) {
    const int bi = infos[i].index[0], bj = infos[i].index[1];
    for (int iy = 0; iy < _BS_; iy++)
      for (int ix = 0; ix < _BS_; ix++)
        a[(size_t)(bj * _BS_ + iy) * g_N + (bi * _BS_ + ix)] = x[i * _BS_ * _BS_ + iy * _BS_ + ix];
  }
  fwrite(a.data(), sizeof(double), a.size(), g_fout);
}
double field_umax() {
  double umax = 0;
  for (auto &info : var.vel->infos)
    for (int j = 0; j < 2 * _BS_ * _BS_; j++) umax = std::max(umax, std::fabs(info.block[j]));
  return umax;
}
// the reference's own calling sequences (main.cpp:6611-6617, 7007-7013, 7022-7027, 7174-7179)
void call_advect() {
  if (var.tmpV->UpdateFluxCorrection) {
    prepare0(var.buf2, &var.tmpV->infos, &var.tmpV->all, &var.tmpV->tree, 2);
    var.tmpV->UpdateFluxCorrection = false;
  }
  computeA<VectorLab>(KernelAdvectDiffuse(), var.vel, 2);
  fillcases(var.buf2, &var.tmpV->tree, 2);
}
void call_rhs() {
  if (var.tmp->UpdateFluxCorrection) {
    prepare0(var.buf1, &var.tmp->infos, &var.tmp->all, &var.tmp->tree, 1);
    var.tmp->UpdateFluxCorrection = false;
  }
  computeB<pressure_rhs, VectorLab, VectorLab>(pressure_rhs(), var.vel, 2, var.tmpV, 2);
  fillcases(var.buf1, &var.tmp->tree, 1);
}
void call_rhs1() {
  if (var.tmp->UpdateFluxCorrection) {
    prepare0(var.buf1, &var.tmp->infos, &var.tmp->all, &var.tmp->tree, 1);
    var.tmp->UpdateFluxCorrection = false;
  }
  computeA<ScalarLab>(pressure_rhs1(), var.pold, 1);
  fillcases(var.buf1, &var.tmp->tree, 1);
}
void call_gradp() {
  if (var.tmp->UpdateFluxCorrection) {
    prepare0(var.buf1, &var.tmp->infos, &var.tmp->all, &var.tmp->tree, 1);
    var.tmp->UpdateFluxCorrection = false;
  }
  computeA<ScalarLab>(pressureCorrectionKernel(), var.pres, 1);
  fillcases(var.buf1, &var.tmp->tree, 1);
}
void rk_update(double c) { // main.cpp:6618-6626 / 6634-6642
  auto &velInfo = var.vel->infos;
#pragma omp parallel for
  for (size_t i = 0; i < velInfo.size(); i++) {
    Real *V = velInfo[i].block, *Vold = var.vold->infos[i].block, *tmpV = var.tmpV->infos[i].block;
    Real ih2 = c / (velInfo[i].h * velInfo[i].h);
    for (int j = 0; j < 2 * _BS_ * _BS_; j++) V[j] = Vold[j] + tmpV[j] * ih2;
  }
}
void corr_update() { // main.cpp:7180-7187
  auto &velInfo = var.vel->infos;
#pragma omp parallel for
  for (size_t i = 0; i < velInfo.size(); i++) {
    Real ih2 = 1.0 / velInfo[i].h / velInfo[i].h;
    Real *V = velInfo[i].block, *tmpV = var.tmpV->infos[i].block;
    for (int j = 0; j < 2 * _BS_ * _BS_; j++) V[j] += tmpV[j] * ih2;
  }
}
void seed_taylor_green() {
  const size_t n2 = (size_t)g_N * g_NY;
  g_input.assign(6 * n2, 0.0);
  for (int iy = 0; iy < g_NY; iy++)
    for (int ix = 0; ix < g_N; ix++) {
      double x = (ix + 0.5) / g_N, y = (iy + 0.5) / g_NY;
      g_input[0 * n2 + (size_t)iy * g_N + ix] = sin(2 * M_PI * x) * cos(2 * M_PI * y);
      g_input[1 * n2 + (size_t)iy * g_N + ix] = -cos(2 * M_PI * x) * sin(2 * M_PI * y);
      g_input[2 * n2 + (size_t)iy * g_N + ix] = cos(2 * M_PI * x) * cos(2 * M_PI * y);
    }
}
struct TimeStat { double min, med, max; };
static TimeStat g_last_stat;
template <class F> double median_time(int reps, F f) {
  std::vector<double> t;
  f(); // warm-up (also builds the cached sync plan, excluded as BASELINE.md §3 says)
  for (int r = 0; r < reps; r++) {
    double t0 = now();
    f();
    t.push_back(now() - t0);
  }
  std::sort(t.begin(), t.end());
  g_last_stat = {t.front(), t[t.size() / 2], t.back()};
  return t[t.size() / 2];
}

void do_order() {
  FILE *f = fopen(g_out.c_str(), "wb");
  for (auto &info : var.vel->infos) {
    int ij[2] = {info.index[0], info.index[1]};
    fwrite(ij, sizeof(int), 2, f);
  }
  fclose(f);
}
void do_ops() {
  const size_t n2 = (size_t)g_N * g_NY;
  read_input(6);
  const double *in = g_input.data();
  sim.nu = g_nu;
  sim.dt = g_dt;
  g_fout = fopen(g_out.c_str(), "wb");
  scatter(var.vel, 2, in, in + n2);
  call_advect();
  write_field(var.tmpV, 2);
  scatter(var.tmpV, 2, in + 4 * n2, in + 5 * n2);
  scatter(var.chi, 1, in + 3 * n2, nullptr);
  call_rhs();
  write_field(var.tmp, 1);
  scatter(var.pold, 1, in + 2 * n2, nullptr);
  call_rhs1();
  write_field(var.tmp, 1);
  scatter(var.pres, 1, in + 2 * n2, nullptr);
  call_gradp();
  write_field(var.tmpV, 2);
  fclose(g_fout);
}
void do_vort() { // adapt()'s tagging input: KernelVorticity on vel -> tmp (main.cpp:3343-3366, 4659)
  const size_t n2 = (size_t)g_N * g_NY;
  read_input(6);
  scatter(var.vel, 2, g_input.data(), g_input.data() + n2);
  computeA<VectorLab>(KernelVorticity(), var.vel, 2);
  g_fout = fopen(g_out.c_str(), "wb");
  write_field(var.tmp, 1);
  fclose(g_fout);
}
void do_tags() { // main.cpp:4659-4660: what adapt() thresholds per block
  const size_t n2 = (size_t)g_N * g_NY;
  read_input(6);
  scatter(var.vel, 2, g_input.data(), g_input.data() + n2);
  scatter(var.chi, 1, g_input.data() + 3 * n2, nullptr);
  sim.Rtol = g_rtol;
  computeA<VectorLab>(KernelVorticity(), var.vel, 2);
  computeA<ScalarLab>(GradChiOnTmp(), var.chi, 1);
  g_fout = fopen(g_out.c_str(), "wb");
  write_field(var.tmp, 1);
  fclose(g_fout);
}
void do_dump() {
  const size_t n2 = (size_t)g_N * g_NY;
  read_input(6);
  scatter(var.vel, 2, g_input.data(), g_input.data() + n2);
  std::vector<char> path(g_out.begin(), g_out.end());
  path.push_back(0);
  dump(g_time, var.vel->infos.size(), var.vel->infos.data(), path.data());
}
void do_time() {
  seed_taylor_green();
  const size_t n2 = (size_t)g_N * g_NY;
  const double *in = g_input.data();
  scatter(var.vel, 2, in, in + n2);
  scatter(var.vold, 2, in, in + n2);
  scatter(var.pres, 1, in + 2 * n2, nullptr);
  scatter(var.pold, 1, in + 2 * n2, nullptr);
  double h = var.vel->infos[0].h, umax = field_umax();
  sim.nu = 1e-3;
  sim.dt = std::min(0.25 * h * h / (sim.nu + 0.25 * h * umax), 0.5 * h / (umax + 1e-8));
  double t_stage = median_time(g_reps, [] { call_advect(); rk_update(0.5); scatter(var.vel, 2, g_input.data(), g_input.data() + (size_t)g_N * g_NY); });
  TimeStat s_stage = g_last_stat;
  double t_scatter = median_time(g_reps, [] { scatter(var.vel, 2, g_input.data(), g_input.data() + (size_t)g_N * g_NY); });
  t_stage -= t_scatter;
  s_stage.min -= t_scatter, s_stage.med -= t_scatter, s_stage.max -= t_scatter;
  double t_rhs = median_time(g_reps, [] { call_rhs(); call_rhs1(); });
  const TimeStat s_rhs = g_last_stat;
  double t_corr = median_time(g_reps, [] { call_gradp(); corr_update(); });
  const TimeStat s_corr = g_last_stat;
  // Poisson: the reference has no CPU solver (cuda.cu is its only implementation); this times the CPU
  // restatement of cuda.cu:403-548 (oracle/ref_spmat_cpu.cpp) over the COO the reference's own assembly
  // loop built during step 0, exactly g_kiter iterations per solve (secondary, clearly-labelled figure).
  cup2d_ref_force_iters = g_kiter;
  double t_solve = median_time(g_reps, [] {
    std::fill(sim.mat->get_x().begin(), sim.mat->get_x().end(), 0.0);
    sim.mat->solveNoUpdate(0, 0, 0);
  });
  const int iters_run = cup2d_ref_fixed_iters > 0 ? cup2d_ref_fixed_iters : g_kiter;
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  const TimeStat s_solve = g_last_stat;
  const double it = iters_run > 0 ? 1.0 / iters_run : 0.0;
  printf("{\"L\": %d, \"N\": %d, \"cells\": %zu, \"threads\": %d, \"reps\": %d, \"t_stage\": %.6e, \"t_rhs\": %.6e, "
         "\"t_correct\": %.6e, \"kiter\": %d, \"t_poisson_iter\": %.6e, \"poisson_solver\": \"%s\", "
         "\"min_med_max\": {\"t_stage\": [%.6e, %.6e, %.6e], \"t_rhs\": [%.6e, %.6e, %.6e], \"t_correct\": [%.6e, %.6e, %.6e], "
         "\"t_poisson_iter\": [%.6e, %.6e, %.6e]}}\n",
         g_L, g_N, n2, nthreads, g_reps, t_stage, t_rhs, t_corr, iters_run, t_solve * it,
         cup2d_ref_fixed_iters > 0 ? "reference cuda.cu (GPU)" : "CPU restatement of cuda.cu",
         s_stage.min, s_stage.med, s_stage.max, s_rhs.min, s_rhs.med, s_rhs.max, s_corr.min, s_corr.med, s_corr.max,
         s_solve.min * it, s_solve.med * it, s_solve.max * it);
}
} // namespace

static void put(double tag, const std::vector<double> &v) {
  const double hdr[2] = {tag, (double)v.size()};
  fwrite(hdr, sizeof(double), 2, g_fout);
  fwrite(v.data(), sizeof(double), v.size(), g_fout);
}
static void put_fields(double tag, std::initializer_list<std::pair<Grid *, int>> gs, std::vector<double> head = {}) {
  const size_t n2 = (size_t)g_N * g_NY;
  for (auto &g : gs) {
    std::vector<double> a(n2), b(n2);
    gather(g.first, g.second, a.data(), b.data());
    head.insert(head.end(), a.begin(), a.end());
    if (g.second == 2) head.insert(head.end(), b.begin(), b.end());
  }
  put(tag, head);
}
// ---- amrlab: ghost assembly and operators of the reference on a real multi-level mesh -----------------------------
template <class LabT, int DIM> struct DumpLab { // a "kernel" that only copies the assembled lab out
  Stencil stencil;
  std::vector<double> *out;
  int nm;
  DumpLab(Stencil st, std::vector<double> *o) : stencil(st), out(o), nm(_BS_ + st.ex - st.sx - 1) {}
  void operator()(LabT *lab, const Info *info) const {
    const size_t n = (size_t)nm * nm * DIM;
    memcpy(out->data() + (size_t)info->id * n, lab->m, n * sizeof(double));
  }
};
static double seeded(int level, int i, int j, int ix, int iy, int comp, double x, double y) {
  // smooth part + a deterministic per-cell perturbation (so that an indexing error cannot hide behind smoothness)
  unsigned long long k = (((((unsigned long long)level * 4099 + i) * 4099 + j) * 67 + ix) * 67 + iy) * 7 + comp;
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  const double noise = (double)(k >> 11) / 9007199254740992.0 - 0.5;
  const double s = comp == 0 ? sin(1.3 * x + 0.4) * cos(2.1 * y) : comp == 1 ? cos(0.9 * x) * sin(1.7 * y + 0.2)
                 : comp == 2 ? cos(1.1 * x - 0.3) * cos(0.8 * y) : comp == 3 ? 0.5 + 0.5 * sin(2.0 * x) * sin(1.5 * y)
                 : comp == 4 ? 0.3 * sin(3 * x + y) : 0.2 * cos(2 * y - x);
  return s + 0.05 * noise;
}
static void seed_grid(Grid *g, int dim, int comp0) {
  for (auto &info : g->infos)
    for (int iy = 0; iy < _BS_; iy++)
      for (int ix = 0; ix < _BS_; ix++)
        for (int c = 0; c < dim; c++)
          info.block[dim * (_BS_ * iy + ix) + c] =
              seeded(info.level, info.index[0], info.index[1], ix, iy, comp0 + c, info.origin[0] + info.h * (ix + 0.5),
                     info.origin[1] + info.h * (iy + 0.5));
}
static void put_blocks(double tag, Grid *g, int dim) {
  std::vector<double> v;
  for (auto &info : g->infos) v.insert(v.end(), info.block, info.block + dim * _BS_ * _BS_);
  put(tag, v);
}
static void do_amrlab() {
  g_fout = fopen(g_out.c_str(), "wb");
  sim.dt = 1e-3; // fixed, so that the operator outputs do not depend on the flow the mesh was grown with
  put(10, {sim.nu, sim.dt, sim.h0, (double)sim.bpdx, (double)sim.bpdy, (double)sim.levelMax});
  std::vector<double> mesh;
  for (auto &info : var.vel->infos) { mesh.push_back(info.level); mesh.push_back(info.index[0]); mesh.push_back(info.index[1]); }
  put(11, mesh);
  seed_grid(var.vel, 2, 0);
  seed_grid(var.pres, 1, 2);
  seed_grid(var.pold, 1, 2);
  seed_grid(var.chi, 1, 3);
  put_blocks(12, var.vel, 2);
  put_blocks(13, var.pres, 1);
  put_blocks(14, var.chi, 1);
  const size_t nb = var.vel->infos.size();
  {
    std::vector<double> lab(nb * 14 * 14 * 2);
    computeA<VectorLab>(DumpLab<VectorLab, 2>(Stencil{-3, -3, 4, 4, true}, &lab), var.vel, 2);
    put(20, lab);
  }
  {
    std::vector<double> lab(nb * 10 * 10 * 2);
    computeA<VectorLab>(DumpLab<VectorLab, 2>(Stencil{-1, -1, 2, 2, false}, &lab), var.vel, 2);
    put(21, lab);
  }
  {
    std::vector<double> lab(nb * 10 * 10);
    computeA<ScalarLab>(DumpLab<ScalarLab, 1>(Stencil{-1, -1, 2, 2, false}, &lab), var.pres, 1);
    put(22, lab);
  }
  call_advect();
  put_blocks(30, var.tmpV, 2);
  seed_grid(var.tmpV, 2, 4); // u_def
  put_blocks(15, var.tmpV, 2);
  call_rhs();
  put_blocks(31, var.tmp, 1);
  call_rhs1();
  put_blocks(32, var.tmp, 1);
  call_gradp();
  put_blocks(33, var.tmpV, 2);
  fclose(g_fout);
}

// ---- atags: what adapt() looks at on a real multi-level mesh with real bodies (main.cpp:4676-4678) ---------------------
static void do_atags() {
  g_fout = fopen(g_out.c_str(), "wb");
  put(10, {sim.Rtol, sim.Ctol, (double)sim.levelMax, sim.h0, (double)sim.bpdx, (double)sim.bpdy});
  std::vector<double> mesh;
  for (auto &info : var.vel->infos) { mesh.push_back(info.level); mesh.push_back(info.index[0]); mesh.push_back(info.index[1]); }
  put(11, mesh);
  put_blocks(12, var.vel, 2);
  put_blocks(14, var.chi, 1);
  const size_t nb = var.vel->infos.size();
  {
    std::vector<double> lab(nb * 16 * 16);
    computeA<ScalarLab>(DumpLab<ScalarLab, 1>(Stencil{-4, -4, 5, 5, true}, &lab), var.chi, 1);
    put(23, lab);
  }
  computeA<VectorLab>(KernelVorticity(), var.vel, 2);
  put_blocks(40, var.tmp, 1);
  computeA<ScalarLab>(GradChiOnTmp(), var.chi, 1);
  put_blocks(41, var.tmp, 1);
  fclose(g_fout);
}

// ---- adump: the reference's dump() on a real multi-level mesh (main.cpp:3367-3467) -------------------------------------
static void do_adump() {
  g_fout = fopen((g_out + ".bin").c_str(), "wb");
  put(10, {0.1875, sim.h0, (double)sim.bpdx, (double)sim.bpdy});
  std::vector<double> mesh;
  for (auto &info : var.vel->infos) { mesh.push_back(info.level); mesh.push_back(info.index[0]); mesh.push_back(info.index[1]); }
  put(11, mesh);
  put_blocks(12, var.vel, 2);
  fclose(g_fout);
  std::vector<char> path(g_out.begin(), g_out.end());
  path.push_back(0);
  dump(0.1875, var.vel->infos.size(), var.vel->infos.data(), path.data());
}

static void penal_hook(int op, void *buf, int count) {
  const int S = (int)sim.shapes.size();
  if (op == MPI_MAX && count == 1) { // dt of a new step (main.cpp:6592)
    if (!g_fout) g_fout = fopen(g_out.c_str(), "wb");
    if (g_step == g_nsteps) { fclose(g_fout); exit(0); }
    g_sum7 = g_sum2 = 0;
    g_step++;
    return;
  }
  if (g_step == 0) return; // initialisation calls before the time loop
  if (op == MPI_SUM && count == 7) {
    const int c = g_sum7++;
    if (c < S) return; // the S calls of ongrid() (main.cpp:4514) come first
    const int k = c - S;
    const auto &shape = sim.shapes[k];
    if (k == 0) put_fields(2, {{var.vel, 2}});
    const double *Q = (const double *)buf;
    std::vector<double> r = {(double)(g_step - 1), (double)k, sim.lambda, sim.dt, shape->centerOfMass[0],
                             shape->centerOfMass[1]};
    r.insert(r.end(), Q, Q + 7);
    std::vector<double> ids, chi, udef;
    const auto &ob = shape->obstacleBlocks;
    for (size_t i = 0; i < ob.size(); i++) {
      if (!ob[i]) continue;
      ids.push_back((double)i);
      const double *c0 = (const double *)ob[i]->chi, *u0 = (const double *)ob[i]->udef;
      chi.insert(chi.end(), c0, c0 + _BS_ * _BS_);
      udef.insert(udef.end(), u0, u0 + 2 * _BS_ * _BS_);
    }
    r.push_back((double)ids.size());
    r.insert(r.end(), ids.begin(), ids.end());
    r.insert(r.end(), chi.begin(), chi.end());
    r.insert(r.end(), udef.begin(), udef.end());
    put(1, r);
    return;
  }
  if (op == MPI_SUM && count == 2 && g_sum2++ == 0) { // main.cpp:7138
    std::vector<double> head;
    for (auto &sh : sim.shapes) { head.push_back(sh->u); head.push_back(sh->v); head.push_back(sh->omega); }
    put_fields(3, {{var.vel, 2}, {var.tmpV, 2}, {var.chi, 1}}, head);
  }
}

void cup2d_ref_hook(int op, void *buf, int count) {
  if (g_mode == PENAL) { penal_hook(op, buf, count); return; }
  if (g_mode == ASTEPS) {
    if (op != MPI_MAX || count != 1) return;
    const int call = g_calls++;
    if (call == 0) { g_fout = fopen(g_out.c_str(), "wb"); return; }
    const double hdr[2] = {sim.dt, (double)var.vel->infos.size()};
    fwrite(hdr, sizeof(double), 2, g_fout);
    std::vector<double> mesh;
    for (auto &info : var.vel->infos) { mesh.push_back(info.level); mesh.push_back(info.index[0]); mesh.push_back(info.index[1]); }
    fwrite(mesh.data(), sizeof(double), mesh.size(), g_fout);
    for (auto &info : var.vel->infos) fwrite(info.block, sizeof(double), 2 * _BS_ * _BS_, g_fout);
    for (auto &info : var.pres->infos) fwrite(info.block, sizeof(double), _BS_ * _BS_, g_fout);
    if (call == g_nsteps) { fclose(g_fout); exit(0); }
    return;
  }
  if (g_mode == FSTEPS) {
    if (op != MPI_MAX || count != 1) return;
    const int call = g_calls++;
    if (call == 0) { g_fout = fopen(g_out.c_str(), "wb"); return; }
    fwrite(&sim.dt, sizeof(double), 1, g_fout);
    write_field(var.vel, 2);
    write_field(var.pres, 1);
    for (auto &sh : sim.shapes) {
      const double r[5] = {sh->centerOfMass[0], sh->centerOfMass[1], sh->u, sh->v, sh->omega};
      fwrite(r, sizeof(double), 5, g_fout);
    }
    if (call == g_nsteps) { fclose(g_fout); exit(0); }
    return;
  }
  if (g_mode == AMRLAB || g_mode == ATAGS || g_mode == ADUMP) {
    if (op != MPI_MAX || count != 1) return;
    if (g_calls++ < g_nsteps) return; // let the reference run nsteps steps first
    if (g_mode == AMRLAB) do_amrlab();
    else if (g_mode == ATAGS) do_atags();
    else do_adump();
    exit(0);
  }
  if (op != MPI_MAX || count != 1) return;
  const int call = g_calls++;
  if (g_mode == ORDER) { do_order(); exit(0); }
  if (g_mode == OPS) { do_ops(); exit(0); }
  if (g_mode == VORT) { do_vort(); exit(0); }
  if (g_mode == TAGS) { do_tags(); exit(0); }
  if (g_mode == DUMP) { do_dump(); exit(0); }
  if (g_mode == TIME) {
    // call 0: seed a Taylor-Green field and let the reference run step 0 (builds the sync plans and the
    // Poisson matrix); call 1: time the operators on the state after that step.
    if (call == 0) {
      seed_taylor_green();
      const size_t n2t = (size_t)g_N * g_NY;
      scatter(var.vel, 2, g_input.data(), g_input.data() + n2t);
      *(double *)buf = field_umax();
      cup2d_ref_force_iters = 2;
      return;
    }
    do_time();
    exit(0);
  }
  if (g_mode == AMR) {
    // the reference's own run.sh case (2 fish, AMR levels 5..8): record the Poisson system of every step
    // as raw block-ordered vectors: int64 nrows, int64 nblocks, double dt, b[nrows], x[nrows],
    // int32 (level,i,j)[nblocks]
    if (call == 0) {
      g_fout = fopen(g_out.c_str(), "wb");
      return;
    }
    const std::vector<double> &b = sim.mat->get_b(), &x = sim.mat->get_x();
    long long nrows = (long long)b.size(), nblk = (long long)var.tmp->infos.size();
    fwrite(&nrows, sizeof nrows, 1, g_fout);
    fwrite(&nblk, sizeof nblk, 1, g_fout);
    fwrite(&sim.dt, sizeof(double), 1, g_fout);
    fwrite(b.data(), sizeof(double), b.size(), g_fout);
    fwrite(x.data(), sizeof(double), x.size(), g_fout);
    for (auto &info : var.tmp->infos) {
      int lij[3] = {info.level, info.index[0], info.index[1]};
      fwrite(lij, sizeof(int), 3, g_fout);
    }
    fprintf(stderr, "ref_harness: amr step %d dt %.6e blocks %lld iters %d err %.3e\n", call - 1, sim.dt, nblk,
            cup2d_ref_last_iters, cup2d_ref_last_err);
    if (call == g_nsteps) {
      fclose(g_fout);
      exit(0);
    }
    return;
  }
  // STEPS: call 0 = before step 0 (seed), call k = after k steps (record)
  const size_t n2 = (size_t)g_N * g_NY;
  if (call == 0) {
    read_input(6);
    const double *in = g_input.data();
    scatter(var.vel, 2, in, in + n2);
    scatter(var.pres, 1, in + 2 * n2, nullptr);
    *(double *)buf = field_umax(); // umax feeds dt (main.cpp:6593-6595)
    g_fout = fopen(g_out.c_str(), "wb");
    return;
  }
  fwrite(&sim.dt, sizeof(double), 1, g_fout);
  write_field(var.vel, 2);
  write_field(var.pres, 1);
  write_vec(sim.mat->get_b());
  write_vec(sim.mat->get_x());
  fprintf(stderr, "ref_harness: step %d dt %.17g poisson iters %d err %.3e\n", call - 1, sim.dt,
          cup2d_ref_last_iters, cup2d_ref_last_err);
  if (call == g_nsteps) {
    fclose(g_fout);
    exit(0);
  }
}

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: ref_harness order|ops|steps|time L ...\n");
    return 2;
  }
  std::string mode = argv[1];
  g_L = atoi(argv[2]);
  g_bx = getenv("CUP2D_REF_BPDX") ? atoi(getenv("CUP2D_REF_BPDX")) : 1;
  g_by = getenv("CUP2D_REF_BPDY") ? atoi(getenv("CUP2D_REF_BPDY")) : 1;
  g_N = (_BS_ << g_L) * g_bx;   // cells per row; g_NY rows
  g_NY = (_BS_ << g_L) * g_by;
  g_nu = 1e-3;
  g_cfl = 0.5;
  if (mode == "order" && argc == 4) { g_mode = ORDER; g_out = argv[3]; }
  else if (mode == "ops" && argc == 7) { g_mode = OPS; g_nu = atof(argv[3]); g_dt = atof(argv[4]); g_in = argv[5]; g_out = argv[6]; }
  else if (mode == "vort" && argc == 5) { g_mode = VORT; g_in = argv[3]; g_out = argv[4]; }
  else if (mode == "tags" && argc == 7) { g_mode = TAGS; g_rtol = atof(argv[3]); g_extra = atoi(argv[4]); g_in = argv[5]; g_out = argv[6]; }
  else if (mode == "dump" && argc == 6) { g_mode = DUMP; g_time = atof(argv[3]); g_in = argv[4]; g_out = argv[5]; }
  else if (mode == "steps" && argc == 9) {
    g_mode = STEPS; g_nu = atof(argv[3]); g_cfl = atof(argv[4]); g_nsteps = atoi(argv[5]);
    g_kiter = atoi(argv[6]); g_in = argv[7]; g_out = argv[8];
    cup2d_ref_force_iters = g_kiter;
  } else if (mode == "time" && argc == 5) { g_mode = TIME; g_reps = atoi(argv[3]); g_kiter = atoi(argv[4]); }
  else if (mode == "amr" && argc == 6) {
    // ref_harness amr <levelMax> <nsteps> <kiter> <out.bin>   (argv[2] is parsed as g_L = levelMax here)
    g_mode = AMR; g_nsteps = atoi(argv[3]); g_kiter = atoi(argv[4]); g_out = argv[5];
    cup2d_ref_force_iters = g_kiter;
    char a_lmax[16];
    snprintf(a_lmax, sizeof a_lmax, "%d", g_L);
    // run.sh:1-22 verbatim except levelMax (argument) and tdump 0 (no output files)
    const char *args[] = {"ref_main", "-AdaptSteps", "20", "-bpdx", "2", "-bpdy", "1", "-CFL", "0.5", "-Ctol", "1",
                          "-extent", "4", "-lambda", "1e7", "-levelMax", a_lmax, "-levelStart", "5",
                          "-maxPoissonIterations", "1000", "-maxPoissonRestarts", "0", "-nu", "0.00004",
                          "-poissonTol", "1e-3", "-poissonTolRel", "1e-2", "-Rtol", "2", "-tdump", "0", "-tend", "10.0",
                          "-shapes", "angle=0 L=0.2 xpos=1.8 ypos=0.8\n angle=180 L=0.2 xpos=1.6 ypos=0.8"};
    return ref_main(sizeof args / sizeof *args, (char **)args);
  }
  else if (((mode == "amrlab" || mode == "atags" || mode == "adump") && argc == 5) || (mode == "asteps" && argc == 6)) {
    if (mode != "asteps") { g_mode = mode == "amrlab" ? AMRLAB : mode == "atags" ? ATAGS : ADUMP; g_nsteps = atoi(argv[3]); g_out = argv[4]; cup2d_ref_force_iters = 5; }
    else { g_mode = ASTEPS; g_nsteps = atoi(argv[3]); g_kiter = atoi(argv[4]); g_out = argv[5]; cup2d_ref_force_iters = g_kiter; }
    char a_lmax[16];
    snprintf(a_lmax, sizeof a_lmax, "%d", g_L);
    const char *args[] = {"ref_main", "-AdaptSteps", "20", "-bpdx", "2", "-bpdy", "1", "-CFL", "0.5", "-Ctol", "1",
                          "-extent", "4", "-lambda", "1e7", "-levelMax", a_lmax, "-levelStart", "5",
                          "-maxPoissonIterations", "1000", "-maxPoissonRestarts", "0", "-nu", "0.00004",
                          "-poissonTol", "1e-3", "-poissonTolRel", "1e-2", "-Rtol", "2", "-tdump", "0", "-tend", "10.0",
                          "-shapes", "angle=0 L=0.2 xpos=1.8 ypos=0.8\n angle=180 L=0.2 xpos=1.6 ypos=0.8"};
    return ref_main(sizeof args / sizeof *args, (char **)args);
  }
  else if ((mode == "penal" || mode == "fsteps") && argc == 6) {
    g_mode = mode == "penal" ? PENAL : FSTEPS; g_nsteps = atoi(argv[3]); g_kiter = atoi(argv[4]); g_out = argv[5];
    cup2d_ref_force_iters = g_kiter;
    char a_l[16], a_lm[16];
    snprintf(a_l, sizeof a_l, "%d", g_L);
    snprintf(a_lm, sizeof a_lm, "%d", g_L + 1);
    const char *args[] = {"ref_main", "-AdaptSteps", "1000000", "-bpdx", "1", "-bpdy", "1", "-CFL", "0.5", "-Ctol", "0",
                          "-extent", "1", "-lambda", "1e7", "-levelMax", a_lm, "-levelStart", a_l,
                          "-maxPoissonIterations", "1000", "-maxPoissonRestarts", "0", "-nu", "0.00004",
                          "-poissonTol", "0", "-poissonTolRel", "0", "-Rtol", "1e300", "-tdump", "0", "-tend", "10.0",
                          "-shapes", getenv("CUP2D_REF_SHAPES") ? getenv("CUP2D_REF_SHAPES")
                                                                : "angle=0 L=0.4 xpos=0.55 ypos=0.4\n angle=180 L=0.4 xpos=0.45 ypos=0.6"};
    return ref_main(sizeof args / sizeof *args, (char **)args);
  }
  else { fprintf(stderr, "ref_harness: bad arguments\n"); return 2; }
  char a_ls[32], a_lm[32], a_nu[64], a_cfl[64];
  snprintf(a_ls, sizeof a_ls, "%d", g_L);
  snprintf(a_lm, sizeof a_lm, "%d", g_L + 1 + g_extra);
  snprintf(a_nu, sizeof a_nu, "%.17g", g_nu);
  snprintf(a_cfl, sizeof a_cfl, "%.17g", g_cfl);
  char a_bx[16], a_by[16];
  snprintf(a_bx, sizeof a_bx, "%d", g_bx);
  snprintf(a_by, sizeof a_by, "%d", g_by);
  const char *args[] = {"ref_main", "-AdaptSteps", "1000000", "-bpdx", a_bx, "-bpdy", a_by, "-CFL", a_cfl,
                        "-Ctol", "0", "-extent", "1", "-lambda", "1e7", "-levelMax", a_lm, "-levelStart", a_ls,
                        "-maxPoissonIterations", "1000", "-maxPoissonRestarts", "0", "-nu", a_nu,
                        "-poissonTol", "0", "-poissonTolRel", "0", "-Rtol", "1e300", "-tdump", "0",
                        "-tend", "1e300", "-shapes", ""};
  int n = sizeof args / sizeof *args;
  return ref_main(n, (char **)args);
}
