// TEST INFRASTRUCTURE — race hunt.  Linked against the emulated product sources built with -fsanitize=thread
// (tests/host_emu/build.py build_tsan): every CUDA thread is an OS thread, so an access pair that is not ordered by a
// __syncthreads / __syncwarp / shuffle / atomic shows up as a ThreadSanitizer report.  Runs one uniform-grid time step and the
// multi-level operators + step (baseline and fast kernels) on a small two-level mesh.
#include "../../include/cup2d_b200.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                                  \
  do {                                                                                            \
    int _rc = (x);                                                                                \
    if (_rc) {                                                                                    \
      fprintf(stderr, "%s -> %d: %s\n", #x, _rc, cup2d_last_error());                             \
      return 1;                                                                                   \
    }                                                                                             \
  } while (0)

// tsan_driver <mesh.bin> <bpdx> <bpdy> <h0>: multi-level steps on a mesh given as int32 (level, i, j) triples
static int run_mesh(const char *path, int bpdx, int bpdy, double h0) {
  FILE *f = fopen(path, "rb");
  if (!f) return 2;
  std::vector<int32_t> b;
  int32_t t[3];
  while (fread(t, sizeof(int32_t), 3, f) == 3) b.insert(b.end(), t, t + 3);
  fclose(f);
  const int64_t n = (int64_t)b.size() / 3;
  cup2d_amr *a = nullptr;
  CHECK(cup2d_amr_create(n, b.data(), bpdx, bpdy, h0, 1e-3, 0, &a));
  std::vector<double> vel(n * 128), pres(n * 64);
  for (size_t i = 0; i < vel.size(); i++) vel[i] = std::sin(0.29 * i) * 0.7;
  for (size_t i = 0; i < pres.size(); i++) pres[i] = std::cos(0.13 * i);
  for (int fast = 0; fast < 2; fast++) {
    CHECK(cup2d_amr_set_fast(a, fast));
    CHECK(cup2d_amr_field_upload(a, CUP2D_VEL, vel.data()));
    CHECK(cup2d_amr_field_upload(a, CUP2D_PRES, pres.data()));
    double dt, err;
    int it;
    CHECK(cup2d_amr_step(a, 0.5, 0.0, 0.0, 0.0, 0, 3, &dt, &it, &err));
    if (!std::isfinite(dt) || !std::isfinite(err)) return 3;
    printf("mesh step (fast=%d): %lld blocks dt %.3e iters %d err %.3e\n", fast, (long long)n, dt, it, err);
  }
  cup2d_amr_destroy(a);
  return 0;
}

int main(int argc, char **argv) {
  if (argc == 5) return run_mesh(argv[1], atoi(argv[2]), atoi(argv[3]), atof(argv[4]));
  { // uniform grid, level 2: 16 blocks
    const int L = 2, nb1 = 1 << L, n = nb1 * nb1;
    std::vector<int32_t> ij(2 * n);
    CHECK(cup2d_block_order(1, 1, L, ij.data()));
    int64_t rb[2] = {0, n};
    cup2d_config cfg = {nb1, nb1, n, ij.data(), 0, 1, rb, 1.0 / (8 * nb1), 1e-3, 0.5, 0, 0};
    cup2d_sim *s = nullptr;
    CHECK(cup2d_create(&cfg, &s));
    std::vector<double> vel(n * 128), pres(n * 64);
    for (size_t i = 0; i < vel.size(); i++) vel[i] = std::sin(0.37 * i) * 0.8;
    for (size_t i = 0; i < pres.size(); i++) pres[i] = std::cos(0.11 * i);
    CHECK(cup2d_field_upload(s, CUP2D_VEL, vel.data()));
    CHECK(cup2d_field_upload(s, CUP2D_PRES, pres.data()));
    double dt, err;
    int it;
    CHECK(cup2d_step(s, 0.0, 0, 0.0, 0.0, 0, 3, &dt, &it, &err));
    std::vector<double> linf(n);
    CHECK(cup2d_adapt_tags(s, 1.0, 4, linf.data()));
    printf("uniform step: dt %.3e iters %d err %.3e\n", dt, it, err);
    cup2d_destroy(s);
  }
  { // two levels: a 4x4 level-2 mesh with block (1,1) and (2,1) refined
    std::vector<int32_t> b;
    for (int j = 0; j < 4; j++)
      for (int i = 0; i < 4; i++) {
        if ((i == 1 || i == 2) && j == 1) {
          for (int c = 0; c < 4; c++) b.insert(b.end(), {3, 2 * i + (c & 1), 2 * j + (c >> 1)});
        } else
          b.insert(b.end(), {2, i, j});
      }
    const int64_t n = (int64_t)b.size() / 3;
    cup2d_amr *a = nullptr;
    CHECK(cup2d_amr_create(n, b.data(), 1, 1, 1.0 / 8, 1e-3, 0, &a));
    std::vector<double> vel(n * 128), pres(n * 64);
    for (size_t i = 0; i < vel.size(); i++) vel[i] = std::sin(0.29 * i) * 0.7;
    for (size_t i = 0; i < pres.size(); i++) pres[i] = std::cos(0.13 * i);
    for (int fast = 0; fast < 2; fast++) {
      CHECK(cup2d_amr_set_fast(a, fast));
      CHECK(cup2d_amr_field_upload(a, CUP2D_VEL, vel.data()));
      CHECK(cup2d_amr_field_upload(a, CUP2D_PRES, pres.data()));
      double dt, err;
      int it;
      CHECK(cup2d_amr_step(a, 0.5, 0.0, 0.0, 0.0, 0, 2, &dt, &it, &err));
      printf("multi-level step (fast=%d): %lld blocks dt %.3e iters %d err %.3e\n", fast, (long long)n, dt, it, err);
    }
    { // bodies, tagging and dump on the same mesh: one shape over blocks of both levels
      std::vector<int32_t> ids = {0, 5, 6, 9, (int32_t)n - 1};
      std::vector<double> X(ids.size() * 64), U(ids.size() * 128), q(7), linf(n);
      for (size_t i = 0; i < X.size(); i++) X[i] = 0.5 + 0.6 * std::sin(0.7 * i);
      for (size_t i = 0; i < U.size(); i++) U[i] = 0.3 * std::cos(0.31 * i);
      CHECK(cup2d_amr_field_upload(a, CUP2D_CHI, pres.data()));
      CHECK(cup2d_amr_shape_set(a, 0, (int)ids.size(), ids.data(), X.data(), U.data()));
      CHECK(cup2d_amr_shape_integrals(a, 0, 1e7, 1e-3, 0.4, 0.3, q.data()));
      CHECK(cup2d_amr_penalize(a, 0, 1e7, 1e-3, 0.4, 0.3, 0.1, -0.2, 0.5));
      CHECK(cup2d_amr_udef_assemble(a));
      CHECK(cup2d_amr_adapt_tags(a, 2.0, 4, linf.data()));
      CHECK(cup2d_amr_dump(a, 0.5, "/tmp/cup2d_sanitizer_dump"));
      if (!std::isfinite(q[0] + q[6] + linf[0])) return 4;
      printf("multi-level bodies / tags / dump: PM %.3e AM %.3e linf[0] %.3e\n", q[0], q[6], linf[0]);
    }
    cup2d_amr_destroy(a);
  }
  return 0;
}
