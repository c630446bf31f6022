/* TEST INFRASTRUCTURE — minimal stand-in for the eight GSL symbols the reference uses
 * at main.cpp:6692-6703 (one 3x3 LU solve for the rigid-body momenta; OUT of the hot
 * path but needed to link the unmodified translation unit).  Partial-pivot LU.
 */
#ifndef CUP2D_ORACLE_GSL_SHIM_H
#define CUP2D_ORACLE_GSL_SHIM_H
#include <cmath>
#include <cstdlib>

typedef struct { size_t size1, size2, tda; double *data; } gsl_matrix;
typedef struct { size_t size, stride; double *data; int owner; } gsl_vector;
typedef struct { gsl_matrix matrix; } gsl_matrix_view;
typedef struct { gsl_vector vector; } gsl_vector_view;
typedef struct { size_t size; size_t *data; } gsl_permutation;

static inline gsl_matrix_view gsl_matrix_view_array(double *a, size_t n1, size_t n2) {
  gsl_matrix_view v;
  v.matrix.size1 = n1; v.matrix.size2 = n2; v.matrix.tda = n2; v.matrix.data = a;
  return v;
}
static inline gsl_vector_view gsl_vector_view_array(double *a, size_t n) {
  gsl_vector_view v;
  v.vector.size = n; v.vector.stride = 1; v.vector.data = a; v.vector.owner = 0;
  return v;
}
static inline gsl_vector *gsl_vector_alloc(size_t n) {
  gsl_vector *v = (gsl_vector *)malloc(sizeof *v);
  v->size = n; v->stride = 1; v->data = (double *)calloc(n, sizeof(double)); v->owner = 1;
  return v;
}
static inline void gsl_vector_free(gsl_vector *v) { free(v->data); free(v); }
static inline double gsl_vector_get(const gsl_vector *v, size_t i) { return v->data[i * v->stride]; }
static inline gsl_permutation *gsl_permutation_alloc(size_t n) {
  gsl_permutation *p = (gsl_permutation *)malloc(sizeof *p);
  p->size = n; p->data = (size_t *)malloc(n * sizeof(size_t));
  for (size_t i = 0; i < n; i++) p->data[i] = i;
  return p;
}
static inline void gsl_permutation_free(gsl_permutation *p) { free(p->data); free(p); }
static inline int gsl_linalg_LU_decomp(gsl_matrix *A, gsl_permutation *p, int *signum) {
  const size_t n = A->size1, lda = A->tda;
  double *a = A->data;
  *signum = 1;
  for (size_t i = 0; i < n; i++) p->data[i] = i;
  for (size_t j = 0; j + 1 < n; j++) {
    size_t piv = j; double big = fabs(a[j * lda + j]);
    for (size_t i = j + 1; i < n; i++)
      if (fabs(a[i * lda + j]) > big) { big = fabs(a[i * lda + j]); piv = i; }
    if (piv != j) {
      for (size_t k = 0; k < n; k++) { double t = a[j * lda + k]; a[j * lda + k] = a[piv * lda + k]; a[piv * lda + k] = t; }
      size_t t = p->data[j]; p->data[j] = p->data[piv]; p->data[piv] = t;
      *signum = -*signum;
    }
    if (a[j * lda + j] != 0.0)
      for (size_t i = j + 1; i < n; i++) {
        double m = a[i * lda + j] / a[j * lda + j];
        a[i * lda + j] = m;
        for (size_t k = j + 1; k < n; k++) a[i * lda + k] -= m * a[j * lda + k];
      }
  }
  return 0;
}
static inline int gsl_linalg_LU_solve(const gsl_matrix *LU, const gsl_permutation *p,
                                      const gsl_vector *b, gsl_vector *x) {
  const size_t n = LU->size1, lda = LU->tda;
  const double *a = LU->data;
  for (size_t i = 0; i < n; i++) x->data[i] = b->data[p->data[i] * b->stride];
  for (size_t i = 0; i < n; i++)
    for (size_t k = 0; k < i; k++) x->data[i] -= a[i * lda + k] * x->data[k];
  for (size_t ii = n; ii-- > 0;) {
    for (size_t k = ii + 1; k < n; k++) x->data[ii] -= a[ii * lda + k] * x->data[k];
    x->data[ii] /= a[ii * lda + ii];
  }
  return 0;
}
#endif
