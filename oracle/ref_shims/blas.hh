// TEST INFRASTRUCTURE (oracle/_ref build glue) -- not part of the product.
// Minimal stand-in for blaspp's <blas.hh> (fetched by the reference's CMake,
// not vendored). Forwards the handful of level-1/3 calls MACIS makes on the CI
// path (solvers/davidson.hpp:189-335, asci/iteration.hpp:162-178) to the LP64
// Fortran BLAS inside scipy's bundled OpenBLAS (symbols carry a scipy_ prefix).
#pragma once
#include <complex>
#include <cstdint>
#include <vector>
extern "C" {
void scipy_dgemm_(const char*, const char*, const int*, const int*, const int*,
                  const double*, const double*, const int*, const double*,
                  const int*, const double*, double*, const int*);
double scipy_dnrm2_(const int*, const double*, const int*);
double scipy_ddot_(const int*, const double*, const int*, const double*,
                   const int*);
void scipy_dscal_(const int*, const double*, double*, const int*);
void scipy_daxpy_(const int*, const double*, const double*, const int*, double*,
                  const int*);
int scipy_idamax_(const int*, const double*, const int*);
}
namespace blas {
enum class Layout : char { ColMajor = 'C', RowMajor = 'R' };
enum class Op : char { NoTrans = 'N', Trans = 'T', ConjTrans = 'C' };
inline void gemm(Layout, Op ta, Op tb, int64_t m, int64_t n, int64_t k,
                 double alpha, const double* A, int64_t lda, const double* B,
                 int64_t ldb, double beta, double* C, int64_t ldc) {
  char cta = (ta == Op::NoTrans) ? 'N' : 'T';
  char ctb = (tb == Op::NoTrans) ? 'N' : 'T';
  int m_ = m, n_ = n, k_ = k, lda_ = lda, ldb_ = ldb, ldc_ = ldc;
  scipy_dgemm_(&cta, &ctb, &m_, &n_, &k_, &alpha, A, &lda_, B, &ldb_, &beta, C,
               &ldc_);
}
inline double nrm2(int64_t n, const double* x, int64_t incx) {
  int n_ = n, i_ = incx;
  return scipy_dnrm2_(&n_, x, &i_);
}
inline double dot(int64_t n, const double* x, int64_t incx, const double* y,
                  int64_t incy) {
  int n_ = n, ix = incx, iy = incy;
  return scipy_ddot_(&n_, x, &ix, y, &iy);
}
inline void scal(int64_t n, double a, double* x, int64_t incx) {
  int n_ = n, ix = incx;
  scipy_dscal_(&n_, &a, x, &ix);
}
inline void axpy(int64_t n, double a, const double* x, int64_t incx, double* y,
                 int64_t incy) {
  int n_ = n, ix = incx, iy = incy;
  scipy_daxpy_(&n_, &a, x, &ix, y, &iy);
}
inline int64_t iamax(int64_t n, const double* x, int64_t incx) {
  int n_ = n, ix = incx;
  return scipy_idamax_(&n_, x, &ix) - 1;
}
}  // namespace blas
