// TEST INFRASTRUCTURE (oracle/_ref build glue) -- not part of the product.
// Minimal stand-in for lapackpp's <lapack.hh>: syev (rayleigh_ritz.hpp:75,
// asci/grow.hpp:185) and gesvd (hamiltonian_generator/base.ipp:85, never
// executed on the CI path) forwarded to scipy's bundled OpenBLAS LAPACK.
#pragma once
#include <cstdint>
#include <vector>
extern "C" {
void scipy_dsyev_(const char*, const char*, const int*, double*, const int*,
                  double*, double*, const int*, int*);
void scipy_dgesvd_(const char*, const char*, const int*, const int*, double*,
                   const int*, double*, double*, const int*, double*,
                   const int*, double*, const int*, int*);
}
namespace lapack {
enum class Job : char {
  NoVec = 'N',
  Vec = 'V',
  UpdateVec = 'U',
  AllVec = 'A',
  SomeVec = 'S',
  OverwriteVec = 'O'
};
enum class Uplo : char { Upper = 'U', Lower = 'L', General = 'G' };
inline int64_t syev(Job job, Uplo uplo, int64_t n, double* A, int64_t lda,
                    double* W) {
  char j = (char)job, u = (char)uplo;
  int n_ = n, lda_ = lda, info = 0, lwork = -1;
  double wq;
  scipy_dsyev_(&j, &u, &n_, A, &lda_, W, &wq, &lwork, &info);
  lwork = (int)wq;
  std::vector<double> work(lwork > 1 ? lwork : 1);
  scipy_dsyev_(&j, &u, &n_, A, &lda_, W, work.data(), &lwork, &info);
  return info;
}
inline int64_t gesvd(Job jobu, Job jobvt, int64_t m, int64_t n, double* A,
                     int64_t lda, double* S, double* U, int64_t ldu,
                     double* VT, int64_t ldvt) {
  char ju = (char)jobu, jv = (char)jobvt;
  int m_ = m, n_ = n, lda_ = lda, ldu_ = ldu, ldvt_ = ldvt, info = 0,
      lwork = -1;
  double wq;
  scipy_dgesvd_(&ju, &jv, &m_, &n_, A, &lda_, S, U, &ldu_, VT, &ldvt_, &wq,
                &lwork, &info);
  lwork = (int)wq;
  std::vector<double> work(lwork > 1 ? lwork : 1);
  scipy_dgesvd_(&ju, &jv, &m_, &n_, A, &lda_, S, U, &ldu_, VT, &ldvt_,
                work.data(), &lwork, &info);
  return info;
}
}  // namespace lapack
