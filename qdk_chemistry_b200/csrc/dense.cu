// Dense ground state of a small resident CSR matrix.
//
// Replaces the dense branch of the MACIS adapters (n <= iterative_solver_dimension_cutoff):
// sparsexx::convert_to_dense + lapack::syev(Vec, Upper) and "eigenvalue 0 / first column"
// (cpp/src/qdk/chemistry/algorithms/microsoft/macis_cas.cpp:89-101, macis_pmc.cpp:98-112).
// The scatter to dense storage is a kernel of this library; the symmetric eigensolver is
// cuSOLVER's Dsyevd (a plain library call, like the reference's LAPACK call), resolved at run
// time so that nothing but this entry point depends on libcusolver.
#include <cusolverDn.h>
#include <dlfcn.h>

#include "common.cuh"

namespace b2ci {
namespace {

struct SolverApi {
  void* handle = nullptr;
  cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*BufferSize)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, const double*,
                                 int, const double*, int*) = nullptr;
  cusolverStatus_t (*Dsyevd)(cusolverDnHandle_t, cusolverEigMode_t, cublasFillMode_t, int, double*, int,
                             double*, double*, int, int*) = nullptr;
};

SolverApi& solver_api() {
  static SolverApi a;
  if (a.handle) return a;
  const char* names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11",
                         "/usr/local/cuda/lib64/libcusolver.so"};
  for (auto nm : names) {
    a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) throw Error(std::string("cannot load libcusolver (dense CI branch): ") + dlerror());
#define B2_SYM(field, sym)                                             \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym)); \
  if (!a.field) throw Error(std::string("libcusolver lacks symbol ") + sym);
  B2_SYM(Create, "cusolverDnCreate")
  B2_SYM(Destroy, "cusolverDnDestroy")
  B2_SYM(SetStream, "cusolverDnSetStream")
  B2_SYM(BufferSize, "cusolverDnDsyevd_bufferSize")
  B2_SYM(Dsyevd, "cusolverDnDsyevd")
#undef B2_SYM
  return a;
}

__global__ void k_csr_to_dense(int64_t nrows, const int64_t* __restrict__ rowptr,
                               const int32_t* __restrict__ colind, const double* __restrict__ nzval,
                               int64_t ld, double* __restrict__ A) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= nrows) return;
  for (int64_t p = rowptr[r] + lane; p < rowptr[r + 1]; p += 32) A[r + int64_t(colind[p]) * ld] = nzval[p];
}

}  // namespace

void dense_ground_state(b2ci_ctx* ctx, const b2ci_csr* m, double* eigval, double* eigvec_host) {
  const int64_t n = m->nrows;
  if (n != m->ncols || m->row_begin != 0) throw Error("b2ci_dense_ground_state: needs the full square matrix");
  if (n < 1 || n > 32768) throw Error("b2ci_dense_ground_state: dimension out of range [1, 32768]");
  SolverApi& S = solver_api();
  cudaStream_t st = ctx->stream;
  DevBuf<double> A(size_t(n) * n), W(n);
  DevBuf<int> info(1);
  B2_CUDA(cudaMemsetAsync(A, 0, size_t(n) * n * 8, st));
  k_csr_to_dense<<<unsigned((n * 32 + 255) / 256), 256, 0, st>>>(n, m->rowptr, m->colind, m->nzval, n, A);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  cusolverDnHandle_t h = nullptr;
  if (S.Create(&h) != CUSOLVER_STATUS_SUCCESS) throw Error("cusolverDnCreate failed");
  try {
    if (S.SetStream(h, st) != CUSOLVER_STATUS_SUCCESS) throw Error("cusolverDnSetStream failed");
    int lwork = 0;
    if (S.BufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, int(n), A, int(n), W, &lwork) !=
        CUSOLVER_STATUS_SUCCESS)
      throw Error("cusolverDnDsyevd_bufferSize failed");
    DevBuf<double> work(lwork > 0 ? lwork : 1);
    if (S.Dsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, int(n), A, int(n), W, work, lwork, info) !=
        CUSOLVER_STATUS_SUCCESS)
      throw Error("cusolverDnDsyevd failed");
    int hinfo = 0;
    B2_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(eigval, W, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(eigvec_host, A, size_t(n) * 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (hinfo != 0) throw Error("cusolverDnDsyevd: info = " + std::to_string(hinfo));
  } catch (...) {
    S.Destroy(h);
    throw;
  }
  S.Destroy(h);
}

}  // namespace b2ci
