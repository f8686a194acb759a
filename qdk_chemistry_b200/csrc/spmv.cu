// sigma = H c : HBM-streaming FP64 CSR SpMV.
//
// Replaces sparsexx::spblas::gespmbv (K = 1, alpha = 1, beta = 0)
// (external/macis/src/sparsexx/include/sparsexx/spblas/spmbv.hpp:49-85).
//
// Algorithmic bytes per call: nnz*(8 + 4) + (nrows + 1)*8 + ncols*8 + nrows*8. The kernel
// is a CSR-vector scheme: a group of TPR lanes owns one row; nzval is streamed with 16-byte
// (2 x f64) and colind with 8-byte (2 x i32) non-coherent loads after a scalar head element
// that brings the row to even alignment; x is gathered through the read-only path (x fits
// L2 for every single-GPU config: N*8 <= 80 MB). Partial sums are combined with shuffles
// in a fixed tree, so the result is deterministic for a given TPR.
#include "common.cuh"

namespace b2ci {
namespace {

__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
               : "=d"(v.x), "=d"(v.y)
               : "l"(p));
  return v;
}
__device__ __forceinline__ int2 ldg_stream_i32x2(const int32_t* p) {
  int2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];"
               : "=r"(v.x), "=r"(v.y)
               : "l"(p));
  return v;
}
__device__ __forceinline__ double ldg_stream_f64(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int32_t ldg_stream_i32(const int32_t* p) {
  int32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

template <int TPR>
__global__ void __launch_bounds__(256)
k_spmv(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
       const double* __restrict__ nzval, const double* __restrict__ x, double* __restrict__ y) {
  const int64_t gtid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row = gtid / TPR;
  const int sub = int(gtid % TPR);
  double acc0 = 0., acc1 = 0.;
  if (row < nrows) {
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    // head: make the vector body start at an even element (16 B aligned nzval, 8 B colind)
    int64_t b = s + (s & 1);
    if (b > e) b = e;
    if (sub == 0 && b > s) acc0 = ldg_stream_f64(nzval + s) * __ldg(x + ldg_stream_i32(colind + s));
    const int64_t npairs = (e - b) >> 1;
    int64_t p = sub;
    // two pairs per lane in flight per iteration
    for (; p + TPR < npairs; p += 2 * TPR) {
      const int64_t k0 = b + 2 * p, k1 = b + 2 * (p + TPR);
      const double2 v0 = ldg_stream_f64x2(nzval + k0);
      const int2 c0 = ldg_stream_i32x2(colind + k0);
      const double2 v1 = ldg_stream_f64x2(nzval + k1);
      const int2 c1 = ldg_stream_i32x2(colind + k1);
      const double x00 = __ldg(x + c0.x), x01 = __ldg(x + c0.y);
      const double x10 = __ldg(x + c1.x), x11 = __ldg(x + c1.y);
      acc0 = fma(v0.x, x00, acc0);
      acc1 = fma(v0.y, x01, acc1);
      acc0 = fma(v1.x, x10, acc0);
      acc1 = fma(v1.y, x11, acc1);
    }
    if (p < npairs) {
      const int64_t k0 = b + 2 * p;
      const double2 v0 = ldg_stream_f64x2(nzval + k0);
      const int2 c0 = ldg_stream_i32x2(colind + k0);
      acc0 = fma(v0.x, __ldg(x + c0.x), acc0);
      acc1 = fma(v0.y, __ldg(x + c0.y), acc1);
    }
    // tail element
    const int64_t tail = b + 2 * npairs;
    if (sub == (TPR > 1 ? 1 : 0) && tail < e)
      acc1 = fma(ldg_stream_f64(nzval + tail), __ldg(x + ldg_stream_i32(colind + tail)), acc1);
  }
  double acc = acc0 + acc1;
#pragma unroll
  for (int d = TPR >> 1; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d, TPR);
  if (row < nrows && sub == 0) y[row] = acc;
}

template <int TPR>
void launch(b2ci_ctx* ctx, const b2ci_csr* m, const double* x, double* y) {
  const int64_t threads = m->nrows * TPR;
  const unsigned grid = unsigned((threads + 255) / 256);
  k_spmv<TPR><<<grid, 256, 0, ctx->stream>>>(m->nrows, m->rowptr, m->colind, m->nzval, x, y);
  ctx->launches++;
  B2_CHECK_LAUNCH();
}

}  // namespace

void spmv_launch(b2ci_ctx* ctx, const b2ci_csr* m, const double* x, double* y) {
  if (m->nrows == 0) return;
  const double mean = double(m->nnz) / double(m->nrows);
  if (mean >= 96.) launch<32>(ctx, m, x, y);
  else if (mean >= 48.) launch<16>(ctx, m, x, y);
  else if (mean >= 24.) launch<8>(ctx, m, x, y);
  else if (mean >= 12.) launch<4>(ctx, m, x, y);
  else launch<2>(ctx, m, x, y);
}

}  // namespace b2ci
