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

// ---- sharded sigma with the exchange overlapped (sparsexx/spblas/pspmbv.hpp:352-387: the reference posts
// the halo exchange and multiplies the diagonal tile meanwhile). Columns ascend inside a row, so the
// columns this rank owns, [col0, col1), are ONE sub-range [lo, hi) of every row (k_local_range, once per
// matrix). PART 0 multiplies that sub-range with the rank's own block of the trial vector while the
// peers' blocks are still arriving; PART 1 adds the two outer sub-ranges once they are there.
template <int TPR>
__device__ __forceinline__ void range_sum(int64_t s, int64_t e, int sub, const int32_t* __restrict__ colind,
                                          const double* __restrict__ nzval, const double* __restrict__ x,
                                          double& acc0, double& acc1) {
  int64_t b = s + (s & 1);
  if (b > e) b = e;
  if (sub == 0 && b > s) acc0 = fma(ldg_stream_f64(nzval + s), __ldg(x + ldg_stream_i32(colind + s)), acc0);
  const int64_t npairs = (e - b) >> 1;
  int64_t p = sub;
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
  const int64_t tail = b + 2 * npairs;
  if (sub == (TPR > 1 ? 1 : 0) && tail < e)
    acc1 = fma(ldg_stream_f64(nzval + tail), __ldg(x + ldg_stream_i32(colind + tail)), acc1);
}
template <int TPR, int PART>
__global__ void __launch_bounds__(256)
k_spmv_part(int64_t nrows, const int64_t* __restrict__ rowptr, const int2* __restrict__ loc,
            const int32_t* __restrict__ colind, const double* __restrict__ nzval, const double* __restrict__ x,
            double* __restrict__ y) {
  const int64_t gtid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row = gtid / TPR;
  const int sub = int(gtid % TPR);
  double acc0 = 0., acc1 = 0.;
  if (row < nrows) {
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    const int2 lh = loc[row];
    if (PART == 0) {
      // x = this rank's block, indexed by the global column: the caller passes (block - col0)
      range_sum<TPR>(s + lh.x, s + lh.y, sub, colind, nzval, x, acc0, acc1);
    } else {
      range_sum<TPR>(s, s + lh.x, sub, colind, nzval, x, acc0, acc1);
      range_sum<TPR>(s + lh.y, e, sub, colind, nzval, x, acc0, acc1);
    }
  }
  double acc = acc0 + acc1;
#pragma unroll
  for (int d = TPR >> 1; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d, TPR);
  if (row < nrows && sub == 0) y[row] = PART == 0 ? acc : y[row] + acc;
}
// own-column sub-range [lo, hi) of every row: lower bounds of col0 and col1 in the row's ascending columns.
// 8 lanes per row narrow the interval 8-fold per step (a thread-per-row binary search spends ~22 dependent
// loads per row: 0.5 ms on a 853,776-row matrix).
__global__ void __launch_bounds__(256)
k_local_range(int64_t nrows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
              int32_t col0, int32_t col1, int2* __restrict__ loc) {
  const int64_t gtid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t row = gtid >> 3;
  const int l = int(gtid & 7);
  const unsigned gmask = 0xFFu << ((threadIdx.x & 31) & ~7);
  const bool valid = row < nrows;
  const int64_t s = valid ? rowptr[row] : 0, e = valid ? rowptr[row + 1] : 0;
  int res[2];
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const int32_t v = which == 0 ? col0 : col1;
    int64_t lo = s, hi = e;  // first position with colind >= v lies in [lo, hi]
    while (__any_sync(0xffffffffu, hi - lo > 0)) {
      // 8 probes split [lo, hi) in 9 pieces; count the probes whose column is below v
      const int64_t len = hi - lo;
      const int64_t pos = lo + (len * (l + 1)) / 9;
      const bool below = len > 0 && pos < hi && colind[pos] < v;
      const unsigned m = __ballot_sync(0xffffffffu, below) & gmask;
      const int nb = __popc(m);  // probes are ascending: the first nb are below
      if (len > 0) {
        const int64_t nlo = nb == 0 ? lo : lo + (len * nb) / 9 + 1;
        const int64_t nhi = nb == 8 ? hi : lo + (len * (nb + 1)) / 9;
        lo = nlo;
        hi = nhi < nlo ? nlo : nhi;
      }
    }
    res[which] = int(lo - s);
  }
  if (valid && l == 0) loc[row] = make_int2(res[0], res[1]);
}
template <int TPR, int PART>
void launch_part(b2ci_ctx* ctx, const b2ci_csr* m, const double* x, double* y) {
  const int64_t threads = m->nrows * TPR;
  const unsigned grid = unsigned((threads + 255) / 256);
  k_spmv_part<TPR, PART><<<grid, 256, 0, ctx->stream>>>(m->nrows, m->rowptr, static_cast<const int2*>(m->loc_range), m->colind,
                                                        m->nzval, x, y);
  ctx->launches++;
  B2_CHECK_LAUNCH();
}

// ---- row-binned product for matrices with skewed row lengths (selected-CI lists: a few hundred to several
// thousand non-zeros per row). One lane count per launch wastes lanes on the short rows and leaves the long
// ones to a single warp; here the rows are sorted into four length classes once per matrix and every class
// is multiplied with its own lanes-per-row, rows of a class side by side in a CTA (same life time).
constexpr int NBINS = 4;
__host__ __device__ __forceinline__ int bin_of(int64_t len) { return len < 48 ? 0 : (len < 384 ? 1 : (len < 3072 ? 2 : 3)); }
__global__ void k_bin_flags(int64_t nrows, const int64_t* __restrict__ rowptr, int32_t* __restrict__ flags /* NBINS x nrows */,
                            unsigned long long* __restrict__ minmax /* [0] min, [1] max */) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int64_t len = rowptr[row + 1] - rowptr[row];
  const int b = bin_of(len);
#pragma unroll
  for (int k = 0; k < NBINS; ++k) flags[int64_t(k) * nrows + row] = k == b ? 1 : 0;
  atomicMin(minmax, (unsigned long long)len);
  atomicMax(minmax + 1, (unsigned long long)len);
}
__global__ void k_bin_scatter(int64_t nrows, const int32_t* __restrict__ flags, const int32_t* __restrict__ excl,
                              int32_t* __restrict__ list) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row < nrows && flags[row]) list[excl[row]] = int32_t(row);
}
template <int TPR>
__global__ void __launch_bounds__(256)
k_spmv_list(int64_t nlist, const int32_t* __restrict__ list, const int64_t* __restrict__ rowptr,
            const int32_t* __restrict__ colind, const double* __restrict__ nzval, const double* __restrict__ x,
            double* __restrict__ y) {
  const int64_t gtid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t li = gtid / TPR;
  const int sub = int(gtid % TPR);
  double acc0 = 0., acc1 = 0.;
  int64_t row = 0;
  if (li < nlist) {
    row = list[li];
    range_sum<TPR>(rowptr[row], rowptr[row + 1], sub, colind, nzval, x, acc0, acc1);
  }
  double acc = acc0 + acc1;
#pragma unroll
  for (int d = TPR >> 1; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d, TPR);
  if (li < nlist && sub == 0) y[row] = acc;
}
// rows of 3072 and more non-zeros: one CTA per row, eight warps on consecutive slices of it
__global__ void __launch_bounds__(256)
k_spmv_long(int64_t nlist, const int32_t* __restrict__ list, const int64_t* __restrict__ rowptr,
            const int32_t* __restrict__ colind, const double* __restrict__ nzval, const double* __restrict__ x,
            double* __restrict__ y) {
  __shared__ double part[8];
  const int row = list[blockIdx.x];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t s = rowptr[row], e = rowptr[row + 1];
  const int64_t chunk = (((e - s) + 7) / 8 + 1) & ~int64_t(1);  // even slice lengths keep the 16-byte alignment pattern
  const int64_t a = min(e, s + w * chunk), b = min(e, a + chunk);
  double acc0 = 0., acc1 = 0.;
  range_sum<32>(a, b, lane, colind, nzval, x, acc0, acc1);
  double acc = acc0 + acc1;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
  if (lane == 0) part[w] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k];
    y[row] = t;
  }
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

void spmv_launch_part(b2ci_ctx* ctx, const b2ci_csr* m, int part, const double* x, double* y);
// length classes of the rows, built at the first product of a matrix. OFF by default (B2CI_SPMV_BINS=1 turns it
// on): measured on the N2-like ASCI(14e,26o) matrix of 1e6 determinants (1.33e9 non-zeros, 100 .. 5000 per row) the
// binned product takes 3.40 ms against 3.27 ms for the single launch -- the row lengths are not what keeps that
// matrix at 0.75 of the HBM peak (full-CI: 0.92); the scattered 8-byte gathers of x are (one 32-byte L2 sector
// per gather, 2.7x the matrix bytes over the L2 -> SM path).
static void prepare_bins(b2ci_ctx* ctx, b2ci_csr* m) {
  m->bins_tried = true;
  const char* env = getenv("B2CI_SPMV_BINS");
  if (m->nrows < 4096 || !env || atoi(env) == 0) return;
  cudaStream_t st = ctx->stream;
  const int64_t n = m->nrows;
  DevBuf<int32_t> flags(size_t(NBINS) * n), excl(n + 1);
  DevBuf<unsigned long long> mm(2);
  const unsigned long long init[2] = {~0ull, 0ull};
  B2_CUDA(cudaMemcpyAsync(mm, init, 16, cudaMemcpyHostToDevice, st));
  const unsigned g = unsigned((n + 255) / 256);
  k_bin_flags<<<g, 256, 0, st>>>(n, m->rowptr, flags, mm);
  ctx->launches++;
  unsigned long long hmm[2];
  B2_CUDA(cudaMemcpyAsync(hmm, mm, 16, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  const double mean = double(m->nnz) / double(n);
  // uniform rows (full-CI matrices): the single-launch kernel is already at the roofline
  if (double(hmm[1]) < 1.5 * mean && double(hmm[0]) > 0.66 * mean) return;
  int32_t* list = static_cast<int32_t*>(dev_alloc(size_t(n) * sizeof(int32_t)));
  int64_t off = 0;
  for (int k = 0; k < NBINS; ++k) {
    exclusive_scan_i32(ctx, flags.p + size_t(k) * n, excl, n);
    int32_t cnt = 0;
    B2_CUDA(cudaMemcpyAsync(&cnt, excl.p + n, 4, cudaMemcpyDeviceToHost, st));
    k_bin_scatter<<<g, 256, 0, st>>>(n, flags.p + size_t(k) * n, excl, list + off);
    ctx->launches++;
    B2_CUDA(cudaStreamSynchronize(st));
    m->bin_off[k] = off;
    off += cnt;
  }
  m->bin_off[NBINS] = off;
  m->bin_list = list;
  B2_CHECK_LAUNCH();
}
void spmv_launch(b2ci_ctx* ctx, const b2ci_csr* m, const double* x, double* y) {
  if (m->nrows == 0) return;
  if (!m->bins_tried) prepare_bins(ctx, const_cast<b2ci_csr*>(m));
  if (m->bin_list) {
    const int32_t* L = static_cast<const int32_t*>(m->bin_list);
    cudaStream_t st = ctx->stream;
    auto nb = [&](int k) { return m->bin_off[k + 1] - m->bin_off[k]; };
    if (nb(0)) k_spmv_list<4><<<unsigned((nb(0) * 4 + 255) / 256), 256, 0, st>>>(nb(0), L + m->bin_off[0], m->rowptr, m->colind, m->nzval, x, y);
    if (nb(1)) k_spmv_list<16><<<unsigned((nb(1) * 16 + 255) / 256), 256, 0, st>>>(nb(1), L + m->bin_off[1], m->rowptr, m->colind, m->nzval, x, y);
    if (nb(2)) k_spmv_list<32><<<unsigned((nb(2) * 32 + 255) / 256), 256, 0, st>>>(nb(2), L + m->bin_off[2], m->rowptr, m->colind, m->nzval, x, y);
    if (nb(3)) k_spmv_long<<<unsigned(nb(3)), 256, 0, st>>>(nb(3), L + m->bin_off[3], m->rowptr, m->colind, m->nzval, x, y);
    ctx->launches += (nb(0) > 0) + (nb(1) > 0) + (nb(2) > 0) + (nb(3) > 0);
    B2_CHECK_LAUNCH();
    return;
  }
  if (getenv("B2CI_SPMV_SPLIT_TEST") && m->row_begin == 0 && m->nrows == m->ncols) {
    // (profiling hook: the two-part product of the sharded sigma on one GPU, columns split in the middle)
    b2ci_csr* mm = const_cast<b2ci_csr*>(m);
    if (!mm->loc_range) {
      mm->loc_range = dev_alloc(size_t(m->nrows) * sizeof(int2));
      k_local_range<<<unsigned((m->nrows * 8 + 255) / 256), 256, 0, ctx->stream>>>(m->nrows, m->rowptr, m->colind, 0,
                                                                                  int32_t(m->ncols / 2), static_cast<int2*>(mm->loc_range));
    }
    const int64_t save = mm->nrows;
    mm->nrows = save;  // part 0 computes its lane count from nrows / ncols: halve it through a temporary view
    b2ci_csr v = *mm;
    v.ncols = m->ncols;
    launch_part<32, 0>(ctx, &v, x, y);
    launch_part<32, 1>(ctx, &v, x, y);
    v.row_offsets.clear();
    return;
  }
  const double mean = double(m->nnz) / double(m->nrows);
  if (const char* env = getenv("B2CI_SPMV_TPR")) {  // (tuning hook: lanes per row)
    switch (atoi(env)) {
      case 2: launch<2>(ctx, m, x, y); return;
      case 4: launch<4>(ctx, m, x, y); return;
      case 8: launch<8>(ctx, m, x, y); return;
      case 16: launch<16>(ctx, m, x, y); return;
      case 32: launch<32>(ctx, m, x, y); return;
      default: break;
    }
  }
  if (mean >= 96.) launch<32>(ctx, m, x, y);
  else if (mean >= 48.) launch<16>(ctx, m, x, y);
  else if (mean >= 24.) launch<8>(ctx, m, x, y);
  else if (mean >= 12.) launch<4>(ctx, m, x, y);
  else launch<2>(ctx, m, x, y);
}

// local-column sub-range of every row (device array of nrows int2, owned by the matrix)
void spmv_prepare_parts(b2ci_ctx* ctx, b2ci_csr* m) {
  if (m->loc_range || m->nrows == 0) return;
  m->loc_range = dev_alloc(size_t(m->nrows) * sizeof(int2));
  k_local_range<<<unsigned((m->nrows * 8 + 255) / 256), 256, 0, ctx->stream>>>(
      m->nrows, m->rowptr, m->colind, int32_t(m->row_begin), int32_t(m->row_begin + m->nrows), static_cast<int2*>(m->loc_range));
  ctx->launches++;
  B2_CHECK_LAUNCH();
}
// part 0: y = A(:, own columns) x_own (x_local = this rank's block); part 1: y += A(:, other columns) x_full
void spmv_launch_part(b2ci_ctx* ctx, const b2ci_csr* m, int part, const double* x, double* y) {
  if (m->nrows == 0) return;
  // the own-column share of a row is ~1 / nranks of it: fewer lanes per row for part 0
  const double mean = double(m->nnz) / double(m->nrows) * (part == 0 ? double(m->nrows) / double(m->ncols) : 1.0);
  const double* xx = part == 0 ? x - m->row_begin : x;
  if (part == 0) {
    if (mean >= 96.) launch_part<32, 0>(ctx, m, xx, y);
    else if (mean >= 24.) launch_part<8, 0>(ctx, m, xx, y);
    else launch_part<2, 0>(ctx, m, xx, y);
  } else {
    if (mean >= 96.) launch_part<32, 1>(ctx, m, xx, y);
    else if (mean >= 24.) launch_part<8, 1>(ctx, m, xx, y);
    else launch_part<2, 1>(ctx, m, xx, y);
  }
}

}  // namespace b2ci
