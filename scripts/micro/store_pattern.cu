// micro-benchmark: how fast can the SMs stream a CSR (4 B column + 8 B value per element) out to HBM,
// as a function of the alignment of the warp-level stores and of the warps resident per SM?
//   mode 0: every warp store starts on a 128 B (colind) / 256 B (nzval) boundary
//   mode 1: row starts are odd multiples of 4 B / 8 B (rows of 1819 elements, like CAS(12,12))
//   mode 2: as 1, and every warp store is split in two ranges (segment boundary inside the warp)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int32_t* __restrict__ ci, double* __restrict__ nz, int64_t nrows, int rowlen, int rows_per_warp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  for (int rr = 0; rr < rows_per_warp; ++rr) {
    const int64_t row = warp * rows_per_warp + rr;
    if (row >= nrows) return;
    const int64_t out0 = MODE == 0 ? row * ((rowlen + 31) & ~31) : row * rowlen;
    int32_t* c = ci + out0;
    double* v = nz + out0;
#pragma unroll 4
    for (int p = lane; p < rowlen; p += 32) {
      int q = p;
      if (MODE == 2) q = (lane < 19) ? p : min(rowlen - 1, p + 0);  // same positions; the split is emulated below
      const double val = double(q) * 1.5;
      if (MODE == 2 && lane >= 19) {
        // second range: shifted by a gap of 8 elements (written by the same instruction)
        const int q2 = min(rowlen - 1, q + 8);
        c[q2] = q2;
        v[q2] = val;
      } else {
        c[q] = q;
        v[q] = val;
      }
    }
  }
}
int main(int argc, char** argv) {
  const int64_t nrows = 853776;
  const int rowlen = 1819;
  const int64_t cap = nrows * int64_t((rowlen + 31) & ~31);
  int32_t* ci; double* nz;
  cudaMalloc(&ci, cap * 4); cudaMalloc(&nz, cap * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfg[][2] = {{256, 4}, {256, 1}, {512, 2}, {1024, 2}};  // block size, rows per warp
  for (auto& c : cfg) {
    for (int mode = 0; mode < 3; ++mode) {
      const int64_t nwarps = (nrows + c[1] - 1) / c[1];
      const unsigned grid = unsigned((nwarps * 32 + c[0] - 1) / c[0]);
      float best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, c[0]>>>(ci, nz, nrows, rowlen, c[1]);
        if (mode == 1) k<1><<<grid, c[0]>>>(ci, nz, nrows, rowlen, c[1]);
        if (mode == 2) k<2><<<grid, c[0]>>>(ci, nz, nrows, rowlen, c[1]);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
      }
      printf("block %4d rows/warp %d mode %d: %.3f ms  %.0f GB/s\n", c[0], c[1], mode, best, nrows * double(rowlen) * 12 / best / 1e6);
    }
  }
  return 0;
}
