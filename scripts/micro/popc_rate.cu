// micro-benchmark: POPC / LOP3 / broadcast-LDS issue rates on one B200 (informs the H-build row scan design)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  __shared__ uint32_t sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 2654435761u + seed;
  __syncthreads();
  uint32_t b = threadIdx.x * 2246822519u + seed, acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 32; u += 4) {
      const uint4 s = *reinterpret_cast<const uint4*>(&sm[(it * 32 + u) & 1023]);
      if (MODE == 0) {  // xor + popc + compare-accumulate (the scan's test)
        acc += (__popc(b ^ s.x) <= 2) + (__popc(b ^ s.y) <= 2) + (__popc(b ^ s.z) <= 2) + (__popc(b ^ s.w) <= 2);
      } else if (MODE == 1) {  // bit trick: clear lowest set bit twice
        uint32_t x0 = b ^ s.x, x1 = b ^ s.y, x2 = b ^ s.z, x3 = b ^ s.w;
        x0 &= x0 - 1; x1 &= x1 - 1; x2 &= x2 - 1; x3 &= x3 - 1;
        x0 &= x0 - 1; x1 &= x1 - 1; x2 &= x2 - 1; x3 &= x3 - 1;
        acc += (x0 == 0) + (x1 == 0) + (x2 == 0) + (x3 == 0);
      } else {  // half popc, half bit trick
        uint32_t x0 = b ^ s.x, x1 = b ^ s.y;
        x0 &= x0 - 1; x1 &= x1 - 1; x0 &= x0 - 1; x1 &= x1 - 1;
        acc += (x0 == 0) + (x1 == 0) + (__popc(b ^ s.z) <= 2) + (__popc(b ^ s.w) <= 2);
      }
    }
    b += acc;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  uint32_t* out;
  const int grid = 148 * 8, block = 256, iters = 4096;
  cudaMalloc(&out, size_t(grid) * block * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, block>>>(out, iters, rep);
      if (mode == 1) k<1><<<grid, block>>>(out, iters, rep);
      if (mode == 2) k<2><<<grid, block>>>(out, iters, rep);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double tests = double(grid) * block * iters * 32.0;
      if (rep == 2) printf("mode %d: %.3f ms, %.3e tests/s, %.2f tests/clk/SM at 1.9 GHz\n", mode, ms, tests / (ms * 1e-3), tests / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
