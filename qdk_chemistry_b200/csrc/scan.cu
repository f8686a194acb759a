// Exclusive prefix sums (row counts -> rowptr, flags -> compaction offsets).
// Hierarchical reduce-then-scan: tiles of 2048 elements per CTA, recursion on the tile
// sums. HBM traffic is 2 reads + 1 write of the input, negligible next to the CSR fill.
#include "common.cuh"

namespace b2ci {

namespace {
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* >= 33 */) {
  // exclusive scan of one value per thread across the CTA; returns the prefix, *total = sum
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    T w = (lane < (SCAN_THREADS >> 5)) ? smem[lane] : T(0);
    T wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      T t = __shfl_up_sync(0xffffffffu, wi, d);
      if (lane >= d) wi += t;
    }
    if (lane < (SCAN_THREADS >> 5)) smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem[32] = wi;                        // block total
  }
  __syncthreads();
  T res = smem[warp] + incl - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS)
k_tile_sums(const Tin* __restrict__ in, int64_t n, Tout* __restrict__ sums) {
  __shared__ Tout sm[40];
  const int64_t base = int64_t(blockIdx.x) * SCAN_TILE;
  Tout s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k * SCAN_THREADS + threadIdx.x;
    if (i < n) s += Tout(in[i]);
  }
  Tout total;
  block_exclusive_scan<Tout>(s, &total, sm);
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS)
k_tile_scan(const Tin* __restrict__ in, int64_t n, const Tout* __restrict__ tile_offsets,
            Tout* __restrict__ out) {
  __shared__ Tout sm[40];
  const int64_t base = int64_t(blockIdx.x) * SCAN_TILE + int64_t(threadIdx.x) * SCAN_ITEMS;
  Tout v[SCAN_ITEMS];
  Tout s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    v[k] = (i < n) ? Tout(in[i]) : Tout(0);
    s += v[k];
  }
  Tout total;
  Tout pre = block_exclusive_scan<Tout>(s, &total, sm);
  pre += tile_offsets ? tile_offsets[blockIdx.x] : Tout(0);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    const int64_t i = base + k;
    if (i < n) out[i] = pre;
    pre += v[k];
    if (i == n - 1) out[n] = pre;  // grand total in the extra slot
  }
}

template <typename Tin, typename Tout>
void scan_impl(b2ci_ctx* ctx, const Tin* in, Tout* out, int64_t n) {
  if (n <= 0) {
    B2_CUDA(cudaMemsetAsync(out, 0, sizeof(Tout), ctx->stream));
    return;
  }
  const int64_t ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (ntiles == 1) {
    k_tile_scan<Tin, Tout><<<1, SCAN_THREADS, 0, ctx->stream>>>(in, n, nullptr, out);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    return;
  }
  DevBuf<Tout> sums(ntiles), offs(ntiles + 1);
  k_tile_sums<Tin, Tout><<<(unsigned)ntiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, sums);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  scan_impl<Tout, Tout>(ctx, sums, offs, ntiles);
  k_tile_scan<Tin, Tout><<<(unsigned)ntiles, SCAN_THREADS, 0, ctx->stream>>>(in, n, offs, out);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  // temporaries are released in stream order (cudaFreeAsync), no synchronisation needed
}
}  // namespace

void exclusive_scan_i32_to_i64(b2ci_ctx* ctx, const int32_t* in, int64_t* out, int64_t n) {
  scan_impl<int32_t, int64_t>(ctx, in, out, n);
}
void exclusive_scan_i32(b2ci_ctx* ctx, const int32_t* in, int32_t* out, int64_t n) {
  scan_impl<int32_t, int32_t>(ctx, in, out, n);
}

}  // namespace b2ci
