// Stable LSD radix sort of (u64 key, u32 payload) pairs on 8-bit digits, plus a radix
// select for the k-th largest FP64 score. Used by the ASCI search to bring equal
// determinant bitstrings together without disturbing their parent order (so the score
// accumulation order is canonical) -- the GPU counterpart of sort_and_accumulate_asci_pairs
// (external/macis/include/macis/asci/determinant_sort.hpp:115-136, 254-303) and of the
// nth_element top-k (asci/determinant_search.hpp:994-1080).
//
// Per pass: tile histogram -> exclusive scan over (digit, tile) -> stable scatter.
// HBM traffic per element and pass: 8 B (histogram read) + 12 B read + 12 B write.
#include "common.cuh"

namespace b2ci {
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 elements per CTA
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_WCHUNK = RS_TILE / RS_WARPS;   // 512 consecutive elements per warp

__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const uint64_t* __restrict__ keys, int64_t n, int shift, int64_t ntiles,
          int32_t* __restrict__ hist /* [256][ntiles] */) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = int64_t(blockIdx.x) * RS_TILE;
#pragma unroll
  for (int k = 0; k < RS_ITEMS; ++k) {
    const int64_t i = base + k * RS_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  hist[int64_t(threadIdx.x) * ntiles + blockIdx.x] = int32_t(h[threadIdx.x]);
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n,
             int shift, int64_t ntiles, const int64_t* __restrict__ offs /* [256][ntiles] */,
             uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ unsigned cnt[RS_WARPS][256];
  __shared__ unsigned long long off[RS_WARPS][256];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_THREADS) (&cnt[0][0])[k] = 0;
  __syncthreads();
  const int64_t wbase = int64_t(blockIdx.x) * RS_TILE + int64_t(w) * RS_WCHUNK;
  uint64_t key[RS_ITEMS];
  uint32_t val[RS_ITEMS];
  unsigned rank[RS_ITEMS];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    const bool valid = i < n;
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    rank[r] = 0;
    if (valid) {
      key[r] = keys[i];
      val[r] = vals[i];
      const unsigned d = unsigned(key[r] >> shift) & 0xFFu;
      const unsigned m = __match_any_sync(act, d);
      const unsigned prev = cnt[w][d];
      __syncwarp(act);
      if ((m & lt) == 0) cnt[w][d] = prev + __popc(m);  // group leader = lowest lane
      __syncwarp(act);
      rank[r] = prev + __popc(m & lt);
    }
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // one digit per thread
    unsigned long long run = (unsigned long long)offs[int64_t(d) * ntiles + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) {
      off[ww][d] = run;
      run += cnt[ww][d];
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    const int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      const unsigned d = unsigned(key[r] >> shift) & 0xFFu;
      const unsigned long long pos = off[w][d] + rank[r];
      keys_out[pos] = key[r];
      vals_out[pos] = val[r];
    }
  }
}

__global__ void k_iota_u32(uint32_t* __restrict__ v, int64_t n) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) v[i] = uint32_t(i);
}

// histogram of one 8-bit digit of the score bit patterns that match `prefix` above it
__global__ void __launch_bounds__(256)
k_select_hist(const double* __restrict__ score, int64_t n, int shift, uint64_t prefix,
              unsigned long long* __restrict__ hist /*256*/) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t b = (uint64_t)__double_as_longlong(score[i]);
    const bool match = (shift >= 56) ? true : ((b >> (shift + 8)) == (prefix >> (shift + 8)));
    if (match) atomicAdd(&h[(b >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

}  // namespace

// Sorts in place semantics: on return keys/vals hold the sorted sequence (the alt buffers are
// scratch). `shifts` lists the digit positions, least significant first.
void radix_sort_pairs(b2ci_ctx* ctx, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals,
                      uint32_t* vals_alt, int64_t n, const std::vector<int>& shifts) {
  if (n <= 1 || shifts.empty()) return;
  if (n >= (int64_t(1) << 32)) throw Error("radix_sort_pairs: more than 2^32 elements");
  cudaStream_t st = ctx->stream;
  const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
  DevBuf<int32_t> hist(size_t(256) * ntiles);
  DevBuf<int64_t> offs(size_t(256) * ntiles + 1);
  uint64_t *kin = keys, *kout = keys_alt;
  uint32_t *vin = vals, *vout = vals_alt;
  for (int shift : shifts) {
    k_rs_hist<<<unsigned(ntiles), RS_THREADS, 0, st>>>(kin, n, shift, ntiles, hist);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32_to_i64(ctx, hist, offs, 256 * ntiles);
    k_rs_scatter<<<unsigned(ntiles), RS_THREADS, 0, st>>>(kin, vin, n, shift, ntiles, offs, kout, vout);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  if (kin != keys) {
    B2_CUDA(cudaMemcpyAsync(keys, kin, size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
    B2_CUDA(cudaMemcpyAsync(vals, vin, size_t(n) * 4, cudaMemcpyDeviceToDevice, st));
  }
  B2_CUDA(cudaStreamSynchronize(st));
}

void iota_u32(b2ci_ctx* ctx, uint32_t* v, int64_t n) {
  if (!n) return;
  k_iota_u32<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(v, n);
  ctx->launches++;
  B2_CHECK_LAUNCH();
}

// k-th largest (1-based) of n non-negative finite doubles; bit patterns order like values.
// distributed: every rank passes its own scores and the same k; the 256-bin histogram of each digit is
// summed over the ranks (2 KiB per digit), so all ranks walk to the same global k-th value -- the radix
// form of the reference's distributed quickselect (util/dist_quickselect.hpp:50-293).
double select_kth_largest(b2ci_ctx* ctx, const double* score, int64_t n, int64_t k, bool distributed) {
  if (!distributed && (k < 1 || k > n)) throw Error("select_kth_largest: k out of range");
  cudaStream_t st = ctx->stream;
  DevBuf<unsigned long long> dh(256);
  unsigned long long hh[256];
  uint64_t prefix = 0;
  int64_t remaining = k;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(int64_t(ctx->sm_count) * 8, (n + 255) / 256));
  for (int shift = 56; shift >= 0; shift -= 8) {
    B2_CUDA(cudaMemsetAsync(dh, 0, 256 * 8, st));
    if (n > 0) {
      k_select_hist<<<grid, 256, 0, st>>>(score, n, shift, prefix, dh);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    if (distributed) comm_allreduce_sum_u64(ctx, dh, 256);
    B2_CUDA(cudaMemcpyAsync(hh, dh, 256 * 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    int d = 255;
    for (; d >= 0; --d) {
      if ((int64_t)hh[d] >= remaining) break;
      remaining -= (int64_t)hh[d];
    }
    if (d < 0) throw Error("select_kth_largest: histogram inconsistent");
    prefix |= uint64_t(d) << shift;
  }
  double out;
  memcpy(&out, &prefix, 8);
  return out;
}

}  // namespace b2ci
