// Determinant-space Hamiltonian build: row-partitioned CSR assembly.
//
// Replaces SortedDoubleLoopHamiltonianGenerator::make_csr_hamiltonian_block_
// (external/macis/include/macis/hamiltonian_generator/sorted_double_loop.hpp:86-451) for
// the symmetric (bra == ket) case used by selected_ci_diag. The reference evaluates the
// upper triangle under OpenMP, mirrors it with atomic row cursors, sorts every row and
// finally drops |h| <= H_thresh. Here every row is owned by one warp:
//
//   1. alpha run-length encoding  (get_unique_alpha, sd_operations.hpp:449-468)
//   2. run adjacency: runs whose alpha strings differ by <= 4 bits (XOR + popcount)
//   3. count pass : warp per row, lanes scan the beta strings of each adjacent run with
//                   XOR/popcount, hits are queued in shared memory and evaluated 32 at a
//                   time (full lanes) with the Slater-Condon rules; |h| > thr is counted
//   4. exclusive scan -> rowptr
//   5. fill pass  : same walk, ballot-compacted ordered writes -> columns ascending, no
//                   per-row sort, no atomics, both triangles computed with the SAME
//                   (bra = lower index, ket = higher index) roles as the reference so the
//                   values are bit-identical to its mirrored entries.
#include "common.cuh"
#include "slater.cuh"

namespace b2ci {
namespace {

constexpr int ROW_WARPS = 8;  // warps (rows) per CTA

__global__ void k_run_flags(const uint64_t* __restrict__ alpha, int64_t n,
                            int32_t* __restrict__ flag) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || alpha[i] != alpha[i - 1]) ? 1 : 0;
}
__global__ void k_run_scatter(const uint64_t* __restrict__ alpha, int64_t n,
                              const int32_t* __restrict__ flag,
                              const int32_t* __restrict__ excl, int32_t* __restrict__ run_of,
                              int64_t* __restrict__ run_start, uint64_t* __restrict__ run_alpha,
                              int32_t nruns) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = excl[i] + flag[i] - 1;
  run_of[i] = r;
  if (flag[i]) {
    run_start[r] = i;
    run_alpha[r] = alpha[i];
  }
  if (i == n - 1) run_start[nruns] = n;
}

// adjacency between alpha runs; entry = (run index << 2) | (popcount / 2)
template <bool FILL>
__global__ void __launch_bounds__(256)
k_run_adjacency(const uint64_t* __restrict__ run_alpha, int32_t nruns,
                int32_t* __restrict__ cnt, const int64_t* __restrict__ adj_ptr,
                uint32_t* __restrict__ adj) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= nruns) return;
  const uint64_t a = run_alpha[r];
  int64_t out = FILL ? adj_ptr[r] : 0;
  int32_t c = 0;
  if (a != 0) {
    for (int32_t r0 = 0; r0 < nruns; r0 += 32) {
      const int32_t r2 = r0 + lane;
      bool ok = false;
      int d = 0;
      if (r2 < nruns) {
        const uint64_t a2 = run_alpha[r2];
        d = __popcll(a ^ a2);
        ok = (a2 != 0) && d <= 4;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (FILL && ok) adj[out + __popc(m & ((1u << lane) - 1u))] = (uint32_t(r2) << 2) | uint32_t(d >> 1);
      out += __popc(m);
      c += __popc(m);
    }
  }
  if (!FILL && lane == 0) cnt[r] = c;
}

struct RowArgs {
  IntsView I;
  const uint64_t* alpha;
  const uint64_t* beta;
  const int32_t* run_of;
  const int64_t* run_start;
  const int64_t* adj_ptr;
  const uint32_t* adj;
  int64_t row_begin;
  int64_t nrows;
  double thr;
  int32_t* row_cnt;       // count pass output
  const int64_t* rowptr;  // fill pass input
  int32_t* colind;
  double* nzval;
};

// Evaluate up to 32 queued column indices (one per lane) and count / emit the survivors.
template <bool FILL, bool EVAL>
__device__ __forceinline__ void process_batch(const RowArgs& A, int64_t i, uint64_t ai,
                                              uint64_t bi, const int32_t* q, int nvalid, int lane,
                                              int64_t& out, int32_t& cnt) {
  const bool valid = lane < nvalid;
  int32_t j = 0;
  double v = 0.;
  bool keep = valid;
  if (valid) {
    j = q[lane];
    if (FILL || EVAL) {
      const uint64_t aj = A.alpha[j], bj = A.beta[j];
      // the reference computes the upper triangle with bra = lower index and mirrors it
      v = (i <= int64_t(j)) ? matel(A.I, ai, bi, aj, bj) : matel(A.I, aj, bj, ai, bi);
      if (EVAL) keep = fabs(v) > A.thr;
    }
  }
  const unsigned km = __ballot_sync(0xffffffffu, keep);
  if (FILL && keep) {
    const int64_t pos = out + __popc(km & ((1u << lane) - 1u));
    A.colind[pos] = j;
    A.nzval[pos] = v;
  }
  out += __popc(km);
  cnt += __popc(km);
}

template <bool FILL, bool EVAL>
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_rows(const RowArgs A) {
  __shared__ int32_t queue[ROW_WARPS][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = int64_t(blockIdx.x) * ROW_WARPS + w;
  if (row >= A.nrows) return;
  const int64_t i = A.row_begin + row;
  const uint64_t ai = A.alpha[i], bi = A.beta[i];
  const int32_t r = A.run_of[i];
  int32_t* q = queue[w];
  int qn = 0;
  int32_t cnt = 0;
  int64_t out = FILL ? A.rowptr[row] : 0;
  const unsigned lt = (1u << lane) - 1u;
  if (ai != 0) {
    const int64_t e0 = A.adj_ptr[r], e1 = A.adj_ptr[r + 1];
    for (int64_t e = e0; e < e1; ++e) {
      const uint32_t pk = A.adj[e];
      const int da = int(pk & 3u) * 2;
      const int64_t ks = A.run_start[pk >> 2], ke = A.run_start[(pk >> 2) + 1];
      for (int64_t j0 = ks; j0 < ke; j0 += 32) {
        const int64_t j = j0 + lane;
        bool hit = false;
        if (j < ke) hit = (da + __popcll(bi ^ A.beta[j])) <= 4;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
          if (hit) q[qn + __popc(m & lt)] = int32_t(j);
          qn += __popc(m);
          __syncwarp();
          if (qn >= 32) {
            process_batch<FILL, EVAL>(A, i, ai, bi, q, 32, lane, out, cnt);
            const int rest = qn - 32;
            const int32_t t = (lane < rest) ? q[32 + lane] : 0;
            __syncwarp();
            if (lane < rest) q[lane] = t;
            qn = rest;
            __syncwarp();
          }
        }
      }
    }
    if (qn > 0) process_batch<FILL, EVAL>(A, i, ai, bi, q, qn, lane, out, cnt);
  }
  if (!FILL && lane == 0) A.row_cnt[row] = cnt;
}

__global__ void k_unpack_dets(const uint64_t* __restrict__ words, int wpd, int64_t n,
                              uint64_t* __restrict__ alpha, uint64_t* __restrict__ beta) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (wpd == 1) {
    const uint64_t w = words[i];
    alpha[i] = w & 0xFFFFFFFFull;
    beta[i] = w >> 32;
  } else {
    alpha[i] = words[2 * i];
    beta[i] = words[2 * i + 1];
  }
}
__global__ void k_pack_dets(const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ beta,
                            int wpd, int64_t n, uint64_t* __restrict__ words) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (wpd == 1) words[i] = (alpha[i] & 0xFFFFFFFFull) | (beta[i] << 32);
  else { words[2 * i] = alpha[i]; words[2 * i + 1] = beta[i]; }
}

// generate_combs order (sd_operations.hpp:305-323): std::prev_permutation of a 0/1 vector
// whose first nset entries are set == combinations in DESCENDING order of the bit-reversed
// string. Thread t unranks combination t directly: walking positions 0..nbits-1, position p
// is set iff t < C(nbits-p-1, remaining-1) (the block of combinations that keep bit p).
__device__ __forceinline__ uint64_t unrank_comb(int nbits, int nset, int64_t t,
                                                const int64_t* __restrict__ binom /*65x65*/) {
  uint64_t s = 0;
  int rem = nset;
  for (int p = 0; p < nbits && rem > 0; ++p) {
    const int64_t with_p = binom[(nbits - p - 1) * 65 + (rem - 1)];
    if (t < with_p) { s |= uint64_t(1) << p; --rem; }
    else t -= with_p;
  }
  return s;
}
__global__ void k_generate_fci(int norb, int na, int nb, int64_t nalpha_str, int64_t nbeta_str,
                               const int64_t* __restrict__ binom, uint64_t* __restrict__ alpha,
                               uint64_t* __restrict__ beta) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nalpha_str * nbeta_str) return;
  alpha[i] = unrank_comb(norb, na, i / nbeta_str, binom);
  beta[i] = unrank_comb(norb, nb, i % nbeta_str, binom);
}

}  // namespace

// ------------------------------------------------------------------------------------
void dets_from_words(b2ci_ctx* ctx, const uint64_t* words_host, int wpd, int64_t n,
                     b2ci_dets* d) {
  if (wpd != 1 && wpd != 2) throw Error("words_per_det must be 1 (wfn_t<64>) or 2 (wfn_t<128>)");
  d->n = n;
  DevBuf<uint64_t> a(n), b(n), w(size_t(n) * wpd);
  if (n) {
    B2_CUDA(cudaMemcpyAsync(w, words_host, size_t(n) * wpd * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_unpack_dets<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(w, wpd, n, a, b);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  d->alpha = a.take();
  d->beta = b.take();
}

void dets_to_words(b2ci_ctx* ctx, const b2ci_dets* d, uint64_t* words_host, int wpd) {
  if (wpd != 1 && wpd != 2) throw Error("words_per_det must be 1 or 2");
  if (!d->n) return;
  DevBuf<uint64_t> w(size_t(d->n) * wpd);
  k_pack_dets<<<unsigned((d->n + 255) / 256), 256, 0, ctx->stream>>>(d->alpha, d->beta, wpd, d->n, w);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  B2_CUDA(cudaMemcpyAsync(words_host, w, size_t(d->n) * wpd * 8, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
}

void dets_generate_fci(b2ci_ctx* ctx, int norb, int na, int nb, b2ci_dets* d) {
  if (norb < 1 || norb > 64 || na < 0 || nb < 0 || na > norb || nb > norb)
    throw Error("generate_hilbert_space: invalid (norb, nalpha, nbeta)");
  std::vector<int64_t> binom(65 * 65, 0);
  for (int n = 0; n <= 64; ++n) {
    binom[n * 65 + 0] = 1;
    for (int k = 1; k <= n; ++k) {
      const __int128 v = (__int128)binom[(n - 1) * 65 + (k - 1)] + (k <= n - 1 ? binom[(n - 1) * 65 + k] : 0);
      binom[n * 65 + k] = v > (__int128)INT64_MAX ? INT64_MAX : (int64_t)v;
    }
  }
  const int64_t nas = binom[norb * 65 + na], nbs = binom[norb * 65 + nb];
  if (nas == INT64_MAX || nbs == INT64_MAX || nas > INT64_MAX / (nbs ? nbs : 1))
    throw Error("generate_hilbert_space: dimension overflows int64");
  const int64_t n = nas * nbs;
  DevBuf<int64_t> dbinom(binom.size());
  DevBuf<uint64_t> a(n), b(n);
  B2_CUDA(cudaMemcpyAsync(dbinom, binom.data(), binom.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  k_generate_fci<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(norb, na, nb, nas, nbs, dbinom, a, b);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  d->n = n;
  d->alpha = a.take();
  d->beta = b.take();
}

void hbuild_csr(b2ci_ctx* ctx, const b2ci_dets* dets, int64_t row_begin, int64_t row_end,
                double thr, b2ci_csr* out) {
  if (!ctx->ints_dev) throw Error("b2ci_hbuild_csr: integrals not uploaded");
  const int64_t n = dets->n;
  if (row_begin < 0 || row_end < row_begin || row_end > n) throw Error("b2ci_hbuild_csr: bad row range");
  if (n >= (int64_t(1) << 30)) throw Error("b2ci_hbuild_csr: more than 2^30 determinants per list");
  if (!(thr >= 0.0)) throw Error("b2ci_hbuild_csr: h_thresh must be >= 0");
  const int64_t nrows = row_end - row_begin;
  cudaStream_t st = ctx->stream;
  ctx->timers["h_build.setup"] = ctx->timers["h_build.count"] = ctx->timers["h_build.fill"] = 0.;

  DevBuf<int32_t> run_of(n > 0 ? n : 1);
  DevBuf<int64_t> run_start, adj_ptr;
  DevBuf<uint64_t> run_alpha;
  DevBuf<uint32_t> adj;
  int32_t nruns = 0;
  out->nrows = nrows;
  out->ncols = n;
  out->row_begin = row_begin;
  DevBuf<int64_t> rowptr(nrows + 1);
  if (n == 0 || nrows == 0) {
    B2_CUDA(cudaMemsetAsync(rowptr, 0, (nrows + 1) * 8, st));
    B2_CUDA(cudaStreamSynchronize(st));
    out->nnz = 0;
    out->rowptr = rowptr.take();
    return;
  }
  {
    ScopedTimer t(ctx, "h_build.setup");
    DevBuf<int32_t> flag(n), excl(n + 1);
    const unsigned gb = unsigned((n + 255) / 256);
    k_run_flags<<<gb, 256, 0, st>>>(dets->alpha, n, flag);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32(ctx, flag, excl, n);
    B2_CUDA(cudaMemcpyAsync(&nruns, excl.p + n, 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    run_start.alloc(nruns + 1);
    run_alpha.alloc(nruns);
    k_run_scatter<<<gb, 256, 0, st>>>(dets->alpha, n, flag, excl, run_of, run_start, run_alpha, nruns);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    // run adjacency (count, scan, fill)
    DevBuf<int32_t> acnt(nruns);
    adj_ptr.alloc(nruns + 1);
    const unsigned ga = unsigned((int64_t(nruns) * 32 + 255) / 256);
    k_run_adjacency<false><<<ga, 256, 0, st>>>(run_alpha, nruns, acnt, nullptr, nullptr);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32_to_i64(ctx, acnt, adj_ptr, nruns);
    int64_t nadj = 0;
    B2_CUDA(cudaMemcpyAsync(&nadj, adj_ptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    adj.alloc(nadj > 0 ? nadj : 1);
    k_run_adjacency<true><<<ga, 256, 0, st>>>(run_alpha, nruns, nullptr, adj_ptr, adj);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  }

  RowArgs A;
  A.I = ctx->ints;
  A.alpha = dets->alpha;
  A.beta = dets->beta;
  A.run_of = run_of;
  A.run_start = run_start;
  A.adj_ptr = adj_ptr;
  A.adj = adj;
  A.row_begin = row_begin;
  A.nrows = nrows;
  A.thr = thr;
  A.row_cnt = nullptr;
  A.rowptr = nullptr;
  A.colind = nullptr;
  A.nzval = nullptr;
  const unsigned grid = unsigned((nrows + ROW_WARPS - 1) / ROW_WARPS);
  int64_t nnz = 0;
  {
    ScopedTimer t(ctx, "h_build.count");
    DevBuf<int32_t> row_cnt(nrows);
    A.row_cnt = row_cnt;
    if (thr > 0.0) k_rows<false, true><<<grid, ROW_WARPS * 32, 0, st>>>(A);
    else k_rows<false, false><<<grid, ROW_WARPS * 32, 0, st>>>(A);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32_to_i64(ctx, row_cnt, rowptr, nrows);
    B2_CUDA(cudaMemcpyAsync(&nnz, rowptr.p + nrows, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
  }
  DevBuf<int32_t> colind(nnz > 0 ? nnz : 1);
  DevBuf<double> nzval(nnz > 0 ? nnz : 1);
  {
    ScopedTimer t(ctx, "h_build.fill");
    A.row_cnt = nullptr;
    A.rowptr = rowptr;
    A.colind = colind;
    A.nzval = nzval;
    if (thr > 0.0) k_rows<true, true><<<grid, ROW_WARPS * 32, 0, st>>>(A);
    else k_rows<true, false><<<grid, ROW_WARPS * 32, 0, st>>>(A);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  }
  B2_CUDA(cudaStreamSynchronize(st));
  out->nnz = nnz;
  out->rowptr = rowptr.take();
  out->colind = colind.take();
  out->nzval = nzval.take();
}

}  // namespace b2ci
