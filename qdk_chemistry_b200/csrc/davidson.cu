// Single-root Davidson eigensolver with all O(N) work on the device.
//
// Replaces macis::davidson (external/macis/include/macis/solvers/davidson.hpp:259-372),
// gram_schmidt (:185-238), lobpcgxx::rayleigh_ritz (external/macis/src/lobpcgxx/include/
// lobpcgxx/rayleigh_ritz.hpp:70-77), diagonal_guess (:106-113) and the guess policy of
// serial_selected_ci_diag (solvers/selected_ci_diag.hpp:111-158). Semantics kept: no
// restart, max_m = min(max_m, N), CGS2 with the canonical-basis fallback, diagonal
// preconditioner with the 1e-12 denominator clamp, non-convergence is an error.
//
// Differences that do not change the mathematics: the Rayleigh-Ritz matrix V^T A V is
// extended by one row per iteration (O(N k)) instead of being recomputed (O(N k^2)); the
// k x k symmetric eigenproblem is solved on the host (Householder + implicit QL) and
// replicated on every rank instead of rank-0 + broadcast (davidson.hpp:501-531).
// Row-sharded operation (one process per GPU): V/AV hold the local rows only, the trial
// vector is all-gathered before each sigma and every inner product is all-reduced.
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace b2ci {

// ---------------------------------------------------------------- host eigen solver
// Symmetric eigenproblem, lower triangle of column-major A (lda), eigenvalues ascending in W,
// eigenvectors returned in the columns of A. Householder tridiagonalisation followed by the
// implicit QL algorithm (the classic EISPACK tred2/tql2 pair).
void sym_eig_lower(int n, double* A, int lda, double* W) {
  if (n <= 0) return;
  std::vector<double> Vm(size_t(n) * n), d(n), e(n);
  auto V = [&](int i, int j) -> double& { return Vm[size_t(i) * n + j]; };
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) V(i, j) = V(j, i) = A[i + size_t(j) * lda];
  // --- tridiagonalise
  for (int j = 0; j < n; ++j) d[j] = V(n - 1, j);
  for (int i = n - 1; i > 0; --i) {
    double scale = 0., h = 0.;
    for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
    if (scale == 0.) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; ++j) { d[j] = V(i - 1, j); V(i, j) = 0.; V(j, i) = 0.; }
    } else {
      for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = std::sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; ++j) e[j] = 0.;
      for (int j = 0; j < i; ++j) {
        f = d[j];
        V(j, i) = f;
        g = e[j] + V(j, j) * f;
        for (int k = j + 1; k <= i - 1; ++k) { g += V(k, j) * d[k]; e[k] += V(k, j) * f; }
        e[j] = g;
      }
      f = 0.;
      for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
      for (int j = 0; j < i; ++j) {
        f = d[j];
        g = e[j];
        for (int k = j; k <= i - 1; ++k) V(k, j) -= (f * e[k] + g * d[k]);
        d[j] = V(i - 1, j);
        V(i, j) = 0.;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; ++i) {
    V(n - 1, i) = V(i, i);
    V(i, i) = 1.;
    const double h = d[i + 1];
    if (h != 0.) {
      for (int k = 0; k <= i; ++k) d[k] = V(k, i + 1) / h;
      for (int j = 0; j <= i; ++j) {
        double g = 0.;
        for (int k = 0; k <= i; ++k) g += V(k, i + 1) * V(k, j);
        for (int k = 0; k <= i; ++k) V(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; ++k) V(k, i + 1) = 0.;
  }
  for (int j = 0; j < n; ++j) { d[j] = V(n - 1, j); V(n - 1, j) = 0.; }
  V(n - 1, n - 1) = 1.;
  e[0] = 0.;
  // --- implicit QL
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.;
  double f = 0., tst1 = 0.;
  const double eps = std::ldexp(1.0, -52);
  for (int l = 0; l < n; ++l) {
    tst1 = std::fmax(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int m = l;
    while (m < n) {
      if (std::fabs(e[m]) <= eps * tst1) break;
      ++m;
    }
    if (m > l) {
      int iter = 0;
      do {
        ++iter;
        double g = d[l];
        double p = (d[l + 1] - g) / (2. * e[l]);
        double r = std::hypot(p, 1.);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[m];
        double c = 1., c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0., s2 = 0.;
        for (int i = m - 1; i >= l; --i) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; ++k) {
            h = V(k, i + 1);
            V(k, i + 1) = s * V(k, i) + c * h;
            V(k, i) = c * V(k, i) - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] += f;
    e[l] = 0.;
  }
  // --- ascending order
  for (int i = 0; i < n - 1; ++i) {
    int k = i;
    double p = d[i];
    for (int j = i + 1; j < n; ++j)
      if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i];
      d[i] = p;
      for (int j = 0; j < n; ++j) std::swap(V(j, i), V(j, k));
    }
  }
  for (int j = 0; j < n; ++j) {
    W[j] = d[j];
    for (int i = 0; i < n; ++i) A[i + size_t(j) * lda] = V(i, j);
  }
}

// Lowest eigenpair only -- what single-root Davidson needs from the Rayleigh-Ritz step. Householder
// reduction to tridiagonal form keeping the reflectors (no accumulation of Q), Sturm bisection for
// the lowest eigenvalue, inverse iteration on the tridiagonal matrix, back-transformation of that
// one vector: ~(4/3) n^3 flops instead of the ~10 n^3 of the full decomposition, which at
// n = max_m = 200 would otherwise cost more than a sigma application of a 10^6-determinant matrix.
void sym_eig_lowest(int n, const double* A, int lda, double* lambda, double* vec) {
  if (n <= 0) return;
  if (n == 1) { *lambda = A[0]; vec[0] = 1.0; return; }
  std::vector<double> a(size_t(n) * n), d(n), e(n, 0.0), tau(n, 0.0), p(n), w(n);
  auto M = [&](int i, int j) -> double& { return a[size_t(j) * n + i]; };  // column-major, lower used
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) M(i, j) = A[i + size_t(j) * lda];
  for (int i = 0; i < n - 1; ++i) {
    const int m = n - i - 1;  // reflector acts on rows i+1 .. n-1
    double* x = &M(i + 1, i);
    const double alpha = x[0];
    // (std::hypot accumulation: slow, but the bits of the Ritz vector decide ties at the ASCI cut in the
    // natural-orbital runs that are pinned to the reference -- keep them)
    double xnorm = 0.0;
    for (int k = 1; k < m; ++k) xnorm = std::hypot(xnorm, x[k]);
    d[i] = M(i, i);
    if (xnorm == 0.0) {
      tau[i] = 0.0;
      e[i] = alpha;
      x[0] = 1.0;
      continue;
    }
    const double beta = -std::copysign(std::hypot(alpha, xnorm), alpha);
    tau[i] = (beta - alpha) / beta;
    const double sc = 1.0 / (alpha - beta);
    for (int k = 1; k < m; ++k) x[k] *= sc;
    x[0] = 1.0;
    e[i] = beta;
    // p = tau * A22 v (A22 symmetric, lower part stored), columnwise for unit stride
    for (int r = 0; r < m; ++r) p[r] = 0.0;
    for (int c = 0; c < m; ++c) {
      const double* col = &M(i + 1, i + 1 + c);  // entries (i+1.., i+1+c); rows >= c valid
      const double vc = x[c];
      double acc = col[c] * vc;
      for (int r = c + 1; r < m; ++r) {
        p[r] += col[r] * vc;
        acc += col[r] * x[r];
      }
      p[c] += acc;
    }
    double pv = 0.0;
    for (int r = 0; r < m; ++r) { p[r] *= tau[i]; pv += p[r] * x[r]; }
    const double K = -0.5 * tau[i] * pv;
    for (int r = 0; r < m; ++r) w[r] = p[r] + K * x[r];
    for (int c = 0; c < m; ++c) {
      double* col = &M(i + 1, i + 1 + c);
      const double vc = x[c], wc = w[c];
      for (int r = c; r < m; ++r) col[r] -= x[r] * wc + w[r] * vc;
    }
  }
  d[n - 1] = M(n - 1, n - 1);
  // ---- lowest eigenvalue of tridiag(d, e) by bisection on the Sturm count
  double lo = d[0], hi = d[0], tnorm = 0.0;
  for (int i = 0; i < n; ++i) {
    const double r = (i > 0 ? std::fabs(e[i - 1]) : 0.0) + (i < n - 1 ? std::fabs(e[i]) : 0.0);
    lo = std::min(lo, d[i] - r);
    hi = std::max(hi, d[i] + r);
    tnorm = std::max(tnorm, std::fabs(d[i]) + r);
  }
  const double eps = std::ldexp(1.0, -52);
  const double tiny = std::max(tnorm, 1.0) * eps * eps;
  auto count_below = [&](double x) {  // number of eigenvalues < x
    int c = 0;
    double q = d[0] - x;
    if (q < 0) ++c;
    for (int i = 1; i < n; ++i) {
      if (std::fabs(q) < tiny) q = q < 0 ? -tiny : tiny;
      q = d[i] - x - e[i - 1] * e[i - 1] / q;
      if (q < 0) ++c;
    }
    return c;
  };
  double a0 = lo, b0 = hi;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (a0 + b0);
    if (mid <= a0 || mid >= b0) break;
    if (count_below(mid) >= 1) b0 = mid; else a0 = mid;
  }
  const double lam = 0.5 * (a0 + b0);
  // ---- inverse iteration: (T - lam I) y = y_prev, tridiagonal LU with partial pivoting
  std::vector<double> dl(n), dd(n), du(n), du2(n), y(n);
  std::vector<int> piv(n);
  for (int i = 0; i < n; ++i) { dd[i] = d[i] - lam; dl[i] = i < n - 1 ? e[i] : 0.0; du[i] = i < n - 1 ? e[i] : 0.0; du2[i] = 0.0; }
  const double pert = std::max(tnorm, 1.0) * eps;
  for (int i = 0; i < n - 1; ++i) {
    if (std::fabs(dd[i]) >= std::fabs(dl[i])) {
      piv[i] = 0;
      if (std::fabs(dd[i]) < pert) dd[i] = dd[i] < 0 ? -pert : pert;
      const double f = dl[i] / dd[i];
      dl[i] = f;
      dd[i + 1] -= f * du[i];
    } else {
      piv[i] = 1;  // swap rows i and i+1
      const double f = dd[i] / dl[i];
      dd[i] = dl[i];
      dl[i] = f;
      const double t = du[i];
      du[i] = dd[i + 1];
      dd[i + 1] = t - f * dd[i + 1];
      if (i < n - 2) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
    }
  }
  if (std::fabs(dd[n - 1]) < pert) dd[n - 1] = dd[n - 1] < 0 ? -pert : pert;
  for (int i = 0; i < n; ++i) y[i] = 1.0 / std::sqrt(double(n)) * (1.0 + 0.01 * ((i * 7919) % 13));
  for (int iter = 0; iter < 4; ++iter) {
    for (int i = 0; i < n - 1; ++i) {  // forward: L^-1 P y
      if (piv[i]) std::swap(y[i], y[i + 1]);
      y[i + 1] -= dl[i] * y[i];
    }
    y[n - 1] /= dd[n - 1];  // backward: U^-1
    if (n > 1) y[n - 2] = (y[n - 2] - du[n - 2] * y[n - 1]) / dd[n - 2];
    for (int i = n - 3; i >= 0; --i) y[i] = (y[i] - du[i] * y[i + 1] - du2[i] * y[i + 2]) / dd[i];
    double nrm = 0.0;
    for (int i = 0; i < n; ++i) nrm = std::hypot(nrm, y[i]);
    for (int i = 0; i < n; ++i) y[i] /= nrm;
  }
  // ---- back-transformation x = H_0 H_1 ... H_{n-2} y
  for (int i = n - 2; i >= 0; --i) {
    if (tau[i] == 0.0) continue;
    const int m = n - i - 1;
    const double* v = &M(i + 1, i);
    double s = 0.0;
    for (int k = 0; k < m; ++k) s += v[k] * y[i + 1 + k];
    s *= tau[i];
    for (int k = 0; k < m; ++k) y[i + 1 + k] -= s * v[k];
  }
  double nrm = 0.0;
  for (int i = 0; i < n; ++i) nrm = std::hypot(nrm, y[i]);
  for (int i = 0; i < n; ++i) vec[i] = y[i] / nrm;
  // Rayleigh quotient of the back-transformed vector: second-order accurate in the vector error
  double num = 0.0;
  for (int j = 0; j < n; ++j) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += (i >= j ? A[i + size_t(j) * lda] : A[j + size_t(i) * lda]) * vec[i];
    num += t * vec[j];
  }
  *lambda = num;
}

namespace {

constexpr int DOT_THREADS = 256;
constexpr int DOT_CG = 8;  // columns per register group

// Last-CTA-done reduction: after a CTA has written its partial sums it takes a ticket; the CTA that draws the
// last one adds the partials up in CTA order (a fixed order: results do not depend on scheduling) and
// re-arms the counter. Saves the separate reduction launch after every dot product / norm of the solver.
__device__ __forceinline__ bool last_cta_done(unsigned int* counter, unsigned int nctas) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    is_last = t == nctas - 1;
    if (is_last) *counter = 0u;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}
// partial[b*k + j] = sum over the rows of LOGICAL CTA b of A[i + j*ld] * w[i]
// Thread t of logical CTA b owns rows b*256 + t + s*stride (s = 0, 1, ...; stride = nlog * 256) and adds their
// products in that order; its warp's tree, the warp-order sum and the CTA-order sum follow: the summation order (and
// with it every bit of the Rayleigh-Ritz matrix) is fixed by that ownership. What is free is how the operands
// arrive and which hardware CTA plays which logical CTA:
//  * one thread issues bulk (TMA) copies of the step's column pieces and of w into a ring of shared-memory stages,
//    signalled on mbarriers -- the bytes in flight live in shared memory, not in a thread's registers;
//  * a hardware CTA plays DOT_LB ADJACENT logical CTAs, so a piece is DOT_LB * 2 KiB of one column. The streaming
//    kernels of this file showed what the piece size is worth on this HBM: 2 KiB per column and step 3.0 TB/s
//    (this kernel with DOT_LB = 1, whether fed through registers or TMA), 4 KiB 4.6-5.3 TB/s.
constexpr int DOT_LB = 2;
constexpr int DOT_STAGES = 2;
constexpr int DOT_PTHREADS = DOT_THREADS * DOT_LB;                // threads of a hardware CTA
constexpr int DOT_STAGE_DOUBLES = (DOT_CG + 1) * DOT_PTHREADS;    // DOT_CG column pieces + the piece of w
constexpr size_t DOT_SMEM_BYTES = size_t(DOT_STAGES) * DOT_STAGE_DOUBLES * 8;
__global__ void __launch_bounds__(DOT_PTHREADS)
k_multi_dot(int64_t N, int k, const double* __restrict__ A, int64_t ld,
            const double* __restrict__ w, double* __restrict__ partial, unsigned int* __restrict__ counter,
            double* __restrict__ out, int nlog) {
  extern __shared__ __align__(128) double dot_smem[];
  __shared__ double red[DOT_CG][DOT_PTHREADS / 32];
  __shared__ __align__(8) uint64_t full[DOT_STAGES];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int lb = t / DOT_THREADS, tl = t - lb * DOT_THREADS;     // logical CTA within the hardware CTA, thread in it
  const int b = int(blockIdx.x) * DOT_LB + lb;                    // logical CTA
  const int nlb = min(DOT_LB, nlog - int(blockIdx.x) * DOT_LB);   // logical CTAs this hardware CTA plays
  const int64_t stride = int64_t(nlog) * DOT_THREADS, row_base = int64_t(blockIdx.x) * DOT_PTHREADS;
  const int niter = row_base < N ? int((N - row_base + stride - 1) / stride) : 0;
  // the columns are split into gridDim.y contiguous ranges, walked DOT_CG columns at a time (a column's sum never
  // depends on the grouping)
  const int per_y = (k + int(gridDim.y) - 1) / int(gridDim.y);
  const int jbeg = int(blockIdx.y) * per_y, jend = min(k, jbeg + per_y);
  const int nchunks = jend > jbeg ? (jend - jbeg + DOT_CG - 1) / DOT_CG : 0;
  const int total = nchunks * niter;
  if (t == 0) {
    for (int s = 0; s < DOT_STAGES; ++s) mbar_init(&full[s], 1);
  }
  __syncthreads();
  auto issue = [&](int q) {  // thread 0: the copies of step q = (chunk, row step)
    const int chunk = q / niter, s = q - chunk * niter;
    const int j0 = jbeg + chunk * DOT_CG, nc = min(DOT_CG, jend - j0);
    const int64_t r0 = row_base + int64_t(s) * stride;
    const int64_t left = N - r0;
    const int cap = nlb * DOT_THREADS;  // (the rows behind belong to logical CTA 0 of the next step)
    const int nr = left < cap ? int(left) : cap;
    const uint32_t bytes = uint32_t((nr + 1) & ~1) * 8u;  // 16-byte granules; an odd tail reads the column's padding
    const int stage = q % DOT_STAGES;
    double* buf = dot_smem + size_t(stage) * DOT_STAGE_DOUBLES;
    mbar_arrive_expect_tx(&full[stage], bytes * uint32_t(nc + 1));
    bulk_g2s(buf + DOT_CG * DOT_PTHREADS, w + r0, bytes, &full[stage]);
    for (int c = 0; c < nc; ++c) bulk_g2s(buf + c * DOT_PTHREADS, A + r0 + int64_t(j0 + c) * ld, bytes, &full[stage]);
  };
  if (t == 0)
    for (int q = 0; q < min(DOT_STAGES, total); ++q) issue(q);
  double acc[DOT_CG];
#pragma unroll
  for (int c = 0; c < DOT_CG; ++c) acc[c] = 0.;
  for (int q = 0; q < total; ++q) {
    const int stage = q % DOT_STAGES;
    mbar_wait(&full[stage], uint32_t(q / DOT_STAGES) & 1u);
    const int chunk = q / niter, s = q - chunk * niter;
    const int j0 = jbeg + chunk * DOT_CG, nc = min(DOT_CG, jend - j0);
    const int64_t r0 = row_base + int64_t(s) * stride;
    const double* buf = dot_smem + size_t(stage) * DOT_STAGE_DOUBLES;
    if (lb < nlb && r0 + t < N) {
      const double wi = buf[DOT_CG * DOT_PTHREADS + t];
#pragma unroll
      for (int c = 0; c < DOT_CG; ++c)
        if (c < nc) acc[c] = fma(buf[c * DOT_PTHREADS + t], wi, acc[c]);
    }
    __syncthreads();  // the stage is read: refill it
    if (t == 0 && q + DOT_STAGES < total) issue(q + DOT_STAGES);
    if (s == niter - 1) {  // last row step of this chunk of columns: reduce over every logical CTA
#pragma unroll
      for (int c = 0; c < DOT_CG; ++c) {
        double v = acc[c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
        if (lane == 0) red[c][warp] = v;
        acc[c] = 0.;
      }
      __syncthreads();
      if (tl < nc && lb < nlb) {
        double sum = 0.;
#pragma unroll
        for (int wv = 0; wv < DOT_THREADS / 32; ++wv) sum += red[tl][lb * (DOT_THREADS / 32) + wv];
        partial[int64_t(b) * k + j0 + tl] = sum;
      }
      __syncthreads();
    }
  }
  if (niter == 0 && lb < nlb)  // (a CTA without rows still owns its partials)
    for (int j = jbeg + tl; j < jend; j += DOT_THREADS) partial[int64_t(b) * k + j] = 0.;
  if (counter && last_cta_done(counter, gridDim.x * gridDim.y)) {
    // sum over the logical CTAs in order (the order of the former reduction kernel: same bits), with the partials
    // of 16 columns at a time staged in shared memory by all threads -- a single thread walking nlog dependent
    // global loads per column was a 30 us tail on every dot product
    constexpr int JB = 16, MAXB = 320;
    static_assert(size_t(JB) * MAXB * 8 <= DOT_SMEM_BYTES, "the staging area reuses the copy ring");
    double* stage = dot_smem;
    const unsigned nb = unsigned(nlog);
    if (nb <= MAXB) {
      for (int j0 = 0; j0 < k; j0 += JB) {
        const int nj = min(JB, k - j0);
        for (unsigned u = threadIdx.x; u < nb * nj; u += DOT_PTHREADS) {
          const unsigned bb = u / nj, jj = u - bb * nj;
          stage[jj * MAXB + bb] = partial[int64_t(bb) * k + j0 + jj];
        }
        __syncthreads();
        if (int(threadIdx.x) < nj) {
          double sum = 0.;
          for (unsigned bb = 0; bb < nb; ++bb) sum += stage[threadIdx.x * MAXB + bb];
          out[j0 + threadIdx.x] = sum;
        }
        __syncthreads();
      }
    } else {
      for (int j = threadIdx.x; j < k; j += DOT_PTHREADS) {
        double sum = 0.;
        for (unsigned bb = 0; bb < nb; ++bb) sum += partial[int64_t(bb) * k + j];
        out[j] = sum;
      }
    }
  }
}
__global__ void k_reduce_partials(int nblocks, int k, const double* __restrict__ partial,
                                  double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  double s = 0.;
  for (int b = 0; b < nblocks; ++b) s += partial[int64_t(b) * k + j];
  out[j] = s;
}
// w[i] -= sum_j V[i + j*ld] * h[j]
// w -= V h; with nrm_partial: also ||w_new||^2 (per-CTA partials, summed by the last CTA into nrm_out[0]).
// Two consecutive rows per thread (16-byte loads; ld is a multiple of 16 doubles so every column is 128-byte
// aligned); each row's sum runs over j in order, whatever the grid.
__global__ void __launch_bounds__(256)
k_project_out(int64_t N, int k, const double* __restrict__ V, int64_t ld,
              const double* __restrict__ h, double* __restrict__ w, double* __restrict__ nrm_partial,
              unsigned int* __restrict__ counter, double* __restrict__ nrm_out) {
  extern __shared__ double hs[];
  __shared__ double red[8];
  for (int j = threadIdx.x; j < k; j += blockDim.x) hs[j] = h[j];
  __syncthreads();
  const int64_t stride = int64_t(gridDim.x) * blockDim.x * 2;
  double nrm = 0.;
  for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < N; i += stride) {
    if (i + 1 < N) {
      double s0 = 0., s1 = 0.;
      int j = 0;
      for (; j + 8 <= k; j += 8) {  // eight columns in flight per thread
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const double2*>(V + i + int64_t(j + u) * ld);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double hj = hs[j + u];
          s0 = fma(v[u].x, hj, s0);
          s1 = fma(v[u].y, hj, s1);
        }
      }
      for (; j < k; ++j) {
        const double2 v = *reinterpret_cast<const double2*>(V + i + int64_t(j) * ld);
        const double hj = hs[j];
        s0 = fma(v.x, hj, s0);
        s1 = fma(v.y, hj, s1);
      }
      double2 wi = *reinterpret_cast<const double2*>(w + i);
      wi.x -= s0;
      wi.y -= s1;
      *reinterpret_cast<double2*>(w + i) = wi;
      nrm = fma(wi.x, wi.x, nrm);
      nrm = fma(wi.y, wi.y, nrm);
    } else {
      double s0 = 0.;
      for (int j = 0; j < k; ++j) s0 = fma(V[i + int64_t(j) * ld], hs[j], s0);
      const double wi = w[i] - s0;
      w[i] = wi;
      nrm = fma(wi, wi, nrm);
    }
  }
  if (!nrm_partial) return;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) nrm += __shfl_down_sync(0xffffffffu, nrm, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nrm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.;
    for (int wv = 0; wv < 8; ++wv) s += red[wv];
    nrm_partial[blockIdx.x] = s;
  }
  if (last_cta_done(counter, gridDim.x) && threadIdx.x == 0) {
    double s = 0.;
    for (unsigned b = 0; b < gridDim.x; ++b) s += nrm_partial[b];
    nrm_out[0] = s;
  }
}
__global__ void k_scale(int64_t N, double a, double* __restrict__ w) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < N; i += stride) w[i] *= a;
}
// w /= sqrt(nrm2[0]) with the squared norm on the device (left alone when the norm is not above min_norm:
// the host sees the same number and takes the reference's fallback)
__global__ void k_scale_by_norm(int64_t N, const double* __restrict__ nrm2, double min_norm, double* __restrict__ w,
                                double* __restrict__ nrm2_copy) {
  if (nrm2_copy && blockIdx.x == 0 && threadIdx.x == 0) nrm2_copy[0] = nrm2[0];
  const double nrm = sqrt(nrm2[0]);
  if (!(nrm > min_norm)) return;
  const double a = 1.0 / nrm;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < N; i += stride) w[i] *= a;
}
__global__ void k_set_unit(int64_t N, int64_t row0, int64_t idx, double* __restrict__ w) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < N; i += stride)
    w[i] = (row0 + i == idx) ? 1.0 : 0.0;
}
// X = V c ; R = AV c - lam X ; partial ||R||^2 ; W = -R / clamp(D - lam)
__global__ void __launch_bounds__(256)
k_residual(int64_t N, int k, const double* __restrict__ V, const double* __restrict__ AV,
           int64_t ld, const double* __restrict__ c, double lam, const double* __restrict__ D,
           double* __restrict__ X, double* __restrict__ Wout, double* __restrict__ partial,
           unsigned int* __restrict__ counter, double* __restrict__ out) {
  extern __shared__ double cs[];
  __shared__ double red[8];
  for (int j = threadIdx.x; j < k; j += blockDim.x) cs[j] = c[j];
  __syncthreads();
  const int64_t stride = int64_t(gridDim.x) * blockDim.x * 2;
  double nrm = 0.;
  auto finish = [&](int64_t i, double x, double ax) {
    const double r = ax - lam * x;
    nrm = fma(r, r, nrm);
    double denom = D[i] - lam;
    if (fabs(denom) < 1e-12) denom = (denom >= 0) ? 1e-12 : -1e-12;
    X[i] = x;
    Wout[i] = -r / denom;
  };
  // two consecutive rows per thread (16-byte loads of the 128-byte aligned columns); per row the sums run over j
  // in order, whatever the grid
  for (int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 2; i < N; i += stride) {
    if (i + 1 < N) {
      double x0 = 0., x1 = 0., a0 = 0., a1 = 0.;
#pragma unroll 4
      for (int j = 0; j < k; ++j) {
        const double2 v = *reinterpret_cast<const double2*>(V + i + int64_t(j) * ld);
        const double2 av = *reinterpret_cast<const double2*>(AV + i + int64_t(j) * ld);
        const double cj = cs[j];
        x0 = fma(v.x, cj, x0);
        x1 = fma(v.y, cj, x1);
        a0 = fma(av.x, cj, a0);
        a1 = fma(av.y, cj, a1);
      }
      finish(i, x0, a0);
      finish(i + 1, x1, a1);
    } else {
      double x = 0., ax = 0.;
      for (int j = 0; j < k; ++j) {
        x = fma(V[i + int64_t(j) * ld], cs[j], x);
        ax = fma(AV[i + int64_t(j) * ld], cs[j], ax);
      }
      finish(i, x, ax);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) nrm += __shfl_down_sync(0xffffffffu, nrm, d);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nrm;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.;
    for (int wv = 0; wv < 8; ++wv) s += red[wv];
    partial[blockIdx.x] = s;
  }
  if (last_cta_done(counter, gridDim.x)) {
    // (fixed-shape tree over the CTAs' partials; ||r|| only decides convergence, its last bit feeds nothing)
    double s = 0.;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += partial[b];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.;
      for (int wv = 0; wv < 8; ++wv) t += red[wv];
      out[0] = t;
    }
  }
}
// extract_diagonal_elements: first entry of the row whose column is the row's global index
__global__ void k_diag(int64_t nrows, int64_t row_begin, const int64_t* __restrict__ rowptr,
                       const int32_t* __restrict__ colind, const double* __restrict__ nzval,
                       double* __restrict__ D) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int64_t g = row_begin + r;
  // columns ascend within a row: lower bound of g (the linear walk to the middle of an 1819-entry full-CI row
  // made this 2.1 ms on Cr2 CAS(12,12))
  int64_t lo = rowptr[r], hi = rowptr[r + 1];
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (int64_t(colind[mid]) < g) lo = mid + 1;
    else hi = mid;
  }
  double d = 0.;
  if (lo < rowptr[r + 1] && int64_t(colind[lo]) == g) {
    d = nzval[lo];
  } else {  // not where a sorted row has it: an uploaded matrix may hold unsorted rows (b2ci_csr_upload)
    for (int64_t p = rowptr[r]; p < rowptr[r + 1]; ++p)
      if (int64_t(colind[p]) == g) { d = nzval[p]; break; }
  }
  D[r] = d;
}

// dynamic shared memory of the kernels that keep k coefficients there: rounded up to whole 32-byte groups, the
// unrolled loops read the coefficients with 16-byte loads (compute-sanitizer: a read 8 bytes past k * 8 for odd k)
inline size_t coeff_smem(int k) { return (size_t(k) + 4) / 4 * 32; }
struct Work {
  b2ci_ctx* ctx;
  int64_t N, ld;
  int nblocks;   // CTAs of the reductions (one partial per CTA)
  int nstream;   // CTAs of the streaming updates
  int npair;     // CTAs of the two-rows-per-thread kernels: every thread the same number of row pairs, one wave
  DevBuf<double> partial, small;  // small: k-sized device scratch
  DevBuf<double> scal;            // [0] ||r||^2 of the residual, [1] ||w||^2 after the second projection
  DevBuf<unsigned int> counter;   // ticket counter of the last-CTA reductions (re-armed by the kernels)
  std::vector<double> host_small;
};

// out (device, k doubles) = A(:, 0:k)^T w, all-reduced over the ranks; host_out: also copied back (synchronises)
void dots(Work& W, int k, const double* A, const double* w, double* host_out) {
  b2ci_ctx* ctx = W.ctx;
  const dim3 grid(unsigned((W.nblocks + DOT_LB - 1) / DOT_LB), unsigned(std::min(8, (k + DOT_CG - 1) / DOT_CG)));
  static bool smem_set[64] = {};  // per device: function attributes belong to the device's instance of the kernel
  const int dev = ctx->device & 63;
  if (!smem_set[dev]) {
    B2_CUDA(cudaFuncSetAttribute(k_multi_dot, cudaFuncAttributeMaxDynamicSharedMemorySize, int(DOT_SMEM_BYTES)));
    smem_set[dev] = true;
  }
  k_multi_dot<<<grid, DOT_PTHREADS, DOT_SMEM_BYTES, ctx->stream>>>(W.N, k, A, W.ld, w, W.partial, W.counter, W.small,
                                                                  W.nblocks);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  if (ctx->nranks > 1) comm_allreduce_sum(ctx, W.small, k);
  if (host_out) {
    B2_CUDA(cudaMemcpyAsync(host_out, W.small, size_t(k) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
  }
}
double norm2(Work& W, const double* w) {
  double s = 0.;
  dots(W, 1, w, w, &s);
  return std::sqrt(s);
}
// w -= V (V^T w); with_norm: ||w_new||^2 lands in W.scal[1] (all-reduced)
void project(Work& W, int k, const double* V, double* w, bool with_norm = false) {
  b2ci_ctx* ctx = W.ctx;
  dots(W, k, V, w, nullptr);  // h stays on the device (W.small)
  k_project_out<<<W.npair, 256, coeff_smem(k), ctx->stream>>>(W.N, k, V, W.ld, W.small, w,
                                                               with_norm ? W.partial.p : nullptr, W.counter,
                                                               W.scal.p + 1);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  if (with_norm && ctx->nranks > 1) comm_allreduce_sum(ctx, W.scal.p + 1, 1);
}
// the CGS2 step of gram_schmidt queued without touching the host: two projections, the norm of the result
// in W.scal[1], the vector scaled on the device when the norm is usable. The caller reads W.scal back at
// its next synchronisation and finishes with gram_schmidt_fallback if the norm was not above min_norm.
constexpr double GS_MIN_NORM = 1e-12;
void gram_schmidt_queue(Work& W, int k, const double* V, double* w) {
  project(W, k, V, w);
  project(W, k, V, w);
  // ||w||^2 by the same dot-product kernel (and summation order) as gram_schmidt's norm2
  dots(W, 1, w, w, nullptr);
  k_scale_by_norm<<<W.nstream, 256, 0, W.ctx->stream>>>(W.N, W.small.p, GS_MIN_NORM, w, W.scal.p + 1);
  W.ctx->launches++;
  B2_CHECK_LAUNCH();
}
void scale(Work& W, double a, double* w) {
  k_scale<<<W.nstream, 256, 0, W.ctx->stream>>>(W.N, a, w);
  W.ctx->launches++;
  B2_CHECK_LAUNCH();
}
void gram_schmidt_fallback(Work& W, int k, const double* V, double* w, int64_t row0, int64_t Nglobal);
// gram_schmidt (davidson.hpp:185-238)
void gram_schmidt(Work& W, int k, const double* V, double* w, int64_t row0, int64_t Nglobal) {
  const double min_norm = 1e-12;
  if (k <= 0) {
    const double nrm = norm2(W, w);
    if (nrm > min_norm) scale(W, 1. / nrm, w);
    return;
  }
  project(W, k, V, w);
  project(W, k, V, w);
  double nrm = norm2(W, w);
  if (nrm > min_norm) {
    scale(W, 1. / nrm, w);
    return;
  }
  gram_schmidt_fallback(W, k, V, w, row0, Nglobal);
}
// davidson.hpp:216-237: canonical unit vectors, one after the other, until one survives the projection
void gram_schmidt_fallback(Work& W, int k, const double* V, double* w, int64_t row0, int64_t Nglobal) {
  const double min_norm = 1e-12;
  double nrm = 0.;
  for (int64_t idx = 0; idx < Nglobal; ++idx) {
    k_set_unit<<<W.nblocks, 256, 0, W.ctx->stream>>>(W.N, row0, idx, w);
    W.ctx->launches++;
    project(W, k, V, w);
    nrm = norm2(W, w);
    if (nrm > min_norm) {
      scale(W, 1. / nrm, w);
      return;
    }
  }
  throw Error("gram_schmidt: Unable to find orthogonal vector - subspace may already span the entire space");
}

}  // namespace

void csr_diagonal_dev(b2ci_ctx* ctx, const b2ci_csr* m, double* D_dev) {
  if (!m->nrows) return;
  k_diag<<<unsigned((m->nrows + 255) / 256), 256, 0, ctx->stream>>>(m->nrows, m->row_begin, m->rowptr,
                                                                   m->colind, m->nzval, D_dev);
  ctx->launches++;
  B2_CHECK_LAUNCH();
}

int davidson(b2ci_ctx* ctx, const b2ci_csr* m, int64_t max_m, double tol, double* X_host,
             int use_guess_policy, int64_t* niter_out, double* eig_out, double* trace) {
  const int64_t Nloc = m->nrows, N = m->ncols, row0 = m->row_begin;
  if (!X_host) throw Error("Davidson: No Guess Provided");
  if (N <= 0) throw Error("Davidson: empty matrix");
  cudaStream_t st = ctx->stream;
  auto& T = ctx->timers;
  T["davidson.OP_DUR"] = T["davidson.RR_DUR"] = T["davidson.RES_DUR"] = T["davidson.GS_DUR"] = 0.;
  T["davidson.OP_CALLS"] = 0.;

  // row offsets of every rank (rank order tiles [0, N))
  std::vector<int64_t> row_offsets;
  if (ctx->nranks > 1 && int(m->row_offsets.size()) == ctx->nranks + 1) {
    row_offsets = m->row_offsets;  // b2ci_csr_set_row_partition / an earlier sigma
  } else if (ctx->nranks > 1) {
    std::vector<int64_t> counts;
    comm_allgather_i64_host(ctx, Nloc, counts);
    row_offsets.assign(ctx->nranks + 1, 0);
    for (int r = 0; r < ctx->nranks; ++r) row_offsets[r + 1] = row_offsets[r] + counts[r];
    if (row_offsets[ctx->rank] != row0 || row_offsets[ctx->nranks] != N)
      throw Error("b2ci_davidson: row blocks of the ranks do not tile [0, ncols) in rank order");
  } else if (Nloc != N || row0 != 0) {
    throw Error("b2ci_davidson: matrix is a row block but no communicator was initialised");
  }

  // diagonal
  DevBuf<double> D(Nloc > 0 ? Nloc : 1);
  csr_diagonal_dev(ctx, m, D);

  // guess policy (selected_ci_diag.hpp:128-142)
  if (use_guess_policy) {
    double max_c = 0.;
    for (int64_t i = 0; i < N; ++i) max_c = std::fmax(max_c, std::fabs(X_host[i]));
    if (!(max_c > 1. / double(N))) {
      DevBuf<double> Dfull;
      const double* dsrc = D;
      if (ctx->nranks > 1) {
        Dfull.alloc(N);
        comm_allgather_rows(ctx, D, Dfull, row_offsets);
        dsrc = Dfull;
      }
      std::vector<double> Dh(N);
      B2_CUDA(cudaMemcpyAsync(Dh.data(), dsrc, size_t(N) * 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      int64_t mi = 0;
      for (int64_t i = 1; i < N; ++i)
        if (Dh[i] < Dh[mi]) mi = i;
      X_host[mi] = 1.;
    }
  }

  max_m = std::min<int64_t>(max_m, N);
  DevBuf<double> xfull(N);  // gathered trial vector / final X
  if (N == 1) {
    X_host[0] = 1.0;
    DevBuf<double> ax(1);
    B2_CUDA(cudaMemcpyAsync(xfull, X_host, 8, cudaMemcpyHostToDevice, st));
    spmv_launch(ctx, m, xfull, ax);
    double AX = 0.;
    B2_CUDA(cudaMemcpyAsync(&AX, ax, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    *niter_out = 0;
    *eig_out = AX;
    return 0;
  }
  if (max_m < 1) throw Error("Davidson: max_m must be >= 1");
  if (max_m < 2) {
    // no iteration can run: report it without touching the caller's guess (the reference throws
    // "Davidson Did Not Converge!" from davidson.hpp:368 with X as it was)
    *niter_out = 0;
    *eig_out = 0.;
    set_error("Davidson Did Not Converge!");
    return B2CI_NOT_CONVERGED;
  }

  Work W;
  W.ctx = ctx;
  W.N = Nloc;
  W.ld = Nloc > 0 ? (Nloc + 15) / 16 * 16 : 16;  // columns start on 128-byte boundaries
  {
    // streaming kernels with two rows per thread: as many CTAs as stay resident together, then trimmed so that
    // every thread walks the same number of row pairs (a last sweep that is 40 % full was 30 % of the time)
    int per_sm = 0;
    B2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_residual, 256, size_t(max_m + 1) * 8));
    const int64_t resident = int64_t(std::max(1, per_sm)) * ctx->sm_count;
    const int64_t pairs = std::max<int64_t>(1, (Nloc + 1) / 2);
    const int64_t sweeps = (pairs + resident * 256 - 1) / (resident * 256);
    W.npair = int(std::max<int64_t>(1, (pairs + sweeps * 256 - 1) / (sweeps * 256)));
  }
  W.nblocks = (int)std::max<int64_t>(1, std::min<int64_t>(int64_t(ctx->sm_count) * 2, (Nloc + 255) / 256));
  W.nstream = (int)std::max<int64_t>(1, std::min<int64_t>(int64_t(ctx->sm_count) * 16, (Nloc + 255) / 256));
  W.partial.alloc(std::max(size_t(W.nblocks) * (max_m + 2), size_t(std::max(W.nstream, W.npair))));
  W.small.alloc(max_m + 2);
  W.scal.alloc(2);
  W.counter.alloc(1);
  B2_CUDA(cudaMemsetAsync(W.counter, 0, sizeof(unsigned int), st));
  const int64_t ld = W.ld;

  // The reference allocates N x (max_m + 1) for V and AV up front (davidson.hpp:293-294).
  size_t free_b = 0, total_b = 0;
  B2_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t need = size_t(ld) * size_t(max_m + 1) * 8 * 2;
  if (need > free_b)
    throw Error("Davidson: subspace storage 2 x N x (max_m+1) doubles (" + std::to_string(need >> 20) +
                " MiB) exceeds free device memory (" + std::to_string(free_b >> 20) + " MiB); lower max_m");
  DevBuf<double> V(size_t(ld) * (max_m + 1)), AV(size_t(ld) * (max_m + 1));
  DevBuf<double> cdev(max_m + 2);
  std::vector<double> C(size_t(max_m + 1) * (max_m + 1), 0.), Cw, LAM(max_m + 1, 0.), crow(max_m + 2);

  DeferredTimers timers(ctx);
  auto sigma = [&](const double* v_local, double* av_local) {
    DeferredScope t(timers, "davidson.OP_DUR");
    if (ctx->nranks > 1) sigma_sharded(ctx, const_cast<b2ci_csr*>(m), row_offsets, v_local, nullptr, av_local);
    else spmv_launch(ctx, m, v_local, av_local);
    T["davidson.OP_CALLS"] += 1.;
  };

  // V(:,0) = X (local rows)
  B2_CUDA(cudaMemcpyAsync(V, X_host + row0, size_t(Nloc) * 8, cudaMemcpyHostToDevice, st));
  sigma(V, AV);
  B2_CUDA(cudaMemcpyAsync(V + ld, AV, size_t(Nloc) * 8, cudaMemcpyDeviceToDevice, st));
  gram_schmidt(W, 1, V, V + ld, row0, N);

  bool converged = false;
  int64_t iter = 1;
  double lam = 0.;
  // C accumulates the lower triangle of V^T A V; row 0 (= V0^T A V0) first
  dots(W, 1, AV, V, crow.data());
  C[0] = crow[0];
  for (int64_t i = 1; i < max_m; ++i, ++iter) {
    const int k = int(i + 1);
    sigma(V + i * ld, AV + i * ld);
    {
      DeferredScope t(timers, "davidson.RR_DUR");
      // new row of the lower triangle: C(i, j) = V_i^T (A V_j), j = 0..i
      dots(W, k, AV, V + i * ld, crow.data());
      for (int j = 0; j < k; ++j) C[i + size_t(j) * (max_m + 1)] = crow[j];
      Cw.assign(size_t(k) * k + k, 0.);
      for (int b = 0; b < k; ++b)
        for (int a = b; a < k; ++a) Cw[a + size_t(b) * k] = C[a + size_t(b) * (max_m + 1)];
      sym_eig_lowest(k, Cw.data(), k, &LAM[0], Cw.data() + size_t(k) * k);
      std::copy(Cw.begin() + size_t(k) * k, Cw.begin() + size_t(k) * k + k, Cw.begin());
      lam = LAM[0];
      B2_CUDA(cudaMemcpyAsync(cdev, Cw.data(), size_t(k) * 8, cudaMemcpyHostToDevice, st));
    }
    // residual, preconditioned correction and its orthogonalisation are queued back to back; ||r||^2 and the
    // norm after the second projection come back in ONE synchronisation (the orthogonalisation of an
    // iteration that turns out converged is a few passes over V of wasted device time, never used)
    double res_nrm, gs_nrm;
    double* R = V + (i + 1) * ld;
    {
      DeferredScope t(timers, "davidson.RES_DUR");
      k_residual<<<W.npair, 256, coeff_smem(k), st>>>(Nloc, k, V, AV, ld, cdev, lam, D,
                                                        xfull + row0, R, W.partial, W.counter, W.scal.p);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      if (ctx->nranks > 1) comm_allreduce_sum(ctx, W.scal, 1);
    }
    {
      DeferredScope t(timers, "davidson.GS_DUR");
      gram_schmidt_queue(W, k, V, R);
      double* pin = reinterpret_cast<double*>(pinned_words(ctx));
      B2_CUDA(cudaMemcpyAsync(pin, W.scal, 16, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      res_nrm = std::sqrt(pin[0]);
      gs_nrm = std::sqrt(pin[1]);
    }
    if (trace) { trace[2 * (i - 1)] = lam; trace[2 * (i - 1) + 1] = res_nrm; }
    if (res_nrm < tol) { converged = true; break; }
    if (!(gs_nrm > GS_MIN_NORM)) {
      DeferredScope t(timers, "davidson.GS_DUR");
      gram_schmidt_fallback(W, k, V, R, row0, N);
    }
  }
  // X (local rows live in xfull + row0) -> host, full vector on every rank
  if (ctx->nranks > 1) {
    DevBuf<double> xl(Nloc > 0 ? Nloc : 1);
    B2_CUDA(cudaMemcpyAsync(xl, xfull + row0, size_t(Nloc) * 8, cudaMemcpyDeviceToDevice, st));
    comm_allgather_rows(ctx, xl, xfull, row_offsets);
  }
  B2_CUDA(cudaMemcpyAsync(X_host, xfull, size_t(N) * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  timers.resolve();
  *niter_out = iter;
  *eig_out = lam;
  if (!converged) {
    set_error("Davidson Did Not Converge!");
    return B2CI_NOT_CONVERGED;
  }
  return 0;
}

}  // namespace b2ci
