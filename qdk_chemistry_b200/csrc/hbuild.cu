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
#include <cstdlib>

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

// adjacency between bit strings (alpha runs, or the beta template of a rectangular list):
// entry = (string index << 2) | (popcount / 2), ascending in the string index. skip_zero drops
// empty strings on either side (the reference skips alpha-empty determinants, beta-empty ones
// are kept). deg_cnt (count pass, optional): per string the number of neighbours at distance
// 0, 2 and 4.
template <bool FILL>
__global__ void __launch_bounds__(256)
k_string_adjacency(const uint64_t* __restrict__ str, int32_t nstr, int maxd, int skip_zero,
                   int32_t* __restrict__ cnt, int32_t* __restrict__ deg_cnt,
                   const int64_t* __restrict__ adj_ptr, uint32_t* __restrict__ adj) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= nstr) return;
  const uint64_t a = str[r];
  int64_t out = FILL ? adj_ptr[r] : 0;
  int32_t c = 0, c0 = 0, c2 = 0, c4 = 0;
  if (!(skip_zero && a == 0)) {
    for (int32_t r0 = 0; r0 < nstr; r0 += 32) {
      const int32_t r2 = r0 + lane;
      bool ok = false;
      int d = 0;
      if (r2 < nstr) {
        const uint64_t a2 = str[r2];
        d = __popcll(a ^ a2);
        ok = !(skip_zero && a2 == 0) && d <= maxd;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (FILL && ok) adj[out + __popc(m & ((1u << lane) - 1u))] = (uint32_t(r2) << 2) | uint32_t(d >> 1);
      out += __popc(m);
      c += __popc(m);
      if (!FILL && deg_cnt) {
        c0 += __popc(__ballot_sync(0xffffffffu, ok && d == 0));
        c2 += __popc(__ballot_sync(0xffffffffu, ok && d == 2));
        c4 += __popc(__ballot_sync(0xffffffffu, ok && d == 4));
      }
    }
  }
  if (!FILL && lane == 0) {
    cnt[r] = c;
    if (deg_cnt) { deg_cnt[3 * r] = c0; deg_cnt[3 * r + 1] = c2; deg_cnt[3 * r + 2] = c4; }
  }
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

// ------------------------------------------------------------------ rectangular lists
// A determinant list is "rectangular" when it is the product of R alpha strings and one common
// sequence of Nb beta strings (index = r * Nb + k) -- every list generate_hilbert_space makes.
// Then no beta scan is needed at all: with alpha-run adjacency A(r) and beta adjacency lists
// B2(k) (distance <= 2) and B4(k) (distance <= 4), row (r, k) is exactly
//   { (r', k') : r' in A(r), k' in B_{4 - d_alpha(r, r')}(k) },
// already in ascending column order. Every lane evaluates a real matrix element.
__global__ void k_check_rect(const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ beta,
                             int64_t n, int64_t nb, int* __restrict__ bad) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t r = i / nb, k = i % nb;
  bool ok = beta[i] == beta[k] && alpha[i] == alpha[r * nb];
  if (k == 0 && r > 0) ok = ok && alpha[i] != alpha[i - 1];
  if (!ok) atomicOr(bad, 1);
}

// Per-pair metadata, computed once per (alpha run pair) / (beta template pair) instead of once
// per matrix element. Orientation everywhere: bra = lower determinant index, as the reference
// (it evaluates the upper triangle and mirrors it).
//   same-spin double : the VALUE  sign * (V(v1,o1,v2,o2) - V(v1,o2,v2,o1))  (matrix_elements.hpp:113-121)
//   single           : (hole, particle, sign) and the leading part of the single-excitation sum
//                      T(v,o) + sum_{p in occ_same(bra), ascending} G_red(p,v,o)   (:176-182);
//                      the kernel appends the other-spin terms V_red(p,v,o) in ascending p,
//                      i.e. the additions happen in exactly the reference's order.
//                      An opposite-spin double through two singles is one integral load:
//                      sign_a * sign_b * V(v1,o1,v2,o2)  (:140-151).
//   meta = o | v << 8 | (sign < 0) << 16 | dead << 17
// dead (h_thresh > 0 only): a same-spin double with |value| <= thr (dropped outright), or an
// alpha single whose integrals V(v,o,*,*) are all <= thr, i.e. every opposite-spin double
// through it would be dropped, so only its k' = k element is enumerated.
__global__ void k_dead_ov(IntsView I, double thr, unsigned char* __restrict__ dead /* n*n */) {
  const int n = I.n;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n) return;
  const size_t n2 = size_t(n) * n;
  bool all_small = thr > 0.0;
  for (size_t pq = 0; pq < n2 && all_small; ++pq) all_small = fabs(I.V[t + pq * n2]) <= thr;
  dead[t] = all_small ? 1 : 0;
}
__device__ __forceinline__ double single_lead_sum(const IntsView& I, uint64_t occ_same, unsigned o,
                                                  unsigned v) {
  const size_t n = I.n;
  double h = ldg(I.T + v + o * n);
  const double* G = I.G + v * n + o * n * n;
  for (uint64_t s = occ_same; s; s &= s - 1) h += ldg(G + lsb64(s));
  return h;
}
// one warp per string: meta / val of every adjacency entry
__global__ void __launch_bounds__(256)
k_pair_meta(IntsView I, const uint64_t* __restrict__ str, int32_t nstr,
            const int64_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj, double thr,
            const unsigned char* __restrict__ dead_ov, uint32_t* __restrict__ meta,
            double* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= nstr) return;
  const uint64_t s = str[r];
  for (int64_t e = adj_ptr[r] + lane; e < adj_ptr[r + 1]; e += 32) {
    const uint32_t pk = adj[e];
    const int64_t r2 = pk >> 2;
    const int dc = int(pk & 3u);
    const uint64_t s2 = str[r2];
    const uint64_t bra = r < r2 ? s : s2, ket = r < r2 ? s2 : s;
    uint32_t m = 0;
    double v = 0.;
    if (dc == 2) {
      v = me4(I, bra, ket, bra ^ ket);
      if (thr > 0.0 && !(fabs(v) > thr)) m = 1u << 17;
    } else if (dc == 1) {
      unsigned o, vv;
      double sg;
      sx_sign_indices(bra, ket, bra ^ ket, o, vv, sg);
      const bool dd = dead_ov && dead_ov[vv + o * I.n] != 0;
      m = o | (vv << 8) | (sg < 0 ? (1u << 16) : 0u) | (dd ? (1u << 17) : 0u);
      v = single_lead_sum(I, bra, o, vv);
    }
    meta[e] = m;
    val[e] = v;
  }
}

// Compacted alpha-run adjacency: entries that enumerate at least one column, in ascending run
// order, plus the list of the run's single excitations (leading sum + meta) in the same order.
struct __align__(8) ARec {
  uint32_t r2t;   // run index << 2 | kind: 0 self, 1 live single, 2 unit (double, or dead single)
  uint32_t meta;  // o | v << 8 | sign << 16 | is_single << 18   (singles)
};
template <bool FILL>
__global__ void __launch_bounds__(256)
k_adj_compact(int32_t nstr, const int64_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
              const uint32_t* __restrict__ meta, const double* __restrict__ val,
              int32_t* __restrict__ ecnt, int32_t* __restrict__ scnt, int32_t* __restrict__ run_cnt,
              const int64_t* __restrict__ cptr, ARec* __restrict__ crec, double* __restrict__ cval,
              const int64_t* __restrict__ sptr, double* __restrict__ slead, uint32_t* __restrict__ smeta) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= nstr) return;
  const unsigned lt = (1u << lane) - 1u;
  int64_t out = FILL ? cptr[r] : 0, sout = FILL ? sptr[r] : 0;
  uint32_t nself = 0, nlive = 0, nunit = 0, nsing = 0;
  const int64_t e0 = adj_ptr[r], e1 = adj_ptr[r + 1];
  for (int64_t eb = e0; eb < e1; eb += 32) {
    const int64_t e = eb + lane;
    const bool ev = e < e1;
    const uint32_t pk = ev ? adj[e] : 0u;
    const uint32_t m = ev ? meta[e] : 0u;
    const int dc = int(pk & 3u);
    const bool dead = (m >> 17) & 1u;
    const bool is_self = ev && dc == 0, is_live = ev && dc == 1 && !dead;
    const bool is_unit = ev && ((dc == 1 && dead) || (dc == 2 && !dead));
    const bool is_sing = ev && dc == 1;
    const unsigned ms = __ballot_sync(0xffffffffu, is_self);
    const unsigned ml = __ballot_sync(0xffffffffu, is_live);
    const unsigned mu = __ballot_sync(0xffffffffu, is_unit);
    const unsigned mg = __ballot_sync(0xffffffffu, is_sing);
    if (FILL && (is_self || is_live || is_unit)) {
      ARec rec;
      rec.r2t = (pk & ~3u) | (is_self ? 0u : (is_live ? 1u : 2u));
      rec.meta = (m & 0x1FFFFu) | (dc == 1 ? (1u << 18) : 0u);
      const int64_t pos = out + __popc((ms | ml | mu) & lt);
      crec[pos] = rec;
      cval[pos] = val[e];
    }
    if (FILL && is_sing) {
      const int64_t pos = sout + __popc(mg & lt);
      slead[pos] = val[e];
      smeta[pos] = m & 0x1FFFFu;
    }
    out += __popc(ms | ml | mu);
    sout += __popc(mg);
    nself += __popc(ms);
    nlive += __popc(ml);
    nunit += __popc(mu);
    nsing += __popc(mg);
  }
  if (!FILL && lane == 0) {
    ecnt[r] = int32_t(nself + nlive + nunit);
    scnt[r] = int32_t(nsing);
    run_cnt[4 * r] = int32_t(nself);
    run_cnt[4 * r + 1] = int32_t(nlive);
    run_cnt[4 * r + 2] = int32_t(nunit);
    run_cnt[4 * r + 3] = int32_t(nsing);
  }
}

// Beta-side record of list B2(k) (distance <= 2): k2, V offset of the (particle, hole) pair in
// both orientations, sign
struct __align__(16) B2Rec {
  uint32_t k2;
  uint32_t offa;  // v2 n^2 + o2 n^3 (bra = lower template index) | sign << 31 | is_self << 30
  uint32_t offb;  // o2 n^2 + v2 n^3 (orientation swapped)        | (k2 > k) << 31
  uint32_t pad;
};
__global__ void __launch_bounds__(256)
k_beta_rec(int n, int32_t nstr, const int64_t* __restrict__ b2_ptr, const uint32_t* __restrict__ b2,
           const uint32_t* __restrict__ b2_meta, B2Rec* __restrict__ rec) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (k >= nstr) return;
  const uint32_t n2 = uint32_t(n) * n, n3 = n2 * n;
  for (int64_t e = b2_ptr[k] + lane; e < b2_ptr[k + 1]; e += 32) {
    const uint32_t pk = b2[e], m = b2_meta[e];
    const uint32_t k2 = pk >> 2;
    const uint32_t o = m & 0xFFu, v = (m >> 8) & 0xFFu;
    B2Rec r;
    r.k2 = k2;
    r.offa = (v * n2 + o * n3) | (((m >> 16) & 1u) << 31) | ((pk & 3u) == 0 ? (1u << 30) : 0u);
    r.offb = (o * n2 + v * n3) | (k2 > uint32_t(k) ? (1u << 31) : 0u);
    r.pad = 0;
    rec[e] = r;
  }
}

struct ProdArgs {
  IntsView I;
  const uint64_t* run_alpha;  // R
  const uint64_t* tmpl_beta;  // Nb
  const int64_t* cptr;        // compacted alpha-run adjacency
  const ARec* crec;
  const double* cval;
  const int64_t* sptr;        // single excitations of every run, in adjacency order
  const double* slead;
  const uint32_t* smeta;
  const int32_t* run_cnt;     // 4 per run: self, live singles, unit, singles
  const int64_t* b2_ptr;      // beta adjacency, distance <= 2
  const B2Rec* b2rec;
  const uint32_t* b2_meta;
  const double* b2_val;
  const int64_t* b4_ptr;      // distance <= 4
  const uint32_t* b4;
  const double* b4_val;
  const double* diag;         // <D|H|D> of every row of the block
  int64_t nb;
  int64_t row_begin;
  int64_t nrows;
  double thr;
  int smem_a;                 // doubles of per-warp scratch for alpha singles
  int smem_b;                 // ... and beta singles
  int32_t* row_cnt;           // structural count (count kernel) / surviving count (fill kernel)
  const int64_t* rowptr;      // slot offsets of the fill kernel
  int32_t* colind;
  double* nzval;
};

__global__ void k_prod_struct_count(const ProdArgs A) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= A.nrows) return;
  const int64_t i = A.row_begin + row;
  const int64_t r = i / A.nb, k = i % A.nb;
  const int64_t l2 = A.b2_ptr[k + 1] - A.b2_ptr[k], l4 = A.b4_ptr[k + 1] - A.b4_ptr[k];
  const int32_t* d = A.run_cnt + 4 * r;
  const int64_t c = int64_t(d[0]) * l4 + int64_t(d[1]) * l2 + int64_t(d[2]);
  A.row_cnt[row] = int32_t(c);
}
// diagonal elements, one thread per row (matrix_elements.hpp:203-230)
__global__ void k_row_diag(const ProdArgs A, double* __restrict__ diag) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= A.nrows) return;
  const int64_t i = A.row_begin + row;
  diag[row] = me_diag(A.I, A.run_alpha[i / A.nb], A.tmpl_beta[i % A.nb]);
}

__device__ __forceinline__ double flip_sign_if(double v, unsigned neg) {
  return __hiloint2double(__double2hiint(v) ^ int((neg & 1u) << 31), __double2loint(v));
}

// One warp per row. The warp walks the compacted adjacency of its alpha run 32 entries at a
// time; inside a window, runs of unit entries are emitted with one lane per entry and every
// list entry (self: B4(k), live single: B2(k)) with one lane per list element. All control flow
// is warp-uniform, output positions are ascending, so survivors are written in order with a
// ballot prefix. Single-excitation elements (leading sum + V_red terms of the other spin, added
// in ascending orbital order) are evaluated once per row with full lanes into shared memory.
template <bool EVAL>
__device__ __forceinline__ void emit(const ProdArgs& A, bool act, int32_t j, double v, unsigned lt,
                                     int64_t& out) {
  const bool keep = act && (EVAL ? (fabs(v) > A.thr) : true);
  const unsigned km = __ballot_sync(0xffffffffu, keep);
  if (keep) {
    const int64_t pos = out + __popc(km & lt);
    A.colind[pos] = j;
    A.nzval[pos] = v;
  }
  out += __popc(km);
}

template <bool EVAL>
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_rows_product(const ProdArgs A) {
  extern __shared__ double sm_singles[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = int64_t(blockIdx.x) * ROW_WARPS + w;
  if (row >= A.nrows) return;
  double* sa = sm_singles + size_t(w) * (A.smem_a + A.smem_b);
  double* sb = sa + A.smem_a;
  const int64_t i = A.row_begin + row;
  const uint32_t r = uint32_t(i / A.nb), k = uint32_t(i % A.nb);
  const uint32_t nb = uint32_t(A.nb);
  const uint64_t ai = A.run_alpha[r], bi = A.tmpl_beta[k];
  const int64_t b2s = A.b2_ptr[k], b4s = A.b4_ptr[k];
  const int len2 = int(A.b2_ptr[k + 1] - b2s), len4 = int(A.b4_ptr[k + 1] - b4s);
  const int n = A.I.n;
  const size_t n2 = size_t(n) * n;
  int64_t out = A.rowptr[row];
  const int64_t out0 = out;
  const unsigned lt = (1u << lane) - 1u;
  if (ai != 0) {
    // ---- single-excitation elements of this row
    {
      const int64_t sp = A.sptr[r];
      const int ns = int(A.sptr[r + 1] - sp);
      for (int s = lane; s < ns; s += 32) {
        const uint32_t m = A.smeta[sp + s];
        double h = A.slead[sp + s];
        const double* Vr = A.I.Vr + ((m >> 8) & 0xFFu) * n + (m & 0xFFu) * n2;
        for (uint64_t q = bi; q; q &= q - 1) h += ldg(Vr + lsb64(q));
        sa[s] = flip_sign_if(h, m >> 16);
      }
      for (int t = lane; t < len2; t += 32) {
        const uint32_t m = A.b2_meta[b2s + t];
        double h = A.b2_val[b2s + t];
        const double* Vr = A.I.Vr + ((m >> 8) & 0xFFu) * n + (m & 0xFFu) * n2;
        for (uint64_t q = ai; q; q &= q - 1) h += ldg(Vr + lsb64(q));
        sb[t] = flip_sign_if(h, m >> 16);  // the self slot is never read
      }
      __syncwarp();
    }
    const int64_t E0 = A.cptr[r], E1 = A.cptr[r + 1];
    int sord0 = 0;  // singles before the current window
    for (int64_t eb = E0; eb < E1; eb += 32) {
      const int nv = int(min(int64_t(32), E1 - eb));
      const bool ev = lane < nv;
      ARec rec;
      rec.r2t = 0; rec.meta = 0;
      double cv = 0.;
      if (ev) { rec = A.crec[eb + lane]; cv = A.cval[eb + lane]; }
      const int kind_l = int(rec.r2t & 3u);
      const bool sing_l = ev && ((rec.meta >> 18) & 1u);
      const unsigned lm = __ballot_sync(0xffffffffu, ev && kind_l != 2);  // list entries
      const unsigned gm = __ballot_sync(0xffffffffu, sing_l);             // singles
      const int sord_l = sord0 + __popc(gm & lt);
      // unit entries: value known per lane
      const int32_t j_unit = int32_t((rec.r2t >> 2) * nb + k);
      const double v_unit = sing_l ? sa[sord_l] : cv;
      int cur = 0;
      while (cur < nv) {
        const unsigned rem = lm & ~((1u << cur) - 1u);
        const int f = rem ? (__ffs(rem) - 1) : nv;
        if (f > cur) emit<EVAL>(A, lane >= cur && lane < f, j_unit, v_unit, lt, out);
        if (f >= nv) break;
        const uint32_t r2t = __shfl_sync(0xffffffffu, rec.r2t, f);
        const uint32_t r2 = r2t >> 2;
        if ((r2t & 3u) == 1u) {
          // live alpha single x B2(k): opposite-spin doubles + the same-beta single
          const uint32_t am = __shfl_sync(0xffffffffu, rec.meta, f);
          const int so = __shfl_sync(0xffffffffu, sord_l, f);
          const double* Va = A.I.V + ((am >> 8) & 0xFFu) + (am & 0xFFu) * n;
          const double vself = sa[so];
          const bool lower = r < r2;  // the row determinant is the bra
          const uint32_t base = r2 * nb;
          for (int t0 = 0; t0 < len2; t0 += 32) {
            const int t = t0 + lane;
            const bool act = t < len2;
            int32_t j = 0;
            double v = 0.;
            if (act) {
              const B2Rec br = A.b2rec[b2s + t];
              j = int32_t(base + br.k2);
              // the stored beta pair has bra = lower template index
              const bool swap_b = lower != bool(br.offb >> 31);
              const uint32_t off = (swap_b ? br.offb : br.offa) & 0x3FFFFFFFu;
              v = flip_sign_if(ldg(Va + off), (am >> 16) ^ (br.offa >> 31));
              if (br.offa & (1u << 30)) v = vself;
            }
            emit<EVAL>(A, act, j, v, lt, out);
          }
        } else {
          // same alpha string x B4(k): diagonal, beta singles, beta doubles
          const uint32_t base = r * nb;
          int t2run = 0;  // position in B2(k) of the next entry at distance <= 2
          for (int t0 = 0; t0 < len4; t0 += 32) {
            const int t = t0 + lane;
            const bool act = t < len4;
            uint32_t bpk = 3u;
            if (act) bpk = A.b4[b4s + t];
            const int db = int(bpk & 3u);
            const unsigned m01 = __ballot_sync(0xffffffffu, act && db <= 1);
            int32_t j = 0;
            double v = 0.;
            if (act) {
              j = int32_t(base + (bpk >> 2));
              if (db == 2) v = A.b4_val[b4s + t];
              else if (db == 1) v = sb[t2run + __popc(m01 & lt)];
              else v = A.diag[row];
            }
            t2run += __popc(m01);
            emit<EVAL>(A, act, j, v, lt, out);
          }
        }
        cur = f + 1;
      }
      sord0 += __popc(gm);
    }
  }
  if (lane == 0) A.row_cnt[row] = int32_t(out - out0);
}

// move the surviving prefix of every structural row slot to its final position
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_compact_rows(int64_t nrows, const int64_t* __restrict__ slot_ptr,
               const int64_t* __restrict__ rowptr, const int32_t* __restrict__ ci_in,
               const double* __restrict__ nz_in, int32_t* __restrict__ ci_out,
               double* __restrict__ nz_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const int64_t src = slot_ptr[row], dst = rowptr[row], len = rowptr[row + 1] - dst;
  for (int64_t t = lane; t < len; t += 32) {
    ci_out[dst + t] = ci_in[src + t];
    nz_out[dst + t] = nz_in[src + t];
  }
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
  ctx->timers["h_build.thresh"] = 0.;

  DevBuf<int32_t> run_of(n > 0 ? n : 1);
  DevBuf<int64_t> run_start, adj_ptr;
  DevBuf<uint64_t> run_alpha;
  DevBuf<uint32_t> adj;
  DevBuf<int32_t> run_deg;  // per run: neighbours at alpha distance 0 / 2 / 4
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
    run_deg.alloc(size_t(nruns) * 3);
    adj_ptr.alloc(nruns + 1);
    const unsigned ga = unsigned((int64_t(nruns) * 32 + 255) / 256);
    k_string_adjacency<false><<<ga, 256, 0, st>>>(run_alpha, nruns, 4, 1, acnt, run_deg, nullptr, nullptr);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32_to_i64(ctx, acnt, adj_ptr, nruns);
    int64_t nadj = 0;
    B2_CUDA(cudaMemcpyAsync(&nadj, adj_ptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    adj.alloc(nadj > 0 ? nadj : 1);
    k_string_adjacency<true><<<ga, 256, 0, st>>>(run_alpha, nruns, 4, 1, nullptr, nullptr, adj_ptr, adj);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  }

  // ---- rectangular (FCI-shaped) lists: product enumeration, no beta scan
  int64_t nb = 0;
  bool rect = false;
  if (nruns > 0 && n % nruns == 0 && !getenv("B2CI_HBUILD_FORCE_SCAN")) {
    nb = n / nruns;
    if (nb < (int64_t(1) << 29)) {
      DevBuf<int> bad(1);
      B2_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
      k_check_rect<<<unsigned((n + 255) / 256), 256, 0, st>>>(dets->alpha, dets->beta, n, nb, bad);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      int hbad = 1;
      B2_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      rect = hbad == 0;
    }
  }
  ctx->timers["h_build.rectangular"] = rect ? 1. : 0.;
  if (rect) {
    DevBuf<int64_t> b2_ptr(nb + 1), b4_ptr(nb + 1);
    DevBuf<uint32_t> b2, b4;
    {
      ScopedTimer t(ctx, "h_build.setup", true);
      const unsigned gb = unsigned((nb * 32 + 255) / 256);
      DevBuf<int32_t> bc(nb);
      for (int pass = 0; pass < 2; ++pass) {
        const int maxd = pass == 0 ? 2 : 4;
        DevBuf<int64_t>& ptr = pass == 0 ? b2_ptr : b4_ptr;
        DevBuf<uint32_t>& lst = pass == 0 ? b2 : b4;
        k_string_adjacency<false><<<gb, 256, 0, st>>>(dets->beta, int32_t(nb), maxd, 0, bc, nullptr, nullptr, nullptr);
        ctx->launches++;
        B2_CHECK_LAUNCH();
        exclusive_scan_i32_to_i64(ctx, bc, ptr, nb);
        int64_t tot = 0;
        B2_CUDA(cudaMemcpyAsync(&tot, ptr.p + nb, 8, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        lst.alloc(tot > 0 ? tot : 1);
        k_string_adjacency<true><<<gb, 256, 0, st>>>(dets->beta, int32_t(nb), maxd, 0, nullptr, nullptr, ptr, lst);
        ctx->launches++;
        B2_CHECK_LAUNCH();
      }
    }
    // per-pair metadata (values of same-spin doubles, hole/particle/sign/leading sum of singles)
    int64_t nadj_h = 0, nb2_h = 0, nb4_h = 0;
    B2_CUDA(cudaMemcpyAsync(&nadj_h, adj_ptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(&nb2_h, b2_ptr.p + nb, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(&nb4_h, b4_ptr.p + nb, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    DevBuf<double> b4_val(nb4_h > 0 ? nb4_h : 1), b2_val(nb2_h > 0 ? nb2_h : 1);
    DevBuf<uint32_t> b2_meta(nb2_h > 0 ? nb2_h : 1);
    DevBuf<B2Rec> b2rec(nb2_h > 0 ? nb2_h : 1);
    DevBuf<int32_t> run_cnt(size_t(nruns) * 4);
    DevBuf<int64_t> cptr(nruns + 1), sptr(nruns + 1);
    DevBuf<ARec> crec;
    DevBuf<double> cval, slead, diag(nrows);
    DevBuf<uint32_t> smeta;
    {
      ScopedTimer t(ctx, "h_build.setup", true);
      DevBuf<unsigned char> dead_ov(size_t(ctx->norb) * ctx->norb);
      DevBuf<uint32_t> a_meta(nadj_h > 0 ? nadj_h : 1), b4_meta(nb4_h > 0 ? nb4_h : 1);
      DevBuf<double> a_val(nadj_h > 0 ? nadj_h : 1);
      DevBuf<int32_t> ecnt(nruns), scnt(nruns);
      const int nn = ctx->norb * ctx->norb;
      const unsigned ga = unsigned((int64_t(nruns) * 32 + 255) / 256);
      const unsigned gb = unsigned((nb * 32 + 255) / 256);
      k_dead_ov<<<(nn + 127) / 128, 128, 0, st>>>(ctx->ints, thr, dead_ov);
      k_pair_meta<<<ga, 256, 0, st>>>(ctx->ints, run_alpha, nruns, adj_ptr, adj, thr, dead_ov, a_meta, a_val);
      k_pair_meta<<<gb, 256, 0, st>>>(ctx->ints, dets->beta, int32_t(nb), b2_ptr, b2, thr, nullptr, b2_meta, b2_val);
      k_pair_meta<<<gb, 256, 0, st>>>(ctx->ints, dets->beta, int32_t(nb), b4_ptr, b4, thr, nullptr, b4_meta, b4_val);
      k_beta_rec<<<gb, 256, 0, st>>>(ctx->norb, int32_t(nb), b2_ptr, b2, b2_meta, b2rec);
      k_adj_compact<false><<<ga, 256, 0, st>>>(nruns, adj_ptr, adj, a_meta, a_val, ecnt, scnt, run_cnt, nullptr,
                                              nullptr, nullptr, nullptr, nullptr, nullptr);
      ctx->launches += 6;
      B2_CHECK_LAUNCH();
      exclusive_scan_i32_to_i64(ctx, ecnt, cptr, nruns);
      exclusive_scan_i32_to_i64(ctx, scnt, sptr, nruns);
      int64_t ncadj = 0, nsing = 0;
      B2_CUDA(cudaMemcpyAsync(&ncadj, cptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(&nsing, sptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      crec.alloc(ncadj > 0 ? ncadj : 1);
      cval.alloc(ncadj > 0 ? ncadj : 1);
      slead.alloc(nsing > 0 ? nsing : 1);
      smeta.alloc(nsing > 0 ? nsing : 1);
      k_adj_compact<true><<<ga, 256, 0, st>>>(nruns, adj_ptr, adj, a_meta, a_val, nullptr, nullptr, nullptr, cptr,
                                             crec, cval, sptr, slead, smeta);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    ProdArgs P;
    P.I = ctx->ints;
    P.run_alpha = run_alpha;
    P.tmpl_beta = dets->beta;
    P.cptr = cptr;
    P.crec = crec;
    P.cval = cval;
    P.sptr = sptr; P.slead = slead; P.smeta = smeta;
    P.run_cnt = run_cnt;
    P.b2_ptr = b2_ptr; P.b2rec = b2rec; P.b2_meta = b2_meta; P.b2_val = b2_val;
    P.b4_ptr = b4_ptr; P.b4 = b4; P.b4_val = b4_val;
    P.diag = diag;
    P.nb = nb;
    P.row_begin = row_begin;
    P.nrows = nrows;
    P.thr = thr;
    // per-warp scratch for the single-excitation elements of a row: at most
    // nocc * nvirt <= floor(n/2) * ceil(n/2) singles per spin, + 1 for the self slot of B2(k)
    P.smem_a = (ctx->norb / 2) * (ctx->norb - ctx->norb / 2) + 1;
    P.smem_b = P.smem_a;
    P.rowptr = nullptr; P.colind = nullptr; P.nzval = nullptr;
    DevBuf<int64_t> slot_ptr(nrows + 1);
    int64_t nslots = 0;
    {
      ScopedTimer t(ctx, "h_build.count");
      DevBuf<int32_t> scnt(nrows);
      P.row_cnt = scnt;
      k_prod_struct_count<<<unsigned((nrows + 255) / 256), 256, 0, st>>>(P);
      k_row_diag<<<unsigned((nrows + 127) / 128), 128, 0, st>>>(P, diag);
      ctx->launches += 2;
      B2_CHECK_LAUNCH();
      exclusive_scan_i32_to_i64(ctx, scnt, slot_ptr, nrows);
      B2_CUDA(cudaMemcpyAsync(&nslots, slot_ptr.p + nrows, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
    }
    DevBuf<int32_t> ci_s(nslots > 0 ? nslots : 1), kept(nrows);
    DevBuf<double> nz_s(nslots > 0 ? nslots : 1);
    {
      ScopedTimer t(ctx, "h_build.fill");
      P.row_cnt = kept;
      P.rowptr = slot_ptr;
      P.colind = ci_s;
      P.nzval = nz_s;
      const unsigned grid = unsigned((nrows + ROW_WARPS - 1) / ROW_WARPS);
      const size_t smem = size_t(ROW_WARPS) * (P.smem_a + P.smem_b) * sizeof(double);
      if (smem > 48 * 1024) {
        B2_CUDA(cudaFuncSetAttribute(k_rows_product<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        B2_CUDA(cudaFuncSetAttribute(k_rows_product<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      }
      if (thr > 0.0) k_rows_product<true><<<grid, ROW_WARPS * 32, smem, st>>>(P);
      else k_rows_product<false><<<grid, ROW_WARPS * 32, smem, st>>>(P);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    int64_t nnz = nslots;
    if (thr > 0.0) {
      // threshold_parallel equivalent: rows were written compacted inside their structural
      // slots; when something was dropped, pack the rows (csr_matrix.hpp:317-370)
      ScopedTimer t(ctx, "h_build.thresh");
      exclusive_scan_i32_to_i64(ctx, kept, rowptr, nrows);
      B2_CUDA(cudaMemcpyAsync(&nnz, rowptr.p + nrows, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      if (nnz != nslots) {
        DevBuf<int32_t> ci_f(nnz > 0 ? nnz : 1);
        DevBuf<double> nz_f(nnz > 0 ? nnz : 1);
        k_compact_rows<<<unsigned((nrows * 32 + ROW_WARPS * 32 - 1) / (ROW_WARPS * 32)), ROW_WARPS * 32, 0, st>>>(
            nrows, slot_ptr, rowptr, ci_s, nz_s, ci_f, nz_f);
        ctx->launches++;
        B2_CHECK_LAUNCH();
        B2_CUDA(cudaStreamSynchronize(st));
        ci_s = std::move(ci_f);
        nz_s = std::move(nz_f);
      }
    } else {
      rowptr = std::move(slot_ptr);
    }
    B2_CUDA(cudaStreamSynchronize(st));
    out->nnz = nnz;
    out->rowptr = rowptr.take();
    out->colind = ci_s.take();
    out->nzval = nz_s.take();
    return;
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
