// Orbital entropies of a CI vector on the device.
//
// Replaces SortedDoubleLoopHamiltonianGenerator::form_entropies (external/macis/include/macis/
// hamiltonian_generator/sorted_double_loop.hpp:760-905): the orbital-RDM intermediates of
// external/macis/include/macis/util/entropies.hpp:62-205 are accumulated on the device over the
// pairs (i <= j) with at most one alpha and one beta excitation and |C_i C_j| > 1e-16 -- walked
// on the structural pattern of the H build like the RDMs -- and the single-orbital entropies,
// two-orbital entropies and mutual information are then assembled on the host (:476-510, 552-724,
// 875-886; O(norb^2) work). Without two-orbital quantities only the diagonal pairs contribute
// (:917-924) and no pattern is built at all.
//
// Intermediate layout (doubles): 3 vectors of n, then 18 n x n matrices M(i, j) at i + j n, in
// the order of enum Ent below (== oracle/port.py ENT_VECS + ENT_MATS).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "common.cuh"
#include "slater.cuh"

namespace b2ci {
void hbuild_csr(b2ci_ctx* ctx, const b2ci_dets* dets, int64_t row_begin, int64_t row_end, double thr,
                b2ci_csr* out);

enum Ent {
  A_IJ = 0, B_IJ, AA_IIJJ, BB_IIJJ, AB_IIJJ, AB_IJJJ, AB_JIJJ, AB_JJIJ, AB_JJJI, AB_IJIJ, AB_IJJI,
  AAB_IIJJJJ, ABB_JJIIJJ, AAB_IIJJII, ABB_IIIIJJ, ABB_IJIIJJ, AAB_IIJJIJ, AABB_IIJJIIJJ, ENT_NMAT
};
size_t entropy_intermediate_doubles(int n, bool need_s2) {
  return size_t(3) * n + (need_s2 ? size_t(ENT_NMAT) * n * n : 0);
}

namespace {

struct EntOut {
  double* v;  // a_ii | b_ii | ab_iiii
  double* m;  // matrices (NULL when only s1 is wanted)
  int n;
};
__device__ __forceinline__ void addm(const EntOut& E, int which, unsigned i, unsigned j, double x) {
  atomicAdd(E.m + (size_t(which) * E.n + j) * E.n + i, x);
}

// orbital_rdm_contrib_diag(_s1) (entropies.hpp:212-356), spread over the lanes of a warp
__device__ void ent_diag(const EntOut& E, uint64_t oa, uint64_t ob, double val, int lane) {
  const int n = E.n;
  if (lane == 0) {
    for (uint64_t s = oa; s; s &= s - 1) atomicAdd(E.v + lsb64(s), val);
    for (uint64_t s = ob; s; s &= s - 1) atomicAdd(E.v + n + lsb64(s), val);
    for (uint64_t s = oa & ob; s; s &= s - 1) atomicAdd(E.v + 2 * n + lsb64(s), val);
  }
  if (!E.m) return;
  int c = 0;
  for (uint64_t sq = oa; sq; sq &= sq - 1)
    for (uint64_t sp = oa; sp; sp &= sp - 1, ++c) {
      if ((c & 31) != lane) continue;
      const unsigned p = lsb64(sp), q = lsb64(sq);
      if (p == q) continue;
      addm(E, AA_IIJJ, p, q, val);
      const bool pb = (ob >> p) & 1, qb = (ob >> q) & 1;
      if (pb) {
        addm(E, AAB_IIJJII, p, q, val);
        if (qb) addm(E, AABB_IIJJIIJJ, p, q, val);
      }
      if (qb) addm(E, AAB_IIJJJJ, p, q, val);
    }
  for (uint64_t sq = ob; sq; sq &= sq - 1)
    for (uint64_t sp = ob; sp; sp &= sp - 1, ++c) {
      if ((c & 31) != lane) continue;
      const unsigned p = lsb64(sp), q = lsb64(sq);
      if (p == q) continue;
      addm(E, BB_IIJJ, p, q, val);
      if ((oa >> p) & 1) addm(E, ABB_JJIIJJ, q, p, val);
      if ((oa >> q) & 1) addm(E, ABB_IIIIJJ, q, p, val);
    }
  for (uint64_t sq = ob; sq; sq &= sq - 1)
    for (uint64_t sp = oa; sp; sp &= sp - 1, ++c) {
      if ((c & 31) != lane) continue;
      addm(E, AB_IIJJ, lsb64(sp), lsb64(sq), val);
    }
}
// orbital_rdm_contrib_2<transpose> (entropies.hpp:359-443)
template <bool TRANSPOSE>
__device__ void ent_single(const EntOut& E, uint64_t bra, uint64_t ket, uint64_t ex, uint64_t occ_os, double val) {
  unsigned o1, v1;
  double sign;
  sx_sign_indices(bra, ket, ex, o1, v1, sign);
  const double sv = sign * val;
  const int x = TRANSPOSE ? B_IJ : A_IJ;
  addm(E, x, v1, o1, sv);
  addm(E, x, o1, v1, sv);
  const int m1 = TRANSPOSE ? AB_JJIJ : AB_IJJJ, m2 = TRANSPOSE ? AB_JJJI : AB_JIJJ;
  const bool oin = (occ_os >> o1) & 1, vin = (occ_os >> v1) & 1;
  if (oin) { addm(E, m1, v1, o1, sv); addm(E, m2, v1, o1, sv); }
  if (vin) { addm(E, m1, o1, v1, sv); addm(E, m2, o1, v1, sv); }
  if (oin && vin) {
    const int m3 = TRANSPOSE ? AAB_IIJJIJ : ABB_IJIIJJ;
    addm(E, m3, o1, v1, sv);
    addm(E, m3, v1, o1, sv);
  }
}

// need_s2: warp per row of the structural pattern, upper triangle
__global__ void __launch_bounds__(256)
k_entropy_pairs(EntOut E, const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ beta,
                const double* __restrict__ C, int64_t row_begin, int64_t nrows,
                const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const int64_t i = row_begin + row;
  const uint64_t ba = alpha[i], bb = beta[i];
  const double ci = C[i];
  const int64_t e0 = rowptr[row], e1 = rowptr[row + 1];
  for (int64_t e = e0 + lane; e < e1; e += 32) {
    const int64_t j = colind[e];
    if (j <= i) continue;
    const uint64_t ka = alpha[j], kb = beta[j];
    const uint64_t exa = ba ^ ka, exb = bb ^ kb;
    const int ca = popc64(exa), cb = popc64(exb);
    if (ca > 2 || cb > 2) continue;
    const double val = ci * C[j];
    if (!(fabs(val) > 1e-16)) continue;
    if (ca == 2 && cb == 0) {
      ent_single<false>(E, ba, ka, exa, bb, val);
    } else if (ca == 0 && cb == 2) {
      ent_single<true>(E, bb, kb, exb, ba, val);
    } else if (ca == 2 && cb == 2) {  // orbital_rdm_contrib_22 (:445-474)
      unsigned o2, v2, o1, v1;
      double sb, sa;
      sx_sign_indices(ba, ka, exa, o2, v2, sb);
      sx_sign_indices(bb, kb, exb, o1, v1, sa);
      const double sv = sa * sb * val;
      if (o1 == o2 && v1 == v2) { addm(E, AB_IJIJ, v1, o1, sv); addm(E, AB_IJIJ, o1, v1, sv); }
      else if (o1 == v2 && v1 == o2) { addm(E, AB_IJJI, v1, o1, sv); addm(E, AB_IJJI, o1, v1, sv); }
    }
  }
  const double vd = ci * ci;
  if (e1 > e0 && fabs(vd) > 1e-16) ent_diag(E, ba, bb, vd, lane);
}
// s1 only: the diagonal pairs, one warp per determinant
__global__ void __launch_bounds__(256)
k_entropy_diag(EntOut E, const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ beta,
               const double* __restrict__ C, int64_t row_begin, int64_t nrows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const int64_t i = row_begin + row;
  const uint64_t ba = alpha[i];
  const double vd = C[i] * C[i];
  if (ba != 0 && fabs(vd) > 1e-16) ent_diag(E, ba, beta[i], vd, lane);
}

// ---- host assembly (entropies.hpp:476-510, 552-724, 875-886)
inline double xlogx(double v) { return v > std::numeric_limits<double>::epsilon() ? -v * std::log(v) : 0.0; }
inline void eig2(double a, double b, double d, double* out) {
  const double hs = 0.5 * (a + d), hd = 0.5 * (a - d), w = std::sqrt(hd * hd + b * b);
  out[0] = hs - w;
  out[1] = hs + w;
}
// cyclic Jacobi on a symmetric 4 x 4 block (what detail::eigenvalues_4x4 does, :33-86)
void eig4(double a[4][4], double* out) {
  const double eps = std::numeric_limits<double>::epsilon();
  for (int sweep = 0; sweep < 50; ++sweep) {
    double off = 0.;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) off += a[p][q] * a[p][q];
    if (off < eps * eps) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        const double apq = a[p][q];
        if (std::abs(apq) < eps) continue;
        const double theta = 0.5 * (a[q][q] - a[p][p]) / apq;
        double t = 1.0 / (std::abs(theta) + std::sqrt(1.0 + theta * theta));
        if (theta < 0.0) t = -t;
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c, tau = s / (1.0 + c);
        a[p][p] -= t * apq;
        a[q][q] += t * apq;
        a[p][q] = a[q][p] = 0.0;
        for (int r = 0; r < 4; ++r) {
          if (r == p || r == q) continue;
          const double arp = a[r][p], arq = a[r][q];
          a[r][p] = a[p][r] = arp - s * (arq + tau * arp);
          a[r][q] = a[q][r] = arq + s * (arp - tau * arq);
        }
      }
  }
  for (int k = 0; k < 4; ++k) out[k] = a[k][k];
}

}  // namespace

void host_entropies_from_intermediates(int n, bool need_s2, const double* I, double* s1, double* s2, double* mi) {
  const double *a = I, *b = I + n, *d = I + 2 * n;
  for (int i = 0; i < n; ++i)
    s1[i] = xlogx(1 - a[i] - b[i] + d[i]) + xlogx(a[i] - d[i]) + xlogx(b[i] - d[i]) + xlogx(d[i]);
  if (!need_s2 || (!s2 && !mi)) return;
  const double* M = I + 3 * size_t(n);
  auto g = [&](int which, int i, int j) { return M[(size_t(which) * n + j) * n + i]; };
  std::vector<double> s2_local;
  if (!s2) { s2_local.assign(size_t(n) * n, 0.0); s2 = s2_local.data(); }
  for (int i = 0; i < n; ++i) {
    s2[i + size_t(i) * n] = 0.0;
    for (int j = i + 1; j < n; ++j) {
      const double AA = g(AA_IIJJ, i, j), BB = g(BB_IIJJ, i, j);
      const double x1 = g(AAB_IIJJJJ, i, j), x2 = g(ABB_JJIIJJ, i, j), x3 = g(AAB_IIJJII, i, j),
                   x4 = g(ABB_IIIIJJ, i, j), x5 = g(AABB_IIJJIIJJ, i, j);
      const double ABij = g(AB_IIJJ, i, j), ABji = g(AB_IIJJ, j, i), ABii = g(AB_IIJJ, i, i), ABjj = g(AB_IIJJ, j, j);
      double e = 0., ev[4];
      e += xlogx(1 - a[i] - b[i] - a[j] - b[j] + d[i] + d[j] + AA + ABij + ABji + BB - x1 - x2 - x3 - x4 + x5);
      eig2(a[j] - ABji - AA - ABjj + x1 + x3 + x2 - x5,
           g(A_IJ, i, j) - g(AB_JIJJ, j, i) - g(AB_IJJJ, i, j) + g(ABB_IJIIJJ, i, j),
           a[i] - ABij - AA - ABii + x1 + x3 + x4 - x5, ev);
      e += xlogx(ev[0]) + xlogx(ev[1]);
      eig2(b[j] - ABij - BB - ABjj + x4 + x1 + x2 - x5,
           g(B_IJ, i, j) - g(AB_JJIJ, j, i) - g(AB_JJJI, i, j) + g(AAB_IIJJIJ, i, j),
           b[i] - ABji - BB - ABii + x2 + x3 + x4 - x5, ev);
      e += xlogx(ev[0]) + xlogx(ev[1]);
      e += xlogx(AA - x3 - x1 + x5);
      e += xlogx(BB - x4 - x2 + x5);
      double B4[4][4];
      B4[0][0] = ABjj - x1 - x2 + x5;
      B4[0][1] = B4[1][0] = g(AB_IJJJ, i, j) - g(ABB_IJIIJJ, i, j);
      B4[0][2] = B4[2][0] = -g(AB_JJIJ, i, j) + g(AAB_IIJJIJ, i, j);
      B4[0][3] = B4[3][0] = g(AB_IJIJ, i, j);
      B4[1][1] = ABij - x4 - x1 + x5;
      B4[1][2] = B4[2][1] = -g(AB_IJJI, j, i);
      B4[1][3] = B4[3][1] = g(AB_JJJI, j, i) - g(AAB_IIJJIJ, i, j);
      B4[2][2] = ABji - x3 - x2 + x5;
      B4[2][3] = B4[3][2] = -g(AB_JIJJ, j, i) + g(ABB_IJIIJJ, i, j);
      B4[3][3] = d[i] - x3 - x4 + x5;
      eig4(B4, ev);
      for (int k = 0; k < 4; ++k) e += xlogx(ev[k]);
      eig2(x1 - x5, -g(AAB_IIJJIJ, i, j), x3 - x5, ev);
      e += xlogx(ev[0]) + xlogx(ev[1]);
      eig2(x2 - x5, -g(ABB_IJIIJJ, i, j), x4 - x5, ev);
      e += xlogx(ev[0]) + xlogx(ev[1]);
      e += xlogx(x5);
      s2[i + size_t(j) * n] = s2[j + size_t(i) * n] = e;
    }
  }
  if (mi)
    for (int i = 0; i < n; ++i) {
      mi[i + size_t(i) * n] = 0.0;
      for (int j = i + 1; j < n; ++j)
        mi[i + size_t(j) * n] = mi[j + size_t(i) * n] = s1[i] + s1[j] - s2[i + size_t(j) * n];
    }
}

// accumulate the intermediates on the device, all-reduce over the ranks, copy to the host
void entropy_intermediates(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C_host, bool need_s2, double* out_host) {
  const int n = ctx->norb;
  if (!ctx->ints_dev) throw Error("b2ci_form_entropies: integrals not uploaded (they define the orbital count)");
  if (!dets || !C_host || !out_host) throw Error("b2ci_form_entropies: bad arguments");
  const int64_t N = dets->n;
  cudaStream_t st = ctx->stream;
  ctx->timers["entropy.pattern"] = ctx->timers["entropy.scatter"] = 0.;
  const size_t total = entropy_intermediate_doubles(n, need_s2);
  DevBuf<double> buf(total), dC(N > 0 ? N : 1);
  B2_CUDA(cudaMemsetAsync(buf, 0, total * 8, st));
  if (N > 0) B2_CUDA(cudaMemcpyAsync(dC, C_host, size_t(N) * 8, cudaMemcpyHostToDevice, st));
  int64_t r0 = 0, r1 = N;
  if (ctx->nranks > 1) {
    const int64_t base = N / ctx->nranks, rem = N % ctx->nranks;
    r0 = ctx->rank * base + std::min<int64_t>(ctx->rank, rem);
    r1 = r0 + base + (ctx->rank < rem ? 1 : 0);
  }
  EntOut E;
  E.n = n;
  E.v = buf;
  E.m = need_s2 ? buf.p + 3 * size_t(n) : nullptr;
  const int64_t nrows = r1 - r0;
  if (need_s2 && nrows > 0) {
    b2ci_csr H;
    {
      ScopedTimer t(ctx, "entropy.pattern");
      hbuild_csr(ctx, dets, r0, r1, 0.0, &H);
    }
    struct PatternGuard {
      b2ci_ctx* ctx;
      b2ci_csr* m;
      ~PatternGuard() {
        dev_free(m->rowptr);
        if (m->colind_cap) big_release(ctx, 0, m->colind, m->colind_cap); else dev_free(m->colind);
        if (m->nzval_cap) big_release(ctx, 1, m->nzval, m->nzval_cap); else dev_free(m->nzval);
      }
    } guard{ctx, &H};
    ScopedTimer t(ctx, "entropy.scatter");
    k_entropy_pairs<<<unsigned((nrows * 32 + 255) / 256), 256, 0, st>>>(E, dets->alpha, dets->beta, dC, r0, nrows,
                                                                        H.rowptr, H.colind);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    B2_CUDA(cudaStreamSynchronize(st));
  } else if (nrows > 0) {
    ScopedTimer t(ctx, "entropy.scatter");
    k_entropy_diag<<<unsigned((nrows * 32 + 255) / 256), 256, 0, st>>>(E, dets->alpha, dets->beta, dC, r0, nrows);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  }
  comm_allreduce_sum(ctx, buf, int64_t(total));
  B2_CUDA(cudaMemcpyAsync(out_host, buf, total * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
}

}  // namespace b2ci
