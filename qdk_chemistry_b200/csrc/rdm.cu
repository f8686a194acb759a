// Reduced density matrices of a CI vector on the device.
//
// Replaces SortedDoubleLoopHamiltonianGenerator::form_rdms / form_rdms_spin_dep
// (external/macis/include/macis/hamiltonian_generator/sorted_double_loop.hpp:512-760) with the
// contribution rules of external/macis/include/macis/util/rdms.hpp (symm = true: one visit per
// unordered pair, bra = the lower index). The pair enumeration is the H build's: the structural
// pattern (h_thresh = 0: every pair at distance <= 4, alpha-empty determinants skipped) of this
// rank's row block, of which the upper triangle is walked; every pair with |C_i C_j| > 1e-16
// scatters into the density matrices with fp64 reductions in L2 (RED.ADD.F64). Rows are sharded
// over the ranks and the matrices all-reduced. Summation order differs from the reference's
// (which is itself unordered: omp atomic), so parity is to rounding, not bit for bit.
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "slater.cuh"

namespace b2ci {
void hbuild_csr(b2ci_ctx* ctx, const b2ci_dets* dets, int64_t row_begin, int64_t row_end, double thr,
                b2ci_csr* out);

namespace {

struct RdmOut {
  double* o1;  // ordm   | ordm_aa
  double* o2;  //        | ordm_bb
  double* t1;  // trdm   | trdm_aaaa
  double* t2;  //        | trdm_bbbb
  double* t3;  //        | trdm_aabb
  int n;
};
__device__ __forceinline__ void add4(double* t, int n, unsigned p, unsigned q, unsigned r, unsigned s, double v) {
  atomicAdd(t + (size_t(p) + size_t(q) * n + (size_t(r) + size_t(s) * n) * n * n), v);
}
__device__ __forceinline__ void add2(double* m, int n, unsigned p, unsigned q, double v) {
  atomicAdd(m + (size_t(p) + size_t(q) * n), v);
}

// rdm_contributions_4<true> (rdms.hpp:36-65)
__device__ void rdm4(double* t, int n, uint64_t bra, uint64_t ket, uint64_t ex, double val) {
  if (!t) return;
  unsigned o1, v1, o2, v2;
  double sign;
  dx_sign_indices(bra, ket, ex, o1, v1, o2, v2, sign);
  val *= sign * 0.5;
  add4(t, n, v1, o1, v2, o2, val);
  add4(t, n, v2, o1, v1, o2, -val);
  add4(t, n, v1, o2, v2, o1, -val);
  add4(t, n, v2, o2, v1, o1, val);
  add4(t, n, o2, v2, o1, v1, val);
  add4(t, n, o2, v1, o1, v2, -val);
  add4(t, n, o1, v2, o2, v1, -val);
  add4(t, n, o1, v1, o2, v2, val);
}
// rdm_contributions_2<true> / rdm_contributions_2_spin_dep<true, transpose> (rdms.hpp:177-335):
// SPIN == 0 spin-traced; 1 alpha single (transpose = false); 2 beta single (transpose = true)
template <int SPIN>
__device__ void rdm2(const RdmOut& R, uint64_t bra, uint64_t ket, uint64_t ex, uint64_t occ_same,
                     uint64_t occ_othr, double val) {
  const int n = R.n;
  unsigned o1, v1;
  double sign;
  sx_sign_indices(bra, ket, ex, o1, v1, sign);
  double* om = SPIN == 2 ? R.o2 : R.o1;
  if (om) {
    add2(om, n, v1, o1, sign * val);
    add2(om, n, o1, v1, sign * val);
  }
  val *= sign * 0.5;
  double* tss = SPIN == 2 ? R.t2 : R.t1;
  if (tss) {
    for (uint64_t s = occ_same; s; s &= s - 1) {
      const unsigned p = lsb64(s);
      add4(tss, n, v1, o1, p, p, val);
      add4(tss, n, p, p, v1, o1, val);
      add4(tss, n, v1, p, p, o1, -val);
      add4(tss, n, p, o1, v1, p, -val);
      add4(tss, n, p, p, o1, v1, val);
      add4(tss, n, o1, v1, p, p, val);
      add4(tss, n, o1, p, p, v1, -val);
      add4(tss, n, p, v1, o1, p, -val);
    }
  }
  double* tos = SPIN == 0 ? R.t1 : R.t3;
  if (tos) {
    for (uint64_t s = occ_othr; s; s &= s - 1) {
      const unsigned p = lsb64(s);
      if (SPIN == 0) {
        add4(tos, n, v1, o1, p, p, val);
        add4(tos, n, p, p, v1, o1, val);
        add4(tos, n, o1, v1, p, p, val);
        add4(tos, n, p, p, o1, v1, val);
      } else if (SPIN == 2) {  // transpose
        add4(tos, n, v1, o1, p, p, val);
        add4(tos, n, o1, v1, p, p, val);
      } else {
        add4(tos, n, p, p, v1, o1, val);
        add4(tos, n, p, p, o1, v1, val);
      }
    }
  }
}
// rdm_contributions_diag / _diag_spin_dep (rdms.hpp:352-460), spread over the lanes of a warp
template <bool SPIN_DEP>
__device__ void rdm_diag(const RdmOut& R, uint64_t oa, uint64_t ob, double val, int lane) {
  const int n = R.n;
  if (lane == 0) {
    for (uint64_t s = oa; s; s &= s - 1) { const unsigned p = lsb64(s); if (R.o1) add2(R.o1, n, p, p, val); }
    double* ob_m = SPIN_DEP ? R.o2 : R.o1;
    for (uint64_t s = ob; s; s &= s - 1) { const unsigned p = lsb64(s); if (ob_m) add2(ob_m, n, p, p, val); }
  }
  val *= 0.5;
  int c = 0;  // pair counter: pair c is handled by lane c % 32
  double* taa = R.t1;
  double* tbb = SPIN_DEP ? R.t2 : R.t1;
  double* tab = SPIN_DEP ? R.t3 : R.t1;
  for (uint64_t sq = oa; sq; sq &= sq - 1)
    for (uint64_t sp = oa; sp; sp &= sp - 1, ++c) {
      if ((c & 31) != lane || !taa) continue;
      const unsigned p = lsb64(sp), q = lsb64(sq);
      add4(taa, n, p, p, q, q, val);
      add4(taa, n, p, q, q, p, -val);
    }
  for (uint64_t sq = ob; sq; sq &= sq - 1)
    for (uint64_t sp = ob; sp; sp &= sp - 1, ++c) {
      if ((c & 31) != lane || !tbb) continue;
      const unsigned p = lsb64(sp), q = lsb64(sq);
      add4(tbb, n, p, p, q, q, val);
      add4(tbb, n, p, q, q, p, -val);
    }
  for (uint64_t sq = ob; sq; sq &= sq - 1)
    for (uint64_t sp = oa; sp; sp &= sp - 1, ++c) {
      if ((c & 31) != lane || !tab) continue;
      const unsigned p = lsb64(sp), q = lsb64(sq);
      if (SPIN_DEP) {
        add4(tab, n, q, q, p, p, val);
      } else {
        add4(tab, n, p, p, q, q, val);
        add4(tab, n, q, q, p, p, val);
      }
    }
}

// one warp per row of the block; lanes stride over the row's upper-triangle entries
template <bool SPIN_DEP>
__global__ void __launch_bounds__(256)
k_rdm(RdmOut R, const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ beta,
      const double* __restrict__ C, int64_t row_begin, int64_t nrows,
      const int64_t* __restrict__ rowptr, const int32_t* __restrict__ colind) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const int64_t i = row_begin + row;
  const uint64_t ba = alpha[i], bb = beta[i];
  const double ci = C[i];
  const int n = R.n;
  const int64_t e0 = rowptr[row], e1 = rowptr[row + 1];
  for (int64_t e = e0 + lane; e < e1; e += 32) {
    const int64_t j = colind[e];
    if (j <= i) continue;  // the diagonal is handled by the whole warp below
    const double val = ci * C[j];
    if (!(fabs(val) > 1e-16)) continue;
    const uint64_t ka = alpha[j], kb = beta[j];
    const uint64_t exa = ba ^ ka, exb = bb ^ kb;
    const int ca = popc64(exa), cb = popc64(exb);
    if (ca == 4) {
      rdm4(R.t1, n, ba, ka, exa, val);
    } else if (cb == 4) {
      rdm4(SPIN_DEP ? R.t2 : R.t1, n, bb, kb, exb, val);
    } else if (ca == 2 && cb == 2) {
      unsigned oa_, va_, ob_, vb_;
      double sa, sb;
      sx_sign_indices(ba, ka, exa, oa_, va_, sa);
      sx_sign_indices(bb, kb, exb, ob_, vb_, sb);
      const double v = val * sa * sb * 0.5;
      if (SPIN_DEP) {  // rdm_contributions_22_spin_dep: (o1,v1) beta, (o2,v2) alpha
        if (R.t3) {
          add4(R.t3, n, vb_, ob_, va_, oa_, v);
          add4(R.t3, n, ob_, vb_, oa_, va_, v);
        }
      } else if (R.t1) {  // rdm_contributions_22: (o1,v1) alpha, (o2,v2) beta
        add4(R.t1, n, va_, oa_, vb_, ob_, v);
        add4(R.t1, n, vb_, ob_, va_, oa_, v);
        add4(R.t1, n, ob_, vb_, oa_, va_, v);
        add4(R.t1, n, oa_, va_, ob_, vb_, v);
      }
    } else if (ca == 2) {
      rdm2<SPIN_DEP ? 1 : 0>(R, ba, ka, exa, ba, bb, val);
    } else if (cb == 2) {
      rdm2<SPIN_DEP ? 2 : 0>(R, bb, kb, exb, bb, ba, val);
    }
  }
  // diagonal pair (i, i): present in every non-empty row of the structural pattern
  const double vd = ci * ci;
  if (e1 > e0 && fabs(vd) > 1e-16) rdm_diag<SPIN_DEP>(R, ba, bb, vd, lane);
}

}  // namespace

// host outputs (may be NULL), accumulated INTO like the reference does; C_host has dets->n entries
void form_rdms(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C_host, bool spin_dep, double* o1,
               double* o2, double* t1, double* t2, double* t3) {
  const int n = ctx->norb;
  if (!ctx->ints_dev) throw Error("b2ci_form_rdms: integrals not uploaded (they define the orbital count)");
  if (!dets || !C_host) throw Error("b2ci_form_rdms: bad arguments");
  const int64_t N = dets->n;
  cudaStream_t st = ctx->stream;
  ctx->timers["rdm.pattern"] = ctx->timers["rdm.scatter"] = 0.;
  const size_t n2 = size_t(n) * n, n4 = n2 * n2;
  double* host[5] = {o1, spin_dep ? o2 : nullptr, t1, spin_dep ? t2 : nullptr, spin_dep ? t3 : nullptr};
  const size_t len[5] = {n2, n2, n4, n4, n4};
  size_t off[5], total = 0;
  for (int k = 0; k < 5; ++k) { off[k] = total; if (host[k]) total += len[k]; }
  if (total == 0 || N == 0) return;
  DevBuf<double> buf(total), dC(N);
  B2_CUDA(cudaMemsetAsync(buf, 0, total * 8, st));
  B2_CUDA(cudaMemcpyAsync(dC, C_host, size_t(N) * 8, cudaMemcpyHostToDevice, st));
  // contiguous row blocks, remainder spread over the first ranks
  int64_t r0 = 0, r1 = N;
  if (ctx->nranks > 1) {
    const int64_t base = N / ctx->nranks, rem = N % ctx->nranks;
    r0 = ctx->rank * base + std::min<int64_t>(ctx->rank, rem);
    r1 = r0 + base + (ctx->rank < rem ? 1 : 0);
  }
  b2ci_csr H;
  {
    ScopedTimer t(ctx, "rdm.pattern");
    hbuild_csr(ctx, dets, r0, r1, 0.0, &H);
  }
  struct PatternGuard {  // the big arrays go back to the context's recycling slots
    b2ci_ctx* ctx;
    b2ci_csr* m;
    ~PatternGuard() {
      dev_free(m->rowptr);
      if (m->colind_cap) big_release(ctx, 0, m->colind, m->colind_cap); else dev_free(m->colind);
      if (m->nzval_cap) big_release(ctx, 1, m->nzval, m->nzval_cap); else dev_free(m->nzval);
    }
  } guard{ctx, &H};
  const int64_t* rp = H.rowptr;
  const int32_t* ci = H.colind;
  RdmOut R;
  R.n = n;
  R.o1 = host[0] ? buf.p + off[0] : nullptr;
  R.o2 = host[1] ? buf.p + off[1] : nullptr;
  R.t1 = host[2] ? buf.p + off[2] : nullptr;
  R.t2 = host[3] ? buf.p + off[3] : nullptr;
  R.t3 = host[4] ? buf.p + off[4] : nullptr;
  {
    ScopedTimer t(ctx, "rdm.scatter");
    const int64_t nrows = H.nrows;
    if (nrows > 0) {
      const unsigned grid = unsigned((nrows * 32 + 255) / 256);
      if (spin_dep) k_rdm<true><<<grid, 256, 0, st>>>(R, dets->alpha, dets->beta, dC, r0, nrows, rp, ci);
      else k_rdm<false><<<grid, 256, 0, st>>>(R, dets->alpha, dets->beta, dC, r0, nrows, rp, ci);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    comm_allreduce_sum(ctx, buf, int64_t(total));
  }
  std::vector<double> tmp(total);
  B2_CUDA(cudaMemcpyAsync(tmp.data(), buf, total * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 5; ++k)
    if (host[k])
      for (size_t q = 0; q < len[k]; ++q) host[k][q] += tmp[off[k] + q];
}

}  // namespace b2ci
