// Slater-Condon rules on bitstring determinants: __host__ __device__ so the very same
// code is exercised by CPU-only tests (b2ci_host_matrix_element) and by the kernels.
//
// Behaviour to match (bit for bit, hence the fixed summation order and --fmad=false):
//   signs / indices   external/macis/include/macis/sd_operations.hpp:40-50, 394-425
//   matrix elements   external/macis/include/macis/hamiltonian_generator/matrix_elements.hpp:65-230
//   fast diagonals    external/macis/src/macis/hamiltonian_generator/fast_diagonals.ipp:15-127
#pragma once
#include "common.cuh"

#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b2ci {

B2_HD int popc64(uint64_t x) {
#ifdef __CUDA_ARCH__
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}
B2_HD int lsb64(uint64_t x) {  // index of the lowest set bit, x != 0
#ifdef __CUDA_ARCH__
  return __ffsll((long long)x) - 1;
#else
  return __builtin_ctzll(x);
#endif
}
B2_HD double ldg(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
B2_HD uint64_t low_mask(int k) { return k >= 64 ? ~uint64_t(0) : ((uint64_t(1) << k) - 1); }

// The excitation arithmetic below is templated on the string word: 32-bit words when norb <= 32 (half the
// integer work of the general-list fill), 64-bit otherwise. The floating-point part is the same either way.
template <typename Wd> B2_HD int popc_w(Wd x) { return sizeof(Wd) == 4 ?
#ifdef __CUDA_ARCH__
  __popc(uint32_t(x))
#else
  __builtin_popcount(uint32_t(x))
#endif
  : popc64(uint64_t(x)); }
template <typename Wd> B2_HD int lsb_w(Wd x) { return sizeof(Wd) == 4 ?
#ifdef __CUDA_ARCH__
  __ffs(int(uint32_t(x))) - 1
#else
  __builtin_ctz(uint32_t(x))
#endif
  : lsb64(uint64_t(x)); }
template <typename Wd> B2_HD Wd low_mask_w(int k) {
  return k >= int(sizeof(Wd) * 8) ? Wd(~Wd(0)) : Wd((Wd(1) << k) - 1);
}

// (-1)^(number of occupied orbitals strictly between p and q)
template <typename Wd> B2_HD double sx_sign(Wd state, unsigned p, unsigned q) {
  const unsigned lo = p < q ? p : q, hi = p < q ? q : p;
  const Wd mask = state & (low_mask_w<Wd>(hi) ^ low_mask_w<Wd>(lo + 1));
  return (popc_w<Wd>(mask) & 1) ? -1. : 1.;
}
template <typename Wd> B2_HD void sx_sign_indices(Wd bra, Wd ket, Wd ex, unsigned& o1, unsigned& v1, double& sign) {
  o1 = lsb_w<Wd>(ket & ex);
  v1 = lsb_w<Wd>(bra & ex);
  sign = sx_sign<Wd>(ket, v1, o1);
}
template <typename Wd> B2_HD void dx_sign_indices(Wd bra, Wd ket, Wd ex, unsigned& o1, unsigned& v1,
                                                  unsigned& o2, unsigned& v2, double& sign) {
  double s1, s2;
  sx_sign_indices<Wd>(bra, ket, ex, o1, v1, s1);
  const Wd flip = Wd(Wd(1) << o1) | Wd(Wd(1) << v1);
  ket ^= flip;
  ex ^= flip;
  sx_sign_indices<Wd>(bra, ket, ex, o2, v2, s2);
  sign = s1 * s2;
}

// same-spin double: sign * (V(v1,o1,v2,o2) - V(v1,o2,v2,o1))
template <typename Wd> B2_HD double me4(const IntsView& I, Wd bra, Wd ket, Wd ex) {
  unsigned o1, v1, o2, v2;
  double sign;
  dx_sign_indices<Wd>(bra, ket, ex, o1, v1, o2, v2, sign);
  const size_t n = I.n, n2 = n * n, n3 = n2 * n;
  const double g = ldg(I.V + v1 + o1 * n + v2 * n2 + o2 * n3) -
                   ldg(I.V + v1 + o2 * n + v2 * n2 + o1 * n3);
  return sign * g;
}
// opposite-spin double: sign_a * sign_b * V(v1,o1,v2,o2)
template <typename Wd> B2_HD double me22(const IntsView& I, Wd bra_a, Wd ket_a, Wd ex_a, Wd bra_b, Wd ket_b, Wd ex_b) {
  unsigned o1, v1, o2, v2;
  double sa, sb;
  sx_sign_indices<Wd>(bra_a, ket_a, ex_a, o1, v1, sa);
  sx_sign_indices<Wd>(bra_b, ket_b, ex_b, o2, v2, sb);
  const size_t n = I.n, n2 = n * n, n3 = n2 * n;
  const double sign = sa * sb;
  return sign * ldg(I.V + v1 + o1 * n + v2 * n2 + o2 * n3);
}
// single: sign * (T(v,o) + sum_{p in occ_same, ascending} G_red(p,v,o)
//                          + sum_{p in occ_other, ascending} V_red(p,v,o))
template <typename Wd> B2_HD double me2(const IntsView& I, Wd bra, Wd ket, Wd ex, Wd occ_same, Wd occ_othr) {
  unsigned o1, v1;
  double sign;
  sx_sign_indices<Wd>(bra, ket, ex, o1, v1, sign);
  const size_t n = I.n, n2 = n * n;
  double h_el = ldg(I.T + v1 + o1 * n);
  const double* G = I.G + v1 * n + o1 * n2;
  for (Wd s = occ_same; s; s &= Wd(s - 1)) h_el += ldg(G + lsb_w<Wd>(s));
  const double* Vr = I.Vr + v1 * n + o1 * n2;
  for (Wd s = occ_othr; s; s &= Wd(s - 1)) h_el += ldg(Vr + lsb_w<Wd>(s));
  return sign * h_el;
}
B2_HD double me_diag(const IntsView& I, uint64_t occ_a, uint64_t occ_b) {
  const size_t n = I.n;
  double e = 0.;
  for (uint64_t s = occ_a; s; s &= s - 1) { const size_t p = lsb64(s); e += ldg(I.T + p + p * n); }
  for (uint64_t s = occ_b; s; s &= s - 1) { const size_t p = lsb64(s); e += ldg(I.T + p + p * n); }
  for (uint64_t sq = occ_a; sq; sq &= sq - 1) {
    const size_t q = lsb64(sq);
    for (uint64_t sp = occ_a; sp; sp &= sp - 1) e += ldg(I.G2 + lsb64(sp) + q * n);
  }
  for (uint64_t sq = occ_b; sq; sq &= sq - 1) {
    const size_t q = lsb64(sq);
    for (uint64_t sp = occ_b; sp; sp &= sp - 1) e += ldg(I.G2 + lsb64(sp) + q * n);
  }
  for (uint64_t sq = occ_b; sq; sq &= sq - 1) {
    const size_t q = lsb64(sq);
    for (uint64_t sp = occ_a; sp; sp &= sp - 1) e += ldg(I.V2 + lsb64(sp) + q * n);
  }
  return e;
}
// dispatcher on (popcount ex_alpha, popcount ex_beta); caller guarantees total <= 4
B2_HD double matel(const IntsView& I, uint64_t bra_a, uint64_t bra_b, uint64_t ket_a,
                   uint64_t ket_b) {
  const uint64_t ex_a = bra_a ^ ket_a, ex_b = bra_b ^ ket_b;
  const int ca = popc64(ex_a), cb = popc64(ex_b);
  if (ca + cb > 4) return 0.;
  if (ca == 4) return me4(I, bra_a, ket_a, ex_a);
  if (cb == 4) return me4(I, bra_b, ket_b, ex_b);
  if (ca == 2 && cb == 2) return me22(I, bra_a, ket_a, ex_a, bra_b, ket_b, ex_b);
  if (ca == 2) return me2(I, bra_a, ket_a, ex_a, bra_a, bra_b);
  if (cb == 2) return me2(I, bra_b, ket_b, ex_b, bra_b, bra_a);
  return me_diag(I, bra_a, bra_b);
}

// single_orbital_ens entry i (fast_diagonals.ipp:29-49)
B2_HD double orbital_energy(const IntsView& I, unsigned i, uint64_t occ_same, uint64_t occ_othr) {
  const size_t n = I.n;
  double e = ldg(I.T + i + i * n);
  for (uint64_t s = occ_same; s; s &= s - 1) {
    const size_t q = lsb64(s);
    e += ldg(I.G2 + i + q * n) + ldg(I.G2 + q + i * n);
  }
  e -= ldg(I.G2 + i + i * n);
  for (uint64_t s = occ_othr; s; s &= s - 1) e += ldg(I.V2 + i + lsb64(s) * n);
  return e;
}

}  // namespace b2ci
