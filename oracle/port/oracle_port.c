/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT. See oracle_port.h.
 *
 * Every function names the reference file:line it restates (paths relative to
 * /root/reference/external/macis). Floating-point expressions keep the
 * reference's evaluation order (compile with -ffp-contract=off) so that
 * threshold decisions (|h| > H_thresh, |c*h| < h_el_tol) fall the same way.
 */
#include "oracle_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct op_ham {
  int n;
  double *T, *V, *G_red, *V_red, *G2_red, *V2_red;
};

static inline int popc(uint64_t x) { return __builtin_popcountll(x); }
static inline int lsb(uint64_t x) { return __builtin_ctzll(x); }
static inline uint64_t low_mask(int k) { return k >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << k) - 1); }

int op_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ integrals
 * src/macis/hamiltonian_generator/base.ipp:37-77 */
op_ham* op_ham_create(int n, const double* T, const double* V) {
  op_ham* h = (op_ham*)calloc(1, sizeof(op_ham));
  size_t n2 = (size_t)n * n, n3 = n2 * n, n4 = n2 * n2;
  h->n = n;
  h->T = (double*)malloc(n2 * 8);
  h->V = (double*)malloc(n4 * 8);
  h->G_red = (double*)malloc(n3 * 8);
  h->V_red = (double*)malloc(n3 * 8);
  h->G2_red = (double*)malloc(n2 * 8);
  h->V2_red = (double*)malloc(n2 * 8);
  memcpy(h->T, T, n2 * 8);
  memcpy(h->V, V, n4 * 8);
#define VV(p, q, r, s) h->V[(size_t)(p) + (size_t)(q) * n + (size_t)(r) * n2 + (size_t)(s) * n3]
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n; ++k) {
        h->G_red[k + (size_t)i * n + (size_t)j * n2] = VV(k, k, i, j) - VV(k, j, i, k);
        h->V_red[k + (size_t)i * n + (size_t)j * n2] = VV(k, k, i, j);
      }
  for (int j = 0; j < n; ++j)
    for (int i = 0; i < n; ++i) {
      h->G2_red[i + (size_t)j * n] = 0.5 * (VV(i, i, j, j) - VV(i, j, j, i));
      h->V2_red[i + (size_t)j * n] = VV(i, i, j, j);
    }
#undef VV
  return h;
}
void op_ham_destroy(op_ham* h) {
  if (!h) return;
  free(h->T); free(h->V); free(h->G_red); free(h->V_red); free(h->G2_red); free(h->V2_red);
  free(h);
}
void op_ham_intermediates(const op_ham* h, double* G, double* Vr, double* G2, double* V2) {
  size_t n = h->n;
  if (G) memcpy(G, h->G_red, n * n * n * 8);
  if (Vr) memcpy(Vr, h->V_red, n * n * n * 8);
  if (G2) memcpy(G2, h->G2_red, n * n * 8);
  if (V2) memcpy(V2, h->V2_red, n * n * 8);
}

/* ------------------------------------------------------------ Hilbert space
 * include/macis/sd_operations.hpp:305-351: combinations in std::prev_permutation
 * order of a bool vector whose first nset entries start true; alpha-outer, beta-inner */
static int prev_permutation(unsigned char* v, int n) {
  /* std::prev_permutation on a 0/1 sequence */
  int i = n - 1;
  if (n < 2) return 0;
  for (;;) {
    int ip1 = i;
    --i;
    if (v[ip1] < v[i]) {
      int j = n - 1;
      while (!(v[j] < v[i])) --j;
      unsigned char t = v[i]; v[i] = v[j]; v[j] = t;
      for (int a = ip1, b = n - 1; a < b; ++a, --b) { t = v[a]; v[a] = v[b]; v[b] = t; }
      return 1;
    }
    if (i == 0) {
      for (int a = 0, b = n - 1; a < b; ++a, --b) { unsigned char t = v[a]; v[a] = v[b]; v[b] = t; }
      return 0;
    }
  }
}
static int64_t gen_combs(int nbits, int nset, uint64_t* out) {
  unsigned char v[64];
  int64_t cnt = 0;
  for (int i = 0; i < nbits; ++i) v[i] = i < nset;
  do {
    uint64_t s = 0;
    for (int i = 0; i < nbits; ++i)
      if (v[i]) s |= (uint64_t)1 << i;
    if (out) out[cnt] = s;
    ++cnt;
  } while (prev_permutation(v, nbits));
  return cnt;
}
int64_t op_generate_hilbert_space(int norb, int na, int nb, uint64_t* alpha, uint64_t* beta) {
  int64_t ca = gen_combs(norb, na, NULL), cb = gen_combs(norb, nb, NULL);
  if (!alpha || !beta) return ca * cb;
  uint64_t* A = (uint64_t*)malloc(ca * 8);
  uint64_t* B = (uint64_t*)malloc(cb * 8);
  gen_combs(norb, na, A);
  gen_combs(norb, nb, B);
  for (int64_t i = 0; i < ca; ++i)
    for (int64_t j = 0; j < cb; ++j) {
      alpha[i * cb + j] = A[i];
      beta[i * cb + j] = B[j];
    }
  free(A); free(B);
  return ca * cb;
}

/* ------------------------------------------------------------ signs / indices
 * include/macis/sd_operations.hpp:40-50 (single_excitation_sign),
 * :394-402 (single_excitation_sign_indices), :413-425 (doubles_sign_indices) */
static inline double sx_sign(uint64_t state, unsigned p, unsigned q) {
  uint64_t mask;
  if (p > q) mask = state & (low_mask(p) ^ low_mask(q + 1));
  else mask = state & (low_mask(q) ^ low_mask(p + 1));
  return (popc(mask) & 1) ? -1. : 1.;
}
static inline void sx_sign_indices(uint64_t bra, uint64_t ket, uint64_t ex, unsigned* o1,
                                   unsigned* v1, double* sign) {
  *o1 = lsb(ket & ex);
  *v1 = lsb(bra & ex);
  *sign = sx_sign(ket, *v1, *o1);
}
static inline void dx_sign_indices(uint64_t bra, uint64_t ket, uint64_t ex, unsigned* o1,
                                   unsigned* v1, unsigned* o2, unsigned* v2, double* sign) {
  double s1, s2;
  sx_sign_indices(bra, ket, ex, o1, v1, &s1);
  uint64_t flip = ((uint64_t)1 << *o1) | ((uint64_t)1 << *v1);
  ket ^= flip;
  ex ^= flip;
  sx_sign_indices(bra, ket, ex, o2, v2, &s2);
  *sign = s1 * s2;
}
static inline int occ_list(uint64_t s, unsigned* occ) {
  int c = 0;
  while (s) { occ[c++] = lsb(s); s &= s - 1; }
  return c;
}

/* ------------------------------------------------------------ matrix elements
 * include/macis/hamiltonian_generator/matrix_elements.hpp:113-230 */
static inline double me4(const op_ham* h, uint64_t bra, uint64_t ket, uint64_t ex) {
  unsigned o1, v1, o2, v2; double sign;
  size_t n = h->n, n2 = n * n, n3 = n2 * n;
  dx_sign_indices(bra, ket, ex, &o1, &v1, &o2, &v2, &sign);
  double g = h->V[v1 + o1 * n + v2 * n2 + o2 * n3] - h->V[v1 + o2 * n + v2 * n2 + o1 * n3];
  return sign * g;
}
static inline double me22(const op_ham* h, uint64_t bra_a, uint64_t ket_a, uint64_t ex_a,
                          uint64_t bra_b, uint64_t ket_b, uint64_t ex_b) {
  unsigned o1, v1, o2, v2; double sa, sb;
  size_t n = h->n, n2 = n * n, n3 = n2 * n;
  sx_sign_indices(bra_a, ket_a, ex_a, &o1, &v1, &sa);
  sx_sign_indices(bra_b, ket_b, ex_b, &o2, &v2, &sb);
  double sign = sa * sb;
  return sign * h->V[v1 + o1 * n + v2 * n2 + o2 * n3];
}
static inline double me2(const op_ham* h, uint64_t bra, uint64_t ket, uint64_t ex,
                         const unsigned* occ_same, int n_same, const unsigned* occ_othr,
                         int n_othr) {
  unsigned o1, v1; double sign;
  size_t n = h->n, n2 = n * n;
  sx_sign_indices(bra, ket, ex, &o1, &v1, &sign);
  double h_el = h->T[v1 + o1 * n];
  const double* G = h->G_red + v1 * n + o1 * n2;
  for (int k = 0; k < n_same; ++k) h_el += G[occ_same[k]];
  const double* Vr = h->V_red + v1 * n + o1 * n2;
  for (int k = 0; k < n_othr; ++k) h_el += Vr[occ_othr[k]];
  return sign * h_el;
}
static inline double me_diag(const op_ham* h, const unsigned* oa, int na, const unsigned* ob,
                             int nb) {
  size_t n = h->n;
  double e = 0;
  for (int k = 0; k < na; ++k) e += h->T[oa[k] + oa[k] * n];
  for (int k = 0; k < nb; ++k) e += h->T[ob[k] + ob[k] * n];
  for (int q = 0; q < na; ++q)
    for (int p = 0; p < na; ++p) e += h->G2_red[oa[p] + oa[q] * n];
  for (int q = 0; q < nb; ++q)
    for (int p = 0; p < nb; ++p) e += h->G2_red[ob[p] + ob[q] * n];
  for (int q = 0; q < nb; ++q)
    for (int p = 0; p < na; ++p) e += h->V2_red[oa[p] + ob[q] * n];
  return e;
}
/* dispatcher, matrix_elements.hpp:65-97 */
static double matel(const op_ham* h, uint64_t bra_a, uint64_t bra_b, uint64_t ket_a,
                    uint64_t ket_b) {
  uint64_t ex_a = bra_a ^ ket_a, ex_b = bra_b ^ ket_b;
  int ca = popc(ex_a), cb = popc(ex_b);
  unsigned oa[64], ob[64];
  if (ca + cb > 4) return 0.;
  if (ca == 4) return me4(h, bra_a, ket_a, ex_a);
  if (cb == 4) return me4(h, bra_b, ket_b, ex_b);
  if (ca == 2 && cb == 2) return me22(h, bra_a, ket_a, ex_a, bra_b, ket_b, ex_b);
  int na = occ_list(bra_a, oa), nb = occ_list(bra_b, ob);
  if (ca == 2) return me2(h, bra_a, ket_a, ex_a, oa, na, ob, nb);
  if (cb == 2) return me2(h, bra_b, ket_b, ex_b, ob, nb, oa, na);
  return me_diag(h, oa, na, ob, nb);
}
double op_matrix_element(const op_ham* h, uint64_t bra_a, uint64_t bra_b, uint64_t ket_a,
                         uint64_t ket_b) {
  return matel(h, bra_a, bra_b, ket_a, ket_b);
}

/* -------------------------------------------------------------------- H build
 * include/macis/hamiltonian_generator/sorted_double_loop.hpp:86-451, symmetric case.
 * The reference evaluates the upper triangle with bra = lower index and mirrors it;
 * pre-filters same-spin doubles with |h| < thr (:169-173,194-198); finally keeps
 * |h| > thr when thr > 0 (:421-424, csr_matrix.hpp:317-370). Determinants whose alpha
 * string is empty are skipped as bra and as ket (:147,156). Here each row is produced
 * independently with the same (bra, ket) roles, so columns come out ascending. */
int64_t op_hbuild_rows(const op_ham* h, const uint64_t* alpha, const uint64_t* beta,
                       int64_t n, int64_t r0, int64_t r1, double thresh, int64_t* rowptr,
                       int64_t* colind, double* nzval) {
  return op_hbuild_rows_gen(h, alpha, beta, n, r0, r1, thresh, 0, rowptr, colind, nzval);
}
/* pair_rule != 0: the CSR the pair-based generators produce (residue_arrays.hpp,
 * dynamic_bit_masking.hpp -> build_csr_from_pairs, connection_build_utils.hpp:125-249): same
 * connected pairs and matrix elements, but the diagonal is always stored (:181-192), an
 * off-diagonal element is dropped only when |h| < thresh (:166,199) and determinants with an
 * empty alpha string are not skipped. */
int64_t op_hbuild_rows_gen(const op_ham* h, const uint64_t* alpha, const uint64_t* beta,
                           int64_t n, int64_t r0, int64_t r1, double thresh, int pair_rule,
                           int64_t* rowptr, int64_t* colind, double* nzval) {
  /* run-length encode alpha strings (sd_operations.hpp:449-468) */
  int64_t nrun = 0;
  int64_t* run_st = (int64_t*)malloc((n + 1) * 8);
  for (int64_t i = 0; i < n; ++i)
    if (i == 0 || alpha[i] != alpha[i - 1]) run_st[nrun++] = i;
  run_st[nrun] = n;
  const int fill = colind != NULL;
  const int64_t nrows = r1 - r0;
  int64_t* cnt = (int64_t*)calloc(nrows + 1, 8);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t i = r0; i < r1; ++i) {
    const uint64_t ai = alpha[i], bi = beta[i];
    int64_t c = 0;
    int64_t w = fill ? rowptr[i - r0] : 0;
    if (ai || pair_rule) {
      for (int64_t r = 0; r < nrun; ++r) {
        const uint64_t ar = alpha[run_st[r]];
        if (!ar && !pair_rule) continue;
        const int ca = popc(ai ^ ar);
        if (ca > 4) continue;
        for (int64_t j = run_st[r]; j < run_st[r + 1]; ++j) {
          const int cb = popc(bi ^ beta[j]);
          if (ca + cb > 4) continue;
          double v;
          if (i <= j) v = matel(h, ai, bi, alpha[j], beta[j]);
          else v = matel(h, alpha[j], beta[j], ai, bi);
          if (pair_rule) {
            if (i != j && fabs(v) < thresh) continue;
          } else if (thresh > 0.0) {
            if (!(fabs(v) > thresh)) continue;
          }
          if (fill) { colind[w] = j; nzval[w] = v; ++w; }
          ++c;
        }
      }
    }
    cnt[i - r0] = c;
  }
  if (!fill) {
    rowptr[0] = 0;
    for (int64_t i = 0; i < nrows; ++i) rowptr[i + 1] = rowptr[i] + cnt[i];
  }
  int64_t nnz = rowptr[nrows];
  free(cnt); free(run_st);
  return nnz;
}
int64_t op_hbuild(const op_ham* h, const uint64_t* alpha, const uint64_t* beta, int64_t n,
                  double thresh, int64_t* rowptr, int64_t* colind, double* nzval) {
  return op_hbuild_rows(h, alpha, beta, n, 0, n, thresh, rowptr, colind, nzval);
}

/* ----------------------------------------------------------------------- sigma
 * src/sparsexx/include/sparsexx/spblas/spmbv.hpp:49-85 (K = 1, alpha = 1, beta = 0) */
void op_spmv(int64_t n, const int64_t* rowptr, const int64_t* colind, const double* nzval,
             const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double av = 0.;
    for (int64_t j = rowptr[i]; j < rowptr[i + 1]; ++j) av += nzval[j] * x[colind[j]];
    y[i] = 1. * av + 0. * 0.;
  }
}
/* src/sparsexx/include/sparsexx/util/submatrix.hpp:354-383 */
void op_extract_diagonal(int64_t n, const int64_t* rowptr, const int64_t* colind,
                         const double* nzval, double* D) {
  for (int64_t i = 0; i < n; ++i) {
    D[i] = 0.;
    for (int64_t j = rowptr[i]; j < rowptr[i + 1]; ++j)
      if (colind[j] == i) { D[i] = nzval[j]; break; }
  }
}

/* ------------------------------------------------- small symmetric eigensolver
 * stands in for lapack::syev(Vec, Lower) (src/lobpcgxx/include/lobpcgxx/
 * rayleigh_ritz.hpp:75; blaspp/lapackpp are not vendored by the reference). Cyclic
 * Jacobi: eigenvalues ascending in W, eigenvectors in the columns of A. */
int op_syev_lower(int n, double* A, int lda, double* W) {
  double* S = (double*)malloc((size_t)n * n * 8);
  double* Q = (double*)calloc((size_t)n * n, 8);
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) S[i + (size_t)j * n] = S[j + (size_t)i * n] = A[i + (size_t)j * lda];
  for (int i = 0; i < n; ++i) Q[i + (size_t)i * n] = 1.0;
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0., dia = 0.;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        double v = S[i + (size_t)j * n];
        if (i == j) dia += v * v; else off += v * v;
      }
    if (off <= 1e-34 * (dia + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        double apq = S[p + (size_t)q * n];
        if (apq == 0.) continue;
        double app = S[p + (size_t)p * n], aqq = S[q + (size_t)q * n];
        double theta = (aqq - app) / (2. * apq);
        double t = (theta >= 0 ? 1. : -1.) / (fabs(theta) + sqrt(theta * theta + 1.));
        double c = 1. / sqrt(t * t + 1.), s = t * c;
        for (int k = 0; k < n; ++k) {
          double skp = S[k + (size_t)p * n], skq = S[k + (size_t)q * n];
          S[k + (size_t)p * n] = c * skp - s * skq;
          S[k + (size_t)q * n] = s * skp + c * skq;
        }
        for (int k = 0; k < n; ++k) {
          double spk = S[p + (size_t)k * n], sqk = S[q + (size_t)k * n];
          S[p + (size_t)k * n] = c * spk - s * sqk;
          S[q + (size_t)k * n] = s * spk + c * sqk;
        }
        for (int k = 0; k < n; ++k) {
          double qkp = Q[k + (size_t)p * n], qkq = Q[k + (size_t)q * n];
          Q[k + (size_t)p * n] = c * qkp - s * qkq;
          Q[k + (size_t)q * n] = s * qkp + c * qkq;
        }
      }
  }
  /* sort ascending */
  int* idx = (int*)malloc(n * sizeof(int));
  for (int i = 0; i < n; ++i) idx[i] = i;
  for (int i = 1; i < n; ++i) {
    int t = idx[i], j = i - 1;
    while (j >= 0 && S[idx[j] + (size_t)idx[j] * n] > S[t + (size_t)t * n]) { idx[j + 1] = idx[j]; --j; }
    idx[j + 1] = t;
  }
  for (int j = 0; j < n; ++j) {
    W[j] = S[idx[j] + (size_t)idx[j] * n];
    for (int i = 0; i < n; ++i) A[i + (size_t)j * lda] = Q[i + (size_t)idx[j] * n];
  }
  free(S); free(Q); free(idx);
  return 0;
}

/* -------------------------------------------------------------------- Davidson
 * include/macis/solvers/davidson.hpp:185-238 (gram_schmidt, CGS2 + canonical-basis
 * fallback) and :259-372 (davidson: single root, no restart, diagonal preconditioner
 * with the 1e-12 denominator clamp). BLAS calls are restated as ordered loops. */
static double dotp(int64_t n, const double* a, const double* b) {
  double s = 0.;
  for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}
static void proj_out(int64_t N, int64_t K, const double* V, int64_t LDV, double* w,
                     double* inner) {
#pragma omp parallel for schedule(static)
  for (int64_t k = 0; k < K; ++k) inner[k] = dotp(N, V + k * LDV, w);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < N; ++i) {
    double s = 0.;
    for (int64_t k = 0; k < K; ++k) s += V[i + k * LDV] * inner[k];
    w[i] = w[i] - s;
  }
}
static int gram_schmidt(int64_t N, int64_t K, const double* V, int64_t LDV, double* w) {
  const double min_norm = 1e-12;
  if (K <= 0) {
    double nrm = sqrt(dotp(N, w, w));
    if (nrm > min_norm) for (int64_t i = 0; i < N; ++i) w[i] *= 1. / nrm;
    return 0;
  }
  double* inner = (double*)malloc(K * 8);
  proj_out(N, K, V, LDV, w, inner);
  proj_out(N, K, V, LDV, w, inner);
  double nrm = sqrt(dotp(N, w, w));
  int rc = 0;
  if (nrm > min_norm) {
    double inv = 1. / nrm;
    for (int64_t i = 0; i < N; ++i) w[i] *= inv;
  } else {
    rc = 2;
    for (int64_t idx = 0; idx < N; ++idx) {
      memset(w, 0, N * 8);
      w[idx] = 1.0;
      proj_out(N, K, V, LDV, w, inner);
      nrm = sqrt(dotp(N, w, w));
      if (nrm > min_norm) {
        double inv = 1. / nrm;
        for (int64_t i = 0; i < N; ++i) w[i] *= inv;
        rc = 0;
        break;
      }
    }
  }
  free(inner);
  return rc;
}

int op_davidson(int64_t N, const int64_t* rowptr, const int64_t* colind, const double* nzval,
                int64_t max_m, double tol, double* X, int64_t* niter, double* eigval,
                double* trace) {
  const double min_abs_denominator = 1e-12;
  if (max_m > N) max_m = N;
  if (N == 1) {
    double AX = 0.;
    X[0] = 1.0;
    op_spmv(1, rowptr, colind, nzval, X, &AX);
    *niter = 0; *eigval = AX;
    return 0;
  }
  double* D = (double*)malloc(N * 8);
  op_extract_diagonal(N, rowptr, colind, nzval, D);
  double* V = (double*)calloc((size_t)N * (max_m + 1), 8);
  double* AV = (double*)calloc((size_t)N * (max_m + 1), 8);
  double* Cm = (double*)calloc((size_t)(max_m + 1) * (max_m + 1), 8);
  double* LAM = (double*)calloc(max_m + 1, 8);
  memcpy(V, X, N * 8);
  op_spmv(N, rowptr, colind, nzval, V, AV);
  memcpy(V + N, AV, N * 8);
  int rc = gram_schmidt(N, 1, V, N, V + N);
  int converged = 0;
  int64_t iter = 1;
  for (int64_t i = 1; i < max_m && rc == 0; ++i, ++iter) {
    const int64_t k = i + 1;
    op_spmv(N, rowptr, colind, nzval, V + i * N, AV + i * N);
    /* rayleigh_ritz: C = V^T AV (full k x k), syev lower */
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < k; ++b)
      for (int64_t a = 0; a < k; ++a) Cm[a + b * k] = dotp(N, V + a * N, AV + b * N);
    op_syev_lower((int)k, Cm, (int)k, LAM);
    double* R = V + (i + 1) * N;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; ++r) {
      double x = 0., ax = 0.;
      for (int64_t a = 0; a < k; ++a) x += V[r + a * N] * Cm[a];
      for (int64_t a = 0; a < k; ++a) ax += AV[r + a * N] * Cm[a];
      X[r] = x;
      R[r] = ax + (-LAM[0]) * x;
    }
    double res_nrm = sqrt(dotp(N, R, R));
    if (trace) { trace[2 * (i - 1)] = LAM[0]; trace[2 * (i - 1) + 1] = res_nrm; }
    if (res_nrm < tol) { converged = 1; break; }
    for (int64_t j = 0; j < N; ++j) {
      double denom = D[j] - LAM[0];
      if (fabs(denom) < min_abs_denominator)
        denom = (denom >= 0) ? min_abs_denominator : -min_abs_denominator;
      R[j] = -R[j] / denom;
    }
    rc = gram_schmidt(N, k, V, N, R);
  }
  *niter = iter;
  *eigval = LAM[0];
  free(D); free(V); free(AV); free(Cm); free(LAM);
  if (rc) return rc;
  return converged ? 0 : 1;
}

/* include/macis/solvers/selected_ci_diag.hpp:111-158 guess policy */
int op_selected_ci_diag(int64_t n, const int64_t* rowptr, const int64_t* colind,
                        const double* nzval, int64_t max_m, double tol, double* C,
                        int64_t* niter, double* eigval) {
  double max_c = 0.;
  for (int64_t i = 0; i < n; ++i) if (fabs(C[i]) > max_c) max_c = fabs(C[i]);
  if (!(max_c > 1. / (double)n)) {
    /* diagonal_guess (davidson.hpp:106-113): X[argmin D] = 1 on top of the passed vector */
    double* D = (double*)malloc(n * 8);
    op_extract_diagonal(n, rowptr, colind, nzval, D);
    int64_t mi = 0;
    for (int64_t i = 1; i < n; ++i) if (D[i] < D[mi]) mi = i;
    C[mi] = 1.;
    free(D);
  }
  return op_davidson(n, rowptr, colind, nzval, max_m, tol, C, niter, eigval, NULL);
}

/* ----------------------------------------------------------------- ASCI search
 * Contribution formulas: include/macis/asci/determinant_contributions.hpp:93-297 (equal,
 * term by term, to the constraint-filtered variants actually called by the search,
 * asci/mask_constraints.hpp:400-670); fast diagonals: src/macis/hamiltonian_generator/
 * fast_diagonals.ipp:15-127; search/top-k: asci/determinant_search.hpp:349-771, 966-1114.
 *
 * Canonical order: contributions to one determinant are summed in ascending parent
 * (core determinant) index and its h_diag is the lowest parent's. The reference sums in
 * whatever order its unstable std::sort leaves duplicates (determinant_sort.hpp:115-136),
 * so the two can differ in the last bits of a score; the selection differs only if the
 * top-k cut falls inside that rounding (stats[2], stats[3] report the gap). */
typedef struct { uint64_t a, b; double cm, hd; int64_t seq; } contrib_t;
typedef struct { contrib_t* p; int64_t n, cap; } cvec;
static void cv_push(cvec* v, uint64_t a, uint64_t b, double cm, double hd) {
  if (v->n == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 1024;
    v->p = (contrib_t*)realloc(v->p, v->cap * sizeof(contrib_t));
  }
  contrib_t* c = &v->p[v->n];
  c->a = a; c->b = b; c->cm = cm; c->hd = hd; c->seq = v->n;
  v->n++;
}
static int cmp_contrib(const void* x, const void* y) {
  const contrib_t *p = (const contrib_t*)x, *q = (const contrib_t*)y;
  if (p->b != q->b) return p->b < q->b ? -1 : 1; /* bitset_less: beta is the high half */
  if (p->a != q->a) return p->a < q->a ? -1 : 1;
  return p->seq < q->seq ? -1 : (p->seq > q->seq);
}
static void orbital_ens(const op_ham* h, const unsigned* ss, int nss, const unsigned* os,
                        int nos, double* ens) {
  size_t n = h->n;
  for (size_t i = 0; i < n; ++i) {
    double e = h->T[i + i * n];
    for (int q = 0; q < nss; ++q) e += h->G2_red[i + ss[q] * n] + h->G2_red[ss[q] + i * n];
    e -= h->G2_red[i + i * n];
    for (int q = 0; q < nos; ++q) e += h->V2_red[i + os[q] * n];
    ens[i] = e;
  }
}
#define G2(p, q) h->G2_red[(p) + (size_t)(q) * n]
#define V2(p, q) h->V2_red[(p) + (size_t)(q) * n]
static void emit_singles(const op_ham* h, cvec* out, double coeff, uint64_t same,
                         uint64_t othr, int same_is_alpha, const unsigned* occ, int nocc,
                         const unsigned* vir, int nvir, const unsigned* occ_o, int nocc_o,
                         const double* eps, double tol, double root, double E0) {
  size_t n = h->n, n2 = n * n;
  for (int ii = 0; ii < nocc; ++ii)
    for (int aa = 0; aa < nvir; ++aa) {
      unsigned i = occ[ii], a = vir[aa];
      double h_el = h->T[a + i * n];
      const double* G = h->G_red + a * n + i * n2;
      const double* Vr = h->V_red + a * n + i * n2;
      for (int p = 0; p < nocc; ++p) h_el += G[occ[p]];
      for (int p = 0; p < nocc_o; ++p) h_el += Vr[occ_o[p]];
      if (fabs(coeff * h_el) < tol) continue;
      uint64_t ex = same ^ ((uint64_t)1 << i) ^ ((uint64_t)1 << a);
      double sign = sx_sign(same, a, i);
      h_el *= sign;
      double h_diag = root + eps[a] - eps[i] - G2(a, i) - G2(i, a);
      if (same_is_alpha) cv_push(out, ex, othr, coeff * h_el, E0 - h_diag);
      else cv_push(out, othr, ex, coeff * h_el, E0 - h_diag);
    }
}
static void emit_ss_doubles(const op_ham* h, cvec* out, double coeff, uint64_t same,
                            uint64_t othr, int same_is_alpha, const unsigned* occ, int nocc,
                            const unsigned* vir, int nvir, const double* eps, double tol,
                            double root, double E0) {
  size_t n = h->n, n2 = n * n;
  for (int ii = 0; ii < nocc; ++ii)
    for (int aa = 0; aa < nvir; ++aa) {
      unsigned i = occ[ii], a = vir[aa];
      const double* V_ai = h->V + (a + i * n) * n2;
      for (int jj = ii + 1; jj < nocc; ++jj)
        for (int bb = aa + 1; bb < nvir; ++bb) {
          unsigned j = occ[jj], b = vir[bb];
          double V_aibj = V_ai[b + j * n];
          double V_ajbi = (h->V + (a + j * n) * n2)[b + i * n];
          double G_aibj = V_aibj - V_ajbi;
          if (fabs(coeff * G_aibj) < tol) continue;
          uint64_t full_ex = ((uint64_t)1 << i) | ((uint64_t)1 << j) | ((uint64_t)1 << a) |
                             ((uint64_t)1 << b);
          uint64_t ex_spin = same ^ full_ex;
          unsigned o1, v1, o2, v2; double sign;
          dx_sign_indices(same, ex_spin, full_ex, &o1, &v1, &o2, &v2, &sign);
          double h_el = sign * G_aibj;
          double h_diag = root + eps[a] + eps[b] - eps[i] - eps[j] + G2(i, j) + G2(j, i) +
                          G2(a, b) + G2(b, a) - G2(a, i) - G2(i, a) - G2(b, i) - G2(i, b) -
                          G2(a, j) - G2(j, a) - G2(b, j) - G2(j, b);
          if (same_is_alpha) cv_push(out, ex_spin, othr, coeff * h_el, E0 - h_diag);
          else cv_push(out, othr, ex_spin, coeff * h_el, E0 - h_diag);
        }
    }
}
static void emit_os_doubles(const op_ham* h, cvec* out, double coeff, uint64_t sa, uint64_t sb,
                            const unsigned* oa, int na, const unsigned* ob, int nb,
                            const unsigned* va, int nva, const unsigned* vb, int nvb,
                            const double* eps_a, const double* eps_b, double tol, double root,
                            double E0) {
  size_t n = h->n, n2 = n * n;
  for (int ii = 0; ii < na; ++ii)
    for (int aa = 0; aa < nva; ++aa) {
      unsigned i = oa[ii], a = va[aa];
      const double* V_ai = h->V + a + i * n;
      double sign_a = sx_sign(sa, a, i);
      for (int jj = 0; jj < nb; ++jj)
        for (int bb = 0; bb < nvb; ++bb) {
          unsigned j = ob[jj], b = vb[bb];
          double V_aibj = V_ai[(b + j * n) * n2];
          if (fabs(coeff * V_aibj) < tol) continue;
          double sign_b = sx_sign(sb, b, j);
          double sign = sign_a * sign_b;
          uint64_t ea = sa ^ ((uint64_t)1 << i) ^ ((uint64_t)1 << a);
          uint64_t eb = sb ^ ((uint64_t)1 << j) ^ ((uint64_t)1 << b);
          double h_el = sign * V_aibj;
          double h_diag = root + eps_a[a] + eps_b[b] - eps_a[i] - eps_b[j] + V2(i, j) +
                          V2(a, b) - G2(a, i) - G2(i, a) - G2(b, j) - G2(j, b) - V2(a, j) -
                          V2(i, b);
          cv_push(out, ea, eb, coeff * h_el, E0 - h_diag);
        }
    }
}
#undef G2
#undef V2

static void gen_contribs(const op_ham* h, const op_asci_search_opts* o, const uint64_t* ca,
                         const uint64_t* cb, const double* coeff, int64_t nc, double E0,
                         cvec* out) {
  int n = h->n;
  unsigned oa[64], ob[64], va[64], vb[64];
  double eps_a[64], eps_b[64];
  for (int64_t i = 0; i < nc; ++i) {
    uint64_t sa = ca[i], sb = cb[i];
    int na = occ_list(sa, oa), nb = occ_list(sb, ob);
    int nva = occ_list(~sa & low_mask(n), va), nvb = occ_list(~sb & low_mask(n), vb);
    orbital_ens(h, oa, na, ob, nb, eps_a);
    orbital_ens(h, ob, nb, oa, na, eps_b);
    double root = me_diag(h, oa, na, ob, nb);
    double c = coeff[i];
    emit_singles(h, out, c, sa, sb, 1, oa, na, va, nva, ob, nb, eps_a, o->h_el_tol, root, E0);
    emit_singles(h, out, c, sb, sa, 0, ob, nb, vb, nvb, oa, na, eps_b, o->h_el_tol, root, E0);
    if (!o->just_singles) {
      emit_ss_doubles(h, out, c, sa, sb, 1, oa, na, va, nva, eps_a, o->h_el_tol, root, E0);
      emit_ss_doubles(h, out, c, sb, sa, 0, ob, nb, vb, nvb, eps_b, o->h_el_tol, root, E0);
      emit_os_doubles(h, out, c, sa, sb, oa, na, ob, nb, va, nva, vb, nvb, eps_a, eps_b,
                      o->h_el_tol, root, E0);
    }
    /* "No excitation (push inf to remove from list)" determinant_search.hpp:659-660 */
    cv_push(out, sa, sb, INFINITY, 1.0);
  }
}
/* sort + accumulate (determinant_sort.hpp:115-136), in place; returns unique count */
static int64_t sort_accumulate(cvec* v) {
  if (!v->n) return 0;
  qsort(v->p, v->n, sizeof(contrib_t), cmp_contrib);
  int64_t w = 0;
  for (int64_t i = 1; i < v->n; ++i) {
    if (v->p[i].a == v->p[w].a && v->p[i].b == v->p[w].b) v->p[w].cm += v->p[i].cm;
    else v->p[++w] = v->p[i];
  }
  v->n = w + 1;
  return v->n;
}
int64_t op_asci_candidates(const op_ham* h, const op_asci_search_opts* o, const uint64_t* ca,
                           const uint64_t* cb, const double* coeff, int64_t nc, double E0,
                           uint64_t* out_a, uint64_t* out_b, double* out_cm, double* out_hd) {
  cvec v = {0, 0, 0};
  gen_contribs(h, o, ca, cb, coeff, nc, E0, &v);
  int64_t nu = sort_accumulate(&v);
  if (out_a)
    for (int64_t i = 0; i < nu; ++i) {
      out_a[i] = v.p[i].a; out_b[i] = v.p[i].b; out_cm[i] = v.p[i].cm; out_hd[i] = v.p[i].hd;
    }
  free(v.p);
  return nu;
}
static int cmp_dbl_desc(const void* x, const void* y) {
  double a = *(const double*)x, b = *(const double*)y;
  return a > b ? -1 : (a < b);
}
int64_t op_asci_search(const op_ham* h, const op_asci_search_opts* o, const uint64_t* ca,
                       const uint64_t* cb, const double* coeff, int64_t nc, double E0,
                       uint64_t* out_a, uint64_t* out_b, int64_t cap, double* stats) {
  cvec v = {0, 0, 0};
  gen_contribs(h, o, ca, cb, coeff, nc, E0, &v);
  int64_t ngen = v.n;
  int64_t nu = sort_accumulate(&v);
  /* prune |rv| <= rv_prune_tol (determinant_search.hpp:676-681), drop rv = inf (:970-974) */
  int64_t m = 0;
  double* score = (double*)malloc((nu + 1) * 8);
  for (int64_t i = 0; i < nu; ++i) {
    double rv = v.p[i].cm / v.p[i].hd;
    if (!(fabs(rv) > o->rv_prune_tol)) continue;
    if (isinf(rv)) continue;
    v.p[m] = v.p[i];
    score[m] = fabs(rv);
    ++m;
  }
  /* top-k with ties retained (determinant_search.hpp:976-1080) */
  int64_t top_k = o->ndets_max - nc;
  double kth = 0., below = 0.;
  int64_t nkeep = m;
  if (o->ndets_max >= nc && m > top_k) {
    double* s2 = (double*)malloc(m * 8);
    memcpy(s2, score, m * 8);
    qsort(s2, m, 8, cmp_dbl_desc);
    /* top_k == 0: the reference's max_element over an empty range dereferences element 0,
     * i.e. the largest score (determinant_search.hpp:1056-1062) */
    kth = s2[top_k > 0 ? top_k - 1 : 0];
    nkeep = 0;
    for (int64_t i = 0; i < m; ++i) {
      if (score[i] >= kth) ++nkeep;
      else if (score[i] > below) below = score[i];
    }
    free(s2);
  }
  if (stats) { stats[0] = (double)ngen; stats[1] = (double)nu; stats[2] = kth; stats[3] = below; stats[4] = (double)nkeep; }
  int64_t total = nkeep + nc;
  if (total > cap) { free(score); free(v.p); return -total; }
  int64_t w = 0;
  for (int64_t i = 0; i < m; ++i)
    if (nkeep == m || score[i] >= kth) { out_a[w] = v.p[i].a; out_b[w] = v.p[i].b; ++w; }
  for (int64_t i = 0; i < nc; ++i) { out_a[w] = ca[i]; out_b[w] = cb[i]; ++w; }
  free(score); free(v.p);
  return w;
}


/* ---------------------------------------------------------------------------------------
 * Reduced density matrices (util/rdms.hpp, symm = true everywhere: the generators call the
 * contribution routines once per unordered pair, bra = the lower index). */
#define T4(t, n, p, q, r, s_) (t)[(size_t)(p) + (size_t)(q) * (n) + (size_t)(r) * (n) * (n) + (size_t)(s_) * (n) * (n) * (n)]

/* rdm_contributions_4<true> (rdms.hpp:36-65) */
static void rdm4(int n, uint64_t bra, uint64_t ket, uint64_t ex, double val, double* trdm) {
  if (!trdm) return;
  unsigned o1, v1, o2, v2;
  double sign;
  dx_sign_indices(bra, ket, ex, &o1, &v1, &o2, &v2, &sign);
  val *= sign * 0.5;
  T4(trdm, n, v1, o1, v2, o2) += val;
  T4(trdm, n, v2, o1, v1, o2) -= val;
  T4(trdm, n, v1, o2, v2, o1) -= val;
  T4(trdm, n, v2, o2, v1, o1) += val;
  T4(trdm, n, o2, v2, o1, v1) += val;
  T4(trdm, n, o2, v1, o1, v2) -= val;
  T4(trdm, n, o1, v2, o2, v1) -= val;
  T4(trdm, n, o1, v1, o2, v2) += val;
}
/* rdm_contributions_22<true> (rdms.hpp:85-113) */
static void rdm22(int n, uint64_t bra_a, uint64_t ket_a, uint64_t ex_a, uint64_t bra_b,
                  uint64_t ket_b, uint64_t ex_b, double val, double* trdm) {
  if (!trdm) return;
  unsigned o1, v1, o2, v2;
  double sa, sb;
  sx_sign_indices(bra_a, ket_a, ex_a, &o1, &v1, &sa);
  sx_sign_indices(bra_b, ket_b, ex_b, &o2, &v2, &sb);
  val *= sa * sb * 0.5;
  T4(trdm, n, v1, o1, v2, o2) += val;
  T4(trdm, n, v2, o2, v1, o1) += val;
  T4(trdm, n, o2, v2, o1, v1) += val;
  T4(trdm, n, o1, v1, o2, v2) += val;
}
/* rdm_contributions_22_spin_dep<true> (rdms.hpp:134-157): NB the roles -- (o2,v2) from the
 * alpha strings, (o1,v1) from the beta strings */
static void rdm22_sd(int n, uint64_t bra_a, uint64_t ket_a, uint64_t ex_a, uint64_t bra_b,
                     uint64_t ket_b, uint64_t ex_b, double val, double* aabb) {
  if (!aabb) return;
  unsigned o1, v1, o2, v2;
  double sa, sb;
  sx_sign_indices(bra_a, ket_a, ex_a, &o2, &v2, &sb);
  sx_sign_indices(bra_b, ket_b, ex_b, &o1, &v1, &sa);
  val *= sa * sb * 0.5;
  T4(aabb, n, v1, o1, v2, o2) += val;
  T4(aabb, n, o1, v1, o2, v2) += val;
}
/* rdm_contributions_2<true> (rdms.hpp:177-244): occ_same = bra occupation of the excited spin */
static void rdm2(int n, uint64_t bra, uint64_t ket, uint64_t ex, const unsigned* occ_same, int ns,
                 const unsigned* occ_othr, int no, double val, double* ordm, double* trdm) {
  unsigned o1, v1;
  double sign;
  sx_sign_indices(bra, ket, ex, &o1, &v1, &sign);
  if (ordm) {
    ordm[v1 + (size_t)o1 * n] += sign * val;
    ordm[o1 + (size_t)v1 * n] += sign * val;
  }
  if (!trdm) return;
  val *= sign * 0.5;
  for (int i = 0; i < ns; ++i) {
    const unsigned p = occ_same[i];
    T4(trdm, n, v1, o1, p, p) += val;
    T4(trdm, n, p, p, v1, o1) += val;
    T4(trdm, n, v1, p, p, o1) -= val;
    T4(trdm, n, p, o1, v1, p) -= val;
  }
  for (int i = 0; i < ns; ++i) {
    const unsigned p = occ_same[i];
    T4(trdm, n, p, p, o1, v1) += val;
    T4(trdm, n, o1, v1, p, p) += val;
    T4(trdm, n, o1, p, p, v1) -= val;
    T4(trdm, n, p, v1, o1, p) -= val;
  }
  for (int i = 0; i < no; ++i) {
    const unsigned p = occ_othr[i];
    T4(trdm, n, v1, o1, p, p) += val;
    T4(trdm, n, p, p, v1, o1) += val;
  }
  for (int i = 0; i < no; ++i) {
    const unsigned p = occ_othr[i];
    T4(trdm, n, o1, v1, p, p) += val;
    T4(trdm, n, p, p, o1, v1) += val;
  }
}
/* rdm_contributions_2_spin_dep<true, transpose> (rdms.hpp:267-335) */
static void rdm2_sd(int n, int transpose, uint64_t bra, uint64_t ket, uint64_t ex,
                    const unsigned* occ_ss, int ns, const unsigned* occ_os, int no, double val,
                    double* ordm_ss, double* trdm_ss, double* trdm_os) {
  unsigned o1, v1;
  double sign;
  sx_sign_indices(bra, ket, ex, &o1, &v1, &sign);
  if (ordm_ss) {
    ordm_ss[v1 + (size_t)o1 * n] += sign * val;
    ordm_ss[o1 + (size_t)v1 * n] += sign * val;
  }
  val *= sign * 0.5;
  if (trdm_ss) {
    for (int i = 0; i < ns; ++i) {
      const unsigned p = occ_ss[i];
      T4(trdm_ss, n, v1, o1, p, p) += val;
      T4(trdm_ss, n, p, p, v1, o1) += val;
      T4(trdm_ss, n, v1, p, p, o1) -= val;
      T4(trdm_ss, n, p, o1, v1, p) -= val;
    }
    for (int i = 0; i < ns; ++i) {
      const unsigned p = occ_ss[i];
      T4(trdm_ss, n, p, p, o1, v1) += val;
      T4(trdm_ss, n, o1, v1, p, p) += val;
      T4(trdm_ss, n, o1, p, p, v1) -= val;
      T4(trdm_ss, n, p, v1, o1, p) -= val;
    }
  }
  if (trdm_os) {
    for (int i = 0; i < no; ++i) {
      const unsigned p = occ_os[i];
      if (transpose) T4(trdm_os, n, v1, o1, p, p) += val;
      else T4(trdm_os, n, p, p, v1, o1) += val;
    }
    for (int i = 0; i < no; ++i) {
      const unsigned p = occ_os[i];
      if (transpose) T4(trdm_os, n, o1, v1, p, p) += val;
      else T4(trdm_os, n, p, p, o1, v1) += val;
    }
  }
}
/* rdm_contributions_diag (rdms.hpp:352-394) */
static void rdm_diag(int n, const unsigned* oa, int na, const unsigned* ob, int nb, double val,
                     double* ordm, double* trdm) {
  if (ordm) {
    for (int i = 0; i < na; ++i) ordm[oa[i] + (size_t)oa[i] * n] += val;
    for (int i = 0; i < nb; ++i) ordm[ob[i] + (size_t)ob[i] * n] += val;
  }
  if (!trdm) return;
  val *= 0.5;
  for (int j = 0; j < na; ++j)
    for (int i = 0; i < na; ++i) {
      T4(trdm, n, oa[i], oa[i], oa[j], oa[j]) += val;
      T4(trdm, n, oa[i], oa[j], oa[j], oa[i]) -= val;
    }
  for (int j = 0; j < nb; ++j)
    for (int i = 0; i < nb; ++i) {
      T4(trdm, n, ob[i], ob[i], ob[j], ob[j]) += val;
      T4(trdm, n, ob[i], ob[j], ob[j], ob[i]) -= val;
    }
  for (int j = 0; j < nb; ++j)
    for (int i = 0; i < na; ++i) {
      T4(trdm, n, oa[i], oa[i], ob[j], ob[j]) += val;
      T4(trdm, n, ob[j], ob[j], oa[i], oa[i]) += val;
    }
}
/* rdm_contributions_diag_spin_dep (rdms.hpp:415-460) */
static void rdm_diag_sd(int n, const unsigned* oa, int na, const unsigned* ob, int nb, double val,
                        double* ordm_aa, double* ordm_bb, double* aaaa, double* bbbb, double* aabb) {
  if (ordm_aa) for (int i = 0; i < na; ++i) ordm_aa[oa[i] + (size_t)oa[i] * n] += val;
  if (ordm_bb) for (int i = 0; i < nb; ++i) ordm_bb[ob[i] + (size_t)ob[i] * n] += val;
  val *= 0.5;
  if (aaaa)
    for (int j = 0; j < na; ++j)
      for (int i = 0; i < na; ++i) {
        T4(aaaa, n, oa[i], oa[i], oa[j], oa[j]) += val;
        T4(aaaa, n, oa[i], oa[j], oa[j], oa[i]) -= val;
      }
  if (bbbb)
    for (int j = 0; j < nb; ++j)
      for (int i = 0; i < nb; ++i) {
        T4(bbbb, n, ob[i], ob[i], ob[j], ob[j]) += val;
        T4(bbbb, n, ob[i], ob[j], ob[j], ob[i]) -= val;
      }
  if (aabb)
    for (int j = 0; j < nb; ++j)
      for (int i = 0; i < na; ++i) T4(aabb, n, ob[j], ob[j], oa[i], oa[i]) += val;
}

/* the pair loop of form_rdms(_spin_dep) for bra == ket (sorted_double_loop.hpp:560-620,
 * 682-744): pairs i <= j, alpha-empty determinants skipped, popcount filter, |C_i C_j| > 1e-16 */
static void rdm_pairs(int norb, const uint64_t* alpha, const uint64_t* beta, int64_t n,
                      const double* C, int spin_dep, double* o1, double* o2, double* t1,
                      double* t2, double* t3) {
  unsigned oa[64], ob[64];
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t ba = alpha[i], bb = beta[i];
    if (!ba) continue;
    const int na = occ_list(ba, oa), nb = occ_list(bb, ob);
    for (int64_t j = i; j < n; ++j) {
      const uint64_t ka = alpha[j], kb = beta[j];
      if (!ka) continue;
      const uint64_t exa = ba ^ ka, exb = bb ^ kb;
      const int ca = popc(exa), cb = popc(exb);
      if (ca + cb > 4) continue;
      const double val = C[i] * C[j];
      if (!(fabs(val) > 1e-16)) continue;
      if (!spin_dep) {
        if (ca == 4) rdm4(norb, ba, ka, exa, val, t1);
        else if (cb == 4) rdm4(norb, bb, kb, exb, val, t1);
        else if (ca == 2 && cb == 2) rdm22(norb, ba, ka, exa, bb, kb, exb, val, t1);
        else if (ca == 2) rdm2(norb, ba, ka, exa, oa, na, ob, nb, val, o1, t1);
        else if (cb == 2) rdm2(norb, bb, kb, exb, ob, nb, oa, na, val, o1, t1);
        else rdm_diag(norb, oa, na, ob, nb, val, o1, t1);
      } else {
        if (ca == 4) rdm4(norb, ba, ka, exa, val, t1);
        else if (cb == 4) rdm4(norb, bb, kb, exb, val, t2);
        else if (ca == 2 && cb == 2) rdm22_sd(norb, ba, ka, exa, bb, kb, exb, val, t3);
        else if (ca == 2) rdm2_sd(norb, 0, ba, ka, exa, oa, na, ob, nb, val, o1, t1, t3);
        else if (cb == 2) rdm2_sd(norb, 1, bb, kb, exb, ob, nb, oa, na, val, o2, t2, t3);
        else rdm_diag_sd(norb, oa, na, ob, nb, val, o1, o2, t1, t2, t3);
      }
    }
  }
}
void op_form_rdms(int norb, const uint64_t* alpha, const uint64_t* beta, int64_t n,
                  const double* C, double* ordm, double* trdm) {
  rdm_pairs(norb, alpha, beta, n, C, 0, ordm, NULL, trdm, NULL, NULL);
}
void op_form_rdms_spin_dep(int norb, const uint64_t* alpha, const uint64_t* beta, int64_t n,
                           const double* C, double* ordm_aa, double* ordm_bb,
                           double* trdm_aaaa, double* trdm_bbbb, double* trdm_aabb) {
  rdm_pairs(norb, alpha, beta, n, C, 1, ordm_aa, ordm_bb, trdm_aaaa, trdm_bbbb, trdm_aabb);
}
