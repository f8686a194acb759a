/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * Plain-C restatement of the reference's (MACIS) configuration-interaction hot
 * path: determinant bit operations, Slater-Condon matrix elements, the
 * SortedDoubleLoop CSR pattern/threshold semantics, the sparse sigma product,
 * the single-root Davidson solver and the ASCI connected-determinant search.
 * It is the checker for the CUDA path (tests/, __graft_entry__.smoke(),
 * bench.py's cpu_baseline leg) and is pinned against the reference's own
 * golden vectors and against oracle/_ref (the reference compiled unmodified)
 * by tests/test_oracle.py. The product never links or loads it.
 *
 * Determinant layout at this boundary: separate alpha / beta occupation words
 * (uint64 each, bit p = orbital p), i.e. the two halves of the reference's
 * wfn_t<N> (external/macis/include/macis/wfn/raw_bitset.hpp:94-106).
 */
#ifndef ORACLE_PORT_H
#define ORACLE_PORT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct op_ham op_ham;

op_ham* op_ham_create(int norb, const double* T, const double* V);
void op_ham_destroy(op_ham* h);
void op_ham_intermediates(const op_ham* h, double* G_red, double* V_red, double* G2_red,
                          double* V2_red);

int64_t op_generate_hilbert_space(int norb, int nalpha, int nbeta, uint64_t* alpha,
                                  uint64_t* beta);

double op_matrix_element(const op_ham* h, uint64_t bra_a, uint64_t bra_b, uint64_t ket_a,
                         uint64_t ket_b);

/* CSR build with SortedDoubleLoop semantics. Two-call protocol: first call with
 * colind == NULL fills rowptr[n+1] and returns nnz; second call fills colind/nzval. */
int64_t op_hbuild(const op_ham* h, const uint64_t* alpha, const uint64_t* beta, int64_t n,
                  double thresh, int64_t* rowptr, int64_t* colind, double* nzval);
/* rows [r0, r1) only (row-block of the symmetric matrix; columns global) */
int64_t op_hbuild_rows(const op_ham* h, const uint64_t* alpha, const uint64_t* beta,
                       int64_t n, int64_t r0, int64_t r1, double thresh, int64_t* rowptr,
                       int64_t* colind, double* nzval);
/* the same under the pair-based generators' rules (residue_arrays / dynamic_bit_masking) */
int64_t op_hbuild_rows_gen(const op_ham* h, const uint64_t* alpha, const uint64_t* beta,
                           int64_t n, int64_t r0, int64_t r1, double thresh, int pair_rule,
                           int64_t* rowptr, int64_t* colind, double* nzval);

void op_spmv(int64_t n, const int64_t* rowptr, const int64_t* colind, const double* nzval,
             const double* x, double* y);
void op_extract_diagonal(int64_t n, const int64_t* rowptr, const int64_t* colind,
                         const double* nzval, double* D);

/* returns 0 converged, 1 not converged ("Davidson Did Not Converge!"), 2 gram-schmidt failure */
int op_davidson(int64_t n, const int64_t* rowptr, const int64_t* colind,
                const double* nzval, int64_t max_m, double tol, double* X, int64_t* niter,
                double* eigval, double* trace /* 2*max_m doubles or NULL: (lambda, rnorm) */);
/* serial_selected_ci_diag guess policy followed by op_davidson */
int op_selected_ci_diag(int64_t n, const int64_t* rowptr, const int64_t* colind,
                        const double* nzval, int64_t max_m, double tol, double* C,
                        int64_t* niter, double* eigval);

/* symmetric eigensolver used for the Rayleigh-Ritz step (lower triangle of A, col-major,
 * eigenvalues ascending, eigenvectors in columns of A) */
int op_syev_lower(int n, double* A, int lda, double* W);

typedef struct {
  int64_t ndets_max;
  double h_el_tol;
  double rv_prune_tol;
  int32_t just_singles;
  int32_t pad;
} op_asci_search_opts;

/* ASCI search in canonical accumulation order (parents ascending). Returns number of
 * determinants written (selected followed by the core dets), or -(needed) if cap is
 * too small. stats (may be NULL): [0] generated contributions, [1] unique candidates,
 * [2] kth |rv| pivot, [3] largest |rv| strictly below the pivot (cut gap), [4] n kept */
int64_t op_asci_search(const op_ham* h, const op_asci_search_opts* o, const uint64_t* calpha,
                       const uint64_t* cbeta, const double* coeff, int64_t ncdets, double E0,
                       uint64_t* out_alpha, uint64_t* out_beta, int64_t cap, double* stats);
/* the accumulated candidate table itself (for kernel-level parity): arrays sized by the
 * return value of a first call with all outputs NULL */
int64_t op_asci_candidates(const op_ham* h, const op_asci_search_opts* o,
                           const uint64_t* calpha, const uint64_t* cbeta,
                           const double* coeff, int64_t ncdets, double E0,
                           uint64_t* out_alpha, uint64_t* out_beta, double* out_cmatel,
                           double* out_hdiag);

/* Reduced density matrices of sum_i C_i |D_i> (SortedDoubleLoopHamiltonianGenerator::
 * form_rdms / form_rdms_spin_dep, sorted_double_loop.hpp:512-760, with the symmetric
 * bra == ket case; contribution rules of util/rdms.hpp). Pairs (i <= j) with
 * |C_i C_j| > 1e-16 contribute. ordm: n*n column-major; trdm: n^4 with
 * (p,q,r,s) at p + q n + r n^2 + s n^3. Outputs are ACCUMULATED INTO (caller zeroes);
 * any output may be NULL. */
void op_form_rdms(int norb, const uint64_t* alpha, const uint64_t* beta, int64_t n,
                  const double* C, double* ordm, double* trdm);
void op_form_rdms_spin_dep(int norb, const uint64_t* alpha, const uint64_t* beta, int64_t n,
                           const double* C, double* ordm_aa, double* ordm_bb,
                           double* trdm_aaaa, double* trdm_bbbb, double* trdm_aabb);

int op_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
