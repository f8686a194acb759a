/* b2ci.h -- C ABI of the B200-native configuration-interaction hot path.
 *
 * Drop-in boundary for the MACIS calls that QDK/Chemistry's CAS / ASCI / PMC adapters
 * make (cpp/src/qdk/chemistry/algorithms/microsoft/macis_{cas,asci,pmc}.cpp). Every entry
 * point names the reference interface it replaces (paths under /root/reference). Plain
 * pointers and sizes only; all functions return 0 on success, non-zero on failure with a
 * message available from b2ci_last_error(). There is no CPU fallback: without a CUDA
 * device b2ci_ctx_create fails.
 *
 * Determinants cross the boundary exactly as the reference stores wfn_t<N>
 * (external/macis/include/macis/wfn/raw_bitset.hpp:94-106):
 *   words_per_det == 1 : wfn_t<64>,  alpha = bits 0..31, beta = bits 32..63
 *   words_per_det == 2 : wfn_t<128>, word 0 = alpha, word 1 = beta
 * Integrals: T is n*n column-major, V is n^4 with V[p + q n + r n^2 + s n^3] = (pq|rs)
 * (macis_cas.cpp:76-82). CSR indices are 0-based; the device keeps int32 column indices
 * and the download presents the reference's int64 (mcscf/cas.hpp:53, macis_asci.cpp:167).
 */
#ifndef B2CI_H
#define B2CI_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2ci_ctx b2ci_ctx;   /* device, stream, integrals, optional NCCL communicator */
typedef struct b2ci_dets b2ci_dets; /* device-resident determinant list (alpha/beta SoA)     */
typedef struct b2ci_csr b2ci_csr;   /* device-resident CSR row block                         */

const char* b2ci_last_error(void);
const char* b2ci_version(void);

/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for the
 * legacy default stream. All kernels of this context are launched on it. */
int b2ci_ctx_create(int device, void* stream, b2ci_ctx** out);
int b2ci_ctx_destroy(b2ci_ctx* ctx);
int b2ci_ctx_synchronize(b2ci_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches claim) */
int64_t b2ci_ctx_launch_count(const b2ci_ctx* ctx);
/* Device memory of this library is drawn from the device's stream-ordered pool and kept
 * there when handles are freed; this returns the cached blocks to the driver. */
int b2ci_ctx_trim(b2ci_ctx* ctx);

/* ---- multi-GPU (one process per GPU). Replaces MACIS' MPI communicator argument
 * (solvers/selected_ci_diag.hpp:181, solvers/davidson.hpp:391-688). The 128-byte id is
 * produced on rank 0 and distributed by the caller (torch.distributed broadcast). */
int b2ci_comm_unique_id(void* id128);
int b2ci_comm_init(b2ci_ctx* ctx, const void* id128, int rank, int nranks);
int b2ci_comm_rank(const b2ci_ctx* ctx, int* rank, int* nranks);

/* ---- integrals: HamiltonianGeneratorBase ctor + generate_integral_intermediates_
 * (external/macis/src/macis/hamiltonian_generator/base.ipp:27-77). Host pointers. */
int b2ci_integrals_upload(b2ci_ctx* ctx, int norb, const double* T, const double* V);
/* Orbital rotation of the uploaded integrals, T <- C^T T C and the four-index analogue for V, followed
 * by the intermediates again: two_index_transform / four_index_transform
 * (external/macis/src/macis/transform.cxx:22-96) + generate_integral_intermediates as asci_grow's
 * natural-orbital step uses them (external/macis/include/macis/asci/grow.hpp:163-215). C: HOST, n x n
 * column-major, columns = new orbitals. T_out (n^2) / V_out (n^4): HOST, may be NULL. */
int b2ci_integrals_rotate(b2ci_ctx* ctx, const double* C, double* T_out, double* V_out);
/* debug/parity: copy G_red, V_red (n^3), G2_red, V2_red (n^2) back to the host */
int b2ci_integrals_download(b2ci_ctx* ctx, double* G_red, double* V_red, double* G2_red,
                            double* V2_red);

/* ---- determinant lists */
int b2ci_dets_upload(b2ci_ctx* ctx, const uint64_t* words, int words_per_det, int64_t n,
                     b2ci_dets** out);
/* generate_hilbert_space (external/macis/include/macis/sd_operations.hpp:333-351) */
int b2ci_dets_generate_fci(b2ci_ctx* ctx, int norb, int nalpha, int nbeta, b2ci_dets** out);
int b2ci_dets_size(const b2ci_dets* d, int64_t* n);
int b2ci_dets_download(b2ci_ctx* ctx, const b2ci_dets* d, uint64_t* words, int words_per_det);
int b2ci_dets_free(b2ci_ctx* ctx, b2ci_dets* d);

/* ---- Hamiltonian build: make_csr_hamiltonian<index_t> with the SortedDoubleLoop
 * generator (external/macis/include/macis/csr_hamiltonian.hpp:74-80,
 * hamiltonian_generator/sorted_double_loop.hpp:86-451). Builds rows [row_begin,row_end)
 * of the symmetric matrix over `dets` (columns are global). Pattern contract:
 * popcount(bra^ket) <= 4, alpha-empty determinants skipped, |h| > h_thresh kept when
 * h_thresh > 0, everything structurally connected kept when h_thresh == 0. */
int b2ci_hbuild_csr(b2ci_ctx* ctx, const b2ci_dets* dets, int64_t row_begin, int64_t row_end,
                    double h_thresh, b2ci_csr** out);
/* Which generator's pattern rules b2ci_hbuild_csr / b2ci_hbuild_csr_patched follow: QDK's
 * hamiltonian_build_algorithm setting (macis_asci.hpp:174-179, macis_asci.cpp:92-118).
 *   SORTED_DOUBLE_LOOP : sorted_double_loop.hpp:86-451 -- |h| > h_thresh kept (diagonal included),
 *                        alpha-empty determinants skipped
 *   RESIDUE_ARRAYS, DYNAMIC_BIT_MASKING : both end in build_csr_from_pairs
 *                        (connection_build_utils.hpp:125-249) -- the diagonal is always stored,
 *                        an off-diagonal element is dropped only when |h| < h_thresh, alpha-empty
 *                        determinants are ordinary determinants. The reference's two generators
 *                        differ only in how the CPU enumerates the connected pairs; on the device
 *                        both take the same row scan. Matrix elements are identical in all three.
 * The selection stays with the context until changed. */
#define B2CI_GEN_SORTED_DOUBLE_LOOP 0
#define B2CI_GEN_RESIDUE_ARRAYS 1
#define B2CI_GEN_DYNAMIC_BIT_MASKING 2
int b2ci_set_hamiltonian_generator(b2ci_ctx* ctx, int generator);
/* Incremental build between ASCI iterations: CachedHamiltonianState + build_patched_operator
 * (external/macis/include/macis/solvers/incremental_h_build.hpp:192-356), called from
 * selected_ci_diag (solvers/selected_ci_diag.hpp:217-256). old_H must be the full square matrix
 * of old_dets built with the same integrals and h_thresh; both lists spin_comparator-sorted
 * (alpha-major, then beta). The kept x kept block is taken from old_H (columns renumbered), only
 * kept x added and added x all are evaluated, and the blocks are merged into one CSR of new_dets
 * that is bit-identical to b2ci_hbuild_csr(new_dets, 0, n, h_thresh). When n_kept / n_new <
 * min_overlap nothing is built and *out is NULL (the reference falls back to the full build the
 * same way, :266-283). n_kept may be NULL. Not available with a communicator, like the
 * reference's MPI builds (selected_ci_diag.hpp:196-202). */
int b2ci_hbuild_csr_patched(b2ci_ctx* ctx, const b2ci_dets* old_dets, const b2ci_csr* old_H,
                            const b2ci_dets* new_dets, double h_thresh, double min_overlap,
                            b2ci_csr** out, int64_t* n_kept);
/* sparsexx::csr_matrix from caller arrays (python/src/pybind11/algorithms/
 * davidson_solver.cpp:60-80); host pointers, int64 indices, square n x n */
int b2ci_csr_upload(b2ci_ctx* ctx, int64_t n, int64_t nnz, const int64_t* rowptr,
                    const int64_t* colind, const double* nzval, b2ci_csr** out);
int b2ci_csr_info(const b2ci_csr* m, int64_t* nrows, int64_t* ncols, int64_t* nnz,
                  int64_t* row_begin);
int b2ci_csr_download(b2ci_ctx* ctx, const b2ci_csr* m, int64_t* rowptr, int64_t* colind,
                      double* nzval);
/* device pointers of the resident CSR (rowptr int64[nrows+1], colind int32[nnz], nzval) */
int b2ci_csr_device_ptrs(const b2ci_csr* m, const int64_t** rowptr, const int32_t** colind,
                         const double** nzval);
int b2ci_csr_free(b2ci_ctx* ctx, b2ci_csr* m);

/* ---- sigma: sparsexx::spblas::gespmbv with K = 1, alpha = 1, beta = 0
 * (external/macis/src/sparsexx/include/sparsexx/spblas/spmbv.hpp:49-85), called from
 * SparseMatrixOperator::operator_action (solvers/davidson.hpp:79-90).
 * x_dev has ncols entries, y_dev nrows entries; DEVICE pointers. */
int b2ci_spmv(b2ci_ctx* ctx, const b2ci_csr* m, const double* x_dev, double* y_dev);
/* same with HOST buffers (copies inside) */
int b2ci_spmv_host(b2ci_ctx* ctx, const b2ci_csr* m, const double* x, double* y);
/* Row-sharded sigma: exchange of the local trial-vector blocks followed by the local SpMV -- the
 * counterpart of sparsexx::spblas::pgespmv (sparsexx/spblas/pspmbv.hpp:316-405). With peer access
 * between the GPUs every rank stores its block directly into the other ranks' exchange buffers
 * over NVLink (one push kernel + a flag wait, no NCCL launch); otherwise ncclAllGather. x_full_dev
 * (ncols entries) may be NULL; when given it also receives the gathered vector. Without a
 * communicator it is b2ci_spmv(x_local_dev). DEVICE pointers. */
int b2ci_sigma_sharded(b2ci_ctx* ctx, const b2ci_csr* m, const double* x_local_dev,
                       double* x_full_dev, double* y_local_dev);
/* Optional: tell a row block how the rows are split over the ranks (row_offsets has nranks + 1
 * entries, block r = [row_offsets[r], row_offsets[r+1])), the information MACIS keeps in
 * dist_sparse_matrix::row_tiling (sparsexx/matrix_types/dist_sparse_matrix.hpp:83-97). Without
 * it the first b2ci_sigma_sharded / b2ci_davidson call on the block exchanges the block sizes
 * (a host-synchronising all-gather). HOST pointer. */
int b2ci_csr_set_row_partition(b2ci_ctx* ctx, b2ci_csr* m, const int64_t* row_offsets, int nranks);
/* Row cuts for a row-sharded build of a selected-CI list: nparts contiguous blocks with about equal numbers of
 * CONNECTIONS (estimated from the exact degree of nsamples evenly spaced determinants against the whole list), not
 * of rows -- the head of a spin-sorted ASCI list holds the determinants with the most partners. The reference
 * tiles rows evenly (sparsexx/matrix_types/dist_sparse_matrix.hpp:83-97, make_dist_csr_hamiltonian); any
 * contiguous tiling is the same matrix. Deterministic: every rank computes the same cuts from the same list, no
 * exchange. offsets: HOST, nparts + 1 entries, offsets[0] = 0, offsets[nparts] = n. */
int b2ci_dets_balanced_partition(b2ci_ctx* ctx, const b2ci_dets* dets, int nparts, int64_t nsamples,
                                 int64_t* offsets);
/* extract_diagonal_elements (sparsexx/util/submatrix.hpp:354-383); host output, nrows */
int b2ci_csr_diagonal(b2ci_ctx* ctx, const b2ci_csr* m, double* D);

/* ---- Davidson: macis::davidson (solvers/davidson.hpp:259-372) preceded, when
 * use_guess_policy != 0, by serial_selected_ci_diag's guess policy
 * (solvers/selected_ci_diag.hpp:111-158). X: HOST vector of the GLOBAL dimension, in/out.
 * With a communicator, every rank passes its own row block and the same X; rows must
 * tile [0, ncols) in rank order. trace (may be NULL): 2*max_m doubles (lambda, rnorm).
 * Returns 0 converged; B2CI_NOT_CONVERGED mirrors "Davidson Did Not Converge!". */
#define B2CI_NOT_CONVERGED 3
int b2ci_davidson(b2ci_ctx* ctx, const b2ci_csr* m, int64_t max_m, double tol, double* X,
                  int use_guess_policy, int64_t* niter, double* eigval, double* trace);

/* ---- dense branch of the adapters for small spaces (n <= iterative_solver_dimension_cutoff):
 * sparsexx::convert_to_dense + lapack::syev, lowest eigenpair (macis_cas.cpp:89-101,
 * macis_pmc.cpp:98-112). eigvec: HOST, nrows entries. Full square matrices only. */
int b2ci_dense_ground_state(b2ci_ctx* ctx, const b2ci_csr* m, double* eigval, double* eigvec);

/* per-phase device timings of the last b2ci_davidson / b2ci_hbuild_csr / b2ci_asci_search
 * call on this context, in milliseconds (CUDA events on the context stream). Names follow
 * the reference's loggers (h_build: setup/count/fill; davidson: OP_DUR, RR_DUR, RES_DUR,
 * GS_DUR; asci_search: PAIR_DUR, SORT_ACC_DUR, TOPK_DUR). Unknown name -> -1. */
double b2ci_timer_ms(const b2ci_ctx* ctx, const char* name);

/* ---- ASCI search: macis::asci_search (external/macis/include/macis/asci/
 * determinant_search.hpp:808-1123). Core determinants must be spin_comparator-sorted
 * (:363-364). Output: selected determinants followed by the core determinants. */
typedef struct {
  int64_t ndets_max;    /* target size including the core determinants            */
  double h_el_tol;      /* ASCISettings::h_el_tol      (QDK key search_matel_tol) */
  double rv_prune_tol;  /* ASCISettings::rv_prune_tol                              */
  int32_t just_singles; /* ASCISettings::just_singles                              */
  int32_t sort_output;  /* 0: selected determinants then the core ones (the reference's order,
                         * determinant_search.hpp:1107-1114); 1: the whole list in spin_comparator
                         * order, i.e. what asci_iter sorts it into next (asci/iteration.hpp:117-119) */
} b2ci_asci_search_opts;
/* stats (may be NULL, 8 doubles): [0] contributions generated, [1] unique candidates,
 * [2] kth |rv| pivot, [3] largest |rv| below the pivot, [4] number selected */
int b2ci_asci_search(b2ci_ctx* ctx, const b2ci_asci_search_opts* opts,
                     const uint64_t* core_words, int words_per_det, const double* core_coeffs,
                     int64_t ncdets, double E0, uint64_t* out_words, int64_t cap,
                     int64_t* n_out, double* stats);
/* the accumulated candidate table (parity of generation + sort + accumulate). First call
 * with NULL outputs returns the count in *n_out. Keys ascending in (beta, alpha). */
int b2ci_asci_candidates(b2ci_ctx* ctx, const b2ci_asci_search_opts* opts,
                         const uint64_t* core_words, int words_per_det,
                         const double* core_coeffs, int64_t ncdets, double E0,
                         uint64_t* out_words, double* out_cmatel, double* out_hdiag,
                         int64_t* n_out);

/* ---- ASCI-PT2: macis::asci_pt2_constraint (external/macis/include/macis/asci/pt2.hpp:61-565).
 * Second-order energy of the determinants outside the (spin-sorted) wavefunction: every single
 * and double excitation with |c*h| >= pt2_tol is generated, contributions to the same
 * determinant are summed, and EPT2 = sum (sum c*h)^2 / (E_asci - <Q|H|Q>). The reference's
 * alpha-constraint partition is replaced by hash partitions of the determinant key; with a
 * communicator the partitions are dealt to the ranks and the partial sums all-reduced.
 * npt2 (may be NULL): number of external determinants that contributed. */
int b2ci_asci_pt2(b2ci_ctx* ctx, const uint64_t* det_words, int words_per_det, const double* coeffs,
                  int64_t ndets, double E_asci, double pt2_tol, double* ept2, int64_t* npt2);

/* ---- reduced density matrices of sum_i C_i |D_i>:
 * HamiltonianGenerator::form_rdms / form_rdms_spin_dep with bra == ket, as the adapters call them
 * (cpp/src/qdk/chemistry/algorithms/microsoft/macis_base.hpp:166-197, macis_pmc.cpp:128-148;
 * external/macis/include/macis/hamiltonian_generator/sorted_double_loop.hpp:512-760; rules of
 * external/macis/include/macis/util/rdms.hpp). C: HOST, one coefficient per determinant of
 * `dets`. Outputs: HOST, column-major n*n and n^4 ((p,q,r,s) at p + q n + r n^2 + s n^3), any
 * may be NULL, ACCUMULATED INTO as the reference does (zero them first). The orbital count is
 * that of the uploaded integrals. With a communicator the pairs are sharded by rows and the
 * matrices all-reduced (every rank gets the result). */
int b2ci_form_rdms(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C, double* ordm, double* trdm);
int b2ci_form_rdms_spin_dep(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C, double* ordm_aa,
                            double* ordm_bb, double* trdm_aaaa, double* trdm_bbbb, double* trdm_aabb);

/* ---- orbital entropies: HamiltonianGenerator::form_entropies with bra == ket
 * (sorted_double_loop.hpp:760-905, util/entropies.hpp; called from macis_base.hpp:219-245 for
 * calculate_single_orbital_entropies / _two_orbital_entropies / _mutual_information).
 * s1: norb single-orbital entropies (required); s2, mi: norb x norb column-major, either may be
 * NULL (both NULL: only the diagonal pairs are visited, as the reference does). HOST pointers. */
int b2ci_form_entropies(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C, double* s1, double* s2, double* mi);
/* the two halves of it, for parity tests: the OrbitalRDMIntermediates (entropies.hpp:62-205;
 * 3 vectors of norb then 18 norb x norb matrices, order of csrc/entropy.cu) accumulated on the
 * device, and the host assembly of the entropies from them (no GPU needed) */
int64_t b2ci_entropy_intermediate_count(int norb, int need_s2);
int b2ci_entropy_intermediates(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C, int need_s2, double* out);
int b2ci_host_entropies_from_intermediates(int norb, int need_s2, const double* intermediates, double* s1,
                                           double* s2, double* mi);

/* ---- host-side evaluation of the SAME device functions (they are __host__ __device__):
 * lets CPU-only tests check the Slater-Condon code against the oracle without a GPU. */
double b2ci_host_matrix_element(int norb, const double* T, const double* V, uint64_t bra_alpha,
                                uint64_t bra_beta, uint64_t ket_alpha, uint64_t ket_beta);

/* host symmetric eigensolver used for the Rayleigh-Ritz step (lower triangle, column-major,
 * eigenvalues ascending, eigenvectors in the columns of A); stands where the reference calls
 * lapack::syev (external/macis/src/lobpcgxx/include/lobpcgxx/rayleigh_ritz.hpp:75) */
int b2ci_host_sym_eig_lower(int n, double* A, int lda, double* W);
/* lowest eigenpair only (what the single-root Davidson uses every iteration): Householder
 * tridiagonalisation + Sturm bisection + inverse iteration; A is read (lower triangle) */
int b2ci_host_sym_eig_lowest(int n, const double* A, int lda, double* lambda, double* vec);

#ifdef __cplusplus
}
#endif
#endif /* B2CI_H */
