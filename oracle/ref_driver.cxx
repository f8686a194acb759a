// TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
//
// C-ABI driver around the UNMODIFIED reference implementation (MACIS, under
// /root/reference/external/macis, only -I included, never copied). It is
// compiled by oracle/Makefile into oracle/_ref/libmacis_ref.so and is used
//   * to pin the C restatement in oracle/port (tests/, golden generation),
//   * as the CPU baseline ("kind": "reference") timed by bench.py.
// Nothing in qdk_chemistry_b200/ may link or load it.
//
// Each entry point is a thin call into the reference function named in its
// comment; determinants cross the boundary as uint64 words
//   nbits == 64 : one word / det, alpha = bits 0..31, beta = bits 32..63
//   nbits == 128: two words / det, word0 = alpha, word1 = beta
// (wfn_t<N> layout, external/macis/include/macis/wfn/raw_bitset.hpp:94-106).
#include <spdlog/sinks/null_sink.h>
#include <spdlog/sinks/stdout_color_sinks.h>
#include <spdlog/spdlog.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <macis/asci/grow.hpp>
#include <macis/asci/refine.hpp>
#include <macis/csr_hamiltonian.hpp>
#include <macis/hamiltonian_generator/double_loop.hpp>
#include <macis/hamiltonian_generator/dynamic_bit_masking.hpp>
#include <macis/hamiltonian_generator/residue_arrays.hpp>
#include <macis/hamiltonian_generator/sorted_double_loop.hpp>
#include <macis/mcscf/cas.hpp>
#include <macis/sd_operations.hpp>
#include <macis/solvers/davidson.hpp>
#include <macis/solvers/selected_ci_diag.hpp>
#include <memory>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local std::string g_err;

template <size_t N>
std::bitset<N> from_words(const uint64_t* w) {
  if constexpr (N == 64) {
    return std::bitset<64>(w[0]);
  } else {
    std::bitset<N> lo(w[0]), hi(w[1]);
    return (hi << (N / 2)) | lo;
  }
}
template <size_t N>
void to_words(const std::bitset<N>& b, uint64_t* w) {
  if constexpr (N == 64) {
    w[0] = b.to_ullong();
  } else {
    const std::bitset<N> mask(~uint64_t(0));
    w[0] = (b & mask).to_ullong();
    w[1] = ((b >> (N / 2)) & mask).to_ullong();
  }
}
template <size_t N>
std::vector<std::bitset<N>> dets_in(const uint64_t* w, int64_t n) {
  constexpr int W = N / 64;
  std::vector<std::bitset<N>> d(n);
  for (int64_t i = 0; i < n; ++i) d[i] = from_words<N>(w + i * W);
  return d;
}
template <size_t N>
void dets_out(const std::vector<std::bitset<N>>& d, uint64_t* w) {
  constexpr int W = N / 64;
  for (size_t i = 0; i < d.size(); ++i) to_words<N>(d[i], w + i * W);
}

void quiet_loggers(int verbose) {
  const char* names[] = {"davidson",    "ci_solver",  "h_build",    "h_build_inc",
                         "asci_search", "asci_grow",  "asci_refine"};
  for (auto n : names) {
    auto l = spdlog::get(n);
    if (!l) l = verbose ? spdlog::stdout_color_mt(n) : spdlog::null_logger_mt(n);
    l->set_level(verbose ? (verbose > 1 ? spdlog::level::trace : spdlog::level::info)
                         : spdlog::level::off);
  }
}

struct HamGenBase {
  int nbits;
  int norb;
  std::vector<double> T, V;
  virtual ~HamGenBase() = default;
};
template <size_t N>
struct HamGen : HamGenBase {
  using wfn = macis::wfn_t<N>;
  std::unique_ptr<macis::SortedDoubleLoopHamiltonianGenerator<wfn>> sdl;
  std::unique_ptr<macis::DoubleLoopHamiltonianGenerator<wfn>> dl;
  std::unique_ptr<macis::ResidueArrayHamiltonianGenerator<wfn>> ra;
  std::unique_ptr<macis::DynamicBitMaskHamiltonianGenerator<wfn>> dbm;
  HamGen(int n, const double* t, const double* v) {
    nbits = N;
    norb = n;
    T.assign(t, t + size_t(n) * n);
    V.assign(v, v + size_t(n) * n * n * n);
    macis::matrix_span<double> Ts(T.data(), n, n);
    macis::rank4_span<double> Vs(V.data(), n, n, n, n);
    sdl = std::make_unique<macis::SortedDoubleLoopHamiltonianGenerator<wfn>>(Ts, Vs);
    dl = std::make_unique<macis::DoubleLoopHamiltonianGenerator<wfn>>(Ts, Vs);
    ra = std::make_unique<macis::ResidueArrayHamiltonianGenerator<wfn>>(Ts, Vs);
    dbm = std::make_unique<macis::DynamicBitMaskHamiltonianGenerator<wfn>>(Ts, Vs);
  }
  macis::HamiltonianGenerator<wfn>& gen(int which) {
    if (which == 1) return *dl;
    if (which == 2) return *ra;   // hamiltonian_build_algorithm = "residue_arrays"
    if (which == 3) return *dbm;  // hamiltonian_build_algorithm = "dynamic_bit_masking"
    return *sdl;
  }
};

using csr_t = sparsexx::csr_matrix<double, int64_t>;

struct AsciOpts {  // mirrors macis::ASCISettings (determinant_search.hpp:95-199)
  int64_t ntdets_max, ntdets_min, ncdets_max;
  int32_t core_selection_strategy;  // 0 = fixed, 1 = percentage
  int32_t just_singles;
  double core_selection_threshold, h_el_tol, rv_prune_tol;
  int64_t pair_size_max;
  double grow_factor, min_grow_factor, growth_backoff_rate, growth_recovery_rate;
  int64_t max_refine_iter;
  double refine_energy_tol;
  int32_t warm_start_davidson, constraint_level;
  double min_warm_start_overlap, min_patch_overlap, grow_ci_residual_tolerance,
      taper_grow_factor;
  // MCSCFSettings part (mcscf.hpp:22-52)
  double ci_res_tol;
  int64_t ci_max_subspace;
  double ci_matel_tol;
  // natural-orbital rotation during growth (grow.hpp:163-215); rotates the generator's integrals
  int64_t grow_with_rot, rot_size_start;
};

macis::ASCISettings to_asci(const AsciOpts& o) {
  macis::ASCISettings s;
  s.ntdets_max = o.ntdets_max;
  s.ntdets_min = o.ntdets_min;
  s.ncdets_max = o.ncdets_max;
  s.core_selection_strategy = o.core_selection_strategy
                                  ? macis::CoreSelectionStrategy::Percentage
                                  : macis::CoreSelectionStrategy::Fixed;
  s.core_selection_threshold = o.core_selection_threshold;
  s.h_el_tol = o.h_el_tol;
  s.rv_prune_tol = o.rv_prune_tol;
  s.pair_size_max = o.pair_size_max;
  s.just_singles = o.just_singles;
  s.grow_factor = o.grow_factor;
  s.min_grow_factor = o.min_grow_factor;
  s.growth_backoff_rate = o.growth_backoff_rate;
  s.growth_recovery_rate = o.growth_recovery_rate;
  s.max_refine_iter = o.max_refine_iter;
  s.refine_energy_tol = o.refine_energy_tol;
  s.warm_start_davidson = o.warm_start_davidson;
  s.constraint_level = o.constraint_level;
  s.min_warm_start_overlap = o.min_warm_start_overlap;
  s.min_patch_overlap = o.min_patch_overlap;
  s.grow_ci_residual_tolerance = o.grow_ci_residual_tolerance;
  s.taper_grow_factor = o.taper_grow_factor;
  s.grow_with_rot = o.grow_with_rot != 0;
  s.rot_size_start = size_t(o.rot_size_start);
  return s;
}
macis::MCSCFSettings to_mcscf(const AsciOpts& o) {
  macis::MCSCFSettings m;
  m.ci_res_tol = o.ci_res_tol;
  m.ci_max_subspace = o.ci_max_subspace;
  m.ci_matel_tol = o.ci_matel_tol;
  return m;
}

struct AsciResult {
  double E;
  std::vector<uint64_t> words;
  std::vector<double> C;
  int64_t n;
};

template <size_t N>
int64_t hilbert_impl(int norb, int na, int nb, uint64_t* out, int64_t cap) {
  auto d = macis::generate_hilbert_space<macis::wfn_t<N>>(norb, na, nb);
  if ((int64_t)d.size() > cap) return -(int64_t)d.size();
  dets_out<N>(d, out);
  return d.size();
}

template <size_t N>
csr_t* hbuild_impl(HamGenBase* b, int which, const uint64_t* bra, int64_t nbra,
                   const uint64_t* ket, int64_t nket, double thresh) {
  auto* hg = static_cast<HamGen<N>*>(b);
  auto bd = dets_in<N>(bra, nbra);
  if (ket == nullptr) {
    return new csr_t(macis::make_csr_hamiltonian<int64_t>(bd.begin(), bd.end(),
                                                          hg->gen(which), thresh));
  }
  auto kd = dets_in<N>(ket, nket);
  return new csr_t(macis::make_csr_hamiltonian_block<int64_t>(
      bd.begin(), bd.end(), kd.begin(), kd.end(), hg->gen(which), thresh));
}

// form_rdms / form_rdms_spin_dep of the chosen generator, bra == ket
// (sorted_double_loop.hpp:512-760, double_loop.hpp); NULL outputs are skipped
template <size_t N>
void rdm_impl(HamGenBase* b, int which, const uint64_t* dets, int64_t n, const double* C, int spin_dep,
              double* o1, double* o2, double* t1, double* t2, double* t3) {
  auto* hg = static_cast<HamGen<N>*>(b);
  auto d = dets_in<N>(dets, n);
  std::vector<double> c(C, C + n);
  const size_t no = b->norb;
  auto ms = [&](double* p) { return macis::matrix_span<double>(p, no, no); };
  auto rs = [&](double* p) { return macis::rank4_span<double>(p, no, no, no, no); };
  if (spin_dep)
    hg->gen(which).form_rdms_spin_dep(d.begin(), d.end(), d.begin(), d.end(), c.data(), ms(o1), ms(o2),
                                      rs(t1), rs(t2), rs(t3));
  else
    hg->gen(which).form_rdms(d.begin(), d.end(), d.begin(), d.end(), c.data(), ms(o1), rs(t1));
}

// form_entropies of the chosen generator, bra == ket (sorted_double_loop.hpp:760-905)
template <size_t N>
void entropy_impl(HamGenBase* b, int which, const uint64_t* dets, int64_t n, const double* C, double* s1,
                  double* s2, double* mi) {
  auto* hg = static_cast<HamGen<N>*>(b);
  auto d = dets_in<N>(dets, n);
  std::vector<double> c(C, C + n);
  const size_t no = b->norb;
  std::vector<double> s1v(no, 0.0);
  hg->gen(which).form_entropies(d.begin(), d.end(), d.begin(), d.end(), c.data(), s1v,
                                macis::matrix_span<double>(s2, no, no), macis::matrix_span<double>(mi, no, no));
  std::copy(s1v.begin(), s1v.end(), s1);
}

template <size_t N>
double sci_diag_impl(HamGenBase* b, const uint64_t* dets, int64_t n, double h_el_tol,
                     int64_t max_m, double res_tol, double* C) {
  auto* hg = static_cast<HamGen<N>*>(b);
  auto d = dets_in<N>(dets, n);
  std::vector<double> c(C, C + n);
  double E = macis::selected_ci_diag<int64_t, macis::wfn_t<N>>(
      d.begin(), d.end(), hg->gen(0), h_el_tol, size_t(max_m), res_tol, c,
      (macis::CachedHamiltonianState<macis::wfn_t<N>, int64_t>*)nullptr, 0.3);
  std::copy(c.begin(), c.end(), C);
  return E;
}

template <size_t N>
int64_t asci_search_impl(HamGenBase* b, const AsciOpts& o, int64_t ndets_max,
                         const uint64_t* cdets, int64_t ncdets, double E0,
                         const double* C, uint64_t* out, int64_t cap) {
  auto* hg = static_cast<HamGen<N>*>(b);
  auto cd = dets_in<N>(cdets, ncdets);
  std::vector<double> c(C, C + ncdets);
  auto& g = hg->gen(0);
  auto nd = macis::asci_search<N>(to_asci(o), size_t(ndets_max), cd.begin(), cd.end(),
                                  E0, c, size_t(hg->norb), g.T(), g.G_red(),
                                  g.V_red(), g.V(), g);
  if ((int64_t)nd.size() > cap) return -(int64_t)nd.size();
  dets_out<N>(nd, out);
  return nd.size();
}

template <size_t N>
AsciResult* asci_run_impl(HamGenBase* b, const AsciOpts& o, int na, int nb,
                          int do_refine) {
  auto* hg = static_cast<HamGen<N>*>(b);
  auto& g = hg->gen(0);
  using wfn = macis::wfn_t<N>;
  std::vector<wfn> dets = {
      macis::wavefunction_traits<wfn>::canonical_hf_determinant(na, nb)};
  double E = g.matrix_element(dets[0], dets[0]);
  std::vector<double> C = {1.0};
  auto as = to_asci(o);
  auto ms = to_mcscf(o);
  std::tie(E, dets, C) = macis::asci_grow<N, int64_t>(as, ms, E, std::move(dets),
                                                      std::move(C), g, hg->norb);
  if (do_refine && as.max_refine_iter)
    std::tie(E, dets, C) = macis::asci_refine<N, int64_t>(
        as, ms, E, std::move(dets), std::move(C), g, hg->norb);
  auto* r = new AsciResult;
  r->E = E;
  r->n = dets.size();
  r->words.resize(dets.size() * (N / 64));
  dets_out<N>(dets, r->words.data());
  r->C = std::move(C);
  return r;
}

}  // namespace

#define TRY try {
#define CATCH(fail)                   \
  }                                   \
  catch (const std::exception& e) {   \
    g_err = e.what();                 \
    return fail;                      \
  }

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

int ref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void ref_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#endif
  (void)n;
}
void ref_set_verbose(int v) { quiet_loggers(v); }

// generate_hilbert_space  (sd_operations.hpp:333-351)
int64_t ref_generate_hilbert_space(int nbits, int norb, int na, int nb, uint64_t* out,
                                   int64_t cap) {
  TRY
  if (nbits == 64) return hilbert_impl<64>(norb, na, nb, out, cap);
  return hilbert_impl<128>(norb, na, nb, out, cap);
  CATCH(INT64_MIN)
}

// HamiltonianGeneratorBase ctor (src/macis/hamiltonian_generator/base.ipp:27-77)
void* ref_hamgen_create(int nbits, int norb, const double* T, const double* V) {
  quiet_loggers(0);
  TRY
  if (nbits == 64) return (void*)new HamGen<64>(norb, T, V);
  return (void*)new HamGen<128>(norb, T, V);
  CATCH(nullptr)
}
void ref_hamgen_destroy(void* h) { delete static_cast<HamGenBase*>(h); }

// intermediates: out G_red, V_red (n^3)   (base.ipp:50-61)
void ref_hamgen_intermediates(void* h, double* G_red, double* V_red) {
  auto* b = static_cast<HamGenBase*>(h);
  size_t n = b->norb;
  auto cp = [&](auto& g) {
    std::memcpy(G_red, g.G_red(), n * n * n * 8);
    std::memcpy(V_red, g.V_red(), n * n * n * 8);
  };
  if (b->nbits == 64)
    cp(static_cast<HamGen<64>*>(b)->gen(0));
  else
    cp(static_cast<HamGen<128>*>(b)->gen(0));
}

// HamiltonianGenerator::matrix_element (hamiltonian_generator/matrix_elements.hpp:27-97)
double ref_matrix_element(void* h, const uint64_t* bra, const uint64_t* ket) {
  auto* b = static_cast<HamGenBase*>(h);
  if (b->nbits == 64)
    return static_cast<HamGen<64>*>(b)->gen(0).matrix_element(from_words<64>(bra),
                                                              from_words<64>(ket));
  return static_cast<HamGen<128>*>(b)->gen(0).matrix_element(from_words<128>(bra),
                                                             from_words<128>(ket));
}

// make_csr_hamiltonian / make_csr_hamiltonian_block (csr_hamiltonian.hpp:41-80)
// which: 0 = SortedDoubleLoop (sorted_double_loop.hpp:86-451), 1 = DoubleLoop
// (double_loop.hpp:64-132). ket == NULL -> symmetric (bra == ket) build.
void* ref_hbuild(void* h, int which, const uint64_t* bra, int64_t nbra,
                 const uint64_t* ket, int64_t nket, double thresh, double* seconds) {
  auto* b = static_cast<HamGenBase*>(h);
  quiet_loggers(0);
  TRY
  auto t0 = std::chrono::high_resolution_clock::now();
  csr_t* m = b->nbits == 64 ? hbuild_impl<64>(b, which, bra, nbra, ket, nket, thresh)
                            : hbuild_impl<128>(b, which, bra, nbra, ket, nket, thresh);
  auto t1 = std::chrono::high_resolution_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return (void*)m;
  CATCH(nullptr)
}
int ref_form_rdms(void* h, int which, const uint64_t* dets, int64_t n, const double* C, int spin_dep,
                  double* o1, double* o2, double* t1, double* t2, double* t3) {
  auto* b = static_cast<HamGenBase*>(h);
  quiet_loggers(0);
  TRY
  if (b->nbits == 64) rdm_impl<64>(b, which, dets, n, C, spin_dep, o1, o2, t1, t2, t3);
  else rdm_impl<128>(b, which, dets, n, C, spin_dep, o1, o2, t1, t2, t3);
  return 0;
  CATCH(1)
}
int ref_form_entropies(void* h, int which, const uint64_t* dets, int64_t n, const double* C, double* s1,
                       double* s2, double* mi) {
  auto* b = static_cast<HamGenBase*>(h);
  quiet_loggers(0);
  TRY
  if (b->nbits == 64) entropy_impl<64>(b, which, dets, n, C, s1, s2, mi);
  else entropy_impl<128>(b, which, dets, n, C, s1, s2, mi);
  return 0;
  CATCH(1)
}
// wrap caller arrays as a reference csr_matrix
// (python/src/pybind11/algorithms/davidson_solver.cpp:60-80)
void* ref_csr_from_arrays(int64_t n, int64_t nnz, const int64_t* rowptr,
                          const int64_t* colind, const double* nzval) {
  TRY
  std::vector<int64_t> rp(rowptr, rowptr + n + 1);
  std::vector<int64_t> ci(colind, colind + nnz);
  std::vector<double> nz(nzval, nzval + nnz);
  return (void*)new csr_t(n, n, std::move(rp), std::move(ci), std::move(nz));
  CATCH(nullptr)
}
int64_t ref_csr_nrows(void* m) { return static_cast<csr_t*>(m)->m(); }
int64_t ref_csr_nnz(void* m) { return static_cast<csr_t*>(m)->nnz(); }
void ref_csr_copy(void* m, int64_t* rowptr, int64_t* colind, double* nzval) {
  auto* c = static_cast<csr_t*>(m);
  std::copy(c->rowptr().begin(), c->rowptr().end(), rowptr);
  std::copy(c->colind().begin(), c->colind().end(), colind);
  std::copy(c->nzval().begin(), c->nzval().end(), nzval);
}
// rows [r0, r1) of a reference csr_matrix: local row pointer (r1 - r0 + 1 entries starting at
// 0), column indices and values -- lets a caller fingerprint a matrix too large to copy whole
void ref_csr_copy_rows(void* m, int64_t r0, int64_t r1, int64_t* rowptr, int64_t* colind,
                       double* nzval) {
  auto* c = static_cast<csr_t*>(m);
  const auto& rp = c->rowptr();
  const int64_t b = rp[r0], e = rp[r1];
  if (rowptr)
    for (int64_t i = r0; i <= r1; ++i) rowptr[i - r0] = rp[i] - b;
  if (colind) std::copy(c->colind().begin() + b, c->colind().begin() + e, colind);
  if (nzval) std::copy(c->nzval().begin() + b, c->nzval().begin() + e, nzval);
}
void ref_csr_free(void* m) { delete static_cast<csr_t*>(m); }

// sparsexx::spblas::gespmbv, K = 1 (spblas/spmbv.hpp:49-85); returns mean seconds
double ref_spmv(void* m, const double* x, double* y, int nrep) {
  auto* c = static_cast<csr_t*>(m);
  const int64_t n = c->m();
  if (nrep < 1) nrep = 1;
  auto t0 = std::chrono::high_resolution_clock::now();
  for (int r = 0; r < nrep; ++r) sparsexx::spblas::gespmbv(1, 1., *c, x, n, 0., y, n);
  auto t1 = std::chrono::high_resolution_clock::now();
  return std::chrono::duration<double>(t1 - t0).count() / nrep;
}
// extract_diagonal_elements (sparsexx/util/submatrix.hpp:354-383)
void ref_csr_diagonal(void* m, double* D) {
  auto* c = static_cast<csr_t*>(m);
  auto d = sparsexx::extract_diagonal_elements(*c);
  std::copy(d.begin(), d.end(), D);
}

// use_guess_policy = 1: serial_selected_ci_diag (selected_ci_diag.hpp:111-158)
// use_guess_policy = 0: diagonal + davidson directly (davidson.hpp:259-372), X in/out.
// Returns 0 ok, 1 = threw (e.g. "Davidson Did Not Converge!").
int ref_davidson(void* m, int64_t max_m, double tol, double* X, int64_t* niter,
                 double* eig, int use_guess_policy) {
  auto* c = static_cast<csr_t*>(m);
  quiet_loggers(0);
  TRY
  const int64_t n = c->m();
  if (use_guess_policy) {
    std::vector<double> C(X, X + n);
    double E = macis::serial_selected_ci_diag(*c, size_t(max_m), tol, C);
    std::copy(C.begin(), C.end(), X);
    *eig = E;
    *niter = -1;
  } else {
    auto D = sparsexx::extract_diagonal_elements(*c);
    macis::SparseMatrixOperator op(*c);
    auto res = macis::davidson(n, max_m, op, D.data(), tol, X);
    *niter = res.first;
    *eig = res.second;
  }
  return 0;
  CATCH(1)
}

// selected_ci_diag<int64_t> full-build path (selected_ci_diag.hpp:174-311)
int ref_selected_ci_diag(void* h, const uint64_t* dets, int64_t n, double h_el_tol,
                         int64_t max_m, double res_tol, double* C, double* E) {
  auto* b = static_cast<HamGenBase*>(h);
  quiet_loggers(0);
  TRY
  *E = b->nbits == 64 ? sci_diag_impl<64>(b, dets, n, h_el_tol, max_m, res_tol, C)
                      : sci_diag_impl<128>(b, dets, n, h_el_tol, max_m, res_tol, C);
  return 0;
  CATCH(1)
}

// asci_search (asci/determinant_search.hpp:808-1123). C holds ncdets coefficients.
int64_t ref_asci_search(void* h, const void* opts, int64_t ndets_max,
                        const uint64_t* cdets, int64_t ncdets, double E0,
                        const double* C, uint64_t* out, int64_t cap) {
  auto* b = static_cast<HamGenBase*>(h);
  quiet_loggers(0);
  const AsciOpts& o = *static_cast<const AsciOpts*>(opts);
  TRY
  if (b->nbits == 64)
    return asci_search_impl<64>(b, o, ndets_max, cdets, ncdets, E0, C, out, cap);
  return asci_search_impl<128>(b, o, ndets_max, cdets, ncdets, E0, C, out, cap);
  CATCH(INT64_MIN)
}

// HF -> asci_grow -> asci_refine, as
// cpp/src/qdk/chemistry/algorithms/microsoft/macis_asci.cpp:160-196
void* ref_asci_run(void* h, const void* opts, int na, int nb, int do_refine) {
  auto* b = static_cast<HamGenBase*>(h);
  quiet_loggers(0);
  const AsciOpts& o = *static_cast<const AsciOpts*>(opts);
  TRY
  if (b->nbits == 64) return (void*)asci_run_impl<64>(b, o, na, nb, do_refine);
  return (void*)asci_run_impl<128>(b, o, na, nb, do_refine);
  CATCH(nullptr)
}
int64_t ref_asci_result_n(void* r) { return static_cast<AsciResult*>(r)->n; }
double ref_asci_result_energy(void* r) { return static_cast<AsciResult*>(r)->E; }
void ref_asci_result_copy(void* r, uint64_t* dets, double* C) {
  auto* a = static_cast<AsciResult*>(r);
  std::copy(a->words.begin(), a->words.end(), dets);
  std::copy(a->C.begin(), a->C.end(), C);
}
void ref_asci_result_free(void* r) { delete static_cast<AsciResult*>(r); }

}  // extern "C"
