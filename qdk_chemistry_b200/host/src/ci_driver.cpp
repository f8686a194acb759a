// Host side of the B200 CI path: the MultiConfigurationCalculator plugins and the orchestration
// that MACIS keeps on the host (guess policies, the ASCI grow / refine loops), driving the CUDA
// library through the C ABI of include/b2ci.h and nothing else. No arithmetic of the hot path
// happens here: Hamiltonian build, sigma, Davidson, dense diagonalisation and the ASCI search
// are b2ci_* calls; this file sorts determinant lists, picks core sets and keeps settings.
//
// Reference call stacks mirrored:
//   B200Cas::_run_impl   cas_helper::impl      macis_cas.cpp:35-148, mcscf/cas.hpp:33-64
//   B200Asci::_run_impl  asci_helper::impl     macis_asci.cpp:44-234
//   asci_iter/grow/refine                      asci/iteration.hpp:50-226, grow.hpp:45-268, refine.hpp:44-237
//   B200Pmc::_run_impl   pmc_helper::impl      macis_pmc.cpp:36-174
//   selected_ci_diag                           solvers/selected_ci_diag.hpp:111-311
#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <numeric>
#include <unordered_map>

#include "b2ci.h"
#include "qdk_b200/mc.hpp"

namespace qdk_b200::algorithms {

std::pair<int64_t, int64_t> row_block(int64_t n, int rank, int nranks);

namespace {

struct Det {
  uint64_t a, b;  // word 0 = alpha, word 1 = beta: the wfn_t<128> layout of the C ABI
  bool operator==(const Det& o) const { return a == o.a && b == o.b; }
};
static_assert(sizeof(Det) == 16, "Det must be two packed words");
struct DetHash {
  size_t operator()(const Det& d) const { return std::hash<uint64_t>()(d.a * 0x9e3779b97f4a7c15ull ^ (d.b + (d.a << 7))); }
};
// spin_comparator: alpha-major, then beta (raw_bitset.hpp:119-141)
inline bool spin_less(const Det& x, const Det& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; }

// One CUDA context (device + stream + optional NCCL communicator) per process, shared by all
// calculator runs: an NCCL unique id can initialise a communicator only once, so the
// communicator is created in set_communicator() and lives until clear_communicator().
struct Runtime {
  std::mutex mutex;  // run() is single-caller in the reference; concurrent runs serialise here
  b2ci_ctx* ctx = nullptr;
  int device = 0;
  int rank = 0, nranks = 1;
  void drop() {
    if (ctx) b2ci_ctx_destroy(ctx);
    ctx = nullptr;
  }
};
Runtime& runtime() {
  // never destroyed: at process exit the CUDA runtime may already be gone, and tearing a
  // context down then can block
  static Runtime* r = new Runtime;
  return *r;
}
thread_local std::map<std::string, double> g_stats;

// Phase log with the reference's logger names (spdlog loggers h_build / ci_solver / davidson / asci_search /
// asci_grow / asci_refine: selected_ci_diag.hpp:204-296, sorted_double_loop.hpp:441-448, davidson.hpp:340-344), so
// that a CPU log of MACIS and a log of this plugin line up. B2CI_LOG=info|trace (stderr); off by default.
int log_level() {
  static const int lvl = [] {
    const char* e = std::getenv("B2CI_LOG");
    if (!e) return 0;
    const std::string v(e);
    return v == "trace" ? 2 : (v == "off" || v == "0" ? 0 : 1);
  }();
  return lvl;
}
void log_line(int level, const char* logger, const char* fmt, ...) {
  if (log_level() < level) return;
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  std::fprintf(stderr, "[%s] %s\n", logger, buf);
}

[[noreturn]] void fail(const std::string& what) {
  throw std::runtime_error(what + ": " + b2ci_last_error());
}
#define B2(call)                    \
  do {                              \
    if ((call) != 0) fail(#call);   \
  } while (0)

struct McscfSettings {  // macis::MCSCFSettings fields that reach the path
  double ci_res_tol, ci_matel_tol;
  int64_t ci_max_subspace;
};
McscfSettings get_mcscf_settings(const data::Settings& s) {  // macis_base.cpp:35-45
  McscfSettings m;
  m.ci_res_tol = s.get<double>("ci_residual_tolerance");
  m.ci_max_subspace = s.get<int64_t>("max_solver_iterations");
  m.ci_matel_tol = s.get<double>("ci_matel_tol");
  return m;
}
struct AsciSettings {  // macis::ASCISettings fields that reach the path
  int64_t ntdets_max, ntdets_min, ncdets_max, max_refine_iter;
  double h_el_tol, rv_prune_tol, grow_factor, min_grow_factor, growth_backoff_rate, growth_recovery_rate,
      refine_energy_tol, core_selection_threshold, min_warm_start_overlap, grow_ci_residual_tolerance,
      taper_grow_factor, min_patch_overlap;
  bool just_singles, warm_start_davidson, fixed_core, grow_with_rot;
  int64_t rot_size_start;
  int generator;  // B2CI_GEN_*: hamiltonian_build_algorithm
};
AsciSettings get_asci_settings(const data::Settings& s) {  // macis_base.cpp:47-134
  AsciSettings a;
  a.ntdets_max = s.get<int64_t>("ntdets_max");
  a.ntdets_min = s.get<int64_t>("ntdets_min");
  a.ncdets_max = s.get<int64_t>("ncdets_max");
  a.h_el_tol = s.get<double>("search_matel_tol");
  a.rv_prune_tol = s.get<double>("rv_prune_tol");
  a.just_singles = s.get<bool>("just_singles");
  a.grow_factor = s.get<double>("grow_factor");
  a.min_grow_factor = s.get<double>("min_grow_factor");
  a.growth_backoff_rate = s.get<double>("growth_backoff_rate");
  a.growth_recovery_rate = s.get<double>("growth_recovery_rate");
  a.max_refine_iter = s.get<int64_t>("max_refine_iter");
  a.refine_energy_tol = s.get<double>("refine_energy_tol");
  const std::string strat = s.get<std::string>("core_selection_strategy");
  if (strat == "fixed") a.fixed_core = true;
  else if (strat == "percentage") a.fixed_core = false;
  else
    throw std::invalid_argument("Invalid core_selection_strategy: '" + strat +
                                "'. Valid options are 'fixed' or 'percentage'.");
  a.core_selection_threshold = s.get<double>("core_selection_threshold");
  a.warm_start_davidson = s.get<bool>("warm_start_davidson");
  a.min_warm_start_overlap = s.get<double>("min_warm_start_overlap");
  a.grow_ci_residual_tolerance = s.get<double>("grow_ci_residual_tolerance");
  a.taper_grow_factor = s.get<double>("taper_grow_factor");
  a.min_patch_overlap = s.get<double>("min_patch_overlap");
  a.grow_with_rot = s.get<bool>("grow_with_rot");
  a.rot_size_start = s.get<int64_t>("rot_size_start");
  if (a.grow_factor <= 1.0) throw std::runtime_error("grow_factor must be > 1.0, got " + std::to_string(a.grow_factor));
  if (a.min_grow_factor <= 1.0)
    throw std::runtime_error("min_grow_factor must be > 1.0, got " + std::to_string(a.min_grow_factor));
  if (a.min_grow_factor > a.grow_factor) throw std::runtime_error("min_grow_factor must be <= grow_factor");
  if (a.growth_backoff_rate <= 0.0 || a.growth_backoff_rate >= 1.0)
    throw std::runtime_error("growth_backoff_rate must be in (0, 1), got " + std::to_string(a.growth_backoff_rate));
  if (a.growth_recovery_rate <= 1.0)
    throw std::runtime_error("growth_recovery_rate must be > 1.0, got " + std::to_string(a.growth_recovery_rate));
  if (!a.fixed_core && (a.core_selection_threshold < std::numeric_limits<double>::epsilon() ||
                        a.core_selection_threshold > 1.0))
    throw std::invalid_argument("core_selection_threshold must be in [epsilon, 1.0], got " +
                                std::to_string(a.core_selection_threshold));
  // macis_asci.cpp:92-118: anything that is not one of the two pair-based generators is the
  // sorted double loop. On the device the three share one enumeration; what the choice selects is
  // the generator's pattern rules (b2ci_set_hamiltonian_generator).
  const std::string algo = s.get<std::string>("hamiltonian_build_algorithm");
  a.generator = algo == "residue_arrays" ? B2CI_GEN_RESIDUE_ARRAYS
                : algo == "dynamic_bit_masking" ? B2CI_GEN_DYNAMIC_BIT_MASKING : B2CI_GEN_SORTED_DOUBLE_LOOP;
  return a;
}

int64_t binomial(int64_t n, int64_t k) {
  if (k < 0 || k > n) return 0;
  k = std::min(k, n - k);
  __int128 r = 1;
  for (int64_t i = 1; i <= k; ++i) {
    r = r * (n - k + i) / i;
    if (r > (__int128)std::numeric_limits<int64_t>::max()) return std::numeric_limits<int64_t>::max();
  }
  return (int64_t)r;
}

// ---------------------------------------------------------------------------------------------
class CiSession {
 public:
  explicit CiSession(const data::Hamiltonian& h)
      : norb_(int(h.num_active_orbitals())), ham_(h), lock_(runtime().mutex) {
    Runtime& rt = runtime();
    if (norb_ > 64) throw std::runtime_error("active spaces with more than 64 orbitals are not built");
    if (!rt.ctx) B2(b2ci_ctx_create(rt.device, nullptr, &rt.ctx));
    ctx_ = rt.ctx;
    rank_ = rt.rank;
    nranks_ = rt.nranks;
    // T is symmetric; the two-body array is handed over as MACIS reinterprets it
    // (macis_cas.cpp:76-82): element (pq|rs) of the QDK layout p n^3 + q n^2 + r n + s is
    // read as column-major V(s,r,q,p) = (sr|qp), equal by symmetry.
    B2(b2ci_integrals_upload(ctx_, norb_, h.get_one_body_integrals().data(), h.get_two_body_integrals().data()));
  }
  CiSession(const CiSession&) = delete;

  int norb() const { return norb_; }
  b2ci_ctx* ctx() const { return ctx_; }

  double diagonal_element(const Det& d) const {  // ham_gen.matrix_element(d, d)
    return b2ci_host_matrix_element(norb_, ham_.get_one_body_integrals().data(),
                                    ham_.get_two_body_integrals().data(), d.a, d.b, d.a, d.b);
  }

  std::pair<int64_t, int64_t> my_rows(int64_t n) const { return row_block(n, rank_, nranks_); }

  // selected_ci_diag on a device-resident list: H build of this rank's rows + Davidson with
  // the guess policy of serial_selected_ci_diag. X: empty or a guess; returns the full vector.
  //
  // cache (ASCI iterations, single rank): CachedHamiltonianState of incremental_h_build.hpp. When
  // the previous list overlaps the new one by min_patch_overlap or more, only the blocks that touch
  // added determinants are evaluated (b2ci_hbuild_csr_patched); the matrix and the device list of
  // this call are then kept for the next one. take_dets: the session owns `dets` afterwards.
  double selected_ci_diag(b2ci_dets* dets, int64_t n, double matel_tol, int64_t max_m, double res_tol,
                          std::vector<double>& X, bool use_cache = false, double min_patch_overlap = 0.3,
                          bool take_dets = false, bool balance_rows = false) {
    if (n == 0) throw std::runtime_error("selected_ci_diag: empty determinant list");
    X.resize(size_t(n), 0.0);
    // row blocks of the ranks: even row counts for full-CI lists (uniform rows); for selected-CI lists cuts that
    // balance the estimated connections -- the spin-sorted head of an ASCI list holds the determinants with the
    // most partners, and with even rows the first rank built and multiplied twice the mean (Cr2 1e7 on 8 GPUs:
    // 688 ms against ~300 on the others). B2CI_EQUAL_ROWS=1 keeps the even split.
    std::vector<int64_t> off(size_t(nranks_) + 1, 0);
    for (int r = 0; r < nranks_; ++r) off[size_t(r) + 1] = row_block(n, r, nranks_).second;
    if (nranks_ > 1 && balance_rows && !getenv("B2CI_EQUAL_ROWS")) {
      B2(b2ci_dets_balanced_partition(ctx_, dets, nranks_, 1024, off.data()));
      g_stats["row_partition_max_over_mean"] = b2ci_timer_ms(ctx_, "h_build.partition_max_over_mean_rows");
    }
    const std::pair<int64_t, int64_t> rows{off[size_t(rank_)], off[size_t(rank_) + 1]};
    b2ci_csr* H = nullptr;
    struct DetsGuard {  // frees a taken-over list unless the cache adopts it
      b2ci_ctx* c;
      b2ci_dets* d;
      ~DetsGuard() { if (d) b2ci_dets_free(c, d); }
    } guard{ctx_, take_dets ? dets : nullptr};
    use_cache = use_cache && nranks_ == 1 && take_dets && !getenv("B2CI_NO_INCREMENTAL");
    if (use_cache && cache_H_ && cache_tol_ == matel_tol) {
      int64_t n_kept = 0;
      // a patch that cannot be built (three matrices do not fit the device) is not an error:
      // the cache is dropped and the full build runs with the memory it held
      if (b2ci_hbuild_csr_patched(ctx_, cache_dets_, cache_H_, dets, matel_tol, min_patch_overlap, &H, &n_kept) != 0) {
        H = nullptr;
        g_stats["h_build_patch_failed"] += 1.0;
      }
      g_stats["h_build_patch_last_overlap"] = double(n_kept) / double(n);
      if (H) g_stats["h_build_patched"] += 1.0;
    }
    if (!H) {
      drop_cache();
      B2(b2ci_hbuild_csr(ctx_, dets, rows.first, rows.second, matel_tol, &H));
    }
    if (nranks_ > 1)  // every rank knows the split: no exchange of block sizes
      B2(b2ci_csr_set_row_partition(ctx_, H, off.data(), nranks_));
    add_timer("h_build_ms", {"h_build.setup", "h_build.count", "h_build.fill", "h_build.thresh"});
    add_timer("h_build_setup_ms", {"h_build.setup"});
    add_timer("h_build_count_ms", {"h_build.count"});
    add_timer("h_build_fill_ms", {"h_build.fill"});
    g_stats["h_build_last_ms"] = b2ci_timer_ms(ctx_, "h_build.setup") + b2ci_timer_ms(ctx_, "h_build.count") +
                                 b2ci_timer_ms(ctx_, "h_build.fill") + b2ci_timer_ms(ctx_, "h_build.thresh");
    int64_t nnz = 0;
    b2ci_csr_info(H, nullptr, nullptr, &nnz, nullptr);
    g_stats["nnz_local"] = double(nnz);
    int64_t niter = 0;
    double E = 0.;
    const int rc = b2ci_davidson(ctx_, H, max_m, res_tol, X.data(), 1, &niter, &E, nullptr);
    const std::string err = rc ? b2ci_last_error() : "";
    add_timer("davidson_sigma_ms", {"davidson.OP_DUR"});
    add_timer("davidson_other_ms", {"davidson.RR_DUR", "davidson.RES_DUR", "davidson.GS_DUR"});
    g_stats["davidson_iterations"] += double(niter);
    g_stats["davidson_calls"] += 1.0;
    log_line(1, "h_build", "SETUP_DUR = %.5e ms, COUNT_DUR = %.5e ms, FILL_DUR = %.5e ms, THRESH_DUR = %.5e ms",
             b2ci_timer_ms(ctx_, "h_build.setup"), b2ci_timer_ms(ctx_, "h_build.count"), b2ci_timer_ms(ctx_, "h_build.fill"),
             b2ci_timer_ms(ctx_, "h_build.thresh"));
    log_line(1, "ci_solver", "NDETS = %lld, NNZ = %lld, H_DUR = %.5e ms, HMEM_LOC = %.2e GiB, H_SPARSE = %.2e %%",
             (long long)n, (long long)nnz, g_stats["h_build_last_ms"], double(nnz) * 12.0 / 1073741824.0,
             100.0 * double(nnz) / (double(rows.second - rows.first) * double(n)));
    log_line(1, "ci_solver", "DAV_NITER = %lld, E0 = %.12e, DAVIDSON_DUR = %.5e ms", (long long)niter, E,
             b2ci_timer_ms(ctx_, "davidson.OP_DUR") + b2ci_timer_ms(ctx_, "davidson.RR_DUR") +
                 b2ci_timer_ms(ctx_, "davidson.RES_DUR") + b2ci_timer_ms(ctx_, "davidson.GS_DUR"));
    log_line(2, "davidson", "OP_DUR = %.5e ms, RR_DUR = %.5e ms, RES_DUR = %.5e ms, GS_DUR = %.5e ms",
             b2ci_timer_ms(ctx_, "davidson.OP_DUR"), b2ci_timer_ms(ctx_, "davidson.RR_DUR"),
             b2ci_timer_ms(ctx_, "davidson.RES_DUR"), b2ci_timer_ms(ctx_, "davidson.GS_DUR"));
    {
      // sigma of the LAST matrix: time per application and its algorithmic bytes (SURVEY 8(d):
      // nnz * 12 + (rows + 1) * 8 + N * 8 + rows * 8), so that callers can quote a roofline fraction
      const double calls = b2ci_timer_ms(ctx_, "davidson.OP_CALLS");
      const int64_t nrows = rows.second - rows.first;
      g_stats["sigma_last_ms"] = calls > 0 ? b2ci_timer_ms(ctx_, "davidson.OP_DUR") / calls : 0.0;
      g_stats["sigma_last_bytes"] = double(nnz) * 12.0 + double(nrows + 1) * 8.0 + double(n) * 8.0 + double(nrows) * 8.0;
      g_stats["sigma_last_rows"] = double(nrows);
    }
    if (use_cache) {
      drop_cache();
      cache_H_ = H;
      cache_dets_ = dets;
      cache_tol_ = matel_tol;
      guard.d = nullptr;
    } else {
      b2ci_csr_free(ctx_, H);
    }
    if (rc != 0) throw std::runtime_error(err);
    return E;
  }
  double selected_ci_diag(const std::vector<Det>& dets, double matel_tol, int64_t max_m, double res_tol,
                          std::vector<double>& X, bool use_cache = false, double min_patch_overlap = 0.3) {
    b2ci_dets* d = nullptr;
    B2(b2ci_dets_upload(ctx_, reinterpret_cast<const uint64_t*>(dets.data()), 2, int64_t(dets.size()), &d));
    // the list is handed over: freed there, or adopted by the cache
    // (host lists are the selected-CI lists of the ASCI loop: connection-balanced row blocks)
    return selected_ci_diag(d, int64_t(dets.size()), matel_tol, max_m, res_tol, X, use_cache, min_patch_overlap, true, true);
  }
  // natural-orbital step of asci_grow (grow.hpp:163-215): spin-traced 1-RDM of the current
  // wavefunction, eigenvectors of -ordm (occupations descending), integrals rotated on the device
  void rotate_to_natural_orbitals(const std::vector<Det>& wfn, const std::vector<double>& X) {
    const size_t n = size_t(norb_);
    std::vector<double> ordm(n * n, 0.0), none, occ(n);
    form_rdms(wfn, X, false, ordm, none, none, none, none);
    for (double& x : ordm) x *= -1.0;
    if (b2ci_host_sym_eig_lower(int(n), ordm.data(), int(n), occ.data()) != 0) fail("b2ci_host_sym_eig_lower");
    double on_sum = 0.;
    for (double o : occ) on_sum -= o;
    B2(b2ci_integrals_rotate(ctx_, ordm.data(), nullptr, nullptr));
    drop_cache();  // the rotation invalidates the cached matrix (grow.hpp:218-219)
    g_stats["natural_orbital_rotations"] += 1.0;
    g_stats["natural_occupation_sum"] = on_sum;
  }
  void set_generator(int g) { B2(b2ci_set_hamiltonian_generator(ctx_, g)); }
  void drop_cache() {  // CachedHamiltonianState::clear
    if (cache_H_) b2ci_csr_free(ctx_, cache_H_);
    if (cache_dets_) b2ci_dets_free(ctx_, cache_dets_);
    cache_H_ = nullptr;
    cache_dets_ = nullptr;
  }
  ~CiSession() {
    drop_cache();
    b2ci_set_hamiltonian_generator(ctx_, B2CI_GEN_SORTED_DOUBLE_LOOP);  // the context is shared by all runs
  }
  // dense branch: full CSR (index pattern as the reference's int32 build) -> lowest eigenpair
  double dense_diag(const b2ci_dets* dets, int64_t n, double matel_tol, std::vector<double>& X) {
    X.assign(size_t(n), 0.0);
    b2ci_csr* H = nullptr;
    B2(b2ci_hbuild_csr(ctx_, dets, 0, n, matel_tol, &H));
    double E = 0.;
    const int rc = b2ci_dense_ground_state(ctx_, H, &E, X.data());
    const std::string err = rc ? b2ci_last_error() : "";
    b2ci_csr_free(ctx_, H);
    if (rc != 0) throw std::runtime_error(err);
    return E;
  }
  double dense_diag(const std::vector<Det>& dets, double matel_tol, std::vector<double>& X) {
    b2ci_dets* d = nullptr;
    B2(b2ci_dets_upload(ctx_, reinterpret_cast<const uint64_t*>(dets.data()), 2, int64_t(dets.size()), &d));
    try {
      const double E = dense_diag(d, int64_t(dets.size()), matel_tol, X);
      b2ci_dets_free(ctx_, d);
      return E;
    } catch (...) {
      b2ci_dets_free(ctx_, d);
      throw;
    }
  }

  std::vector<Det> asci_search(const Det* core, const double* coeffs, int64_t ncore, double E0, int64_t ndets_max,
                               const AsciSettings& a) {
    b2ci_asci_search_opts o;
    o.ndets_max = ndets_max;
    o.h_el_tol = a.h_el_tol;
    o.rv_prune_tol = a.rv_prune_tol;
    o.just_singles = a.just_singles ? 1 : 0;
    o.sort_output = 1;  // spin_comparator order, made on the device
    int64_t cap = ndets_max + ncore + 4096, n_out = 0;
    std::vector<Det> out;
    double stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (;;) {
      out.resize(size_t(cap));
      const int rc = b2ci_asci_search(ctx_, &o, reinterpret_cast<const uint64_t*>(core), 2, coeffs, ncore, E0,
                                      reinterpret_cast<uint64_t*>(out.data()), cap, &n_out, stats);
      if (rc == 4 && n_out > cap) {  // ties at the cut exceed the capacity: retry with the exact size
        cap = n_out;
        continue;
      }
      if (rc != 0) fail("b2ci_asci_search");
      break;
    }
    add_timer("asci_search_ms", {"asci_search.PAIR_DUR", "asci_search.SORT_ACC_DUR", "asci_search.TOPK_DUR"});
    add_timer("asci_pair_ms", {"asci_search.PAIR_DUR"});
    add_timer("asci_sort_acc_ms", {"asci_search.SORT_ACC_DUR"});
    add_timer("asci_topk_ms", {"asci_search.TOPK_DUR"});
    g_stats["asci_last_ncore"] = double(ncore);
    g_stats["asci_contributions"] += stats[0];       // (determinant, c*h) records generated, all searches
    g_stats["asci_unique_candidates"] += stats[1];   // after sort + accumulate
    g_stats["asci_last_key_partitions"] = stats[5];
    g_stats["asci_search_calls"] += 1.0;
    log_line(1, "asci_search", "NCDETS = %lld, NDETS_MAX = %lld, NCONTRIB = %.0f, NUNIQ = %.0f, NKEEP = %lld, PAIR_DUR = %.5e ms, "
             "SORT_ACC_DUR = %.5e ms, TOPK_DUR = %.5e ms", (long long)ncore, (long long)ndets_max, stats[0], stats[1],
             (long long)n_out, b2ci_timer_ms(ctx_, "asci_search.PAIR_DUR"), b2ci_timer_ms(ctx_, "asci_search.SORT_ACC_DUR"),
             b2ci_timer_ms(ctx_, "asci_search.TOPK_DUR"));
    out.resize(size_t(n_out));
    return out;
  }

  // HamiltonianGenerator::form_entropies on the solver-ordered list (macis_base.hpp:219-245)
  void form_entropies(const std::vector<Det>& dets, const std::vector<double>& C, std::vector<double>& s1,
                      std::vector<double>& s2, std::vector<double>& mi) {
    b2ci_dets* d = nullptr;
    B2(b2ci_dets_upload(ctx_, reinterpret_cast<const uint64_t*>(dets.data()), 2, int64_t(dets.size()), &d));
    const int rc = b2ci_form_entropies(ctx_, d, C.data(), s1.data(), s2.empty() ? nullptr : s2.data(),
                                       mi.empty() ? nullptr : mi.data());
    b2ci_dets_free(ctx_, d);
    if (rc != 0) throw std::runtime_error(b2ci_last_error());
    g_stats["entropy_pattern_ms"] = b2ci_timer_ms(ctx_, "entropy.pattern");
    g_stats["entropy_scatter_ms"] = b2ci_timer_ms(ctx_, "entropy.scatter");
  }
  // HamiltonianGenerator::form_rdms_spin_dep / form_rdms on the solver-ordered list
  // (macis_base.hpp:166-215, macis_pmc.cpp:128-160); outputs sized by the caller, empty = skip
  void form_rdms(const std::vector<Det>& dets, const std::vector<double>& C, bool spin_dep, std::vector<double>& o1,
                 std::vector<double>& o2, std::vector<double>& t1, std::vector<double>& t2, std::vector<double>& t3) {
    b2ci_dets* d = nullptr;
    B2(b2ci_dets_upload(ctx_, reinterpret_cast<const uint64_t*>(dets.data()), 2, int64_t(dets.size()), &d));
    auto p = [](std::vector<double>& v) { return v.empty() ? nullptr : v.data(); };
    const int rc = spin_dep ? b2ci_form_rdms_spin_dep(ctx_, d, C.data(), p(o1), p(o2), p(t1), p(t2), p(t3))
                            : b2ci_form_rdms(ctx_, d, C.data(), p(o1), p(t1));
    b2ci_dets_free(ctx_, d);
    if (rc != 0) throw std::runtime_error(b2ci_last_error());
    g_stats["rdm_pattern_ms"] = b2ci_timer_ms(ctx_, "rdm.pattern");
    g_stats["rdm_scatter_ms"] = b2ci_timer_ms(ctx_, "rdm.scatter");
  }

 private:
  void add_timer(const char* key, std::initializer_list<const char*> names) {
    double t = 0.;
    for (const char* nm : names) {
      const double v = b2ci_timer_ms(ctx_, nm);
      if (v > 0) t += v;
    }
    g_stats[key] += t;
  }
  int norb_;
  const data::Hamiltonian& ham_;
  std::unique_lock<std::mutex> lock_;
  b2ci_ctx* ctx_ = nullptr;  // borrowed from the process runtime
  int rank_ = 0, nranks_ = 1;
  b2ci_csr* cache_H_ = nullptr;  // matrix and list of the last cached selected_ci_diag
  b2ci_dets* cache_dets_ = nullptr;
  double cache_tol_ = 0.;
};

std::shared_ptr<data::Wavefunction> make_wavefunction(const std::vector<Det>& dets, std::vector<double> C,
                                                      size_t norb) {
  std::vector<data::Configuration> cfg;
  cfg.reserve(dets.size());
  for (const Det& d : dets) cfg.emplace_back(d.a, d.b, norb);
  return std::make_shared<data::Wavefunction>(std::move(C), std::move(cfg), norb);
}

// build_wavefunction's RDM step (macis_base.hpp:155-215): spin-dependent blocks, two-body ones
// scaled by 2 as the adapter stores them
void attach_rdms_spin_dependent(CiSession& S, const data::Settings& st, const std::vector<Det>& dets,
                                const std::vector<double>& C, data::Wavefunction& w) {
  const bool one = st.get<bool>("calculate_one_rdm"), two = st.get<bool>("calculate_two_rdm");
  if (!one && !two) return;
  const size_t n = w.num_active_orbitals(), n2 = n * n, n4 = n2 * n2;
  std::vector<double> aa(one ? n2 : 0, 0.0), bb(one ? n2 : 0, 0.0), aaaa(two ? n4 : 0, 0.0), bbbb(two ? n4 : 0, 0.0),
      aabb(two ? n4 : 0, 0.0);
  S.form_rdms(dets, C, true, aa, bb, aaaa, bbbb, aabb);
  for (auto* v : {&aaaa, &bbbb, &aabb})
    for (double& x : *v) x *= 2.0;
  w.set_rdms_spin_dependent(std::move(aa), std::move(bb), std::move(aaaa), std::move(aabb), std::move(bbbb));
}

// build_wavefunction's entropy step (macis_base.hpp:219-260): s1 whenever any of the three is
// asked for, the two-orbital matrix / mutual information only where requested
void attach_entropies(CiSession& S, const data::Settings& st, const std::vector<Det>& dets,
                      const std::vector<double>& C, data::Wavefunction& w) {
  const bool e1 = st.get<bool>("calculate_single_orbital_entropies");
  const bool e2 = st.get<bool>("calculate_two_orbital_entropies");
  const bool emi = st.get<bool>("calculate_mutual_information");
  if (!e1 && !e2 && !emi) return;
  const size_t n = w.num_active_orbitals();
  std::vector<double> s1(n, 0.0), s2(e2 ? n * n : 0, 0.0), mi(emi ? n * n : 0, 0.0);
  S.form_entropies(dets, C, s1, s2, mi);
  if (!e1) s1.clear();
  w.set_entropies(std::move(s1), std::move(s2), std::move(mi));
}

void check_hamiltonian(const data::Hamiltonian& h, const char* who, unsigned na, unsigned nb) {
  if (h.is_unrestricted())
    throw std::runtime_error(std::string(who) + " does not support unrestricted orbitals. "
                                                "Only restricted orbitals are supported.");
  if (na > h.num_active_orbitals() || nb > h.num_active_orbitals())
    throw std::invalid_argument(std::string(who) + ": more electrons of one spin than active orbitals");
}

// CASCI on the device-generated Hilbert space (cas_helper::impl / compute_casci_rdms)
double casci(CiSession& S, const data::Settings& st, unsigned na, unsigned nb, std::vector<Det>& dets,
             std::vector<double>& C) {
  const McscfSettings m = get_mcscf_settings(st);
  const int64_t cutoff = st.get<int64_t>("iterative_solver_dimension_cutoff");
  b2ci_dets* d = nullptr;
  B2(b2ci_dets_generate_fci(S.ctx(), S.norb(), int(na), int(nb), &d));
  int64_t n = 0;
  b2ci_dets_size(d, &n);
  double E = 0.;
  try {
    dets.resize(size_t(n));
    B2(b2ci_dets_download(S.ctx(), d, reinterpret_cast<uint64_t*>(dets.data()), 2));
    if (n == 1) {
      E = S.diagonal_element(dets[0]);
      C = {1.0};
    } else if (n <= cutoff) {
      E = S.dense_diag(d, n, m.ci_matel_tol, C);
    } else {
      C.clear();
      E = S.selected_ci_diag(d, n, m.ci_matel_tol, m.ci_max_subspace, m.ci_res_tol, C);
    }
  } catch (...) {
    b2ci_dets_free(S.ctx(), d);
    throw;
  }
  b2ci_dets_free(S.ctx(), d);
  g_stats["ndets"] = double(n);
  return E;
}

// ---- ASCI outer loop -------------------------------------------------------------------------
// reorder_ci_on_coeff (determinant_sort.hpp:45-64, |c| descending) is folded into asci_iter's core
// selection below. The reference's std::sort leaves the order of equal |c| unspecified; ties keep
// their current (spin-sorted) order here, the same rule as oracle/port.py, so the core set is
// reproducible.
void reorder_ci_on_alpha(std::vector<Det>& wfn, std::vector<double>& X, size_t nkeep) {  // :78-101
  std::vector<int64_t> idx(nkeep);
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int64_t i, int64_t j) { return spin_less(wfn[i], wfn[j]); });
  std::vector<Det> w2(nkeep);
  std::vector<double> x2(nkeep);
  for (size_t i = 0; i < nkeep; ++i) { w2[i] = wfn[idx[i]]; x2[i] = X[idx[i]]; }
  std::copy(w2.begin(), w2.end(), wfn.begin());
  std::copy(x2.begin(), x2.end(), X.begin());
}

}  // namespace

namespace {
void occ_vir(size_t norb, uint64_t s, std::vector<unsigned>& occ, std::vector<unsigned>& vir) {
  occ.clear();
  vir.clear();
  for (unsigned p = 0; p < norb && p < 64; ++p) ((s >> p) & 1u ? occ : vir).push_back(p);
}
std::vector<uint64_t> spin_singles(size_t norb, uint64_t s) {  // append_singles, sd_operations.hpp:60-75
  std::vector<unsigned> occ, vir;
  occ_vir(norb, s, occ, vir);
  std::vector<uint64_t> out;
  out.reserve(occ.size() * vir.size());
  for (unsigned a : vir)
    for (unsigned i : occ) out.push_back(s ^ (uint64_t(1) << i) ^ (uint64_t(1) << a));
  return out;
}
std::vector<uint64_t> spin_doubles(size_t norb, uint64_t s) {  // append_doubles, sd_operations.hpp:77-98
  std::vector<unsigned> occ, vir;
  occ_vir(norb, s, occ, vir);
  std::vector<uint64_t> out;
  for (size_t a = 0; a < vir.size(); ++a)
    for (size_t i = 0; i < occ.size(); ++i)
      for (size_t b = a + 1; b < vir.size(); ++b)
        for (size_t j = i + 1; j < occ.size(); ++j)
          out.push_back(s ^ (uint64_t(1) << occ[i]) ^ (uint64_t(1) << occ[j]) ^ (uint64_t(1) << vir[a]) ^
                        (uint64_t(1) << vir[b]));
  return out;
}
}  // namespace

std::vector<std::pair<uint64_t, uint64_t>> generate_cis_hilbert_space(size_t norb, uint64_t alpha, uint64_t beta) {
  std::vector<std::pair<uint64_t, uint64_t>> dets{{alpha, beta}};
  for (uint64_t a : spin_singles(norb, alpha)) dets.emplace_back(a, beta);
  for (uint64_t b : spin_singles(norb, beta)) dets.emplace_back(alpha, b);
  return dets;
}
std::vector<std::pair<uint64_t, uint64_t>> generate_cisd_hilbert_space(size_t norb, uint64_t alpha, uint64_t beta) {
  const std::vector<uint64_t> sa = spin_singles(norb, alpha), sb = spin_singles(norb, beta);
  std::vector<std::pair<uint64_t, uint64_t>> dets = generate_cis_hilbert_space(norb, alpha, beta);
  for (uint64_t a : spin_doubles(norb, alpha)) dets.emplace_back(a, beta);
  for (uint64_t b : spin_doubles(norb, beta)) dets.emplace_back(alpha, b);
  for (uint64_t a : sa)
    for (uint64_t b : sb) dets.emplace_back(a, b);
  return dets;
}

// Indices of the core determinants in order of decreasing |c| (ties: lower index first): the prefix
// reorder_ci_on_coeff + the core-selection rule of asci_iter keep (asci/iteration.hpp:62-100;
// fixed: the ncdets_max largest, percentage: the shortest prefix whose weight reaches the threshold).
std::vector<int64_t> select_core_indices(const std::vector<double>& X, bool fixed_core, size_t ncdets_max,
                                         double core_selection_threshold) {
  const size_t n = X.size();
  std::vector<int64_t> idx(n);
  std::iota(idx.begin(), idx.end(), 0);
  auto by_coeff = [&](int64_t i, int64_t j) {
    const double ci = std::abs(X[size_t(i)]), cj = std::abs(X[size_t(j)]);
    return ci > cj || (ci == cj && i < j);
  };
  size_t K = fixed_core ? std::min<size_t>(ncdets_max, n) : std::min<size_t>(n, 4096);
  size_t nkeep = 0;
  for (;;) {
    if (K < n) std::nth_element(idx.begin(), idx.begin() + int64_t(K), idx.end(), by_coeff);
    std::sort(idx.begin(), idx.begin() + int64_t(K), by_coeff);
    if (fixed_core) { nkeep = K; break; }
    double w = 0.0;
    nkeep = 0;
    bool reached = false;
    for (size_t i = 0; i < K; ++i) {
      w += X[size_t(idx[i])] * X[size_t(idx[i])];
      nkeep++;
      if (w >= core_selection_threshold) { reached = true; break; }
    }
    if (reached || K == n) break;
    K = std::min(n, K * 4);
  }
  idx.resize(nkeep);
  return idx;
}

namespace {
struct WallTimer {  // host wall clock of one phase, accumulated into the run statistics
  const char* key;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  explicit WallTimer(const char* k) : key(k) {}
  ~WallTimer() { g_stats[key] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

double asci_iter(CiSession& S, const AsciSettings& a, const McscfSettings& m, int64_t ndets_max, double E0,
                 std::vector<Det>& wfn, std::vector<double>& X) {  // iteration.hpp:50-226
  std::unique_ptr<WallTimer> wt(new WallTimer("wall_core_selection_ms"));
  // warm start (iteration.hpp:132-180) needs old determinant -> old coefficient. The list arrives
  // spin-sorted from the previous iteration, so a copy taken here turns the lookup into a merge
  // of two sorted lists; only an unsorted input (first call of a user-supplied list) is hashed.
  std::vector<Det> old_dets;
  std::vector<double> old_X;
  bool old_sorted = false;
  if (a.warm_start_davidson) {
    old_sorted = std::is_sorted(wfn.begin(), wfn.end(), spin_less);
    if (old_sorted) { old_dets = wfn; old_X = X; }
  }
  // Core selection (iteration.hpp:62-100). The reference sorts the whole wavefunction by |c| and
  // keeps a prefix; only that prefix is ever used, so it is found by selection (nth_element) and
  // sorted alone -- with ties broken by position, the same prefix and order as the stable full
  // sort (1.2 s of host time per iteration at 1e7 determinants otherwise).
  const std::vector<int64_t> top = select_core_indices(X, a.fixed_core, size_t(a.ncdets_max), a.core_selection_threshold);
  const size_t nkeep = top.size();
  std::vector<Det> core(nkeep);
  std::vector<double> core_X(nkeep);
  for (size_t i = 0; i < nkeep; ++i) { core[i] = wfn[size_t(top[i])]; core_X[i] = X[size_t(top[i])]; }
  if (nkeep > 1) reorder_ci_on_alpha(core, core_X, nkeep);
  std::unordered_map<Det, double, DetHash> old;
  if (a.warm_start_davidson && !old_sorted) {
    old.reserve(wfn.size());
    for (size_t i = 0; i < wfn.size(); ++i) old.emplace(wfn[i], X[i]);
  }
  wt.reset(new WallTimer("wall_search_ms"));
  wfn = S.asci_search(core.data(), core_X.data(), int64_t(nkeep), E0, ndets_max, a);
  wt.reset(new WallTimer("wall_sort_warmstart_ms"));
  if (!std::is_sorted(wfn.begin(), wfn.end(), spin_less)) std::sort(wfn.begin(), wfn.end(), spin_less);
  std::vector<double> X_local;
  if (a.warm_start_davidson && (!old.empty() || !old_dets.empty())) {
    X_local.assign(wfn.size(), 0.0);
    if (old_sorted) {
      size_t j = 0;
      for (size_t i = 0; i < wfn.size() && j < old_dets.size(); ++i) {
        while (j < old_dets.size() && spin_less(old_dets[j], wfn[i])) ++j;
        if (j < old_dets.size() && old_dets[j] == wfn[i]) X_local[i] = old_X[j];
      }
    } else {
      for (size_t i = 0; i < wfn.size(); ++i) {
        auto it = old.find(wfn[i]);
        if (it != old.end()) X_local[i] = it->second;
      }
    }
    double nrm = 0.;
    for (double x : X_local) nrm += x * x;
    nrm = std::sqrt(nrm);
    if (nrm < std::max(a.min_warm_start_overlap, std::numeric_limits<double>::epsilon())) {
      X_local.clear();  // diagonal guess
    } else {
      const double inv = 1.0 / nrm;
      for (double& x : X_local) x *= inv;
    }
  }
  wt.reset(new WallTimer("wall_selected_ci_diag_ms"));
  const double E = S.selected_ci_diag(wfn, m.ci_matel_tol, m.ci_max_subspace, m.ci_res_tol, X_local, true,
                                      a.min_patch_overlap);
  wt.reset();
  X = std::move(X_local);
  g_stats["asci_iterations"] += 1.0;
  return E;
}

double asci_grow(CiSession& S, const AsciSettings& a, const McscfSettings& m, double E0, std::vector<Det>& wfn,
                 std::vector<double>& X) {  // grow.hpp:45-268
  McscfSettings gm = m;
  if (a.grow_ci_residual_tolerance > 0) gm.ci_res_tol = a.grow_ci_residual_tolerance;
  size_t prev = wfn.size();
  double gf = a.grow_factor;
  const size_t ntmax = size_t(a.ntdets_max);
  while (wfn.size() < ntmax) {
    double eff = gf;
    if (a.taper_grow_factor > 0 && size_t(std::ceil(wfn.size() * gf)) > ntmax)
      eff = std::max(a.min_grow_factor, a.taper_grow_factor);
    size_t ndets_new = std::min(std::max(size_t(a.ntdets_min), size_t(std::ceil(wfn.size() * eff))), ntmax);
    if (ndets_new <= wfn.size()) {
      ndets_new = std::min(wfn.size() + 1, ntmax);
      if (ndets_new <= wfn.size()) break;
    }
    const double E = asci_iter(S, a, gm, int64_t(ndets_new), E0, wfn, X);
    log_line(1, "asci_grow", "NDETS = %zu (requested %zu), E0 = %.12e, dE = %.5e, GROW_FACTOR = %.4f", wfn.size(), ndets_new, E,
             E - E0, gf);
    if (wfn.size() < ndets_new) {
      gf = std::max(a.min_grow_factor, gf * a.growth_backoff_rate);
      if (wfn.size() <= prev) break;  // (the reference leaves E0 at its previous value here)
    } else {
      gf = std::min(a.grow_factor, gf * a.growth_recovery_rate);
    }
    prev = wfn.size();
    if (a.grow_with_rot && wfn.size() >= size_t(a.rot_size_start)) {
      // rotate to natural orbitals and rediagonalise in the rotated basis with a diagonal guess and
      // the final (not the grow) tolerances (grow.hpp:163-258); the energy carried on is the one
      // from before the rotation, as in the reference
      S.rotate_to_natural_orbitals(wfn, X);
      std::vector<double> X_local;
      S.selected_ci_diag(wfn, m.ci_matel_tol, m.ci_max_subspace, m.ci_res_tol, X_local);
      X = std::move(X_local);
    }
    E0 = E;
  }
  return E0;
}

double asci_refine(CiSession& S, const AsciSettings& a, const McscfSettings& m, double E0, std::vector<Det>& wfn,
                   std::vector<double>& X) {  // refine.hpp:44-237
  size_t ndets = wfn.size();
  bool converged = false;
  double prev_dE = 0.0;
  int oscillation = 0;
  size_t max_iter = size_t(a.max_refine_iter), total_ext = 0;
  const size_t max_ext = size_t(a.max_refine_iter);
  std::vector<Det> prev_wfn;
  for (size_t iter = 0; iter < max_iter; ++iter) {
    const double E = asci_iter(S, a, m, int64_t(ndets), E0, wfn, X);
    if (wfn.size() != ndets) {
      ndets = wfn.size();
      if (wfn.size() < size_t(a.ntdets_min)) break;
    }
    const double dE = E - E0;
    log_line(1, "asci_refine", "ITER = %zu, NDETS = %zu, E0 = %.12e, dE = %.5e", iter, wfn.size(), E, dE);
    if (std::abs(dE) < a.refine_energy_tol) {
      E0 = E;
      converged = true;
      break;
    }
    if (iter > 0 && prev_dE * dE < 0 && std::abs(prev_dE + dE) < a.refine_energy_tol) {
      oscillation++;
      if (oscillation >= 2 && !prev_wfn.empty()) {
        // stabilise by diagonalising in the union of the last two determinant sets
        std::vector<Det> uni;
        uni.reserve(wfn.size() + prev_wfn.size());
        std::set_union(prev_wfn.begin(), prev_wfn.end(), wfn.begin(), wfn.end(), std::back_inserter(uni), spin_less);
        std::vector<double> Xu(uni.size(), 0.0);
        for (size_t i = 0; i < wfn.size(); ++i) {
          auto it = std::lower_bound(uni.begin(), uni.end(), wfn[i], spin_less);
          if (it != uni.end() && *it == wfn[i]) Xu[size_t(it - uni.begin())] = X[i];
        }
        const double Eu = S.selected_ci_diag(uni, m.ci_matel_tol, m.ci_max_subspace, m.ci_res_tol, Xu, true,
                                             a.min_patch_overlap);
        const size_t ext = std::min(size_t(oscillation), max_ext - total_ext);
        if (ext > 0) { max_iter += ext; total_ext += ext; }
        g_stats["asci_refine_unions"] += 1.0;
        g_stats["asci_refine_extensions"] = double(total_ext);
        wfn = std::move(uni);
        X = std::move(Xu);
        ndets = wfn.size();
        E0 = Eu;
        prev_wfn.clear();
        oscillation = 0;
        prev_dE = 0.0;
        continue;
      }
    } else {
      oscillation = 0;
    }
    prev_dE = dE;
    prev_wfn = wfn;
    E0 = E;
  }
  if (!converged) {
    std::string msg = "ASCI Refine did not converge";
    if (total_ext > 0)
      msg += " (oscillation detected, " + std::to_string(total_ext) +
             " extra iterations granted). Consider using percentage core_selection_strategy, increasing "
             "ncdets_max, or loosening refine_energy_tol.";
    throw std::runtime_error(msg);
  }
  return E0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
MultiConfigurationSettings::MultiConfigurationSettings() {  // mc.hpp:26-61
  set_default<bool>("calculate_one_rdm", false);
  set_default<bool>("calculate_two_rdm", false);
  set_default<bool>("calculate_single_orbital_entropies", false);
  set_default<bool>("calculate_two_orbital_entropies", false);
  set_default<bool>("calculate_mutual_information", false);
  set_default<double>("ci_residual_tolerance", 1.0e-6, "CI residual convergence tolerance",
                      data::BoundConstraint<double>{0.0, 1.0});
  set_default<int64_t>("max_solver_iterations", 200, "Maximum number of Davidson iterations",
                       data::BoundConstraint<int64_t>{1, std::numeric_limits<int64_t>::max()});
  set_default<int64_t>("iterative_solver_dimension_cutoff", 2000,
                       "Matrix size cutoff for using iterative eigensolver",
                       data::BoundConstraint<int64_t>{1, std::numeric_limits<int64_t>::max()});
}
B200CiSettings::B200CiSettings() {  // MacisSettings, macis_base.hpp:27-37
  set_default<double>("ci_matel_tol", std::numeric_limits<double>::epsilon(),
                      "Hamiltonian matrix element sparsification threshold", data::BoundConstraint<double>{0.0, 1.0});
}
B200AsciSettings::B200AsciSettings() {  // MacisAsciSettings, macis_asci.hpp:34-183
  const int64_t imax = std::numeric_limits<int64_t>::max();
  const double dmax = std::numeric_limits<double>::max();
  using BI = data::BoundConstraint<int64_t>;
  using BD = data::BoundConstraint<double>;
  set_default<int64_t>("ntdets_max", 100000, "Maximum number of trial determinants in the variational space", BI{1, imax});
  set_default<int64_t>("ntdets_min", 100, "Minimum number of trial determinants required", BI{1, imax});
  set_default<int64_t>("ncdets_max", 100, "Maximum number of core determinants", BI{1, imax});
  set_default<double>("search_matel_tol", 1e-8, "Hamiltonian matrix element magnitude threshold", BD{0.0, 1.0});
  set_default<double>("rv_prune_tol", 1e-8, "Ratio value pruning threshold", BD{0.0, 1.0});
  set_default<double>("pt2_tol", 1e-16, "PT2 correction tolerance", BD{0.0, 1.0});
  set_default<double>("refine_energy_tol", 1e-6, "Energy convergence tolerance for refinement", BD{0.0, 1.0});
  set_default<int64_t>("pt2_reserve_count", 70000000, "Reserve count for PT2 calculations", BI{0, imax});
  set_default<bool>("pt2_prune", false);
  set_default<bool>("pt2_precompute_eps", false);
  set_default<bool>("pt2_precompute_idx", false);
  set_default<bool>("pt2_print_progress", false);
  set_default<int64_t>("pt2_bigcon_thresh", 250, "Threshold for using bigcon PT2 algorithm", BI{0, imax});
  set_default<int64_t>("pair_size_max", 500000000, "Maximum number of ASCI contribution pairs to store in memory", BI{1, imax});
  set_default<int64_t>("nxtval_bcount_thresh", 1000, "Threshold for next value batch count", BI{1, imax});
  set_default<int64_t>("nxtval_bcount_inc", 10, "Increment for next value batch count", BI{1, imax});
  set_default<bool>("just_singles", false);
  set_default<double>("grow_factor", 8.0, "Factor by which to grow the variational space", BD{1.0, dmax});
  set_default<double>("min_grow_factor", 1.01, "Minimum allowed growth factor", BD{1.0, dmax});
  set_default<double>("growth_backoff_rate", 0.5, "Rate to reduce grow_factor on failure", BD{0.0, 1.0});
  set_default<double>("growth_recovery_rate", 1.1, "Rate to restore grow_factor on success", BD{1.0, dmax});
  set_default<int64_t>("max_refine_iter", 6, "Maximum number of refinement iterations", BI{0, imax});
  set_default<bool>("grow_with_rot", false);
  set_default<int64_t>("rot_size_start", 1000, "Starting size for rotations", BI{1, imax});
  set_default<int64_t>("constraint_level", 2, "Constraint level for excitation generation", BI{0, imax});
  set_default<int64_t>("pt2_max_constraint_level", 5, "Maximum constraint level for PT2 calculations", BI{0, imax});
  set_default<int64_t>("pt2_min_constraint_level", 0, "Minimum constraint level for PT2 calculations", BI{0, imax});
  set_default<int64_t>("pt2_constraint_refine_force", 0, "Force constraint refinement for PT2 calculations", BI{0, imax});
  set_default("core_selection_strategy", std::string("percentage"),
              "Core determinant selection strategy: 'percentage' uses cumulative weight threshold, 'fixed' "
              "uses a fixed number of determinants",
              data::ListConstraint<std::string>{{"percentage", "fixed"}});
  set_default<double>("core_selection_threshold", 0.95, "Cumulative weight threshold for core selection",
                      BD{std::numeric_limits<double>::epsilon(), 1.0});
  set_default<bool>("warm_start_davidson", true, "Warm-start Davidson from previous eigenvector");
  set_default<double>("min_warm_start_overlap", 0.5, "Minimum projected vector norm for warm-start Davidson", BD{0.0, 1.0});
  set_default<double>("min_patch_overlap", 0.3, "Minimum determinant overlap for incremental H build", BD{0.0, 1.0});
  set_default<double>("grow_ci_residual_tolerance", 0.0, "CI residual tolerance during grow phase (0 = use refine tolerance)");
  set_default<double>("taper_grow_factor", 0.0, "Growth factor for final expansion near ntdets_max (0 = disabled)");
  set_default<std::string>("hamiltonian_build_algorithm", std::string(""),
                           "Algorithm for diagonal Hamiltonian construction: '' or 'sorted_double_loop' (default), "
                           "'residue_arrays', 'dynamic_bit_masking'");
  // a tuning knob of the reference's CPU pair enumeration: accepted, no effect on the device enumeration
  set_default<int64_t>("dynamic_bit_masking_num_masks", 0,
                       "Number of bit masks for dynamic_bit_masking generator (0 = use generator default)");
}

std::string MultiConfigurationCalculator::hash(std::shared_ptr<data::Hamiltonian> h, unsigned na, unsigned nb) const {
  return hash_hex(type_name() + "|" + name() + "|" + settings().content_hash() + "|" +
                  (h ? h->content_hash() : std::string("null")) + "|" + std::to_string(na) + "|" + std::to_string(nb));
}

void MultiConfigurationCalculatorFactory::register_default_instances() {  // mc.cpp:26-31
  register_instance([]() { return std::make_unique<B200Cas>(); });
  register_instance([]() { return std::make_unique<B200Asci>(); });
}
void ProjectedMultiConfigurationCalculatorFactory::register_default_instances() {
  register_instance([]() { return std::make_unique<B200Pmc>(); });
}

McResult B200Cas::_run_impl(std::shared_ptr<data::Hamiltonian> h, unsigned na, unsigned nb) const {
  if (!h) throw std::invalid_argument("B200Cas: null Hamiltonian");
  check_hamiltonian(*h, "B200Cas", na, nb);
  g_stats.clear();
  const auto t0 = std::chrono::steady_clock::now();
  CiSession S(*h);
  const double launches0 = double(b2ci_ctx_launch_count(S.ctx()));
  std::vector<Det> dets;
  std::vector<double> C;
  const double E = casci(S, *_settings, na, nb, dets, C);
  auto w = make_wavefunction(dets, C, h->num_active_orbitals());
  attach_rdms_spin_dependent(S, *_settings, dets, C, *w);
  attach_entropies(S, *_settings, dets, C, *w);
  g_stats["wall_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  g_stats["launches"] = double(b2ci_ctx_launch_count(S.ctx())) - launches0;
  return {E + h->get_core_energy(), w};
}

McResult B200Asci::_run_impl(std::shared_ptr<data::Hamiltonian> h, unsigned na, unsigned nb) const {
  if (!h) throw std::invalid_argument("B200Asci: null Hamiltonian");
  check_hamiltonian(*h, "B200Asci", na, nb);
  if (_settings->get<bool>("grow_with_rot") && runtime().nranks > 1)
    throw std::runtime_error("grow_with_rot is not available with a communicator in this build");
  const McscfSettings m = get_mcscf_settings(*_settings);
  const AsciSettings a = get_asci_settings(*_settings);
  g_stats.clear();
  const auto t0 = std::chrono::steady_clock::now();
  CiSession S(*h);
  const double launches0 = double(b2ci_ctx_launch_count(S.ctx()));
  const int norb = S.norb();
  // residue arrays hold O(n_e^2) residues per determinant: beyond 60 electrons the reference falls
  // back to the sorted double loop (macis_asci.cpp:32,96-108) and so do its pattern rules
  int generator = a.generator;
  if (generator == B2CI_GEN_RESIDUE_ARRAYS && na + nb > 60) generator = B2CI_GEN_SORTED_DOUBLE_LOOP;
  S.set_generator(generator);
  g_stats["hamiltonian_generator"] = double(generator);
  std::vector<Det> dets;
  std::vector<double> C;
  double E = 0.;
  const int64_t fci_dim_a = binomial(norb, na), fci_dim_b = binomial(norb, nb);
  const bool fits = fci_dim_a != std::numeric_limits<int64_t>::max() && fci_dim_b != std::numeric_limits<int64_t>::max() &&
                    (fci_dim_b == 0 || fci_dim_a <= std::numeric_limits<int64_t>::max() / std::max<int64_t>(1, fci_dim_b));
  if (fits && a.ntdets_max > fci_dim_a * fci_dim_b) {
    // more determinants requested than the space holds: CASCI (macis_asci.cpp:125-157)
    E = casci(S, *_settings, na, nb, dets, C);
  } else {
    // canonical HF determinant (wavefunction_traits::canonical_hf_determinant)
    Det hf{na >= 64 ? ~uint64_t(0) : ((uint64_t(1) << na) - 1), nb >= 64 ? ~uint64_t(0) : ((uint64_t(1) << nb) - 1)};
    dets = {hf};
    C = {1.0};
    E = S.diagonal_element(hf);
    E = asci_grow(S, a, m, E, dets, C);
    g_stats["ndets_after_grow"] = double(dets.size());
    if (a.max_refine_iter) E = asci_refine(S, a, m, E, dets, C);
    g_stats["ndets"] = double(dets.size());
  }
  auto w = make_wavefunction(dets, C, h->num_active_orbitals());
  attach_rdms_spin_dependent(S, *_settings, dets, C, *w);
  attach_entropies(S, *_settings, dets, C, *w);
  g_stats["wall_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  g_stats["launches"] = double(b2ci_ctx_launch_count(S.ctx())) - launches0;
  return {E + h->get_core_energy(), w};
}

McResult B200Pmc::_run_impl(std::shared_ptr<data::Hamiltonian> h,
                            const std::vector<data::Configuration>& configurations) const {
  if (!h) throw std::invalid_argument("B200Pmc: null Hamiltonian");
  if (h->is_unrestricted())
    throw std::runtime_error("B200Pmc does not support unrestricted orbitals. Only restricted orbitals are supported.");
  const McscfSettings m = get_mcscf_settings(*_settings);
  const int64_t cutoff = _settings->get<int64_t>("iterative_solver_dimension_cutoff");
  const size_t norb = h->num_active_orbitals();
  if (!configurations.empty() && configurations[0].get_orbital_capacity() < norb)
    throw std::runtime_error("Configuration orbital capacity does not match Hamiltonian active space.");
  if (configurations.empty()) throw std::runtime_error("Configuration basis cannot be empty");
  std::vector<Det> dets;
  dets.reserve(configurations.size());
  for (const auto& c : configurations) dets.push_back(Det{c.alpha_word(), c.beta_word()});
  g_stats.clear();
  const auto t0 = std::chrono::steady_clock::now();
  CiSession S(*h);
  const double launches0 = double(b2ci_ctx_launch_count(S.ctx()));
  std::vector<double> C;
  double E = 0.;
  const int64_t n = int64_t(dets.size());
  if (n == 1) {
    E = S.diagonal_element(dets[0]);
    C = {1.0};
  } else if (n <= cutoff) {
    E = S.dense_diag(dets, m.ci_matel_tol, C);
  } else {
    E = S.selected_ci_diag(dets, m.ci_matel_tol, m.ci_max_subspace, m.ci_res_tol, C);
  }
  g_stats["ndets"] = double(n);
  auto w = make_wavefunction(dets, C, norb);
  if (_settings->get<bool>("calculate_one_rdm") || _settings->get<bool>("calculate_two_rdm")) {
    // the PMC adapter always forms both spin-traced matrices (macis_pmc.cpp:128-160)
    const size_t n2 = norb * norb;
    std::vector<double> one(n2, 0.0), two(n2 * n2, 0.0), none;
    S.form_rdms(dets, C, false, one, none, two, none, none);
    w->set_rdms_spin_traced(std::move(one), std::move(two));
  }
  g_stats["wall_ms"] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  g_stats["launches"] = double(b2ci_ctx_launch_count(S.ctx())) - launches0;
  return {E + h->get_core_energy(), w};
}

double compute_casci_rdms(const MCSCFSettings& st, size_t norb, size_t nalpha, size_t nbeta, const double* T,
                          const double* V, double* ORDM, double* TRDM, std::vector<double>& C) {  // mcscf/cas.hpp:33-64
  if (!T || !V) throw std::invalid_argument("compute_casci_rdms: null integrals");
  const size_t n2 = norb * norb;
  data::Hamiltonian ham(norb, std::vector<double>(T, T + n2), std::vector<double>(V, V + n2 * n2), 0.0);
  check_hamiltonian(ham, "compute_casci_rdms", unsigned(nalpha), unsigned(nbeta));
  CiSession S(ham);
  b2ci_dets* d = nullptr;
  B2(b2ci_dets_generate_fci(S.ctx(), int(norb), int(nalpha), int(nbeta), &d));
  int64_t n = 0;
  b2ci_dets_size(d, &n);
  std::vector<Det> dets;
  if (ORDM && TRDM) {
    dets.resize(size_t(n));
    if (b2ci_dets_download(S.ctx(), d, reinterpret_cast<uint64_t*>(dets.data()), 2) != 0) {
      b2ci_dets_free(S.ctx(), d);
      fail("b2ci_dets_download");
    }
  }
  // (the session frees the list: take_dets)
  const double E0 = S.selected_ci_diag(d, n, st.ci_matel_tol, int64_t(st.ci_max_subspace), st.ci_res_tol, C, false, 0.3,
                                       true);
  if (ORDM && TRDM) {
    std::vector<double> o1(n2, 0.0), t1(n2 * n2, 0.0), none;
    S.form_rdms(dets, C, false, o1, none, t1, none, none);
    for (size_t i = 0; i < n2; ++i) ORDM[i] += o1[i];
    for (size_t i = 0; i < n2 * n2; ++i) TRDM[i] += t1[i];
  }
  return E0;
}

std::pair<int64_t, int64_t> row_block(int64_t n, int rank, int nranks) {
  if (n < 0 || nranks < 1 || rank < 0 || rank >= nranks) throw std::invalid_argument("row_block: bad arguments");
  const int64_t base = n / nranks, rem = n % nranks;
  const int64_t r0 = rank * base + std::min<int64_t>(rank, rem);
  return {r0, r0 + base + (rank < rem ? 1 : 0)};
}
void set_device(int device) {
  Runtime& rt = runtime();
  std::lock_guard<std::mutex> g(rt.mutex);
  if (device != rt.device) {
    if (rt.nranks > 1) throw std::runtime_error("set_device: clear_communicator() first");
    rt.drop();
    rt.device = device;
  }
}
void set_communicator(const std::string& id, int rank, int nranks) {
  if (id.size() != 128) throw std::invalid_argument("set_communicator: the NCCL unique id has 128 bytes");
  if (nranks < 1 || rank < 0 || rank >= nranks) throw std::invalid_argument("set_communicator: bad rank / nranks");
  Runtime& rt = runtime();
  std::lock_guard<std::mutex> g(rt.mutex);
  rt.drop();
  rt.rank = 0;
  rt.nranks = 1;
  B2(b2ci_ctx_create(rt.device, nullptr, &rt.ctx));
  if (nranks > 1) {
    if (b2ci_comm_init(rt.ctx, id.data(), rank, nranks) != 0) {
      const std::string err = b2ci_last_error();
      rt.drop();
      throw std::runtime_error("b2ci_comm_init: " + err);
    }
  }
  rt.rank = rank;
  rt.nranks = nranks;
}
void clear_communicator() {
  Runtime& rt = runtime();
  std::lock_guard<std::mutex> g(rt.mutex);
  rt.drop();
  rt.rank = 0;
  rt.nranks = 1;
}
std::map<std::string, double> last_run_stats() { return g_stats; }

std::pair<double, std::vector<double>> davidson_solver(int64_t n, const int64_t* rowptr, const int64_t* colind,
                                                       const double* nzval, double tol, int64_t max_m) {
  if (n < 1) throw std::invalid_argument("davidson_solver: empty matrix");
  b2ci_ctx* ctx = nullptr;
  B2(b2ci_ctx_create(runtime().device, nullptr, &ctx));
  b2ci_csr* H = nullptr;
  std::vector<double> X(size_t(n), 0.0);
  double E = 0.;
  int64_t niter = 0;
  try {
    B2(b2ci_csr_upload(ctx, n, rowptr[n] - rowptr[0], rowptr, colind, nzval, &H));
    // diagonal_guess: unit vector at the smallest diagonal element (davidson.hpp:106-113)
    std::vector<double> D(static_cast<size_t>(n), 0.0);
    B2(b2ci_csr_diagonal(ctx, H, D.data()));
    X[size_t(std::min_element(D.begin(), D.end()) - D.begin())] = 1.0;
    const int rc = b2ci_davidson(ctx, H, max_m, tol, X.data(), 0, &niter, &E, nullptr);
    if (rc != 0) throw std::runtime_error(b2ci_last_error());
  } catch (...) {
    if (H) b2ci_csr_free(ctx, H);
    b2ci_ctx_destroy(ctx);
    throw;
  }
  b2ci_csr_free(ctx, H);
  b2ci_ctx_destroy(ctx);
  return {E, std::move(X)};
}

}  // namespace qdk_b200::algorithms
