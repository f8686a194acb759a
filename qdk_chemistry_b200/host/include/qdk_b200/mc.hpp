// MultiConfigurationCalculator / ProjectedMultiConfigurationCalculator interfaces and factories
// (cpp/include/qdk/chemistry/algorithms/mc.hpp:26-214, pmc.hpp:65-190), and the B200 entries
// that stand where the reference registers MacisCas / MacisAsci / MacisPmc (mc.cpp:26-31,
// macis_cas.hpp:24-66, macis_asci.hpp:34-235, macis_pmc.hpp).
#pragma once
#include <limits>
#include <utility>

#include "algorithm.hpp"

namespace qdk_b200::algorithms {

class MultiConfigurationSettings : public data::Settings {
 public:
  MultiConfigurationSettings();
};

using McResult = std::pair<double, std::shared_ptr<data::Wavefunction>>;

class MultiConfigurationCalculator
    : public Algorithm<MultiConfigurationCalculator, McResult, std::shared_ptr<data::Hamiltonian>,
                       unsigned int, unsigned int> {
 public:
  std::string type_name() const final { return "multi_configuration_calculator"; }
  std::string hash(std::shared_ptr<data::Hamiltonian> h, unsigned na, unsigned nb) const;
};

struct MultiConfigurationCalculatorFactory
    : public AlgorithmFactory<MultiConfigurationCalculator, MultiConfigurationCalculatorFactory> {
  static std::string algorithm_type_name() { return "multi_configuration_calculator"; }
  static void register_default_instances();
  // the reference's default is "macis_cas"; this build's default is its drop-in
  static std::string default_algorithm_name() { return "b200_cas"; }
};

class ProjectedMultiConfigurationCalculator
    : public Algorithm<ProjectedMultiConfigurationCalculator, McResult,
                       std::shared_ptr<data::Hamiltonian>, const std::vector<data::Configuration>&> {
 public:
  std::string type_name() const final { return "projected_multi_configuration_calculator"; }
};

struct ProjectedMultiConfigurationCalculatorFactory
    : public AlgorithmFactory<ProjectedMultiConfigurationCalculator,
                              ProjectedMultiConfigurationCalculatorFactory> {
  static std::string algorithm_type_name() { return "projected_multi_configuration_calculator"; }
  static void register_default_instances();
  static std::string default_algorithm_name() { return "b200_pmc"; }
};

// ---- settings (MacisSettings, MacisAsciSettings: same keys, defaults and bounds)
class B200CiSettings : public MultiConfigurationSettings {
 public:
  B200CiSettings();
};
class B200AsciSettings : public B200CiSettings {
 public:
  B200AsciSettings();
};

// ---- concrete calculators
class B200Cas : public MultiConfigurationCalculator {
 public:
  B200Cas() { _settings = std::make_unique<B200CiSettings>(); }
  std::string name() const override { return "b200_cas"; }

 protected:
  McResult _run_impl(std::shared_ptr<data::Hamiltonian> hamiltonian, unsigned int nalpha,
                     unsigned int nbeta) const override;
};
class B200Asci : public MultiConfigurationCalculator {
 public:
  B200Asci() { _settings = std::make_unique<B200AsciSettings>(); }
  std::string name() const override { return "b200_asci"; }

 protected:
  McResult _run_impl(std::shared_ptr<data::Hamiltonian> hamiltonian, unsigned int nalpha,
                     unsigned int nbeta) const override;
};
class B200Pmc : public ProjectedMultiConfigurationCalculator {
 public:
  B200Pmc() { _settings = std::make_unique<B200CiSettings>(); }
  std::string name() const override { return "b200_pmc"; }

 protected:
  McResult _run_impl(std::shared_ptr<data::Hamiltonian> hamiltonian,
                     const std::vector<data::Configuration>& configurations) const override;
};

// ---- the CASCI functor MCSCF drivers call every macro-iteration: macis::compute_casci_rdms and its
// CASRDMFunctor wrapper (external/macis/include/macis/mcscf/cas.hpp:33-88). T (n x n) and V (n^4)
// are column-major and only read; full-CI space in generate_hilbert_space order, always the
// iterative solver (selected_ci_diag), C in/out (a non-trivial C is the Davidson guess). When both
// ORDM and TRDM are given they receive (accumulate, like the reference) the spin-traced RDMs.
struct MCSCFSettings {  // macis::MCSCFSettings fields the functor reads (mcscf.hpp:22-52)
  double ci_res_tol = 1e-8;
  size_t ci_max_subspace = 20;
  double ci_matel_tol = std::numeric_limits<double>::epsilon();
};
double compute_casci_rdms(const MCSCFSettings& settings, size_t norb, size_t nalpha, size_t nbeta, const double* T,
                          const double* V, double* ORDM, double* TRDM, std::vector<double>& C);
struct CASRDMFunctor {
  template <typename... Args>
  static auto rdms(Args&&... args) {
    return compute_casci_rdms(std::forward<Args>(args)...);
  }
};

// CIS / CISD determinant spaces of a reference determinant, in the reference's generation order
// (sd_operations.hpp:60-383: generate_cis_hilbert_space, generate_cisd_hilbert_space): the reference itself,
// alpha singles, beta singles, [alpha doubles, beta doubles, alpha single x beta single]; inside a spin the
// virtual index runs outermost. (alpha, beta) occupation words, bit p = orbital p.
std::vector<std::pair<uint64_t, uint64_t>> generate_cis_hilbert_space(size_t norb, uint64_t alpha, uint64_t beta);
std::vector<std::pair<uint64_t, uint64_t>> generate_cisd_hilbert_space(size_t norb, uint64_t alpha, uint64_t beta);

// core determinants of an ASCI iteration (asci/iteration.hpp:62-100): indices in order of decreasing |c|
std::vector<int64_t> select_core_indices(const std::vector<double>& X, bool fixed_core, size_t ncdets_max,
                                         double core_selection_threshold);

// ---- multi-GPU: one process per GPU. Call once per process before run(); the 128-byte id comes
// from b2ci_comm_unique_id on rank 0 and is distributed by the caller (torch.distributed).
void set_device(int device);
void set_communicator(const std::string& unique_id128, int rank, int nranks);
void clear_communicator();

// contiguous row block [r0, r1) of a rank: equal blocks, the remainder spread over the first ranks
std::pair<int64_t, int64_t> row_block(int64_t n, int rank, int nranks);

// last-run statistics of the calling thread (phase timings in ms, sizes, Davidson iterations)
std::map<std::string, double> last_run_stats();

// davidson_solver(csr, tol, max_m) -> (eigenvalue, eigenvector): python/src/pybind11/algorithms/
// davidson_solver.cpp:60-80 (diagonal guess, int64 CSR, throws on non-convergence)
std::pair<double, std::vector<double>> davidson_solver(int64_t n, const int64_t* rowptr,
                                                       const int64_t* colind, const double* nzval,
                                                       double tol, int64_t max_m);

}  // namespace qdk_b200::algorithms
