// Minimal same-shaped stand-ins for the QDK/Chemistry data classes that cross the
// MultiConfigurationCalculator boundary (the full data model -- HDF5/JSON serialisation,
// basis sets, orbitals -- is out of scope):
//   Settings      cpp/include/qdk/chemistry/data/settings.hpp (typed key/value store with
//                 defaults, bound/list constraints and locking at run())
//   Hamiltonian   cpp/include/qdk/chemistry/data/hamiltonian.hpp (active-space integrals as the
//                 MACIS adapters read them: macis_cas.cpp:58-82)
//   Configuration cpp/include/qdk/chemistry/data/configuration.hpp:64-100 (2 bits per orbital,
//                 printed as '2' 'u' 'd' '0')
//   Wavefunction  the StateVectorContainer view: coefficients + determinants in solver order
//                 (macis_base.hpp:268-282)
#pragma once
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <variant>
#include <tuple>
#include <vector>

namespace qdk_b200::data {

class SettingsAreLocked : public std::runtime_error {
 public:
  SettingsAreLocked() : std::runtime_error("Settings are locked and cannot be modified") {}
};
class SettingNotFound : public std::runtime_error {
 public:
  explicit SettingNotFound(const std::string& key) : std::runtime_error("Setting not found: " + key) {}
};
class SettingTypeMismatch : public std::runtime_error {
 public:
  SettingTypeMismatch(const std::string& key, const std::string& expected)
      : std::runtime_error("Type mismatch for setting '" + key + "'. Expected type: " + expected) {}
};

using SettingValue = std::variant<bool, int64_t, double, std::string>;

template <typename T>
struct BoundConstraint {
  T min, max;
};
template <typename T>
struct ListConstraint {
  std::vector<T> allowed;
};

class Settings {
 public:
  virtual ~Settings() = default;

  // -- modification (throws SettingsAreLocked / SettingNotFound / SettingTypeMismatch /
  //    std::invalid_argument when a constraint is violated)
  void set(const std::string& key, const SettingValue& value);
  void set(const std::string& key, const char* value) { set(key, SettingValue(std::string(value))); }
  template <typename T, std::enable_if_t<std::is_integral_v<T> && !std::is_same_v<T, bool>, int> = 0>
  void set(const std::string& key, T value) { set(key, SettingValue(static_cast<int64_t>(value))); }
  void update(const std::map<std::string, SettingValue>& overrides) {
    for (const auto& [k, v] : overrides) set(k, v);
  }

  template <typename T>
  T get(const std::string& key) const;
  template <typename T>
  T get_or_default(const std::string& key, const T& def) const {
    return has(key) ? get<T>(key) : def;
  }
  const SettingValue& get_raw(const std::string& key) const;
  bool has(const std::string& key) const { return values_.count(key) != 0; }
  std::vector<std::string> keys() const;
  size_t size() const { return values_.size(); }
  bool empty() const { return values_.empty(); }
  std::string get_as_string(const std::string& key) const;
  std::string get_type_name(const std::string& key) const;
  bool has_description(const std::string& key) const { return desc_.count(key) != 0; }
  std::string get_description(const std::string& key) const;
  void lock() const { locked_ = true; }
  bool is_locked() const { return locked_; }
  // stable digest of (key, value) pairs, part of Algorithm::hash
  std::string content_hash() const;

 protected:
  template <typename T>
  void set_default(const std::string& key, const T& value, const std::string& description = "") {
    if constexpr (std::is_integral_v<T> && !std::is_same_v<T, bool>)
      values_[key] = static_cast<int64_t>(value);
    else
      values_[key] = value;
    if (!description.empty()) desc_[key] = description;
  }
  template <typename T>
  void set_default(const std::string& key, const T& value, const std::string& description,
                   const BoundConstraint<T>& c) {
    set_default<T>(key, value, description);
    bounds_[key] = {static_cast<double>(c.min), static_cast<double>(c.max)};
  }
  void set_default(const std::string& key, const std::string& value, const std::string& description,
                   const ListConstraint<std::string>& c) {
    set_default<std::string>(key, value, description);
    lists_[key] = c.allowed;
  }

 private:
  std::map<std::string, SettingValue> values_;
  std::map<std::string, std::string> desc_;
  std::map<std::string, std::pair<double, double>> bounds_;
  std::map<std::string, std::vector<std::string>> lists_;
  mutable bool locked_ = false;
};

template <typename T>
T Settings::get(const std::string& key) const {
  auto it = values_.find(key);
  if (it == values_.end()) throw SettingNotFound(key);
  const SettingValue& v = it->second;
  if constexpr (std::is_same_v<T, bool>) {
    if (auto p = std::get_if<bool>(&v)) return *p;
    throw SettingTypeMismatch(key, "bool");
  } else if constexpr (std::is_integral_v<T>) {
    if (auto p = std::get_if<int64_t>(&v)) {
      if constexpr (std::is_unsigned_v<T>)
        if (*p < 0) throw SettingTypeMismatch(key, "unsigned integer");
      return static_cast<T>(*p);
    }
    throw SettingTypeMismatch(key, "integer");
  } else if constexpr (std::is_floating_point_v<T>) {
    if (auto p = std::get_if<double>(&v)) return static_cast<T>(*p);
    if (auto p = std::get_if<int64_t>(&v)) return static_cast<T>(*p);
    throw SettingTypeMismatch(key, "double");
  } else {
    if (auto p = std::get_if<std::string>(&v)) return *p;
    throw SettingTypeMismatch(key, "string");
  }
}

// ---------------------------------------------------------------------------------------------
// Active-space Hamiltonian: restricted integrals as MACIS consumes them.
//   one_body : n x n (symmetric; row- or column-major is the same matrix)
//   two_body : n^4, element (pq|rs) at p n^3 + q n^2 + r n + s  -- identical, by the 8-fold
//              symmetry, to MACIS' column-major V(p,q,r,s) at p + q n + r n^2 + s n^3
//              (canonical_four_center.cpp:148-154, macis_cas.cpp:76-82)
class Hamiltonian {
 public:
  Hamiltonian(size_t norb, std::vector<double> one_body, std::vector<double> two_body, double core_energy,
              bool unrestricted = false);
  size_t num_active_orbitals() const { return norb_; }
  const std::vector<double>& get_one_body_integrals() const { return one_body_; }
  const std::vector<double>& get_two_body_integrals() const { return two_body_; }
  double get_core_energy() const { return core_energy_; }
  bool is_unrestricted() const { return unrestricted_; }
  std::string content_hash() const;

 private:
  size_t norb_;
  std::vector<double> one_body_, two_body_;
  double core_energy_;
  bool unrestricted_;
};

// One Slater determinant: alpha / beta occupation words (bit p = active orbital p), the two
// halves of macis::wfn_t<N> (external/macis/include/macis/wfn/raw_bitset.hpp:94-106).
class Configuration {
 public:
  Configuration() = default;
  Configuration(uint64_t alpha, uint64_t beta, size_t norb) : alpha_(alpha), beta_(beta), norb_(norb) {}
  // "2ud0..." one character per orbital, orbital 0 first (configuration.hpp to_string)
  explicit Configuration(const std::string& occupation);
  static Configuration from_spin_half_words(uint64_t alpha, uint64_t beta, size_t norb) {
    return Configuration(alpha, beta, norb);
  }
  std::string to_string() const;
  uint64_t alpha_word() const { return alpha_; }
  uint64_t beta_word() const { return beta_; }
  size_t get_orbital_capacity() const { return norb_; }
  std::pair<size_t, size_t> get_n_electrons() const {
    return {size_t(__builtin_popcountll(alpha_)), size_t(__builtin_popcountll(beta_))};
  }
  bool operator==(const Configuration& o) const { return alpha_ == o.alpha_ && beta_ == o.beta_; }

 private:
  uint64_t alpha_ = 0, beta_ = 0;
  size_t norb_ = 0;
};

class Wavefunction {
 public:
  Wavefunction(std::vector<double> coeffs, std::vector<Configuration> dets, size_t norb)
      : coeffs_(std::move(coeffs)), dets_(std::move(dets)), norb_(norb) {}
  size_t size() const { return dets_.size(); }
  size_t num_active_orbitals() const { return norb_; }
  const std::vector<double>& get_coefficients() const { return coeffs_; }
  const std::vector<Configuration>& get_active_determinants() const { return dets_; }
  double norm() const;
  // <this|other> over the determinants both hold
  double overlap(const Wavefunction& other) const;

  // ---- reduced density matrices over the active orbitals, as the reference's containers hold
  // them (cpp/include/qdk/chemistry/data/wavefunction.hpp:338-370,556-577): matrices n*n
  // column-major, two-body tensors n^4 with (p,q,r,s) at p + q n + r n^2 + s n^3.
  // Spin-dependent set (macis_base.hpp:199-215; the two-body blocks carry the adapter's factor 2).
  void set_rdms_spin_dependent(std::vector<double> one_aa, std::vector<double> one_bb, std::vector<double> two_aaaa,
                               std::vector<double> two_aabb, std::vector<double> two_bbbb);
  // Spin-traced set (macis_pmc.cpp:128-160)
  void set_rdms_spin_traced(std::vector<double> one, std::vector<double> two);
  bool has_one_rdm_spin_dependent() const { return !one_aa_.empty(); }
  bool has_two_rdm_spin_dependent() const { return !two_aaaa_.empty(); }
  bool has_one_rdm_spin_traced() const { return !one_st_.empty() || has_one_rdm_spin_dependent(); }
  bool has_two_rdm_spin_traced() const { return !two_st_.empty() || has_two_rdm_spin_dependent(); }
  // (aa, bb)
  std::pair<std::vector<double>, std::vector<double>> get_active_one_rdm_spin_dependent() const;
  // (aaaa, aabb, bbbb)
  std::tuple<std::vector<double>, std::vector<double>, std::vector<double>> get_active_two_rdm_spin_dependent() const;
  // stored, or derived: aa + bb;  aaaa + bbbb + aabb + aabb^T(pq<->rs)  (wavefunction.cpp:255-340)
  std::vector<double> get_active_one_rdm_spin_traced() const;
  std::vector<double> get_active_two_rdm_spin_traced() const;

  // ---- orbital entropies (OrbitalEntropies, wavefunction.hpp:165-180,380-423): vector of n,
  // matrices n x n column-major; empty = not computed
  void set_entropies(std::vector<double> single_orbital, std::vector<double> two_orbital,
                     std::vector<double> mutual_information);
  bool has_single_orbital_entropies() const { return !s1_.empty(); }
  bool has_two_orbital_entropies() const { return !s2_.empty(); }
  bool has_mutual_information() const { return !mi_.empty(); }
  const std::vector<double>& get_single_orbital_entropies() const;
  const std::vector<double>& get_two_orbital_entropies() const;
  const std::vector<double>& get_mutual_information() const;

 private:
  std::vector<double> coeffs_;
  std::vector<Configuration> dets_;
  size_t norb_;
  std::vector<double> one_aa_, one_bb_, two_aaaa_, two_aabb_, two_bbbb_, one_st_, two_st_;
  std::vector<double> s1_, s2_, mi_;
};

}  // namespace qdk_b200::data
