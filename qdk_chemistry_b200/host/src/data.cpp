// Settings / Hamiltonian / Configuration / Wavefunction stand-ins (see include/qdk_b200/data.hpp).
#include <stdexcept>
#include "qdk_b200/data.hpp"

#include <cmath>
#include <cstring>
#include <sstream>
#include <unordered_map>

#include "qdk_b200/algorithm.hpp"

namespace qdk_b200::algorithms {
// 2 x 64-bit FNV-1a lanes over the byte string; stable across runs and platforms
std::string hash_hex(const std::string& bytes) {
  uint64_t a = 0xcbf29ce484222325ull, b = 0x84222325cbf29ce4ull;
  for (unsigned char c : bytes) {
    a = (a ^ c) * 0x100000001b3ull;
    b = (b ^ (c + 0x9eu)) * 0x100000001b3ull;
    b = (b << 13) | (b >> 51);
  }
  char buf[33];
  snprintf(buf, sizeof(buf), "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
  return buf;
}
}  // namespace qdk_b200::algorithms

namespace qdk_b200::data {

namespace {
const char* type_name_of(const SettingValue& v) {
  switch (v.index()) {
    case 0: return "bool";
    case 1: return "int64";
    case 2: return "double";
    default: return "string";
  }
}
}  // namespace

void Settings::set(const std::string& key, const SettingValue& value) {
  if (locked_) throw SettingsAreLocked();
  auto it = values_.find(key);
  if (it == values_.end()) throw SettingNotFound(key);
  SettingValue v = value;
  // integers are accepted for floating-point settings (Python ints), nothing else converts
  if (it->second.index() == 2 && v.index() == 1) v = static_cast<double>(std::get<int64_t>(v));
  if (it->second.index() != v.index()) throw SettingTypeMismatch(key, type_name_of(it->second));
  auto b = bounds_.find(key);
  if (b != bounds_.end()) {
    const double x = v.index() == 1 ? static_cast<double>(std::get<int64_t>(v)) : std::get<double>(v);
    if (!(x >= b->second.first && x <= b->second.second)) {
      std::ostringstream os;
      os << "Value " << x << " for setting '" << key << "' is outside [" << b->second.first << ", "
         << b->second.second << "]";
      throw std::invalid_argument(os.str());
    }
  }
  auto l = lists_.find(key);
  if (l != lists_.end()) {
    const std::string& sv = std::get<std::string>(v);
    bool ok = false;
    for (const auto& a : l->second) ok = ok || a == sv;
    if (!ok) throw std::invalid_argument("Value '" + sv + "' is not allowed for setting '" + key + "'");
  }
  it->second = v;
}

const SettingValue& Settings::get_raw(const std::string& key) const {
  auto it = values_.find(key);
  if (it == values_.end()) throw SettingNotFound(key);
  return it->second;
}
std::vector<std::string> Settings::keys() const {
  std::vector<std::string> k;
  for (const auto& kv : values_) k.push_back(kv.first);
  return k;
}
std::string Settings::get_as_string(const std::string& key) const {
  const SettingValue& v = get_raw(key);
  std::ostringstream os;
  os.precision(17);
  switch (v.index()) {
    case 0: os << (std::get<bool>(v) ? "true" : "false"); break;
    case 1: os << std::get<int64_t>(v); break;
    case 2: os << std::get<double>(v); break;
    default: os << std::get<std::string>(v);
  }
  return os.str();
}
std::string Settings::get_type_name(const std::string& key) const { return type_name_of(get_raw(key)); }
std::string Settings::get_description(const std::string& key) const {
  auto it = desc_.find(key);
  if (it == desc_.end()) throw SettingNotFound(key);
  return it->second;
}
std::string Settings::content_hash() const {
  std::string bytes;
  for (const auto& kv : values_) bytes += kv.first + "=" + get_as_string(kv.first) + ";";
  return algorithms::hash_hex(bytes);
}

Hamiltonian::Hamiltonian(size_t norb, std::vector<double> one_body, std::vector<double> two_body,
                         double core_energy, bool unrestricted)
    : norb_(norb), one_body_(std::move(one_body)), two_body_(std::move(two_body)),
      core_energy_(core_energy), unrestricted_(unrestricted) {
  if (norb_ == 0) throw std::invalid_argument("Hamiltonian: no active orbitals");
  if (one_body_.size() != norb_ * norb_)
    throw std::invalid_argument("Hamiltonian: one-body integrals must have norb^2 entries");
  if (two_body_.size() != norb_ * norb_ * norb_ * norb_)
    throw std::invalid_argument("Hamiltonian: two-body integrals must have norb^4 entries");
}
std::string Hamiltonian::content_hash() const {
  std::string bytes(reinterpret_cast<const char*>(one_body_.data()), one_body_.size() * 8);
  bytes.append(reinterpret_cast<const char*>(two_body_.data()), two_body_.size() * 8);
  bytes.append(reinterpret_cast<const char*>(&core_energy_), 8);
  return algorithms::hash_hex(bytes);
}

Configuration::Configuration(const std::string& occ) : norb_(occ.size()) {
  if (occ.size() > 64) throw std::invalid_argument("Configuration: more than 64 orbitals");
  for (size_t p = 0; p < occ.size(); ++p) {
    switch (occ[p]) {
      case '2': alpha_ |= uint64_t(1) << p; beta_ |= uint64_t(1) << p; break;
      case 'u': alpha_ |= uint64_t(1) << p; break;
      case 'd': beta_ |= uint64_t(1) << p; break;
      case '0': break;
      default: throw std::invalid_argument("Configuration: occupation characters are '2', 'u', 'd', '0'");
    }
  }
}
std::string Configuration::to_string() const {
  std::string s(norb_, '0');
  for (size_t p = 0; p < norb_; ++p) {
    const bool a = (alpha_ >> p) & 1, b = (beta_ >> p) & 1;
    s[p] = a && b ? '2' : (a ? 'u' : (b ? 'd' : '0'));
  }
  return s;
}

double Wavefunction::norm() const {
  double s = 0.;
  for (double c : coeffs_) s += c * c;
  return std::sqrt(s);
}
double Wavefunction::overlap(const Wavefunction& other) const {
  struct H {
    size_t operator()(const std::pair<uint64_t, uint64_t>& p) const {
      return std::hash<uint64_t>()(p.first * 0x9e3779b97f4a7c15ull ^ p.second);
    }
  };
  std::unordered_map<std::pair<uint64_t, uint64_t>, double, H> m;
  m.reserve(dets_.size());
  for (size_t i = 0; i < dets_.size(); ++i) m[{dets_[i].alpha_word(), dets_[i].beta_word()}] = coeffs_[i];
  double s = 0.;
  for (size_t i = 0; i < other.dets_.size(); ++i) {
    auto it = m.find({other.dets_[i].alpha_word(), other.dets_[i].beta_word()});
    if (it != m.end()) s += it->second * other.coeffs_[i];
  }
  return s;
}

void Wavefunction::set_rdms_spin_dependent(std::vector<double> one_aa, std::vector<double> one_bb,
                                           std::vector<double> two_aaaa, std::vector<double> two_aabb,
                                           std::vector<double> two_bbbb) {
  one_aa_ = std::move(one_aa); one_bb_ = std::move(one_bb);
  two_aaaa_ = std::move(two_aaaa); two_aabb_ = std::move(two_aabb); two_bbbb_ = std::move(two_bbbb);
}
void Wavefunction::set_rdms_spin_traced(std::vector<double> one, std::vector<double> two) {
  one_st_ = std::move(one);
  two_st_ = std::move(two);
}
std::pair<std::vector<double>, std::vector<double>> Wavefunction::get_active_one_rdm_spin_dependent() const {
  if (!has_one_rdm_spin_dependent()) throw std::runtime_error("Spin-dependent one-body RDM not set");
  return {one_aa_, one_bb_};
}
std::tuple<std::vector<double>, std::vector<double>, std::vector<double>>
Wavefunction::get_active_two_rdm_spin_dependent() const {
  if (!has_two_rdm_spin_dependent()) throw std::runtime_error("Spin-dependent two-body RDM not set");
  return {two_aaaa_, two_aabb_, two_bbbb_};
}
std::vector<double> Wavefunction::get_active_one_rdm_spin_traced() const {
  if (!one_st_.empty()) return one_st_;
  if (!has_one_rdm_spin_dependent()) throw std::runtime_error("Spin-traced one-body RDM not set");
  std::vector<double> r(one_aa_);
  for (size_t i = 0; i < r.size(); ++i) r[i] += one_bb_[i];
  return r;
}
std::vector<double> Wavefunction::get_active_two_rdm_spin_traced() const {
  if (!two_st_.empty()) return two_st_;
  if (!has_two_rdm_spin_dependent()) throw std::runtime_error("Spin-traced two-body RDM not set");
  const size_t n = norb_, n2 = n * n;
  std::vector<double> r(two_aaaa_);
  for (size_t pq = 0; pq < n2; ++pq)
    for (size_t rs = 0; rs < n2; ++rs)
      r[pq + rs * n2] += two_bbbb_[pq + rs * n2] + two_aabb_[pq + rs * n2] + two_aabb_[rs + pq * n2];
  return r;
}

void Wavefunction::set_entropies(std::vector<double> single_orbital, std::vector<double> two_orbital,
                                 std::vector<double> mutual_information) {
  s1_ = std::move(single_orbital);
  s2_ = std::move(two_orbital);
  mi_ = std::move(mutual_information);
}
const std::vector<double>& Wavefunction::get_single_orbital_entropies() const {
  if (s1_.empty()) throw std::runtime_error("Single orbital entropies not available");
  return s1_;
}
const std::vector<double>& Wavefunction::get_two_orbital_entropies() const {
  if (s2_.empty()) throw std::runtime_error("Two orbital entropies not available");
  return s2_;
}
const std::vector<double>& Wavefunction::get_mutual_information() const {
  if (mi_.empty()) throw std::runtime_error("Mutual information not available");
  return mi_;
}

}  // namespace qdk_b200::data
