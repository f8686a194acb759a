// Algorithm base class and per-type factory -- the plugin API of QDK/Chemistry
// (cpp/include/qdk/chemistry/algorithms/algorithm.hpp:53-131 Algorithm, :232-417
// AlgorithmFactory). Same member names, same registration rules: an instance is registered
// under name() and every alias, the type name is checked, duplicates throw, create("") returns
// the default algorithm, run() locks the settings before delegating to _run_impl().
#pragma once
#include <functional>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "data.hpp"

namespace qdk_b200::algorithms {

class DuplicateRegistrationError : public std::runtime_error {
 public:
  using std::runtime_error::runtime_error;
};

std::string hash_hex(const std::string& bytes);  // FNV-1a 128-bit style digest, hex

template <typename Derived, typename ReturnType, typename... Args>
class Algorithm {
 public:
  Algorithm() = default;
  virtual ~Algorithm() = default;

  virtual ReturnType run(Args... args) const {
    this->lock_settings();
    return this->_run_impl(std::forward<Args>(args)...);
  }
  data::Settings& settings() { return *_settings; }
  const data::Settings& settings() const { return *_settings; }
  virtual std::string name() const = 0;
  virtual std::vector<std::string> aliases() const { return {this->name()}; }
  virtual std::string type_name() const = 0;

 protected:
  void lock_settings() const { this->_settings->lock(); }
  virtual ReturnType _run_impl(Args... args) const = 0;
  std::unique_ptr<data::Settings> _settings = std::make_unique<data::Settings>();
};

template <typename BaseAlgorithmType, typename Derived>
class AlgorithmFactory {
 public:
  using return_type = std::unique_ptr<BaseAlgorithmType>;
  using functor_type = std::function<return_type(void)>;

  static return_type create(const std::string& name = "") {
    std::string key = name.empty() ? Derived::default_algorithm_name() : name;
    auto& reg = registry();
    auto it = reg.find(key);
    if (it == reg.end()) {
      std::string avail;
      for (const auto& [k, _] : reg) avail += (avail.empty() ? "" : ", ") + k;
      throw std::runtime_error("Algorithm factory for " + Derived::algorithm_type_name() +
                               ": Algorithm with name '" + key +
                               "' not found in registry, available options are: " + avail);
    }
    auto instance = it->second();
    if (!instance)
      throw std::runtime_error("Algorithm factory for " + Derived::algorithm_type_name() +
                               ": Algorithm with name '" + key + "' returned nullptr");
    return instance;
  }
  static void register_instance(functor_type func) {
    auto& reg = registry();
    auto tmp = func();
    if (!tmp) throw std::runtime_error("register_instance: functor returned nullptr");
    if (tmp->type_name() != Derived::algorithm_type_name())
      throw std::runtime_error("Algorithm factory for " + Derived::algorithm_type_name() +
                               ": algorithm with name '" + tmp->name() +
                               "' has incorrect algorithm type: " + tmp->type_name() +
                               " expected is: " + Derived::algorithm_type_name());
    auto aliases = tmp->aliases();
    for (const auto& a : aliases)
      if (reg.find(a) != reg.end())
        throw DuplicateRegistrationError("Algorithm factory for " + Derived::algorithm_type_name() +
                                         ": algorithm with name/alias '" + a +
                                         "' already exists in registry");
    for (const auto& a : aliases) reg[a] = func;
  }
  static bool unregister_instance(const std::string& key) { return registry().erase(key) != 0; }
  static std::vector<std::string> available() {
    std::vector<std::string> keys;
    for (const auto& [k, _] : registry()) keys.push_back(k);
    return keys;
  }
  static bool has(const std::string& key) { return registry().count(key) != 0; }
  static void clear() { registry().clear(); }

 protected:
  static std::unordered_map<std::string, functor_type>& registry() {
    static std::unordered_map<std::string, functor_type> instance;
    static bool initialized = false;
    if (!initialized) {
      initialized = true;
      Derived::register_default_instances();
    }
    return instance;
  }
};

}  // namespace qdk_b200::algorithms
