// pybind11 surface mirroring the part of qdk_chemistry._core that belongs to the CI path:
//   algorithms.MultiConfigurationCalculator{run, settings, name, type_name, hash} with a
//     trampoline so Python subclasses can be registered   (python/src/pybind11/algorithms/mc.cpp:20-163)
//   algorithms.MultiConfigurationCalculatorFactory statics (factory_bindings.hpp:111-268)
//   algorithms.ProjectedMultiConfigurationCalculator(+Factory) (pmc.cpp)
//   algorithms.davidson_solver(csr, tol=1e-8, max_m=20)      (davidson_solver.cpp:60-107)
//   data.Settings / Hamiltonian / Configuration / Wavefunction stand-ins
#include <pybind11/functional.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "qdk_b200/fcidump.hpp"
#include "qdk_b200/mc.hpp"

namespace py = pybind11;
using namespace qdk_b200;
using namespace qdk_b200::algorithms;

namespace {

py::object setting_to_python(const data::SettingValue& v) {
  switch (v.index()) {
    case 0: return py::bool_(std::get<bool>(v));
    case 1: return py::int_(std::get<int64_t>(v));
    case 2: return py::float_(std::get<double>(v));
    default: return py::str(std::get<std::string>(v));
  }
}
data::SettingValue setting_from_python(const py::handle& o) {
  if (py::isinstance<py::bool_>(o)) return o.cast<bool>();
  if (py::isinstance<py::int_>(o)) return o.cast<int64_t>();
  if (py::isinstance<py::float_>(o)) return o.cast<double>();
  if (py::isinstance<py::str>(o)) return o.cast<std::string>();
  throw py::type_error("settings values are bool, int, float or str");
}

// Python subclasses: settings come from a plain Settings object that can be replaced via _settings
class PySettings : public data::Settings {
 public:
  void declare(const std::string& key, const data::SettingValue& v, const std::string& desc) {
    std::visit([&](const auto& x) { this->set_default(key, x, desc); }, v);
  }
};

class MultiConfigurationCalculatorBase : public MultiConfigurationCalculator, public py::trampoline_self_life_support {
 public:
  MultiConfigurationCalculatorBase() { _settings = std::make_unique<PySettings>(); }
  std::string name() const override { PYBIND11_OVERRIDE_PURE(std::string, MultiConfigurationCalculator, name); }
  std::vector<std::string> aliases() const override {
    PYBIND11_OVERRIDE(std::vector<std::string>, MultiConfigurationCalculator, aliases);
  }
  McResult _run_impl(std::shared_ptr<data::Hamiltonian> h, unsigned na, unsigned nb) const override {
    PYBIND11_OVERRIDE_PURE(McResult, MultiConfigurationCalculator, _run_impl, h, na, nb);
  }
  using MultiConfigurationCalculator::_run_impl;
};

}  // namespace

PYBIND11_MODULE(_core, m) {
  m.doc() = "B200-native CI path behind the QDK/Chemistry MultiConfigurationCalculator plugin API";
  auto dmod = m.def_submodule("data");
  auto amod = m.def_submodule("algorithms");

  py::register_exception<data::SettingsAreLocked>(dmod, "SettingsAreLocked", PyExc_RuntimeError);
  py::register_exception<data::SettingNotFound>(dmod, "SettingNotFound", PyExc_KeyError);
  py::register_exception<data::SettingTypeMismatch>(dmod, "SettingTypeMismatch", PyExc_TypeError);
  py::register_exception<DuplicateRegistrationError>(amod, "DuplicateRegistrationError", PyExc_RuntimeError);

  py::class_<data::Settings>(dmod, "Settings")
      .def("set", [](data::Settings& s, const std::string& k, py::object v) { s.set(k, setting_from_python(v)); })
      .def("get", [](const data::Settings& s, const std::string& k) { return setting_to_python(s.get_raw(k)); })
      .def("get_or_default",
           [](const data::Settings& s, const std::string& k, py::object d) {
             return s.has(k) ? setting_to_python(s.get_raw(k)) : d;
           })
      .def("update",
           [](data::Settings& s, const py::dict& d) {
             for (auto kv : d) s.set(kv.first.cast<std::string>(), setting_from_python(kv.second));
           })
      .def("has", &data::Settings::has)
      .def("keys", &data::Settings::keys)
      .def("size", &data::Settings::size)
      .def("empty", &data::Settings::empty)
      .def("get_as_string", &data::Settings::get_as_string)
      .def("get_type_name", &data::Settings::get_type_name)
      .def("has_description", &data::Settings::has_description)
      .def("get_description", &data::Settings::get_description)
      .def("lock", &data::Settings::lock)
      .def("is_locked", &data::Settings::is_locked)
      .def("to_dict",
           [](const data::Settings& s) {
             py::dict d;
             for (const auto& k : s.keys()) d[py::str(k)] = setting_to_python(s.get_raw(k));
             return d;
           })
      .def("__contains__", &data::Settings::has)
      .def("__getitem__", [](const data::Settings& s, const std::string& k) { return setting_to_python(s.get_raw(k)); })
      .def("__setitem__", [](data::Settings& s, const std::string& k, py::object v) { s.set(k, setting_from_python(v)); })
      .def("__len__", &data::Settings::size);
  py::class_<PySettings, data::Settings>(dmod, "UserSettings")
      .def(py::init<>())
      .def("set_default",
           [](PySettings& s, const std::string& k, py::object v, const std::string& desc) {
             s.declare(k, setting_from_python(v), desc);
           },
           py::arg("key"), py::arg("value"), py::arg("description") = "");

  py::class_<data::Hamiltonian, std::shared_ptr<data::Hamiltonian>>(dmod, "Hamiltonian")
      .def(py::init([](py::array_t<double, py::array::c_style | py::array::forcecast> one,
                       py::array_t<double, py::array::c_style | py::array::forcecast> two, double core, bool unres) {
             const size_t n = size_t(std::llround(std::sqrt(double(one.size()))));
             std::vector<double> T(one.data(), one.data() + one.size()), V(two.data(), two.data() + two.size());
             return std::make_shared<data::Hamiltonian>(n, std::move(T), std::move(V), core, unres);
           }),
           py::arg("one_body_integrals"), py::arg("two_body_integrals"), py::arg("core_energy") = 0.0,
           py::arg("unrestricted") = false)
      .def("num_active_orbitals", &data::Hamiltonian::num_active_orbitals)
      .def("get_core_energy", &data::Hamiltonian::get_core_energy)
      .def("is_unrestricted", &data::Hamiltonian::is_unrestricted)
      .def("get_one_body_integrals",
           [](const data::Hamiltonian& h) {
             const auto& v = h.get_one_body_integrals();
             const py::ssize_t n = py::ssize_t(h.num_active_orbitals());
             py::array_t<double> a({n, n});
             std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
             return a;
           })
      .def("get_two_body_integrals", [](const data::Hamiltonian& h) {
        const auto& v = h.get_two_body_integrals();
        py::array_t<double> a(py::ssize_t(v.size()));
        std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
        return a;
      });

  py::class_<data::Configuration>(dmod, "Configuration")
      .def(py::init<const std::string&>())
      .def(py::init<uint64_t, uint64_t, size_t>(), py::arg("alpha"), py::arg("beta"), py::arg("num_orbitals"))
      .def("to_string", &data::Configuration::to_string)
      .def("alpha_word", &data::Configuration::alpha_word)
      .def("beta_word", &data::Configuration::beta_word)
      .def("get_orbital_capacity", &data::Configuration::get_orbital_capacity)
      .def("get_n_electrons", &data::Configuration::get_n_electrons)
      .def("__eq__", &data::Configuration::operator==)
      .def("__repr__", [](const data::Configuration& c) { return "Configuration('" + c.to_string() + "')"; });

  py::class_<data::Wavefunction, std::shared_ptr<data::Wavefunction>>(dmod, "Wavefunction")
      .def("size", &data::Wavefunction::size)
      .def("__len__", &data::Wavefunction::size)
      .def("num_active_orbitals", &data::Wavefunction::num_active_orbitals)
      .def("norm", &data::Wavefunction::norm)
      .def("overlap", &data::Wavefunction::overlap)
      .def("get_coefficients",
           [](const data::Wavefunction& w) {
             const auto& c = w.get_coefficients();
             py::array_t<double> a(py::ssize_t(c.size()));
             std::memcpy(a.mutable_data(), c.data(), c.size() * 8);
             return a;
           })
      .def("has_one_rdm_spin_dependent", &data::Wavefunction::has_one_rdm_spin_dependent)
      .def("has_two_rdm_spin_dependent", &data::Wavefunction::has_two_rdm_spin_dependent)
      .def("has_one_rdm_spin_traced", &data::Wavefunction::has_one_rdm_spin_traced)
      .def("has_two_rdm_spin_traced", &data::Wavefunction::has_two_rdm_spin_traced)
      .def("get_active_one_rdm_spin_dependent",
           [](const data::Wavefunction& w) {
             auto r = w.get_active_one_rdm_spin_dependent();
             const py::ssize_t n = py::ssize_t(w.num_active_orbitals());
             auto mk = [n](const std::vector<double>& v) {  // column-major n x n
               py::array_t<double, py::array::f_style> a({n, n});
               std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
               return a;
             };
             return py::make_tuple(mk(r.first), mk(r.second));
           })
      .def("get_active_two_rdm_spin_dependent",
           [](const data::Wavefunction& w) {  // (aaaa, aabb, bbbb), flat n^4 like the reference's VectorXd
             auto r = w.get_active_two_rdm_spin_dependent();
             auto mk = [](const std::vector<double>& v) {
               py::array_t<double> a{py::ssize_t(v.size())};
               std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
               return a;
             };
             return py::make_tuple(mk(std::get<0>(r)), mk(std::get<1>(r)), mk(std::get<2>(r)));
           })
      .def("get_active_one_rdm_spin_traced",
           [](const data::Wavefunction& w) {
             auto v = w.get_active_one_rdm_spin_traced();
             const py::ssize_t n = py::ssize_t(w.num_active_orbitals());
             py::array_t<double, py::array::f_style> a({n, n});
             std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
             return a;
           })
      .def("get_active_two_rdm_spin_traced",
           [](const data::Wavefunction& w) {
             auto v = w.get_active_two_rdm_spin_traced();
             py::array_t<double> a{py::ssize_t(v.size())};
             std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
             return a;
           })
      .def("has_single_orbital_entropies", &data::Wavefunction::has_single_orbital_entropies)
      .def("has_two_orbital_entropies", &data::Wavefunction::has_two_orbital_entropies)
      .def("has_mutual_information", &data::Wavefunction::has_mutual_information)
      .def("get_single_orbital_entropies",
           [](const data::Wavefunction& w) {
             const auto& v = w.get_single_orbital_entropies();
             py::array_t<double> a{py::ssize_t(v.size())};
             std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
             return a;
           })
      .def("get_two_orbital_entropies",
           [](const data::Wavefunction& w) {
             const auto& v = w.get_two_orbital_entropies();
             const py::ssize_t n = py::ssize_t(w.num_active_orbitals());
             py::array_t<double, py::array::f_style> a({n, n});
             std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
             return a;
           })
      .def("get_mutual_information",
           [](const data::Wavefunction& w) {
             const auto& v = w.get_mutual_information();
             const py::ssize_t n = py::ssize_t(w.num_active_orbitals());
             py::array_t<double, py::array::f_style> a({n, n});
             std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
             return a;
           })
      .def("get_active_determinants", &data::Wavefunction::get_active_determinants)
      .def("determinant_words", [](const data::Wavefunction& w) {
        // (alpha, beta) occupation words, shape (n, 2)
        const auto& d = w.get_active_determinants();
        py::array_t<uint64_t> a({py::ssize_t(d.size()), py::ssize_t(2)});
        auto r = a.mutable_unchecked<2>();
        for (py::ssize_t i = 0; i < py::ssize_t(d.size()); ++i) { r(i, 0) = d[i].alpha_word(); r(i, 1) = d[i].beta_word(); }
        return a;
      });

  // ---- MultiConfigurationCalculator
  py::class_<MultiConfigurationCalculator, MultiConfigurationCalculatorBase, py::smart_holder>(
      amod, "MultiConfigurationCalculator")
      .def(py::init<>())
      .def("run", &MultiConfigurationCalculator::run, py::arg("hamiltonian"), py::arg("n_active_alpha_electrons"),
           py::arg("n_active_beta_electrons"), py::call_guard<py::gil_scoped_release>())
      .def("settings", [](MultiConfigurationCalculator& c) -> data::Settings& { return c.settings(); },
           py::return_value_policy::reference_internal)
      .def("name", &MultiConfigurationCalculator::name)
      .def("aliases", &MultiConfigurationCalculator::aliases)
      .def("type_name", &MultiConfigurationCalculator::type_name)
      .def("hash", &MultiConfigurationCalculator::hash, py::arg("hamiltonian"), py::arg("n_active_alpha_electrons"),
           py::arg("n_active_beta_electrons"))
      .def("__repr__", [](const MultiConfigurationCalculator&) {
        return "<qdk_chemistry_b200.algorithms.MultiConfigurationCalculator>";
      });
  py::class_<B200Cas, MultiConfigurationCalculator, py::smart_holder>(amod, "B200Cas").def(py::init<>());
  py::class_<B200Asci, MultiConfigurationCalculator, py::smart_holder>(amod, "B200Asci").def(py::init<>());

  using F = MultiConfigurationCalculatorFactory;
  py::class_<F>(amod, "MultiConfigurationCalculatorFactory")
      .def_static("create", [](const std::string& name) { return F::create(name); }, py::arg("name") = "")
      .def_static("available", &F::available)
      .def_static("register_instance",
                  [](py::function fn) {
                    // the Python callable returns a MultiConfigurationCalculator (subclass) instance
                    F::register_instance([fn]() -> std::unique_ptr<MultiConfigurationCalculator> {
                      py::gil_scoped_acquire gil;
                      py::object obj = fn();
                      return obj.cast<std::unique_ptr<MultiConfigurationCalculator>>();
                    });
                  },
                  py::arg("func"))
      .def_static("unregister_instance", &F::unregister_instance, py::arg("key"))
      .def_static("algorithm_type_name", &F::algorithm_type_name)
      .def_static("default_algorithm_name", &F::default_algorithm_name)
      .def_static("clear", &F::clear)
      .def_static("has", &F::has, py::arg("key"));

  // ---- ProjectedMultiConfigurationCalculator
  py::class_<ProjectedMultiConfigurationCalculator, py::smart_holder>(amod, "ProjectedMultiConfigurationCalculator")
      .def("run", &ProjectedMultiConfigurationCalculator::run, py::arg("hamiltonian"), py::arg("configurations"),
           py::call_guard<py::gil_scoped_release>())
      .def("settings", [](ProjectedMultiConfigurationCalculator& c) -> data::Settings& { return c.settings(); },
           py::return_value_policy::reference_internal)
      .def("name", &ProjectedMultiConfigurationCalculator::name)
      .def("type_name", &ProjectedMultiConfigurationCalculator::type_name);
  py::class_<B200Pmc, ProjectedMultiConfigurationCalculator, py::smart_holder>(amod, "B200Pmc").def(py::init<>());
  using PF = ProjectedMultiConfigurationCalculatorFactory;
  py::class_<PF>(amod, "ProjectedMultiConfigurationCalculatorFactory")
      .def_static("create", [](const std::string& name) { return PF::create(name); }, py::arg("name") = "")
      .def_static("available", &PF::available)
      .def_static("unregister_instance", &PF::unregister_instance, py::arg("key"))
      .def_static("algorithm_type_name", &PF::algorithm_type_name)
      .def_static("default_algorithm_name", &PF::default_algorithm_name)
      .def_static("clear", &PF::clear)
      .def_static("has", &PF::has, py::arg("key"));

  amod.def(
      "davidson_solver",
      [](const py::object& csr, double tol, int64_t max_m) {
        auto data = csr.attr("data").cast<py::array_t<double, py::array::c_style | py::array::forcecast>>();
        auto indices = csr.attr("indices").cast<py::array_t<int64_t, py::array::c_style | py::array::forcecast>>();
        auto indptr = csr.attr("indptr").cast<py::array_t<int64_t, py::array::c_style | py::array::forcecast>>();
        auto shape = csr.attr("shape").cast<std::pair<int64_t, int64_t>>();
        if (shape.first != shape.second) throw std::invalid_argument("davidson_solver: matrix must be square");
        if (indptr.size() != shape.first + 1) throw std::invalid_argument("davidson_solver: bad indptr length");
        std::pair<double, std::vector<double>> r;
        {
          py::gil_scoped_release nogil;
          r = davidson_solver(shape.first, indptr.data(), indices.data(), data.data(), tol, max_m);
        }
        py::array_t<double> x(py::ssize_t(r.second.size()));
        std::memcpy(x.mutable_data(), r.second.data(), r.second.size() * 8);
        return py::make_tuple(r.first, x);
      },
      py::arg("csr_matrix"), py::arg("tol") = 1e-8, py::arg("max_m") = 20);

  // ---- compute_casci_rdms (mcscf/cas.hpp:33-64): (E0, C, ordm | None, trdm | None)
  amod.def(
      "compute_casci_rdms",
      [](size_t norb, size_t nalpha, size_t nbeta, py::array_t<double, py::array::f_style | py::array::forcecast> T,
         py::array_t<double, py::array::f_style | py::array::forcecast> V, bool rdms, double ci_res_tol,
         size_t ci_max_subspace, double ci_matel_tol) {
        if (size_t(T.size()) != norb * norb || size_t(V.size()) != norb * norb * norb * norb)
          throw std::invalid_argument("compute_casci_rdms: T must hold norb^2 and V norb^4 elements");
        MCSCFSettings st;
        st.ci_res_tol = ci_res_tol;
        st.ci_max_subspace = ci_max_subspace;
        st.ci_matel_tol = ci_matel_tol;
        std::vector<double> C, o1, t1;
        if (rdms) { o1.assign(norb * norb, 0.0); t1.assign(norb * norb * norb * norb, 0.0); }
        double E0;
        {
          py::gil_scoped_release nogil;
          E0 = CASRDMFunctor::rdms(st, norb, nalpha, nbeta, T.data(), V.data(), rdms ? o1.data() : nullptr,
                                   rdms ? t1.data() : nullptr, C);
        }
        auto arr = [](const std::vector<double>& v) {
          py::array_t<double> a{py::ssize_t(v.size())};
          std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
          return a;
        };
        return py::make_tuple(E0, arr(C), rdms ? py::object(arr(o1)) : py::object(py::none()),
                              rdms ? py::object(arr(t1)) : py::object(py::none()));
      },
      py::arg("norb"), py::arg("nalpha"), py::arg("nbeta"), py::arg("T"), py::arg("V"), py::arg("rdms") = true,
      py::arg("ci_res_tol") = 1e-8, py::arg("ci_max_subspace") = 20,
      py::arg("ci_matel_tol") = std::numeric_limits<double>::epsilon());

  // ---- FCIDUMP / binary RDM files (macis/util/fcidump.hpp)
  py::module_ io = m.def_submodule("io", "FCIDUMP and RDM file formats shared with MACIS");
  io.def("fcidump_read_header", [](const std::string& f) {
    const io::FCIDumpHeader h = io::fcidump_read_header(f);
    py::dict d;
    d["norb"] = h.norb; d["nelec"] = h.nelec; d["ms2"] = h.ms2; d["isym"] = h.isym; d["orbsym"] = h.orbsym;
    return d;
  });
  io.def("read_fcidump_norb", &io::read_fcidump_norb);
  io.def("read_fcidump_core", &io::read_fcidump_core);
  io.def("read_fcidump_all", [](const std::string& f) {
    const size_t n = io::read_fcidump_norb(f);
    py::array_t<double> T{py::ssize_t(n * n)}, V{py::ssize_t(n * n * n * n)};
    double core = 0.;
    io::read_fcidump_all(f, T.mutable_data(), n, V.mutable_data(), n, core);
    return py::make_tuple(T, V, core);   // flat, column-major
  });
  io.def("read_fcidump_1body", [](const std::string& f) {
    const size_t n = io::read_fcidump_norb(f);
    py::array_t<double> T{py::ssize_t(n * n)};
    std::fill(T.mutable_data(), T.mutable_data() + n * n, 0.0);
    io::read_fcidump_1body(f, T.mutable_data(), n);
    return T;
  });
  io.def("read_fcidump_2body", [](const std::string& f) {
    const size_t n = io::read_fcidump_norb(f);
    py::array_t<double> V{py::ssize_t(n * n * n * n)};
    std::fill(V.mutable_data(), V.mutable_data() + n * n * n * n, 0.0);
    io::read_fcidump_2body(f, V.mutable_data(), n);
    return V;
  });
  io.def(
      "write_fcidump",
      [](const std::string& f, uint32_t norb, uint32_t nelec, int32_t ms2,
         py::array_t<double, py::array::f_style | py::array::forcecast> T,
         py::array_t<double, py::array::f_style | py::array::forcecast> V, double core, double threshold) {
        if (size_t(T.size()) != size_t(norb) * norb || size_t(V.size()) != size_t(norb) * norb * norb * norb)
          throw std::invalid_argument("write_fcidump: T must hold norb^2 and V norb^4 elements");
        io::FCIDumpHeader h;
        h.norb = norb; h.nelec = nelec; h.ms2 = ms2; h.isym = 1;
        h.orbsym.assign(norb, 1);
        io::write_fcidump(f, h, T.data(), norb, V.data(), norb, core, threshold);
      },
      py::arg("fname"), py::arg("norb"), py::arg("nelec"), py::arg("ms2"), py::arg("T"), py::arg("V"),
      py::arg("core_energy"), py::arg("threshold") = 1e-15);
  io.def("read_rdms_binary", [](const std::string& f, size_t norb) {
    py::array_t<double> o{py::ssize_t(norb * norb)}, t{py::ssize_t(norb * norb * norb * norb)};
    io::read_rdms_binary(f, norb, o.mutable_data(), norb, t.mutable_data(), norb);
    return py::make_tuple(o, t);
  });
  io.def("write_rdms_binary", [](const std::string& f, size_t norb,
                                 py::array_t<double, py::array::f_style | py::array::forcecast> o,
                                 py::array_t<double, py::array::f_style | py::array::forcecast> t) {
    if (size_t(o.size()) != norb * norb || size_t(t.size()) != norb * norb * norb * norb)
      throw std::invalid_argument("write_rdms_binary: bad array sizes");
    io::write_rdms_binary(f, norb, o.data(), norb, t.data(), norb);
  });

  io.def("to_canonical_string", &io::to_canonical_string, py::arg("alpha"), py::arg("beta"), py::arg("norb"));
  io.def("from_canonical_string", &io::from_canonical_string, py::arg("string"));
  io.def("read_wavefunction", [](const std::string& f) {
    const io::WavefunctionFile w = io::read_wavefunction(f);
    auto u64 = [](const std::vector<uint64_t>& v) {
      py::array_t<uint64_t> a{py::ssize_t(v.size())};
      std::memcpy(a.mutable_data(), v.data(), v.size() * 8);
      return a;
    };
    py::array_t<double> c{py::ssize_t(w.coeffs.size())};
    std::memcpy(c.mutable_data(), w.coeffs.data(), w.coeffs.size() * 8);
    return py::make_tuple(u64(w.alpha), u64(w.beta), c, py::make_tuple(w.nstates, w.norb, w.nalpha, w.nbeta));
  });
  io.def("write_wavefunction", [](const std::string& f, size_t norb, const std::vector<uint64_t>& a,
                                  const std::vector<uint64_t>& b, const std::vector<double>& c) {
    io::write_wavefunction(f, norb, a, b, c);
  }, py::arg("fname"), py::arg("norb"), py::arg("alpha"), py::arg("beta"), py::arg("coeffs"));
  // Hamiltonian of an FCIDUMP file (what pymacis users start from): (Hamiltonian, nalpha, nbeta)
  io.def("hamiltonian_from_fcidump", [](const std::string& f) {
    const io::FCIDumpHeader h = io::fcidump_read_header(f);
    if (h.norb == 0) throw std::runtime_error("NORB not found or is zero in FCIDUMP header");
    const size_t n = h.norb;
    std::vector<double> T(n * n), V(n * n * n * n);
    double core = 0.;
    io::read_fcidump_all(f, T.data(), n, V.data(), n, core);
    auto ham = std::make_shared<data::Hamiltonian>(n, std::move(T), std::move(V), core);
    const int na = (int(h.nelec) + h.ms2) / 2, nb = (int(h.nelec) - h.ms2) / 2;
    return py::make_tuple(ham, na, nb);
  });

  auto det_array = [](const std::vector<std::pair<uint64_t, uint64_t>>& d) {
    py::array_t<uint64_t> a({py::ssize_t(d.size()), py::ssize_t(2)});
    auto r = a.mutable_unchecked<2>();
    for (py::ssize_t i = 0; i < py::ssize_t(d.size()); ++i) { r(i, 0) = d[size_t(i)].first; r(i, 1) = d[size_t(i)].second; }
    return a;
  };
  amod.def("generate_cis_hilbert_space", [det_array](size_t norb, uint64_t a, uint64_t b) {
    return det_array(generate_cis_hilbert_space(norb, a, b));
  }, py::arg("norb"), py::arg("alpha"), py::arg("beta"), "(n, 2) array of (alpha, beta) occupation words");
  amod.def("generate_cisd_hilbert_space", [det_array](size_t norb, uint64_t a, uint64_t b) {
    return det_array(generate_cisd_hilbert_space(norb, a, b));
  }, py::arg("norb"), py::arg("alpha"), py::arg("beta"), "(n, 2) array of (alpha, beta) occupation words");
  amod.def("select_core_indices", &select_core_indices, py::arg("coefficients"), py::arg("fixed_core"),
           py::arg("ncdets_max"), py::arg("core_selection_threshold"));
  amod.def("set_device", &set_device, py::arg("device"));
  amod.def("set_communicator", [](const py::bytes& id, int rank, int nranks) { set_communicator(std::string(id), rank, nranks); },
           py::arg("unique_id"), py::arg("rank"), py::arg("nranks"));
  amod.def("clear_communicator", &clear_communicator);
  amod.def("last_run_stats", &last_run_stats);
  amod.def("row_block", &row_block, py::arg("n"), py::arg("rank"), py::arg("nranks"),
           "contiguous row block [r0, r1) of rank `rank` (the partition every sharded run uses)");
}
