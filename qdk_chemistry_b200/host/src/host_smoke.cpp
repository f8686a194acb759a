// C++ consumer of the plugin API: creates the calculators through the factory exactly as QDK
// code does (MultiConfigurationCalculatorFactory::create("...")->run(ham, na, nb)) on a tiny
// Hubbard dimer whose FCI energy is known in closed form. Needs a GPU; used by tests/.
#include <cmath>
#include <cstdio>

#include "qdk_b200/mc.hpp"

using namespace qdk_b200;
using namespace qdk_b200::algorithms;

int main() {
  // two-site Hubbard, t = 1, U = 4: E0 = U/2 - sqrt(U^2/4 + 4 t^2)
  const size_t n = 2;
  std::vector<double> T = {0., -1., -1., 0.}, V(16, 0.);
  V[0] = 4.;
  V[15] = 4.;
  auto ham = std::make_shared<data::Hamiltonian>(n, T, V, 0.0);
  const double exact = 2.0 - std::sqrt(4.0 + 4.0);
  int bad = 0;
  for (const char* name : {"b200_cas", "b200_asci"}) {
    auto calc = MultiConfigurationCalculatorFactory::create(name);
    if (std::string(name) == "b200_cas") calc->settings().set("iterative_solver_dimension_cutoff", 1);
    auto [E, wfn] = calc->run(ham, 1, 1);
    std::printf("%s: E = %.12f (exact %.12f), %zu determinants, first = %s\n", name, E, exact, wfn->size(),
                wfn->get_active_determinants()[0].to_string().c_str());
    if (std::abs(E - exact) > 1e-8) bad = 1;
    try {
      calc->settings().set("ci_residual_tolerance", 1e-9);
      std::printf("settings were not locked after run()\n");
      bad = 1;
    } catch (const data::SettingsAreLocked&) {
    }
  }
  return bad;
}
