"""End-to-end parity at BASELINE configs[3] scale (collected last: one plugin run of ~2 s).

The N2-like ASCI(14e,26o) wavefunction is grown from the HF determinant to 2e5 determinants (five search +
diagonalisation iterations: 100, 800, 6400, 51200, 200000) and the energy compared with the unmodified reference's,
which took 49.6 s on 8 CPU threads (profiles/r01_reference_cpu_asci_n2_14e26o.json; measured difference 9.6e-14 Eh)."""
import pytest

from qdk_chemistry_b200 import algorithms as alg
from qdk_chemistry_b200 import data
from qdk_chemistry_b200 import workloads as W

pytestmark = pytest.mark.gpu

REFERENCE_E_2E5 = -21.516493341749076   # oracle/_ref: asci_grow, ntdets_max = 200000, QDK defaults otherwise


def test_n2_asci26_growth_to_2e5_determinants_matches_reference_energy():
    sp = W.config("n2_asci26")
    E, w = alg.create("multi_configuration_calculator", "macis_asci", ntdets_max=200000, max_refine_iter=0,
                      ci_residual_tolerance=1e-8).run(data.Hamiltonian(sp.T, sp.V, sp.core_energy), sp.nalpha, sp.nbeta)
    assert w.size() == 200000 and abs(w.norm() - 1) < 1e-12
    assert abs(E - sp.core_energy - REFERENCE_E_2E5) < 1e-8      # north_star: energies to 1e-8 Eh
    st = alg.last_run_stats()
    assert st["asci_iterations"] == 5 and st["ndets_after_grow"] == 200000
