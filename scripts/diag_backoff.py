"""Diagnostic for tests/test_gpu_host.py::test_asci_growth_backoff_scenarios_match_reference[fractional_grow_factor]:
runs that scenario with and without patched builds and with the row scan instead of the product kernel.
All four give the same energy (-85.313123201404, 1.7e-4 Eh above the reference's -85.313291280309): the
difference is a spin-flip-partner swap at the tiny cuts, not the H build.
    python scripts/diag_backoff.py        (on a B200 box, from the repository root)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qdk_chemistry_b200 import algorithms as alg, data, workloads as W
water = W.load_sparse_npz("tests/golden/h2o_ccpvdz.ints.npz")
ham = data.Hamiltonian(water.T, water.V, water.core_energy)
m = json.load(open("tests/golden/backoff_meta.json"))["fractional_grow_factor"]
kw = dict(m["settings"]); kw["core_selection_strategy"] = "fixed"
for env in ({}, {"B2CI_NO_INCREMENTAL": "1"}, {"B2CI_HBUILD_FORCE_SCAN": "1"}, {"B2CI_NO_INCREMENTAL": "1", "B2CI_HBUILD_FORCE_SCAN": "1"}):
    for k in ("B2CI_NO_INCREMENTAL", "B2CI_HBUILD_FORCE_SCAN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    E, w = alg.create("multi_configuration_calculator", "macis_asci", max_refine_iter=0, ci_residual_tolerance=1e-8, **kw).run(ham, 5, 5)
    st = alg.last_run_stats()
    print(env, w.size(), repr(E - water.core_energy), "ref", m["E"], "patched", st.get("h_build_patched"), "iters", st.get("asci_iterations"), flush=True)
