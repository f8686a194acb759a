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
