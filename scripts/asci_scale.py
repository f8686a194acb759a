#!/usr/bin/env python
"""Run the b200_asci plugin on a named synthetic workload and print per-phase statistics.
    python scripts/asci_scale.py n2_asci26 100000 [key=value ...]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qdk_chemistry_b200 import algorithms as alg, data, workloads as W  # noqa: E402

name, nt = sys.argv[1], int(float(sys.argv[2]))
kw = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    try:
        kw[k] = int(v)
    except ValueError:
        try:
            kw[k] = float(v)
        except ValueError:
            kw[k] = v
sp = W.config(name)
ham = data.Hamiltonian(sp.T, sp.V, sp.core_energy)
c = alg.create("multi_configuration_calculator", "b200_asci", ntdets_max=nt, ci_residual_tolerance=1e-8, **kw)
t0 = time.perf_counter()
try:
    E, w = c.run(ham, sp.nalpha, sp.nbeta)
    out = {"E": E, "ndets": w.size()}
except Exception as e:  # report what failed and the statistics so far
    out = {"error": str(e)[:300]}
out["wall_s"] = time.perf_counter() - t0
out.update(alg.last_run_stats())
print(json.dumps(out))
