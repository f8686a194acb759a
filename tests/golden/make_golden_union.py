#!/usr/bin/env python
"""Golden outcomes of asci_refine's oscillation handling (union stabilisation + extended iteration budget,
asci/refine.hpp:118-205, 222-231) on the 36-orbital synthetic space, from the compiled reference (oracle/_ref,
wfn_t<128>): a run that gives up after 2 granted extra iterations, and one that converges after 15 unions on an
enlarged (624-determinant) set.
    python tests/golden/make_golden_union.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

sp = W.config("wide36")
base = dict(ntdets_max=600, ntdets_min=50, ncdets_max=20, refine_energy_tol=1e-4, core_selection_strategy=0)
meta = {"settings": base, "runs": {}}
for mri in (20, 80):
    hg = ref.HamGen(sp.norb, sp.T, sp.V, nbits=128)
    try:
        E, d, C = hg.asci_run(ref.AsciOpts(max_refine_iter=mri, **base), sp.nalpha, sp.nbeta, refine=True)
        meta["runs"][str(mri)] = dict(converged=True, E=E, n=len(C))
    except RuntimeError as e:
        meta["runs"][str(mri)] = dict(converged=False, message=str(e))
    print(mri, meta["runs"][str(mri)])
with open(os.path.join(HERE, "union_meta.json"), "w") as fh:
    json.dump(meta, fh, indent=1)
