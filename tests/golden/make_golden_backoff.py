#!/usr/bin/env python
"""Outcomes of the reference's growth back-off scenarios (external/macis/tests/asci.cxx:577-733) on water /
cc-pVDZ, produced by the compiled reference (oracle/_ref): final size and energy of asci_grow.
    python tests/golden/make_golden_backoff.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

water = W.load_sparse_npz(os.path.join(HERE, "h2o_ccpvdz.ints.npz"))
CASES = {
    # name: settings (strategy 0 = fixed, 1 = percentage -- the MACIS default)
    "fractional_grow_factor": dict(grow_factor=2.5, ntdets_max=100, ntdets_min=10, core_selection_strategy=0),
    "forced_backoff": dict(grow_factor=10000.0, ntdets_max=10000, ntdets_min=5, ncdets_max=1, core_selection_strategy=0),
    "minimum_grow_factor": dict(grow_factor=10.0, ntdets_max=1000, ntdets_min=5, ncdets_max=5, core_selection_strategy=0),
    "normal_growth": dict(grow_factor=8.0, ntdets_max=1000, ntdets_min=100, ncdets_max=1000, core_selection_strategy=0),
    "taper": dict(grow_factor=8.0, taper_grow_factor=2.0, ntdets_max=3000, ntdets_min=100, core_selection_strategy=0),
    # core-selection strategies (asci.cxx:736-840)
    "fixed_core_5000": dict(ntdets_max=5000, ncdets_max=100, core_selection_strategy=0),
    "percentage_core_5000": dict(ntdets_max=5000, ncdets_max=100, core_selection_strategy=1, core_selection_threshold=0.95),
    "percentage_70": dict(ntdets_max=2000, core_selection_strategy=1, core_selection_threshold=0.7),
    "percentage_99": dict(ntdets_max=2000, core_selection_strategy=1, core_selection_threshold=0.99),
}
meta = {}
for name, kw in CASES.items():
    hg = ref.HamGen(water.norb, water.T, water.V)
    E, d, C = hg.asci_run(ref.AsciOpts(max_refine_iter=0, **kw), 5, 5, refine=False)
    meta[name] = dict(settings=kw, n=len(C), E=E, norm=float(C @ C))
    if len(C) <= 100:  # small enough to keep the wavefunction itself: determinant words (wfn_t<64>) and coefficients
        meta[name]["dets"] = [int(x) for x in d]
        meta[name]["coeffs"] = [float(x) for x in C]
    print(name, len(C), E)
with open(os.path.join(HERE, "backoff_meta.json"), "w") as fh:
    json.dump(meta, fh, indent=1)
