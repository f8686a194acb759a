#!/usr/bin/env python
"""Golden data for ASCI-PT2 (run in the build container, where /root/reference exists).

The reference's asci_pt2_constraint is compiled only with MPI (asci/pt2.hpp:17), which this
image lacks, so the pin is the reference's own known answer: asci.cxx:562-570 evaluates PT2 on
the water/cc-pVDZ wavefunction after asci_grow + asci_refine to 10,000 determinants and expects
-5.701535028967e-03. This script makes that wavefunction with the compiled reference
(oracle/_ref), stores it spin-sorted, evaluates the oracle port on it and records both numbers.
    python tests/golden/make_golden_pt2.py
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port, ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

KNOWN = -5.701535028967e-03  # external/macis/tests/asci.cxx:569

water = W.load_sparse_npz(os.path.join(HERE, "h2o_ccpvdz.ints.npz"))
norb, T, V = water.norb, water.T, water.V
hg = ref.HamGen(norb, T, V)
o = ref.AsciOpts(core_selection_strategy=0, ntdets_max=10000)
E, d, C = hg.asci_run(o, 5, 5, refine=True)
a, b = port.unpack(d)
order = port.spin_sort_order(a, b)
a, b, C = a[order], b[order], C[order]
ham = port.Ham(norb, T, V)
t0 = time.time()
ept2, npt2 = ham.asci_pt2(a, b, C, E, 1e-16)
print("full PT2", ept2, npt2, "known", KNOWN, "diff", ept2 - KNOWN, f"{time.time() - t0:.1f}s")
assert abs(ept2 - KNOWN) < 1e-8, "oracle port disagrees with the reference's known answer"
# a small case for the CPU test suite: the 200 largest coefficients, renormalised
top = np.sort(np.argsort(-np.abs(C), kind="stable")[:200])
cs = C[top] / np.linalg.norm(C[top])
e_small, n_small = ham.asci_pt2(a[top], b[top], cs, E, 1e-16)
np.savez_compressed(os.path.join(HERE, "water_refined_wfn.npz"), alpha=a, beta=b, C=C)
meta = dict(E_asci=E, known_answer=KNOWN, port_full=ept2, port_full_npt2=npt2, small_n=200,
            port_small=e_small, port_small_npt2=n_small, pt2_tol=1e-16)
with open(os.path.join(HERE, "pt2_meta.json"), "w") as fh:
    json.dump(meta, fh, indent=1)
print(meta)
