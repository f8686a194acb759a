#!/usr/bin/env python
"""Golden energies for grow_with_rot (natural-orbital rotation of the integrals during asci_grow,
asci/grow.hpp:163-258), made with the compiled reference (oracle/_ref) on water / cc-pVDZ.
    python tests/golden/make_golden_rot.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

water = W.load_sparse_npz(os.path.join(HERE, "h2o_ccpvdz.ints.npz"))
meta = {}
for tag, refine in (("grow", 0), ("refine", 6)):
    hg = ref.HamGen(water.norb, water.T, water.V)   # a fresh generator: the run rotates its integrals
    o = ref.AsciOpts(core_selection_strategy=0, grow_with_rot=1, ntdets_max=2000, rot_size_start=200,
                     max_refine_iter=refine)
    E, d, C = hg.asci_run(o, 5, 5, refine=refine > 0)
    meta[tag] = dict(E=E, n=len(C), ntdets_max=2000, rot_size_start=200, max_refine_iter=refine)
    print(tag, E, len(C))
with open(os.path.join(HERE, "rot_meta.json"), "w") as fh:
    json.dump(meta, fh, indent=1)
