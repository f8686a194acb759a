#!/usr/bin/env python
"""Golden CSR fingerprints of the reference's pair-based generators (hamiltonian_build_algorithm =
residue_arrays / dynamic_bit_masking), made with the compiled reference (oracle/_ref) in the build
container. Cases where their rules differ from sorted_double_loop are included on purpose: exact
zero diagonals (extended Hubbard), thresholds that cut real elements, alpha-empty determinants.
    python tests/golden/make_golden_generators.py
"""
import hashlib
import itertools
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port, ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

EPS = float(np.finfo(np.float64).eps)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def subset(name, n, seed):
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    pick = np.sort(np.random.default_rng(seed).choice(len(a), size=n, replace=False))
    o = port.spin_sort_order(a[pick], b[pick])
    return sp, a[pick][o], b[pick][o]


cases = {}
for tag, (name, n, seed) in {"hubbard_4x2_s600": ("hubbard_4x2", 600, 5), "small_cas8_s900": ("small_cas8", 900, 6),
                             "n2_cas10_s2500": ("n2_cas10", 2500, 7)}.items():
    sp, a, b = subset(name, n, seed)
    cases[tag] = (sp, a, b, dict(workload=name, n=n, seed=seed))
# alpha-empty determinants: 0 alpha + 3 beta electrons in the small_cas8 integrals
sp8 = W.config("small_cas8")
bs = np.array([sum(1 << i for i in c) for c in itertools.combinations(range(8), 3)], dtype=np.uint64)
cases["alpha_empty_8o"] = (sp8, np.zeros(len(bs), dtype=np.uint64), bs, dict(workload="small_cas8", nalpha=0, nbeta=3))

meta = {}
for tag, (sp, a, b, desc) in cases.items():
    hg = ref.HamGen(sp.norb, sp.T, sp.V)
    ham = port.Ham(sp.norb, sp.T, sp.V)
    for thr_tag, thr in (("eps", EPS), ("zero", 0.0), ("1e-2", 1e-2)):
        rec = {}
        for g in ("sorted_double_loop", "residue_arrays", "dynamic_bit_masking"):
            rp, ci, nz = hg.hbuild(port.pack(a, b), thr, generator=g)[0].arrays()
            prp, pci, pnz = ham.hbuild(a, b, thr, generator=g)
            assert np.array_equal(rp, prp) and np.array_equal(ci, pci) and np.array_equal(nz, pnz), (tag, thr_tag, g)
            rec[g] = dict(nnz=int(len(ci)), rowptr_sha=sha(rp.astype(np.int64)), colind_sha=sha(ci.astype(np.int64)),
                          nzval_sha=sha(nz))
        assert rec["residue_arrays"] == rec["dynamic_bit_masking"]
        meta[f"{tag}.{thr_tag}"] = dict(desc, thr=thr, **rec)
        print(tag, thr_tag, {g: r["nnz"] for g, r in rec.items()})
with open(os.path.join(HERE, "generators_meta.json"), "w") as fh:
    json.dump(meta, fh, indent=1)
