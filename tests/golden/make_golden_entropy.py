#!/usr/bin/env python
"""Golden orbital entropies from the compiled reference (oracle/_ref, form_entropies of the
SortedDoubleLoop generator) on small seeded wavefunctions. Run in the build container:
    python tests/golden/make_golden_entropy.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port, ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

out = {}
for name, m, seed in (("tiny_cas6", 400, 0), ("small_cas8", 600, 1), ("hubbard_4x2", 900, 2)):
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    rng = np.random.default_rng(seed)
    if m < len(a):
        idx = np.sort(rng.choice(len(a), m, replace=False))
        a, b = a[idx], b[idx]
    C = rng.normal(size=len(a)) * np.exp(-rng.uniform(0, 5, size=len(a)))
    C /= np.linalg.norm(C)
    s1, s2, mi = ref.HamGen(sp.norb, sp.T, sp.V).form_entropies(port.pack(a, b), C)
    for k, v in (("alpha", a), ("beta", b), ("C", C), ("s1", s1), ("s2", s2), ("mi", mi)):
        out[f"{name}.{k}"] = v
    print(name, len(a), s1[:3], float(mi.max()))
np.savez_compressed(os.path.join(HERE, "entropy_golden.npz"), **out)
