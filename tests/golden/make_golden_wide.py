#!/usr/bin/env python
"""Golden data for 32 < norb < 64 (wfn_t<128> determinants, 128-bit ASCI keys), made with the
compiled reference (oracle/_ref, nbits = 128) in the build container:
  * one asci_search from a 40-determinant core set of the synthetic `wide36` space
    (determinant_search.hpp:808-1123) -> the selected determinants
  * one HF -> asci_grow -> asci_refine run to 600 determinants (macis_asci.cpp:160-196) -> energy
    python tests/golden/make_golden_wide.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port, ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

sp = W.config("wide36")
hg = ref.HamGen(sp.norb, sp.T, sp.V, nbits=128)
ham = port.Ham(sp.norb, sp.T, sp.V)

# ---- a core set with occupied orbitals on both sides of bit 32 and in both spins
rng = np.random.default_rng(77)
dets = set()
while len(dets) < 40:
    a = sum(1 << int(i) for i in rng.choice(sp.norb, sp.nalpha, replace=False))
    b = sum(1 << int(i) for i in rng.choice(sp.norb, sp.nbeta, replace=False))
    dets.add((a, b))
dets = sorted(dets)
ca = np.array([d[0] for d in dets], dtype=np.uint64)
cb = np.array([d[1] for d in dets], dtype=np.uint64)
cc = rng.normal(size=len(ca)) * np.exp(-np.arange(len(ca)) / 9.0)
cc /= np.linalg.norm(cc)
E0 = -4.0
ndets_max = 3000
o = ref.AsciOpts(ntdets_max=ndets_max)
sel = hg.asci_search(o, ndets_max, port.pack(ca, cb, 128), cc, E0).reshape(-1, 2)
sa, sb, stats = ham.asci_search(ca, cb, cc, E0, ndets_max, h_el_tol=o.h_el_tol, rv_prune_tol=o.rv_prune_tol)
ref_set = sorted(map(tuple, sel.tolist()))
port_set = sorted(zip(sa.tolist(), sb.tolist()))
print("search: reference", len(ref_set), "port", len(port_set), "same set:", ref_set == port_set)
assert ref_set == port_set, "oracle port disagrees with the reference on 128-bit keys"
assert (sel[:, 0] >> 32).any() and (sel[:, 1] >> 32).any()

# ---- a whole run
ro = ref.AsciOpts(ntdets_max=600, ntdets_min=50, max_refine_iter=30)
E, d, C = hg.asci_run(ro, sp.nalpha, sp.nbeta, refine=True)
print("asci run: E", E, "n", len(C))
np.savez_compressed(os.path.join(HERE, "wide36_golden.npz"), core_alpha=ca, core_beta=cb, core_C=cc,
                    selected=np.array(ref_set, dtype=np.uint64), run_dets=d.reshape(-1, 2), run_C=C)
meta = dict(E0=E0, ndets_max=ndets_max, n_selected=len(ref_set), run_E=E, run_n=len(C),
            run_opts=dict(ntdets_max=600, ntdets_min=50, max_refine_iter=30))
with open(os.path.join(HERE, "wide36_meta.json"), "w") as fh:
    json.dump(meta, fh, indent=1)
print(meta)
