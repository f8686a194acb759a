#!/usr/bin/env python
"""Golden H build / RDM / entropy data for wfn_t<128> determinants (32 < norb < 64), made with the
compiled reference (oracle/_ref, nbits = 128) on the 600-determinant wavefunction of
make_golden_wide.py (spin-sorted): CSR fingerprints of the three generators, the one-body RDMs,
checksums and a strided sample of the two-body RDMs, and the orbital entropies.
    python tests/golden/make_golden_wide_props.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port, ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

EPS = float(np.finfo(np.float64).eps)
sha = lambda x: hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest()
sp = W.config("wide36")
z = np.load(os.path.join(HERE, "wide36_golden.npz"))
a, b = z["run_dets"][:, 0].copy(), z["run_dets"][:, 1].copy()
o = port.spin_sort_order(a, b)
a, b, C = a[o], b[o], z["run_C"][o]
hg = ref.HamGen(sp.norb, sp.T, sp.V, nbits=128)
words = port.pack(a, b, 128)
meta = {"n": int(len(a)), "csr": {}}
for g in ("sorted_double_loop", "residue_arrays", "dynamic_bit_masking"):
    rp, ci, nz = hg.hbuild(words, EPS, generator=g)[0].arrays()
    meta["csr"][g] = dict(nnz=int(len(ci)), rowptr_sha=sha(rp.astype(np.int64)), colind_sha=sha(ci.astype(np.int64)),
                          nzval_sha=sha(nz))
aa, bb, aaaa, bbbb, aabb = hg.form_rdms(words, C, spin_dep=True)
ordm, trdm = hg.form_rdms(words, C, spin_dep=False)
s1, s2, mi = hg.form_entropies(words, C)
arrays = dict(ordm_aa=aa, ordm_bb=bb, ordm=ordm, s1=s1, s2=s2, mi=mi)
STEP = 997
for name, t in (("aaaa", aaaa), ("bbbb", bbbb), ("aabb", aabb), ("trdm", trdm)):
    flat = np.asarray(t).reshape(-1, order="F")
    arrays[f"{name}_sample"] = flat[::STEP].copy()
    meta[f"{name}_sum"] = float(flat.sum())
    meta[f"{name}_sumsq"] = float((flat * flat).sum())
meta["sample_step"] = STEP
np.savez_compressed(os.path.join(HERE, "wide36_props.npz"), **arrays)
with open(os.path.join(HERE, "wide36_props.json"), "w") as fh:
    json.dump(meta, fh, indent=1)
print(json.dumps(meta)[:600])
