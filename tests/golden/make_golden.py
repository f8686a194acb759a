"""Generate the committed golden fixtures from the REFERENCE itself.

Run in the build container (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Everything written here comes from (a) files of the reference's own test suite
(FCIDUMPs, the CISD rowptr blob) or (b) outputs of oracle/_ref/libmacis_ref.so, i.e. the
unmodified reference code. The GPU box has no /root/reference, so tests read only these
fixtures. Large arrays are stored as SHA-256 digests plus strided samples.
"""
from __future__ import annotations

import hashlib
import itertools
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import port, ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

REF = "/root/reference/external/macis/"
EPS = float(np.finfo(np.float64).eps)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cisd_space(norb, na, nb):
    """CISD determinants of the canonical HF reference, spin_comparator-sorted
    (external/macis/tests/csr_hamiltonian.cxx:44-47)."""
    occ_a, vir_a = list(range(na)), list(range(na, norb))
    occ_b, vir_b = list(range(nb)), list(range(nb, norb))
    hfa, hfb = (1 << na) - 1, (1 << nb) - 1

    def singles(s, occ, vir):
        return [s ^ (1 << i) ^ (1 << a) for i in occ for a in vir]

    def doubles(s, occ, vir):
        return [s ^ (1 << i) ^ (1 << j) ^ (1 << a) ^ (1 << b)
                for i, j in itertools.combinations(occ, 2)
                for a, b in itertools.combinations(vir, 2)]

    sa, sb = singles(hfa, occ_a, vir_a), singles(hfb, occ_b, vir_b)
    da, db = doubles(hfa, occ_a, vir_a), doubles(hfb, occ_b, vir_b)
    dets = ([(hfa, hfb)] + [(a, hfb) for a in sa] + [(hfa, b) for b in sb] +
            [(a, hfb) for a in da] + [(hfa, b) for b in db] + [(a, b) for a in sa for b in sb])
    dets = sorted(set(dets))
    return (np.array([a for a, _ in dets], dtype=np.uint64),
            np.array([b for _, b in dets], dtype=np.uint64))


def csr_record(hg: ref.HamGen, a, b, thresh, davidson=None):
    H, _ = hg.hbuild(port.pack(a, b), thresh)
    rp, ci, nz = H.arrays()
    rec = dict(n=int(len(a)), nnz=int(rp[-1]), thresh=thresh, rowptr_sha=sha(rp),
               colind_sha=sha(ci), nzval_sha=sha(nz),
               nzval_sum=float(np.sum(nz)), nzval_abs_sum=float(np.sum(np.abs(nz))))
    arrays = dict(rowptr=rp, nzval_sample=nz[::max(1, len(nz) // 4096)],
                  colind_sample=ci[::max(1, len(ci) // 4096)])
    if davidson:
        max_m, tol = davidson
        E, X, nit = H.davidson(max_m, tol, guess_policy=False)
        rec.update(davidson=dict(max_m=max_m, tol=tol, E=E, niter=int(nit)))
        arrays["davidson_X"] = X
    return rec, arrays


def main():
    meta = {}
    # ---- reference test inputs -> compact fixtures
    water = W.read_fcidump(REF + "tests/ref_data/h2o.ccpvdz.fci.dat", name="h2o_ccpvdz")
    W.save_sparse_npz(os.path.join(HERE, "h2o_ccpvdz.ints.npz"), water)
    n2_18 = W.read_fcidump(REF + "python/tests/data/n2_full_14e18o.hamiltonian.fcidump",
                           name="n2_14e18o")
    W.save_sparse_npz(os.path.join(HERE, "n2_14e18o.ints.npz"), n2_18)
    n2_6 = W.read_fcidump(REF + "python/tests/data/n2_selected_6e6o.hamiltonian.fcidump",
                          name="n2_6e6o")
    W.save_sparse_npz(os.path.join(HERE, "n2_6e6o.ints.npz"), n2_6)
    shutil.copyfile(REF + "tests/ref_data/h2o.ccpvdz.cisd.rowptr.bin",
                    os.path.join(HERE, "h2o.ccpvdz.cisd.rowptr.bin"))

    # ---- the reference's published known answers (file:line in BASELINE.md)
    meta["known_answers"] = dict(
        water_hf_total=-76.0267803489191, water_cisd_n=12636, water_cisd_nnz=3517816,
        water_cisd_davidson_total=-76.23197835987, water_asci_grow=-85.42926580489,
        water_asci_refine=-85.42926585527, n2_6e6o_casci=-9.155573,
        n2_14e18o_asci2000=-120.5187165264)

    # ---- water CISD (csr_hamiltonian.cxx:44-99, davidson.cxx:20-75)
    hgw = ref.HamGen(water.norb, water.T, water.V)
    a, b = cisd_space(24, 5, 5)
    arrays = {}
    for tag, thr in (("1e-16", 1e-16), ("eps", EPS), ("zero", 0.0)):
        rec, arr = csr_record(hgw, a, b, thr, davidson=(15, 1e-8) if tag == "1e-16" else None)
        meta[f"water_cisd_{tag}"] = rec
        for k, v in arr.items():
            arrays[f"water_cisd_{tag}.{k}"] = v
    meta["water_core"] = water.core_energy

    # ---- water ASCI: single searches + full runs (asci.cxx:541-558)
    o = ref.AsciOpts(core_selection_strategy=0, ntdets_max=10000)
    E, d, C = hgw.asci_run(o, 5, 5, refine=False)
    meta["water_asci_grow_ref"] = dict(E=E, n=int(len(C)))
    da, db = port.unpack(d)
    arrays["water_asci_grow.alpha"], arrays["water_asci_grow.beta"] = da, db
    arrays["water_asci_grow.C"] = C
    E2, d2, C2 = hgw.asci_run(o, 5, 5, refine=True)
    meta["water_asci_refine_ref"] = dict(E=E2, n=int(len(C2)), dets_sha=sha(np.sort(d2)))
    # one search step from the grown wavefunction (100 fixed core dets) to 10000 dets
    order = np.argsort(-np.abs(C), kind="stable")
    ca, cb, cx = da[order][:100], db[order][:100], C[order][:100]
    o2 = port.spin_sort_order(ca, cb)
    ca, cb, cx = ca[o2], cb[o2], cx[o2]
    sel = hgw.asci_search(o, 10000, port.pack(ca, cb), cx, E)
    arrays["water_search.core_alpha"], arrays["water_search.core_beta"] = ca, cb
    arrays["water_search.core_C"] = cx
    arrays["water_search.selected"] = np.sort(sel)
    meta["water_search"] = dict(E0=E, ndets_max=10000, n_selected=int(len(sel)))
    o_small = ref.AsciOpts(core_selection_strategy=0, ntdets_max=1000, ncdets_max=50)
    E3, d3, C3 = hgw.asci_run(o_small, 5, 5, refine=True)
    meta["water_asci_1000"] = dict(E=E3, n=int(len(C3)))
    arrays["water_asci_1000.dets"] = np.sort(d3)

    # ---- N2 goldens of test_pymacis.py:114-188
    hg6 = ref.HamGen(n2_6.norb, n2_6.T, n2_6.V)
    d6 = ref.generate_hilbert_space(6, 3, 3)
    E6, C6 = hg6.selected_ci_diag(d6, EPS, 200, 1e-8)
    meta["n2_6e6o_casci_ref"] = dict(E=E6, n=int(len(d6)))
    arrays["n2_6e6o.C"] = C6
    hg18 = ref.HamGen(n2_18.norb, n2_18.T, n2_18.V)
    o18 = ref.AsciOpts(ntdets_max=2000, grow_factor=2.0, max_refine_iter=15,
                       ci_max_subspace=1000, ci_res_tol=1e-8)
    try:
        E18, d18, C18 = hg18.asci_run(o18, 7, 7, refine=True)
        meta["n2_14e18o_asci2000_ref"] = dict(E=E18, n=int(len(C18)), opts="percentage core")
        arrays["n2_14e18o_asci2000.dets"] = np.sort(d18)
    except RuntimeError as e:  # recorded, not hidden
        meta["n2_14e18o_asci2000_ref"] = dict(error=str(e))

    # ---- small synthetic workloads: full CSR + Davidson from the reference
    for name in ("tiny_cas6", "small_cas8", "hubbard_3x2", "hubbard_4x2", "n2_cas10"):
        sp = W.config(name)
        hg = ref.HamGen(sp.norb, sp.T, sp.V)
        d = ref.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
        aa, bb = port.unpack(d)
        for tag, thr in (("eps", EPS), ("zero", 0.0)):
            if name == "n2_cas10" and tag == "zero":
                continue
            rec, arr = csr_record(hg, aa, bb, thr, davidson=(200, 1e-8) if tag == "eps" else None)
            meta[f"{name}_{tag}"] = rec
            for k, v in arr.items():
                if name == "n2_cas10" and k == "davidson_X":
                    v = v[:: 16]
                arrays[f"{name}_{tag}.{k}"] = v
        meta[f"{name}_dets_sha"] = sha(d)

    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **arrays)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True)[:3000])


if __name__ == "__main__":
    main()
