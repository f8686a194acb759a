"""Shared helpers for the parity tests (test infrastructure)."""
import hashlib
import itertools

import numpy as np

EPS = float(np.finfo(np.float64).eps)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def cisd_space(norb, na, nb):
    """CISD space of the canonical HF determinant, spin_comparator order
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


def check_csr_against_golden(meta, arrays, key, rp, ci, nz, bit_exact_values=True, rtol=1e-12):
    rec = meta[key]
    assert len(rp) - 1 == rec["n"]
    assert int(rp[-1]) == rec["nnz"]
    assert np.array_equal(np.asarray(rp, dtype=np.int64), arrays[f"{key}.rowptr"])
    assert sha(np.asarray(ci, dtype=np.int64)) == rec["colind_sha"]
    if bit_exact_values:
        assert sha(np.asarray(nz, dtype=np.float64)) == rec["nzval_sha"]
    else:
        step = max(1, len(nz) // 4096)
        ref = arrays[f"{key}.nzval_sample"]
        got = np.asarray(nz)[::step]
        assert np.allclose(got, ref, rtol=rtol, atol=0.0)
        assert abs(float(np.sum(nz)) - rec["nzval_sum"]) <= 1e-9 * max(1.0, rec["nzval_abs_sum"])


def generator_case(tag):
    """(ActiveSpace, alpha, beta) of a tests/golden/generators_meta.json case (make_golden_generators.py)."""
    from oracle import port
    from qdk_chemistry_b200 import workloads as W
    if tag == "alpha_empty_8o":
        sp = W.config("small_cas8")
        bs = np.array([sum(1 << i for i in c) for c in itertools.combinations(range(8), 3)], dtype=np.uint64)
        return sp, np.zeros(len(bs), dtype=np.uint64), bs
    name, n, seed = {"hubbard_4x2_s600": ("hubbard_4x2", 600, 5), "small_cas8_s900": ("small_cas8", 900, 6),
                     "n2_cas10_s2500": ("n2_cas10", 2500, 7)}[tag]
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    pick = np.sort(np.random.default_rng(seed).choice(len(a), size=n, replace=False))
    o = port.spin_sort_order(a[pick], b[pick])
    return sp, a[pick][o], b[pick][o]


def check_generator_golden(rec, rp, ci, nz):
    assert int(rp[-1]) == rec["nnz"]
    assert sha(np.asarray(rp, dtype=np.int64)) == rec["rowptr_sha"]
    assert sha(np.asarray(ci, dtype=np.int64)) == rec["colind_sha"]
    assert sha(np.asarray(nz, dtype=np.float64)) == rec["nzval_sha"]
