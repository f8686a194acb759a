"""GPU parity of the reduced density matrices (b2ci_form_rdms / b2ci_form_rdms_spin_dep) against
the oracle port, which tests/test_oracle.py pins to the compiled reference. Summation order on
the device differs (fp64 reductions in L2), so the bar is rounding: 1e-12 absolute on entries of
magnitude <= 2."""
import numpy as np
import pytest

from oracle import port
from qdk_chemistry_b200 import device
from qdk_chemistry_b200 import workloads as W

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = device.Context(0)
    yield c
    c.close()


def _subset(sp, m, seed):
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    rng = np.random.default_rng(seed)
    if m < len(a):
        idx = np.sort(rng.choice(len(a), m, replace=False))
        a, b = a[idx], b[idx]
    C = rng.normal(size=len(a)) * np.exp(-rng.uniform(0, 6, size=len(a)))
    return a, b, C / np.linalg.norm(C)


@pytest.mark.parametrize("name,m,seed", [("tiny_cas6", 400, 0), ("small_cas8", 700, 1), ("hubbard_4x2", 4900, 2),
                                         ("small_cas8", 3920, 3)])
def test_rdms_match_oracle(ctx, name, m, seed):
    sp = W.config(name)
    a, b, C = _subset(sp, m, seed)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.upload_dets(port.pack(a, b), 1)
    o, t = ctx.form_rdms(dets, C)
    po, pt = port.form_rdms(sp.norb, a, b, C)
    assert np.abs(o - po).max() < TOL and np.abs(t - pt).max() < TOL
    g5 = ctx.form_rdms(dets, C, spin_dep=True)
    p5 = port.form_rdms(sp.norb, a, b, C, spin_dep=True)
    for x, y in zip(g5, p5):
        assert np.abs(x - y).max() < TOL
    # relations the reference's own test checks (double_loop.cxx:330-375)
    aa, bb, aaaa, bbbb, aabb = g5
    assert np.abs(o - (aa + bb)).max() < TOL
    assert np.abs(t - (aaaa + bbbb + aabb + aabb.transpose(2, 3, 0, 1))).max() < TOL
    assert abs(np.trace(o) - (sp.nalpha + sp.nbeta)) < 1e-11
    assert np.abs(o - o.T).max() < TOL
    dets.free()


def test_rdm_energy_equals_rayleigh_quotient_full_ci(ctx):
    """E = sum ordm*T + sum trdm*V (double_loop.cxx:198-237) at the converged CASCI vector."""
    sp = W.config("small_cas8")
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.generate_fci(sp.norb, sp.nalpha, sp.nbeta)
    H = ctx.hbuild(dets, 0.0)
    E, X, _, _ = H.davidson(200, 1e-10)
    o, t = ctx.form_rdms(dets, X)
    n = sp.norb
    Er = np.sum(o * sp.T.reshape(n, n, order="F")) + np.sum(t * sp.V.reshape((n,) * 4, order="F"))
    assert abs(Er - E) < 1e-9
    # requesting only one of the two leaves the other untouched
    o1, t1 = ctx.form_rdms(dets, X, two=False)
    assert t1 is None and np.abs(o1 - o).max() < TOL
    o2, t2 = ctx.form_rdms(dets, X, one=False)
    assert o2 is None and np.abs(t2 - t).max() < TOL
    H.free()
    dets.free()


def test_rdm_hf_determinant(ctx):
    """single determinant: ordm(i,i) = 2, trdm(i,i,j,j) = 2, trdm(i,j,j,i) = -1, trdm(i,i,i,i) = 1
    (double_loop.cxx:266-300)"""
    sp = W.config("tiny_cas6")
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    hf = np.array([(1 << sp.nalpha) - 1], dtype=np.uint64)
    dets = ctx.upload_dets(port.pack(hf, hf), 1)
    o, t = ctx.form_rdms(dets, np.array([1.0]))
    for i in range(sp.nalpha):
        assert o[i, i] == 2.0 and t[i, i, i, i] == 1.0
        for j in range(sp.nalpha):
            if i != j:
                assert t[i, i, j, j] == 2.0 and t[i, j, j, i] == -1.0
    assert np.count_nonzero(o) == sp.nalpha
    dets.free()


# ---- orbital entropies (form_entropies) ---------------------------------------------------------
def _flat(I, need_s2=True):
    parts = [I[k] for k in port.ENT_VECS]
    if need_s2:
        parts += [I[k].reshape(-1, order="F") for k in port.ENT_MATS]
    return np.concatenate(parts)


@pytest.mark.parametrize("name", ["tiny_cas6", "small_cas8", "hubbard_4x2"])
def test_entropies_match_reference_golden(ctx, name):
    """s1 / s2 / mutual information against the compiled reference's form_entropies
    (tests/golden/entropy_golden.npz) and the device intermediates against the python oracle."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "entropy_golden.npz"))
    sp = W.config(name)
    a, b, C = g[f"{name}.alpha"], g[f"{name}.beta"], g[f"{name}.C"]
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.upload_dets(port.pack(a, b), 1)
    s1, s2, mi = ctx.form_entropies(dets, C)
    assert np.abs(s1 - g[f"{name}.s1"]).max() < 1e-12
    assert np.abs(s2 - g[f"{name}.s2"]).max() < 1e-11 and np.abs(mi - g[f"{name}.mi"]).max() < 1e-11
    I = ctx.entropy_intermediates(dets, C)
    assert np.abs(I - _flat(port.entropy_intermediates(sp.norb, a, b, C))).max() < 1e-13
    # single-orbital entropies alone: diagonal pairs only, no pattern build
    o1, n2, nmi = ctx.form_entropies(dets, C, two_orbital=False, mutual_information=False)
    assert n2 is None and nmi is None and np.abs(o1 - g[f"{name}.s1"]).max() < 1e-12
    I1 = ctx.entropy_intermediates(dets, C, need_s2=False)
    assert np.abs(I1 - _flat(port.entropy_intermediates(sp.norb, a, b, C, need_s2=False), False)).max() < 1e-13
    # mutual information without the two-orbital matrix (the reference builds s2 internally)
    m1, m2, mmi = ctx.form_entropies(dets, C, two_orbital=False, mutual_information=True)
    assert m2 is None and np.abs(mmi - g[f"{name}.mi"]).max() < 1e-11
    dets.free()


def test_entropies_full_ci_vector_properties(ctx):
    """On the converged CASCI vector of small_cas8 (3,920 determinants): s2 symmetric with zero
    diagonal, mutual information = s1_i + s1_j - s2_ij >= -1e-12, s1 <= ln 4."""
    sp = W.config("small_cas8")
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.generate_fci(sp.norb, sp.nalpha, sp.nbeta)
    H = ctx.hbuild(dets, 0.0)
    _, X, _, _ = H.davidson(200, 1e-10)
    s1, s2, mi = ctx.form_entropies(dets, X)
    a, b = port.unpack(dets.download(1))
    p1, p2, pmi = port.form_entropies(sp.norb, a, b, X)
    assert np.abs(s1 - p1).max() < 1e-12 and np.abs(s2 - p2).max() < 1e-11 and np.abs(mi - pmi).max() < 1e-11
    assert np.all(s1 <= np.log(4) + 1e-12) and np.all(s1 >= 0)
    assert np.abs(s2 - s2.T).max() == 0 and np.all(np.diag(s2) == 0) and mi.min() > -1e-12
    H.free()
    dets.free()


def test_wfn128_hbuild_rdms_entropies_match_reference_golden():
    """36 orbitals, wfn_t<128> words: CSR fingerprints (all three generators), RDMs and orbital
    entropies against data made with the compiled reference's 128-bit instantiation
    (tests/golden/make_golden_wide_props.py)."""
    import json
    import os
    from helpers import EPS, check_generator_golden
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z, props = np.load(os.path.join(g, "wide36_golden.npz")), np.load(os.path.join(g, "wide36_props.npz"))
    with open(os.path.join(g, "wide36_props.json")) as fh:
        meta = json.load(fh)
    sp = W.config("wide36")
    a, b = z["run_dets"][:, 0].copy(), z["run_dets"][:, 1].copy()
    o = port.spin_sort_order(a, b)
    a, b, C = a[o], b[o], z["run_C"][o]
    c = device.Context(0)
    try:
        c.upload_integrals(sp.norb, sp.T, sp.V)
        d = c.upload_dets(port.pack(a, b, 128), 2)
        for gen, rec in meta["csr"].items():
            c.set_hamiltonian_generator(gen)
            check_generator_golden(rec, *c.hbuild(d, EPS).download())
        c.set_hamiltonian_generator("")
        aa, bb, aaaa, bbbb, aabb = c.form_rdms(d, C, spin_dep=True)
        ordm, trdm = c.form_rdms(d, C, spin_dep=False)
        for got, key in ((aa, "ordm_aa"), (bb, "ordm_bb"), (ordm, "ordm")):
            assert np.abs(got - props[key]).max() < 1e-12
        step = meta["sample_step"]
        for got, key in ((aaaa, "aaaa"), (bbbb, "bbbb"), (aabb, "aabb"), (trdm, "trdm")):
            flat = np.asarray(got).reshape(-1, order="F")
            assert np.abs(flat[::step] - props[f"{key}_sample"]).max() < 1e-12
            assert abs(flat.sum() - meta[f"{key}_sum"]) < 1e-9
        s1, s2, mi = c.form_entropies(d, C)
        assert np.abs(s1 - props["s1"]).max() < 1e-11 and np.abs(s2 - props["s2"]).max() < 1e-11
        assert np.abs(mi - props["mi"]).max() < 1e-11
    finally:
        c.close()
