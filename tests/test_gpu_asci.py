"""GPU parity of the ASCI search (generation, sort/accumulate, top-k) through the C ABI."""
import numpy as np
import pytest

from oracle import port
from qdk_chemistry_b200 import device
from qdk_chemistry_b200 import workloads as W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = device.Context(0)
    yield c
    c.close()


def _core_from_fci(sp, ncore, seed=0):
    """A spin-sorted core set with a normalised random-sign coefficient vector."""
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(len(a), size=min(ncore, len(a)), replace=False))
    ca, cb = a[pick], b[pick]
    o = port.spin_sort_order(ca, cb)
    ca, cb = ca[o], cb[o]
    c = rng.normal(size=len(ca)) * np.exp(-np.arange(len(ca)) / 7.0)
    return ca, cb, c / np.linalg.norm(c)


@pytest.mark.parametrize("name,ncore", [("small_cas8", 1), ("small_cas8", 37), ("hubbard_4x2", 20),
                                        ("n2_cas10", 64)])
def test_candidate_table_bit_exact(ctx, name, ncore):
    sp = W.config(name)
    ca, cb, c = _core_from_fci(sp, ncore)
    E0 = -3.25
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    words, cm, hd = ctx.asci_candidates(port.pack(ca, cb), c, E0, h_el_tol=1e-8)
    oa, ob, ocm, ohd = port.Ham(sp.norb, sp.T, sp.V).asci_candidates(ca, cb, c, E0, 1e-8)
    assert np.array_equal(words, port.pack(oa, ob))     # same keys, same (beta, alpha) order
    assert np.array_equal(hd, ohd)                       # first-parent h_diag, bit exact
    fin = np.isfinite(ocm)
    assert np.array_equal(np.isfinite(cm), fin)
    assert np.array_equal(cm[fin], ocm[fin])            # parent-ordered sums, bit exact


def test_water_search_selection_matches_reference(ctx, water, golden_meta, golden_arrays):
    # inputs and expected selection were produced by the compiled reference (make_golden.py)
    m = golden_meta["water_search"]
    ca, cb = golden_arrays["water_search.core_alpha"], golden_arrays["water_search.core_beta"]
    ctx.upload_integrals(water.norb, water.T, water.V)
    out, stats = ctx.asci_search(port.pack(ca, cb), golden_arrays["water_search.core_C"], m["E0"],
                                 m["ndets_max"])
    assert len(out) == m["n_selected"]
    assert np.array_equal(np.sort(out), golden_arrays["water_search.selected"])
    # core determinants are appended last, in the order given (determinant_search.hpp:1107-1114)
    assert np.array_equal(out[-len(ca):], port.pack(ca, cb))
    sa, sb, ostats = port.Ham(water.norb, water.T, water.V).asci_search(
        ca, cb, golden_arrays["water_search.core_C"], m["E0"], m["ndets_max"])
    assert np.array_equal(stats[:5], ostats[:5])         # counts, pivot and gap identical


@pytest.mark.parametrize("ndets_max", [30, 31, 200, 100000])
def test_topk_edge_cases(ctx, ndets_max):
    sp = W.config("small_cas8")
    ca, cb, c = _core_from_fci(sp, 30, seed=3)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    out, stats = ctx.asci_search(port.pack(ca, cb), c, -2.0, ndets_max)
    sa, sb, ostats = port.Ham(sp.norb, sp.T, sp.V).asci_search(ca, cb, c, -2.0, ndets_max)
    assert np.array_equal(out, port.pack(sa, sb))
    assert np.array_equal(stats[:5], ostats[:5])


def test_just_singles_and_tolerances(ctx):
    sp = W.config("small_cas8")
    ca, cb, c = _core_from_fci(sp, 25, seed=4)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    for kw in (dict(just_singles=True), dict(h_el_tol=1e-3), dict(rv_prune_tol=1e-2)):
        out, _ = ctx.asci_search(port.pack(ca, cb), c, -2.0, 500, **kw)
        sa, sb, _ = port.Ham(sp.norb, sp.T, sp.V).asci_search(ca, cb, c, -2.0, 500, **kw)
        assert np.array_equal(out, port.pack(sa, sb))


def test_unsorted_core_is_rejected(ctx):
    sp = W.config("small_cas8")
    ca, cb, c = _core_from_fci(sp, 10, seed=5)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    with pytest.raises(device.B2ciError) as e:
        ctx.asci_search(port.pack(ca[::-1], cb[::-1]), c, -2.0, 100)
    assert "Sorted" in str(e.value)


@pytest.mark.parametrize("parts", [2, 3, 7])
def test_key_partitioned_search_is_identical(ctx, water, golden_meta, golden_arrays, parts, monkeypatch):
    """Large searches are split into key partitions (hash of the determinant word); every
    determinant's contributions stay together and in parent order, so selection, pivot and
    counts are those of the single-pass search."""
    m = golden_meta["water_search"]
    ca, cb = golden_arrays["water_search.core_alpha"], golden_arrays["water_search.core_beta"]
    cc = golden_arrays["water_search.core_C"]
    ctx.upload_integrals(water.norb, water.T, water.V)
    out1, st1 = ctx.asci_search(port.pack(ca, cb), cc, m["E0"], m["ndets_max"])
    assert st1[5] == 1
    monkeypatch.setenv("B2CI_ASCI_PARTS", str(parts))
    outp, stp = ctx.asci_search(port.pack(ca, cb), cc, m["E0"], m["ndets_max"])
    assert stp[5] == parts
    assert np.array_equal(np.sort(outp), np.sort(out1))
    assert np.array_equal(np.sort(outp), golden_arrays["water_search.selected"])
    assert np.array_equal(stp[:5], st1[:5])        # contributions, unique candidates, pivot, gap, kept
    assert np.array_equal(outp[-len(ca):], port.pack(ca, cb))
    # a tiny memory budget picks the partition count by itself
    monkeypatch.delenv("B2CI_ASCI_PARTS")
    monkeypatch.setenv("B2CI_ASCI_BUDGET", str(int(st1[0] // 5)))
    outb, stb = ctx.asci_search(port.pack(ca, cb), cc, m["E0"], m["ndets_max"])
    assert stb[5] >= 6 and np.array_equal(np.sort(outb), np.sort(out1))


# ---- ASCI-PT2 (asci/pt2.hpp; known answer asci.cxx:562-570) -----------------------------------
@pytest.fixture(scope="module")
def water_refined():
    import json, os
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(g, "water_refined_wfn.npz"))
    with open(os.path.join(g, "pt2_meta.json")) as fh:
        return z["alpha"], z["beta"], z["C"], json.load(fh)


def test_pt2_water_known_answer(ctx, water, water_refined):
    a, b, C, m = water_refined
    ctx.upload_integrals(water.norb, water.T, water.V)
    e, npt2 = ctx.asci_pt2(port.pack(a, b), C, m["E_asci"], m["pt2_tol"])
    assert abs(e - m["known_answer"]) < 1e-8            # the reference's own tolerance (asci.cxx:569)
    assert abs(e - m["port_full"]) < 1e-12 * abs(m["port_full"]) * 1e2   # summation order only
    assert npt2 == m["port_full_npt2"]                   # same external determinants, exactly


@pytest.mark.parametrize("parts", [1, 3])
def test_pt2_matches_oracle_small(ctx, water, water_refined, parts, monkeypatch):
    a, b, C, m = water_refined
    top = np.sort(np.argsort(-np.abs(C), kind="stable")[: m["small_n"]])
    cs = C[top] / np.linalg.norm(C[top])
    eo, no = port.Ham(water.norb, water.T, water.V).asci_pt2(a[top], b[top], cs, m["E_asci"], 1e-16)
    ctx.upload_integrals(water.norb, water.T, water.V)
    monkeypatch.setenv("B2CI_ASCI_PARTS", str(parts))
    e, npt2 = ctx.asci_pt2(port.pack(a[top], b[top]), cs, m["E_asci"], 1e-16)
    assert npt2 == no
    assert abs(e - eo) <= 1e-13 * abs(eo)
    # a looser generation tolerance drops contributions on both sides alike
    e3, n3 = ctx.asci_pt2(port.pack(a[top], b[top]), cs, m["E_asci"], 1e-5)
    eo3, no3 = port.Ham(water.norb, water.T, water.V).asci_pt2(a[top], b[top], cs, m["E_asci"], 1e-5)
    assert n3 == no3 and n3 < no and abs(e3 - eo3) <= 1e-13 * abs(eo3)


# ---- 32 < norb < 64: wfn_t<128> determinants, 128-bit keys (two-phase radix sort) -------------
@pytest.fixture(scope="module")
def wide():
    import json, os
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(g, "wide36_meta.json")) as fh:
        return W.config("wide36"), np.load(os.path.join(g, "wide36_golden.npz")), json.load(fh)


def test_wide_keys_candidate_table_bit_exact(ctx, wide):
    sp, z, m = wide
    ca, cb, cc = z["core_alpha"], z["core_beta"], z["core_C"]
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    words, cm, hd = ctx.asci_candidates(port.pack(ca, cb, 128), cc, m["E0"], h_el_tol=1e-8, words_per_det=2)
    oa, ob, ocm, ohd = port.Ham(sp.norb, sp.T, sp.V).asci_candidates(ca, cb, cc, m["E0"], 1e-8)
    assert np.array_equal(words, port.pack(oa, ob, 128))   # 128-bit numeric order: beta word, then alpha
    assert np.array_equal(hd, ohd)
    fin = np.isfinite(ocm)
    assert np.array_equal(np.isfinite(cm), fin) and np.array_equal(cm[fin], ocm[fin])


@pytest.mark.parametrize("parts", [1, 3])
def test_wide_keys_search_matches_reference(ctx, wide, parts, monkeypatch):
    sp, z, m = wide
    ca, cb, cc = z["core_alpha"], z["core_beta"], z["core_C"]
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    monkeypatch.setenv("B2CI_ASCI_PARTS", str(parts))
    out, stats = ctx.asci_search(port.pack(ca, cb, 128), cc, m["E0"], m["ndets_max"], words_per_det=2)
    out = out.reshape(-1, 2)
    assert stats[5] == parts and len(out) == m["n_selected"]
    assert sorted(map(tuple, out.tolist())) == sorted(map(tuple, z["selected"].tolist()))
    assert np.array_equal(out[-len(ca):, 0], ca) and np.array_equal(out[-len(ca):, 1], cb)
    _, _, ostats = port.Ham(sp.norb, sp.T, sp.V).asci_search(ca, cb, cc, m["E0"], m["ndets_max"])
    assert np.array_equal(stats[:5], ostats[:5])


def test_wide_keys_pt2_matches_oracle(ctx, wide):
    sp, z, m = wide
    ca, cb, cc = z["core_alpha"], z["core_beta"], z["core_C"]
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    e, npt2 = ctx.asci_pt2(port.pack(ca, cb, 128), cc, m["E0"], 1e-16, words_per_det=2)
    eo, no = port.Ham(sp.norb, sp.T, sp.V).asci_pt2(ca, cb, cc, m["E0"], 1e-16)
    assert npt2 == no and abs(e - eo) <= 1e-13 * abs(eo)


@pytest.mark.parametrize("name", ["small_cas8", "wide36"])
def test_sorted_output_is_the_spin_sorted_reference_output(ctx, name, wide):
    """sort_output = 1 returns the list asci_iter would sort it into (iteration.hpp:117-119):
    spin_comparator order, alpha-major then beta, for one- and two-word determinants."""
    if name == "wide36":
        sp, z, m = wide
        ca, cb, cc, E0, nmax, wpd, nbits = z["core_alpha"], z["core_beta"], z["core_C"], m["E0"], m["ndets_max"], 2, 128
    else:
        sp = W.config(name)
        ca, cb, cc = _core_from_fci(sp, 30, seed=3)
        E0, nmax, wpd, nbits = -2.0, 400, 1, 64
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    ref_out, _ = ctx.asci_search(port.pack(ca, cb, nbits), cc, E0, nmax, words_per_det=wpd)
    out, _ = ctx.asci_search(port.pack(ca, cb, nbits), cc, E0, nmax, words_per_det=wpd, sort_output=True)
    ra, rb = port.unpack(ref_out, nbits)
    o = port.spin_sort_order(ra, rb)
    assert np.array_equal(out, port.pack(ra[o], rb[o], nbits))
