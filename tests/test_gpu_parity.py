"""GPU parity tests proper: every call goes through the C ABI (libb2ci.so) and is compared
with the plain-C oracle on the same inputs and with the fixtures generated from the compiled
reference. Bit-exact for patterns / indices / matrix elements; FP tolerances are written
next to each assertion."""
import os

import numpy as np
import pytest

from oracle import port
from qdk_chemistry_b200 import device
from qdk_chemistry_b200 import workloads as W
from helpers import EPS, check_csr_against_golden, cisd_space, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = device.Context(0)
    yield c
    c.close()


def _build(ctx, sp, a, b, thr, rows=None):
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.upload_dets(port.pack(a, b), 1)
    H = ctx.hbuild(dets, thr, rows)
    return dets, H


def test_intermediates_bit_exact(ctx):
    sp = W.config("small_cas8")
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    got = ctx.download_intermediates()
    ref = port.Ham(sp.norb, sp.T, sp.V).intermediates()
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)


@pytest.mark.parametrize("cfg", [(6, 3, 3), (8, 4, 3), (7, 0, 3), (5, 5, 2), (12, 6, 6), (26, 2, 1)])
def test_generate_hilbert_space_order(ctx, cfg):
    norb, na, nb = cfg
    a, b = port.generate_hilbert_space(norb, na, nb)
    d = ctx.generate_fci(norb, na, nb)
    assert np.array_equal(d.download(1), port.pack(a, b))
    assert np.array_equal(d.download(2), port.pack(a, b, 128))


@pytest.mark.parametrize("name", ["tiny_cas6", "small_cas8", "hubbard_3x2", "hubbard_4x2"])
@pytest.mark.parametrize("tag,thr", [("eps", EPS), ("zero", 0.0)])
def test_fci_csr_bit_exact(ctx, golden_meta, golden_arrays, name, tag, thr):
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    dets, H = _build(ctx, sp, a, b, thr)
    rp, ci, nz = H.download()
    # against the reference-made fixture (pattern AND values bit-exact)
    check_csr_against_golden(golden_meta, golden_arrays, f"{name}_{tag}", rp, ci, nz)
    # and against the oracle run now
    orp, oci, onz = port.Ham(sp.norb, sp.T, sp.V).hbuild(a, b, thr)
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(nz, onz)


def test_water_cisd_csr_matches_reference_golden(ctx, water, golden_meta, golden_arrays):
    # external/macis/tests/csr_hamiltonian.cxx:76-99
    a, b = cisd_space(24, 5, 5)
    dets, H = _build(ctx, water, a, b, 1e-16)
    rp, ci, nz = H.download()
    blob = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", "h2o.ccpvdz.cisd.rowptr.bin"),
                       dtype=np.int32)
    assert H.nrows == 12636 and H.nnz == 3517816
    assert np.array_equal(rp, blob.astype(np.int64))
    check_csr_against_golden(golden_meta, golden_arrays, "water_cisd_1e-16", rp, ci, nz)


@pytest.mark.parametrize("tag,thr", [("eps", EPS), ("zero", 0.0)])
def test_water_cisd_threshold_semantics(ctx, water, golden_meta, golden_arrays, tag, thr):
    a, b = cisd_space(24, 5, 5)
    dets, H = _build(ctx, water, a, b, thr)
    rp, ci, nz = H.download()
    check_csr_against_golden(golden_meta, golden_arrays, f"water_cisd_{tag}", rp, ci, nz)


def test_n2_cas10_csr_bit_exact_config0(ctx, golden_meta, golden_arrays):
    # BASELINE.json configs[0] shape: CAS(10e,10o), 63,504 determinants
    sp = W.config("n2_cas10")
    d = ctx.generate_fci(10, 5, 5)
    assert sha(d.download(1)) == golden_meta["n2_cas10_dets_sha"]
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    H = ctx.hbuild(d, EPS)
    rp, ci, nz = H.download()
    check_csr_against_golden(golden_meta, golden_arrays, "n2_cas10_eps", rp, ci, nz)


def test_row_block_build_equals_full(ctx):
    sp = W.config("small_cas8")
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    dets, H = _build(ctx, sp, a, b, EPS)
    rp, ci, nz = H.download()
    for r0, r1 in [(0, 1), (137, 2001), (3919, 3920), (500, 500)]:
        Hb = ctx.hbuild(dets, EPS, (r0, r1))
        rpb, cib, nzb = Hb.download()
        assert Hb.row_begin == r0 and Hb.ncols == len(a)
        assert np.array_equal(rpb, rp[r0:r1 + 1] - rp[r0])
        assert np.array_equal(cib, ci[rp[r0]:rp[r1]]) and np.array_equal(nzb, nz[rp[r0]:rp[r1]])


def test_edge_cases(ctx):
    # empty list, single determinant, alpha-empty determinants (rows skipped by SDL)
    sp = W.synthetic_molecular("x", 5, 0, 2, 5, 6)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    a, b = port.generate_hilbert_space(5, 0, 2)
    H = ctx.hbuild(ctx.upload_dets(port.pack(a, b)), 0.0)
    assert H.nnz == 0 and np.all(H.download()[0] == 0)
    H1 = ctx.hbuild(ctx.upload_dets(np.array([0b00111 | (0b00011 << 32)], dtype=np.uint64)), 0.0)
    rp, ci, nz = H1.download()
    assert H1.nnz == 1 and ci[0] == 0
    assert nz[0] == port.Ham(sp.norb, sp.T, sp.V).matrix_element(7, 3, 7, 3)
    H0 = ctx.hbuild(ctx.upload_dets(np.zeros(0, dtype=np.uint64)), 0.0)
    assert H0.nrows == 0 and H0.nnz == 0
    # unsorted / repeated alpha runs still give the pairwise pattern (run-length encoder)
    sp2 = W.config("tiny_cas6")
    a2, b2 = port.generate_hilbert_space(6, 3, 3)
    perm = np.random.default_rng(5).permutation(len(a2))
    ctx.upload_integrals(sp2.norb, sp2.T, sp2.V)
    Hp = ctx.hbuild(ctx.upload_dets(port.pack(a2[perm], b2[perm])), EPS)
    orp, oci, onz = port.Ham(sp2.norb, sp2.T, sp2.V).hbuild(a2[perm], b2[perm], EPS)
    rp, ci, nz = Hp.download()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(nz, onz)


def test_wfn128_words(ctx):
    # two-word determinants (wfn_t<128> layout) give the same matrix
    sp = W.config("tiny_cas6")
    a, b = port.generate_hilbert_space(6, 3, 3)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    H1 = ctx.hbuild(ctx.upload_dets(port.pack(a, b), 1), EPS).download()
    H2 = ctx.hbuild(ctx.upload_dets(port.pack(a, b, 128), 2), EPS).download()
    for x, y in zip(H1, H2):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("name", ["small_cas8", "hubbard_4x2"])
def test_sigma_matches_oracle(ctx, name):
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    dets, H = _build(ctx, sp, a, b, EPS)
    rp, ci, nz = H.download()
    rng = np.random.default_rng(1)
    for _ in range(3):
        x = rng.normal(size=len(a))
        y = H.spmv(x)
        yo = port.spmv(rp, ci, nz, x)
        # FP64 sums in a different association order: |dy| <= 1e-13 * sum|h||x| per row
        bound = 1e-13 * port.spmv(rp, ci, np.abs(nz), np.abs(x)) + 1e-300
        assert np.all(np.abs(y - yo) <= bound)
    # linearity (size-independent property)
    x1, x2 = rng.normal(size=len(a)), rng.normal(size=len(a))
    assert np.allclose(H.spmv(2.0 * x1 - 3.0 * x2), 2.0 * H.spmv(x1) - 3.0 * H.spmv(x2), rtol=0, atol=1e-10)


def test_sigma_on_uploaded_csr_all_row_shapes(ctx):
    # ragged rows incl. empty ones: exercises every threads-per-row variant + head/tail paths
    rng = np.random.default_rng(9)
    n = 3000
    for mean in (1, 5, 17, 40, 70, 200):
        lens = rng.poisson(mean, size=n)
        lens[rng.integers(0, n, 50)] = 0
        rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        ci = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens] + [np.zeros(0, int)]).astype(np.int64)
        nz = rng.normal(size=rp[-1])
        M = ctx.upload_csr(rp, ci, nz)
        x = rng.normal(size=n)
        y = M.spmv(x)
        yo = port.spmv(rp, ci, nz, x)
        assert np.allclose(y, yo, rtol=0, atol=1e-11)
        rp2, ci2, nz2 = M.download()
        assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and np.array_equal(nz2, nz)


def test_sigma_row_binned_for_skewed_matrices(ctx, monkeypatch):
    # selected-CI matrices have skewed row lengths: rows are sorted into four length classes (< 48, < 384, < 3072,
    # longer) and every class gets its own lanes-per-row (one CTA per row for the longest). Same result as the
    # oracle and as the single-launch kernel, empty rows and all classes present.
    rng = np.random.default_rng(11)
    n = 9000
    lens = np.concatenate([rng.poisson(7, 3000), rng.poisson(150, 3000), rng.poisson(1200, 2900), rng.integers(3072, 6000, 100)])
    rng.shuffle(lens)
    lens[rng.integers(0, n, 40)] = 0
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ci = np.concatenate([np.sort(rng.choice(n, size=l, replace=False)) for l in lens]).astype(np.int64)
    nz = rng.normal(size=rp[-1])
    x = rng.normal(size=n)
    yo = port.spmv(rp, ci, nz, x)
    bound = 1e-13 * port.spmv(rp, ci, np.abs(nz), np.abs(x)) + 1e-300
    monkeypatch.setenv("B2CI_SPMV_BINS", "1")   # off by default (no gain measured on the ASCI matrices, see spmv.cu)
    M = ctx.upload_csr(rp, ci, nz)
    launches0 = ctx.launch_count
    y = M.spmv(x)
    y2 = M.spmv(x)
    assert np.all(np.abs(y - yo) <= bound) and np.array_equal(y, y2)
    monkeypatch.setenv("B2CI_SPMV_BINS", "0")
    M1 = ctx.upload_csr(rp, ci, nz)
    l1 = ctx.launch_count
    y1 = M1.spmv(x)
    assert ctx.launch_count - l1 == 1                  # the single-launch kernel
    assert np.all(np.abs(y1 - yo) <= bound)
    assert ctx.launch_count - launches0 > 8           # binned: preparation + four launches per product


def test_diagonal(ctx):
    sp = W.config("small_cas8")
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    dets, H = _build(ctx, sp, a, b, EPS)
    rp, ci, nz = H.download()
    assert np.array_equal(H.diagonal(), port.extract_diagonal(rp, ci, nz))


def test_diagonal_of_uploaded_matrices_with_unsorted_or_missing_entries(ctx):
    """extract_diagonal_elements (sparsexx/util/submatrix.hpp:354-383) looks the diagonal up wherever it sits in the
    row. The device kernel finds it by binary search in the sorted rows every build makes and must still find it in
    an uploaded matrix whose rows are not sorted; a row without a diagonal entry gives 0."""
    rng = np.random.default_rng(11)
    n = 300
    rp, ci, nz = [0], [], []
    for i in range(n):
        cols = set(int(c) for c in rng.choice(n, size=int(rng.integers(1, 40)), replace=False))
        if i % 7 != 3:
            cols.add(i)
        else:
            cols.discard(i)
        cols = list(cols)
        rng.shuffle(cols)                       # unsorted rows
        ci += cols
        nz += [float(rng.normal()) for _ in cols]
        rp.append(len(ci))
    M = ctx.upload_csr(rp, ci, nz)
    assert np.array_equal(M.diagonal(), port.extract_diagonal(np.array(rp), np.array(ci), np.array(nz)))
    assert all(M.diagonal()[i] == 0.0 for i in range(3, n, 7))


@pytest.mark.parametrize("name", ["tiny_cas6", "small_cas8", "hubbard_3x2", "hubbard_4x2"])
def test_davidson_energy_and_iterations(ctx, golden_meta, name):
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    dets, H = _build(ctx, sp, a, b, EPS)
    E, X, niter, trace = H.davidson(200, 1e-8)
    rec = golden_meta[f"{name}_eps"]["davidson"]
    assert abs(E - rec["E"]) < 1e-8          # north_star: energies within 1e-8 Eh
    assert abs(niter - rec["niter"]) <= 1     # same iteration count as the reference (+-1)
    assert abs(X @ X - 1.0) < 1e-12           # davidson.cxx:62-72 properties
    rp, ci, nz = H.download()
    assert abs(X @ port.spmv(rp, ci, nz, X) - E) < 1e-11
    Eo, Xo, nito, _ = port.davidson(rp, ci, nz, 200, 1e-8)
    assert abs(E - Eo) < 1e-9 and abs(niter - nito) <= 1


def test_davidson_water_cisd_golden(ctx, water, golden_meta, golden_arrays):
    # external/macis/tests/davidson.cxx:20-75 (max_m = 15, tol = 1e-8)
    a, b = cisd_space(24, 5, 5)
    dets, H = _build(ctx, water, a, b, 1e-16)
    E, X, niter, trace = H.davidson(15, 1e-8)
    ka = golden_meta["known_answers"]
    assert abs(E + water.core_energy - ka["water_cisd_davidson_total"]) < 1e-8
    rec = golden_meta["water_cisd_1e-16"]["davidson"]
    assert abs(E - rec["E"]) < 1e-9 and niter == rec["niter"]
    assert abs(abs(X @ golden_arrays["water_cisd_1e-16.davidson_X"]) - 1.0) < 1e-9


def test_davidson_not_converged_raises(ctx):
    sp = W.config("small_cas8")
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    dets, H = _build(ctx, sp, a, b, EPS)
    with pytest.raises(device.B2ciError) as e:
        H.davidson(3, 1e-12)
    assert "Davidson Did Not Converge!" in str(e.value)


def test_davidson_uploaded_matrices(ctx):
    # python/tests/test_davidson_solver.py: 6x6 tridiagonal with analytic eigenpair, 1x1 case
    n = 6
    rp, ci, nz = [0], [], []
    for i in range(n):
        for j in (i - 1, i, i + 1):
            if 0 <= j < n:
                ci.append(j)
                nz.append(2.0 if i == j else -1.0)
        rp.append(len(ci))
    M = ctx.upload_csr(rp, ci, nz)
    E, X, niter, _ = M.davidson(20, 1e-10)
    assert abs(E - (2 - 2 * np.cos(np.pi / (n + 1)))) < 1e-10
    M1 = ctx.upload_csr([0, 1], [0], [-3.5])
    E1, X1, nit1, _ = M1.davidson(20, 1e-8)
    assert E1 == -3.5 and X1[0] == 1.0 and nit1 == 0


@pytest.mark.parametrize("name", ["small_cas8", "hubbard_4x2", "tiny_cas6"])
@pytest.mark.parametrize("thr", [EPS, 0.0, 1e-3])
def test_scan_and_product_paths_agree(ctx, name, thr):
    """FCI-shaped lists take the product enumeration; B2CI_HBUILD_FORCE_SCAN=1 forces the
    general XOR/popcount beta scan. Both must give the oracle's matrix bit for bit."""
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.upload_dets(port.pack(a, b), 1)
    Hp = ctx.hbuild(dets, thr)
    assert ctx.timer_ms("h_build.rectangular") == 1.0
    os.environ["B2CI_HBUILD_FORCE_SCAN"] = "1"
    try:
        Hs = ctx.hbuild(dets, thr)
        assert ctx.timer_ms("h_build.rectangular") == 0.0
    finally:
        del os.environ["B2CI_HBUILD_FORCE_SCAN"]
    orp, oci, onz = port.Ham(sp.norb, sp.T, sp.V).hbuild(a, b, thr)
    for H in (Hp, Hs):
        rp, ci, nz = H.download()
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(nz, onz)
    # row blocks through the product path
    r0, r1 = len(a) // 3, len(a) // 3 + 777
    rpb, cib, nzb = ctx.hbuild(dets, thr, (r0, min(r1, len(a)))).download()
    r1 = min(r1, len(a))
    assert np.array_equal(rpb, orp[r0:r1 + 1] - orp[r0])
    assert np.array_equal(cib, oci[orp[r0]:orp[r1]]) and np.array_equal(nzb, onz[orp[r0]:orp[r1]])


@pytest.mark.parametrize("name,frac,seed", [("small_cas8", 0.3, 1), ("small_cas8", 0.05, 2), ("hubbard_4x2", 0.5, 3),
                                             ("n2_cas10", 0.02, 4)])
@pytest.mark.parametrize("thr", [EPS, 0.0])
def test_irregular_lists_scan_path(ctx, name, frac, seed, thr):
    """Random subsets of an FCI space (selected-CI shaped: ragged alpha runs, beta strings shared
    between runs) go through the run-scan + beta-group merge; sorted and shuffled-run order."""
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    rng = np.random.default_rng(seed)
    pick = np.sort(rng.choice(len(a), size=max(2, int(frac * len(a))), replace=False))
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    h = port.Ham(sp.norb, sp.T, sp.V)
    for order in ("sorted", "runs_shuffled"):
        sa, sb = a[pick], b[pick]
        if order == "runs_shuffled":
            # keep determinants of one alpha string together but permute the runs and the beta
            # order inside them (the run-length encoder needs grouping, not sorting)
            keys = rng.permutation(len(np.unique(sa)))
            rank = keys[np.searchsorted(np.unique(sa), sa)]
            o = np.lexsort((rng.random(len(sa)), rank))
            sa, sb = sa[o], sb[o]
        H = ctx.hbuild(ctx.upload_dets(port.pack(sa, sb), 1), thr)
        assert ctx.timer_ms("h_build.rectangular") == 0.0
        rp, ci, nz = H.download()
        orp, oci, onz = h.hbuild(sa, sb, thr)
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(nz, onz)
        r0, r1 = len(sa) // 4, len(sa) // 4 + 50
        rpb, cib, nzb = ctx.hbuild(ctx.upload_dets(port.pack(sa, sb), 1), thr, (r0, min(r1, len(sa)))).download()
        r1 = min(r1, len(sa))
        assert np.array_equal(rpb, orp[r0:r1 + 1] - orp[r0]) and np.array_equal(cib, oci[orp[r0]:orp[r1]])


def test_dense_ground_state_matches_davidson(ctx):
    sp = W.config("tiny_cas6")
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    dets, H = _build(ctx, sp, a, b, EPS)
    Ed, xd = H.dense_ground_state()
    E, X, niter, _ = H.davidson(200, 1e-10)
    rp, ci, nz = H.download()
    assert abs(Ed - E) < 1e-9 and abs(abs(xd @ X) - 1) < 1e-8
    assert np.allclose(port.spmv(rp, ci, nz, xd), Ed * xd, atol=1e-10)   # an eigenpair to rounding
    import scipy.sparse as sps
    M = sps.csr_matrix((nz, ci, rp), shape=(len(a), len(a))).toarray()
    assert abs(Ed - np.linalg.eigvalsh(M)[0]) < 1e-11


# ---- patched (incremental) build: CachedHamiltonianState / build_patched_operator ---------------
def _spin_sorted_subset(a, b, pick):
    sa, sb = a[pick], b[pick]
    o = port.spin_sort_order(sa, sb)
    return sa[o], sb[o]


@pytest.mark.parametrize("name,n_old,n_new,n_common,seed", [
    ("small_cas8", 600, 700, 500, 1),     # ordinary refine step: most determinants kept
    ("small_cas8", 300, 900, 300, 2),     # pure growth: old list is a subset
    ("small_cas8", 800, 500, 500, 3),     # pure shrink: nothing added, rows and columns dropped
    ("hubbard_4x2", 1500, 1600, 1200, 4),  # exact zeros at the threshold
    ("n2_cas10", 3000, 3300, 2500, 5),
    ("wide36", 400, 450, 300, 6),         # two-word determinants
])
@pytest.mark.parametrize("thr", [EPS, 0.0])
def test_patched_build_is_bit_identical_to_full_build(ctx, name, n_old, n_new, n_common, seed, thr):
    """The merged patch (kept x kept from the cached matrix, kept x added and added x all freshly
    evaluated) equals the full build of the new list bit for bit, and equals the oracle's CSR."""
    sp = W.config(name)
    rng = np.random.default_rng(seed)
    if name == "wide36":
        pool = set()
        while len(pool) < n_old + n_new:
            pool.add((sum(1 << int(i) for i in rng.choice(sp.norb, sp.nalpha, replace=False)),
                      sum(1 << int(i) for i in rng.choice(sp.norb, sp.nbeta, replace=False))))
        pool = sorted(pool)
        a = np.array([p[0] for p in pool], dtype=np.uint64)
        b = np.array([p[1] for p in pool], dtype=np.uint64)
        nbits = 128
    else:
        a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
        nbits = 64
    perm = rng.permutation(len(a))
    common, only_old, only_new = perm[:n_common], perm[n_common:n_old], perm[n_old:n_old + n_new - n_common]
    oa, ob = _spin_sorted_subset(a, b, np.concatenate([common, only_old]))
    na, nb = _spin_sorted_subset(a, b, np.concatenate([common, only_new]))
    wpd = nbits // 64
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    d_old = ctx.upload_dets(port.pack(oa, ob, nbits), wpd)
    d_new = ctx.upload_dets(port.pack(na, nb, nbits), wpd)
    H_old = ctx.hbuild(d_old, thr)
    H_pat, n_kept = ctx.hbuild_patched(d_old, H_old, d_new, thr, min_overlap=0.3)
    assert n_kept == n_common and H_pat is not None
    rp, ci, nz = H_pat.download()
    frp, fci, fnz = ctx.hbuild(d_new, thr).download()
    assert np.array_equal(rp, frp) and np.array_equal(ci, fci) and np.array_equal(nz, fnz)
    orp, oci, onz = port.Ham(sp.norb, sp.T, sp.V).hbuild(na, nb, thr)
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(nz, onz)
    # a patched matrix can seed the next patch (here: back to the old list)
    H_back, _ = ctx.hbuild_patched(d_new, H_pat, d_old, thr, min_overlap=0.0)
    brp, bci, bnz = H_back.download()
    o2 = H_old.download()
    assert np.array_equal(brp, o2[0]) and np.array_equal(bci, o2[1]) and np.array_equal(bnz, o2[2])


def test_patched_build_overlap_gate_and_errors(ctx):
    sp = W.config("small_cas8")
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    o = port.spin_sort_order(a, b)      # both lists must be spin_comparator-sorted
    a, b = a[o], b[o]
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    d_old = ctx.upload_dets(port.pack(a[:100], b[:100]), 1)
    d_new = ctx.upload_dets(port.pack(a[50:1000], b[50:1000]), 1)
    H_old = ctx.hbuild(d_old, EPS)
    H, n_kept = ctx.hbuild_patched(d_old, H_old, d_new, EPS, min_overlap=0.3)   # 50 / 950 kept
    assert H is None and n_kept == 50
    H, _ = ctx.hbuild_patched(d_old, H_old, d_new, EPS, min_overlap=0.05)
    frp, fci, fnz = ctx.hbuild(d_new, EPS).download()
    rp, ci, nz = H.download()
    assert np.array_equal(rp, frp) and np.array_equal(ci, fci) and np.array_equal(nz, fnz)
    with pytest.raises(device.B2ciError, match="full square"):
        ctx.hbuild_patched(d_old, ctx.hbuild(d_old, EPS, (10, 60)), d_new, EPS)


# ---- hamiltonian_build_algorithm: the pair-based generators' pattern rules ---------------------------
@pytest.mark.parametrize("tag", ["hubbard_4x2_s600", "small_cas8_s900", "n2_cas10_s2500", "alpha_empty_8o"])
@pytest.mark.parametrize("gen", ["residue_arrays", "dynamic_bit_masking", "sorted_double_loop"])
def test_generator_rules_match_reference_golden(ctx, tag, gen):
    """build_csr_from_pairs semantics (diagonal always stored, |h| < thr dropped, alpha-empty
    determinants kept) against fingerprints of the compiled reference's generators."""
    import json
    from helpers import check_generator_golden, generator_case
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "generators_meta.json")) as fh:
        meta = json.load(fh)
    sp, a, b = generator_case(tag)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    ctx.set_hamiltonian_generator(gen)
    try:
        d = ctx.upload_dets(port.pack(a, b), 1)
        for thr_tag in ("eps", "zero", "1e-2"):
            m = meta[f"{tag}.{thr_tag}"]
            H = ctx.hbuild(d, m["thr"])
            check_generator_golden(m[gen], *H.download())
            # row block and patched build follow the same rules
            n = len(a)
            r0, r1 = n // 3, min(n, n // 3 + 40)
            rpb, cib, nzb = ctx.hbuild(d, m["thr"], (r0, r1)).download()
            rp, ci, nz = H.download()
            assert np.array_equal(rpb, rp[r0:r1 + 1] - rp[r0]) and np.array_equal(cib, ci[rp[r0]:rp[r1]])
            # (the patch needs spin-sorted lists; the alpha-empty fixture is in combination order)
            o = port.spin_sort_order(a, b)
            sa, sb = a[o], b[o]
            keep = np.sort(np.random.default_rng(1).choice(n, size=(3 * n) // 4, replace=False))
            d_old = ctx.upload_dets(port.pack(sa[keep], sb[keep]), 1)
            d_new = ctx.upload_dets(port.pack(sa, sb), 1)
            Hp, nk = ctx.hbuild_patched(d_old, ctx.hbuild(d_old, m["thr"]), d_new, m["thr"], 0.3)
            assert nk == len(keep)
            prp, pci, pnz = Hp.download()
            if np.array_equal(o, np.arange(n)):
                check_generator_golden(m[gen], prp, pci, pnz)
            orp, oci, onz = port.Ham(sp.norb, sp.T, sp.V).hbuild(sa, sb, m["thr"], generator=gen)
            assert np.array_equal(prp, orp) and np.array_equal(pci, oci) and np.array_equal(pnz, onz)
    finally:
        ctx.set_hamiltonian_generator("")


def test_fci_list_under_pair_rules_takes_the_scan_path(ctx):
    sp = W.config("hubbard_3x2")
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    ctx.set_hamiltonian_generator("residue_arrays")
    try:
        H = ctx.hbuild(ctx.upload_dets(port.pack(a, b), 1), EPS)
        assert ctx.timer_ms("h_build.rectangular") == 0.0
        rp, ci, nz = H.download()
    finally:
        ctx.set_hamiltonian_generator("")
    orp, oci, onz = port.Ham(sp.norb, sp.T, sp.V).hbuild(a, b, EPS, generator="residue_arrays")
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(nz, onz)
    srp, _, _ = port.Ham(sp.norb, sp.T, sp.V).hbuild(a, b, EPS)
    assert rp[-1] >= srp[-1]


@pytest.mark.parametrize("mode", ["hits", "overflow", "tile_overflow", "no_tile", "tile_all", "wide", "off",
                                  "flat_fill"])
def test_hit_list_fill_equals_rescan_fill(ctx, mode, monkeypatch):
    """General lists scan once: the count pass stores the connections it finds (the tiled scan for units of
    4+ rows of one alpha run, the warp-per-row scan for the rest) and the fill pass evaluates them from the
    store; a store that turns out too small falls back to a second scan. Every way gives the oracle's CSR,
    for the plain build, a row block and the patched build."""
    sp, a, b = None, None, None
    from helpers import generator_case
    sp, a, b = generator_case("n2_cas10_s2500")
    monkeypatch.setenv("B2CI_HBUILD_HITLIST_MIN", "1")
    if mode == "overflow":  # (units below 12 rows to the warp-per-row scan, so that its store has enough to overflow)
        monkeypatch.setenv("B2CI_HBUILD_HITLIST_CAP", "100")
        monkeypatch.setenv("B2CI_HBUILD_TILE_MIN", "12")
    if mode == "tile_overflow":
        monkeypatch.setenv("B2CI_HBUILD_TILE_CAP", "1")
    if mode == "no_tile":
        monkeypatch.setenv("B2CI_HBUILD_NO_TILE", "1")
    if mode == "tile_all":
        monkeypatch.setenv("B2CI_HBUILD_TILE_MIN", "1")
    if mode == "wide":
        monkeypatch.setenv("B2CI_HBUILD_WIDE_STRINGS", "1")
    if mode == "off":
        monkeypatch.setenv("B2CI_HBUILD_NO_HITLIST", "1")
    if mode == "flat_fill":  # connections evaluated in arrival order instead of queued per excitation class
        monkeypatch.setenv("B2CI_HBUILD_FLAT_FILL", "1")
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    h = port.Ham(sp.norb, sp.T, sp.V)
    d = ctx.upload_dets(port.pack(a, b), 1)
    for thr in (EPS, 0.0, 1e-2):
        H = ctx.hbuild(d, thr)
        assert ctx.timer_ms("h_build.hit_lists") == (0.0 if mode in ("overflow", "tile_overflow", "off") else 1.0)
        if mode in ("hits", "wide", "flat_fill"):
            assert ctx.timer_ms("h_build.tile_units") > 0 and ctx.timer_ms("h_build.scan_rows") > 0
        if mode == "tile_all":
            assert ctx.timer_ms("h_build.scan_rows") == 0
        if mode == "no_tile":
            assert ctx.timer_ms("h_build.tile_units") == 0
        rp, ci, nz = H.download()
        orp, oci, onz = h.hbuild(a, b, thr)
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci) and np.array_equal(nz, onz)
        r0, r1 = 700, 1900
        rpb, cib, nzb = ctx.hbuild(d, thr, (r0, r1)).download()
        assert np.array_equal(rpb, orp[r0:r1 + 1] - orp[r0]) and np.array_equal(cib, oci[orp[r0]:orp[r1]])
        assert np.array_equal(nzb, onz[orp[r0]:orp[r1]])
        keep = np.sort(np.random.default_rng(2).choice(len(a), size=2000, replace=False))
        d_old = ctx.upload_dets(port.pack(a[keep], b[keep]), 1)
        Hp, _ = ctx.hbuild_patched(d_old, ctx.hbuild(d_old, thr), d, thr, 0.3)
        prp, pci, pnz = Hp.download()
        assert np.array_equal(prp, orp) and np.array_equal(pci, oci) and np.array_equal(pnz, onz)


def test_balanced_row_partition_evens_out_connections(ctx, monkeypatch):
    """b2ci_dets_balanced_partition: contiguous row blocks with about equal numbers of connections (the row blocks
    of a sharded selected-CI build). The cuts tile [0, n), are reproducible (every rank computes its own copy),
    and the blocks' true nnz -- from the oracle's CSR of the same list -- are closer to the mean than with even
    row counts. The matrix assembled from the blocks is the oracle's, whatever the cuts."""
    from helpers import generator_case
    sp, a, b = generator_case("n2_cas10_s2500")
    monkeypatch.setenv("B2CI_BALANCE_MIN", "64")
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    d = ctx.upload_dets(port.pack(a, b), 1)
    n = len(a)
    orp, oci, onz = port.Ham(sp.norb, sp.T, sp.V).hbuild(a, b, EPS)
    for nparts in (2, 4, 8):
        off = ctx.balanced_partition(d, nparts, 256)
        assert off[0] == 0 and off[-1] == n and np.all(np.diff(off) >= 0)
        assert np.array_equal(off, ctx.balanced_partition(d, nparts, 256))
        nnz_bal = np.diff(orp[off])
        even = np.array([r * (n // nparts) + min(r, n % nparts) for r in range(nparts + 1)])
        nnz_even = np.diff(orp[even])
        assert nnz_bal.max() / nnz_bal.mean() <= max(1.10, 0.999 * nnz_even.max() / nnz_even.mean()), (nnz_bal, nnz_even)
        rp, ci, nz = [], [], []
        for r in range(nparts):
            brp, bci, bnz = ctx.hbuild(d, EPS, (int(off[r]), int(off[r + 1]))).download()
            assert np.array_equal(brp, orp[off[r]:off[r + 1] + 1] - orp[off[r]])
            ci.append(bci)
            nz.append(bnz)
        assert np.array_equal(np.concatenate(ci), oci) and np.array_equal(np.concatenate(nz), onz)
    # a list too short to be worth balancing: even row counts
    monkeypatch.setenv("B2CI_BALANCE_MIN", "100000")
    assert np.array_equal(ctx.balanced_partition(d, 4, 256), [0, 625, 1250, 1875, 2500])
