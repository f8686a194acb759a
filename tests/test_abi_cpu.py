"""CPU-side checks of the product library: it loads, exports every symbol that include/b2ci.h
declares, fails loudly without a GPU, and its __host__ __device__ Slater-Condon / eigen code
agrees with the oracle when evaluated on the host. No kernel runs here."""
import ctypes as C

import numpy as np
import pytest

from oracle import port
from qdk_chemistry_b200 import _lib, device
from qdk_chemistry_b200 import workloads as W


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_ctx_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.B2ciError) as e:
        device.Context(0)
    assert "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("name", ["tiny_cas6", "hubbard_3x2"])
def test_host_matrix_elements_bit_equal_oracle(name):
    sp = W.config(name)
    a, b = port.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    h = port.Ham(sp.norb, sp.T, sp.V)
    rng = np.random.default_rng(7)
    n = len(a)
    pairs = [(i, i) for i in rng.integers(0, n, 20)]
    pairs += [(int(i), int(j)) for i, j in rng.integers(0, n, (4000, 2))]
    checked = 0
    for i, j in pairs:
        if bin(int(a[i] ^ a[j])).count("1") + bin(int(b[i] ^ b[j])).count("1") > 4:
            continue
        ref = h.matrix_element(a[i], b[i], a[j], b[j])
        got = device.host_matrix_element(sp.norb, sp.T, sp.V, a[i], b[i], a[j], b[j])
        assert got == ref  # bit exact
        checked += 1
    assert checked > 200


def test_host_matrix_elements_water_bit_equal(water):
    from helpers import cisd_space
    a, b = cisd_space(24, 5, 5)
    h = port.Ham(water.norb, water.T, water.V)
    rng = np.random.default_rng(11)
    checked = 0
    for i, j in rng.integers(0, len(a), (6000, 2)):
        if bin(int(a[i] ^ a[j])).count("1") + bin(int(b[i] ^ b[j])).count("1") > 4:
            continue
        assert device.host_matrix_element(24, water.T, water.V, a[i], b[i], a[j], b[j]) == \
            h.matrix_element(a[i], b[i], a[j], b[j])
        checked += 1
    assert checked > 100


@pytest.mark.parametrize("n", [1, 2, 3, 8, 37, 120])
def test_host_sym_eig(n):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n))
    A = A + A.T
    Wv, Q = device.host_sym_eig_lower(A)
    assert np.allclose(Wv, np.linalg.eigvalsh(A), atol=1e-11 * max(1, n))
    assert np.allclose(Q.T @ A @ Q, np.diag(Wv), atol=1e-10 * max(1, n))
    assert np.allclose(Q.T @ Q, np.eye(n), atol=1e-12 * max(1, n))


def test_host_sym_eig_degenerate_and_graded():
    A = np.diag([1.0, 1.0, 1.0, 2.0])
    Wv, Q = device.host_sym_eig_lower(A)
    assert np.allclose(Wv, [1, 1, 1, 2])
    B = np.diag([1e-12, 1.0, 1e6]) + 1e-3
    Wv, Q = device.host_sym_eig_lower(B)
    assert np.allclose(Wv, np.linalg.eigvalsh(B), rtol=1e-10)


def test_host_lowest_eigenpair_solver():
    """Rayleigh-Ritz helper (host side, no GPU needed): lowest eigenpair of the projected matrix."""
    import numpy as np
    from qdk_chemistry_b200 import device
    rng = np.random.default_rng(7)
    for k in (1, 2, 3, 7, 33, 64, 150):
        for kind in range(3):
            A = rng.normal(size=(k, k))
            A = A + A.T
            if kind == 1:    # Davidson-shaped: dominant diagonal, weak coupling
                A = np.diag(np.arange(k) * 0.3 - 5.0) + 0.05 * A
            elif kind == 2:  # (nearly) diagonal, including exactly zero off-diagonals for small k
                A = np.diag(np.sort(rng.normal(size=k))) + (1e-9 * A if k > 3 else 0.0)
            lam, v = device.host_sym_eig_lowest(A)
            w = np.linalg.eigvalsh(A)
            scale = max(1.0, np.abs(w).max())
            assert abs(lam - w[0]) <= 1e-13 * scale
            assert np.linalg.norm(A @ v - lam * v) <= 1e-12 * scale
            assert abs(np.linalg.norm(v) - 1.0) < 1e-14


def test_water_known_answers_of_the_reference_matrix_element_tests(water):
    """external/macis/tests/double_loop.cxx:59-195 on the water / cc-pVDZ integrals, evaluated with
    the oracle port AND with the product's own __host__ __device__ Slater-Condon code on the host:
    HF energy, excited diagonals, Brillouin zeros + hermiticity, MP2 from the same integrals."""
    n, nocc = water.norb, 5
    T = np.asarray(water.T).reshape(n, n, order="F")
    V = np.asarray(water.V).reshape(n, n, n, n, order="F")
    h = port.Ham(n, water.T, water.V)
    dev = lambda ba, bb, ka, kb: device.host_matrix_element(n, water.T, water.V, ba, bb, ka, kb)
    hf = (1 << nocc) - 1
    text_tol = 1e-6      # testing::ascii_text_tolerance (FCIDUMP text precision)
    for me in (h.matrix_element, dev):
        EHF = me(hf, hf, hf, hf)
        assert abs(EHF + water.core_energy - (-76.0267803489191)) < text_tol
        s = hf ^ 1 ^ (1 << nocc)                                  # alpha 0 -> nocc
        assert abs(me(s, hf, s, hf) - (-6.488097259228e+01)) < text_tol
        d = s ^ 2 ^ (1 << (nocc + 1))                             # same-spin double
        assert abs(me(d, hf, d, hf) - (-6.314093508151e+01)) < text_tol
        sb = hf ^ 2 ^ (1 << (nocc + 1))                           # opposite-spin double: alpha 0->5, beta 1->6
        assert abs(me(s, sb, s, sb) - (-6.304547887231e+01)) < text_tol
        for i in range(nocc):                                     # Brillouin + hermiticity
            for a in range(nocc, n):
                x = hf ^ (1 << i) ^ (1 << a)
                e1, e2 = me(hf, hf, x, hf), me(x, hf, hf, hf)
                assert abs(e1) < text_tol and abs(e1 - e2) < 1e-12
                e1, e2 = me(hf, hf, hf, x), me(hf, x, hf, hf)
                assert abs(e1) < text_tol and abs(e1 - e2) < 1e-12
    # MP2 (double_loop.cxx:142-195) needs only the integrals: a check of the (pq|rs) index convention
    eps = np.array([T[p, p] + sum(2 * V[p, p, i, i] - V[p, i, i, p] for i in range(nocc)) for p in range(n)])
    emp2 = 0.0
    for i in range(nocc):
        for a in range(nocc, n):
            for j in range(nocc):
                for b in range(nocc, n):
                    emp2 -= V[a, i, b, j] * (2 * V[a, i, b, j] - V[b, i, a, j]) / (eps[a] + eps[b] - eps[i] - eps[j])
    assert abs(emp2 - (-0.203989305096243)) < text_tol


def test_fast_diagonals_of_the_search_equal_full_diagonals(water):
    """fast_diag_single / _ss_double / _os_double (fast_diagonals.ipp:51-127; double_loop.cxx:65-112 asserts
    equality with matrix_element to 1e-12): every candidate of an ASCI search from the HF determinant
    carries E0 - <Q|H|Q> computed the fast way."""
    n = water.norb
    h = port.Ham(n, water.T, water.V)
    hf = np.array([31], dtype=np.uint64)
    E0 = -80.0
    ka, kb, cm, hd = h.asci_candidates(hf, hf, np.array([1.0]), E0, h_el_tol=1e-10)
    assert len(ka) > 1000
    real = np.nonzero(np.isfinite(cm))[0]          # the core determinant itself carries the inf sentinel
    assert len(real) == len(ka) - 1
    for k in np.random.default_rng(3).choice(real, 400, replace=False):
        full = h.matrix_element(ka[k], kb[k], ka[k], kb[k])
        assert abs((E0 - hd[k]) - full) < 1e-11
        assert device.host_matrix_element(n, water.T, water.V, ka[k], kb[k], ka[k], kb[k]) == full


def test_determinant_width_dispatch():
    """dispatch_by_norb (macis_base.hpp:80-100, cpp/tests/test_macis_dispatch.cpp:37-49): wfn_t<64> below 32
    orbitals, wfn_t<128> below 64; wider ladders are not instantiated by this build and say so."""
    f = device.words_per_det_for_norb
    assert [f(n) for n in (1, 12, 31)] == [1, 1, 1]
    assert [f(n) for n in (32, 36, 63)] == [2, 2, 2]
    for n in (64, 127, 2049):
        with pytest.raises(ValueError, match="wfn_t<256"):
            f(n)
