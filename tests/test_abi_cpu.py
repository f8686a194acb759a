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
