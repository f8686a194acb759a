"""TEST INFRASTRUCTURE -- ctypes wrapper for oracle/liboracle_port.so (the plain-C
restatement in oracle/port/oracle_port.c) plus the numpy restatement of the ASCI outer
loop (asci_iter / asci_grow / asci_refine).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module; the product (qdk_chemistry_b200/) never does.

Pinned (tests/test_oracle.py): CSR pattern of the water CISD space against the reference's
golden rowptr (external/macis/tests/csr_hamiltonian.cxx:76-99), Davidson / ASCI energies
against external/macis/tests/{davidson,asci}.cxx, and piecewise against oracle/_ref where
that library is built.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_port.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "port", "oracle_port.c")
    hdr = os.path.join(_HERE, "port", "oracle_port.h")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
            os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(
            ["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fopenmp",
             "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


class _SearchOpts(C.Structure):
    _fields_ = [("ndets_max", C.c_int64), ("h_el_tol", C.c_double),
                ("rv_prune_tol", C.c_double), ("just_singles", C.c_int32), ("pad", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
        L.op_ham_create.restype = vp
        L.op_ham_create.argtypes = [i32, vp, vp]
        L.op_ham_destroy.argtypes = [vp]
        L.op_ham_intermediates.argtypes = [vp, vp, vp, vp, vp]
        L.op_generate_hilbert_space.restype = i64
        L.op_generate_hilbert_space.argtypes = [i32, i32, i32, vp, vp]
        L.op_matrix_element.restype = dbl
        L.op_matrix_element.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]
        L.op_hbuild_rows.restype = i64
        L.op_hbuild_rows.argtypes = [vp, vp, vp, i64, i64, i64, dbl, vp, vp, vp]
        L.op_spmv.argtypes = [i64, vp, vp, vp, vp, vp]
        L.op_extract_diagonal.argtypes = [i64, vp, vp, vp, vp]
        L.op_davidson.restype = i32
        L.op_davidson.argtypes = [i64, vp, vp, vp, i64, dbl, vp, vp, vp, vp]
        L.op_selected_ci_diag.restype = i32
        L.op_selected_ci_diag.argtypes = [i64, vp, vp, vp, i64, dbl, vp, vp, vp]
        L.op_syev_lower.argtypes = [i32, vp, i32, vp]
        L.op_asci_search.restype = i64
        L.op_asci_search.argtypes = [vp, vp, vp, vp, vp, i64, dbl, vp, vp, i64, vp]
        L.op_asci_candidates.restype = i64
        L.op_asci_candidates.argtypes = [vp, vp, vp, vp, vp, i64, dbl, vp, vp, vp, vp]
        L.op_form_rdms.argtypes = [i32, vp, vp, i64, vp, vp, vp]
        L.op_form_rdms_spin_dep.argtypes = [i32, vp, vp, i64, vp, vp, vp, vp, vp, vp]
        L.op_num_threads.restype = i32
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


# ---------------------------------------------------------------------------------------
# determinant packing helpers (wfn_t<64>: alpha = bits 0..31, beta = bits 32..63;
# wfn_t<128>: word0 = alpha, word1 = beta) -- raw_bitset.hpp:94-106,157-168
# ---------------------------------------------------------------------------------------
def pack(alpha: np.ndarray, beta: np.ndarray, nbits: int = 64) -> np.ndarray:
    a, b = _u64(alpha), _u64(beta)
    if nbits == 64:
        return a | (b << np.uint64(32))
    out = np.empty(2 * a.size, dtype=np.uint64)
    out[0::2] = a
    out[1::2] = b
    return out


def unpack(dets: np.ndarray, nbits: int = 64) -> Tuple[np.ndarray, np.ndarray]:
    d = _u64(dets)
    if nbits == 64:
        return d & np.uint64(0xFFFFFFFF), d >> np.uint64(32)
    return d[0::2].copy(), d[1::2].copy()


def spin_sort_order(alpha: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """argsort by spin_comparator (alpha-major, then beta) -- raw_bitset.hpp:119-141."""
    return np.lexsort((_u64(beta), _u64(alpha)))


def generate_hilbert_space(norb: int, na: int, nb: int) -> Tuple[np.ndarray, np.ndarray]:
    n = math.comb(norb, na) * math.comb(norb, nb)
    a = np.empty(n, dtype=np.uint64)
    b = np.empty(n, dtype=np.uint64)
    r = lib().op_generate_hilbert_space(norb, na, nb, _p(a), _p(b))
    assert r == n
    return a, b


class Ham:
    def __init__(self, norb: int, T: np.ndarray, V: np.ndarray):
        self.norb = norb
        self.T = np.ascontiguousarray(T, dtype=np.float64)
        self.V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1)
        assert self.T.size == norb * norb and self.V.size == norb ** 4
        self.h = lib().op_ham_create(norb, _p(self.T), _p(self.V))

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.op_ham_destroy(self.h)
            self.h = None

    def intermediates(self):
        n = self.norb
        G, Vr, G2, V2 = np.empty(n ** 3), np.empty(n ** 3), np.empty(n * n), np.empty(n * n)
        lib().op_ham_intermediates(self.h, _p(G), _p(Vr), _p(G2), _p(V2))
        return G, Vr, G2, V2

    def matrix_element(self, bra_a, bra_b, ket_a, ket_b) -> float:
        return lib().op_matrix_element(self.h, int(bra_a), int(bra_b), int(ket_a), int(ket_b))

    def hbuild(self, alpha, beta, thresh: float, rows: Optional[Tuple[int, int]] = None):
        a, b = _u64(alpha), _u64(beta)
        n = a.size
        r0, r1 = rows if rows is not None else (0, n)
        rp = np.zeros(r1 - r0 + 1, dtype=np.int64)
        nnz = lib().op_hbuild_rows(self.h, _p(a), _p(b), n, r0, r1, thresh, _p(rp), None, None)
        ci = np.empty(nnz, dtype=np.int64)
        nz = np.empty(nnz, dtype=np.float64)
        lib().op_hbuild_rows(self.h, _p(a), _p(b), n, r0, r1, thresh, _p(rp), _p(ci), _p(nz))
        return rp, ci, nz

    def asci_candidates(self, calpha, cbeta, coeff, E0, h_el_tol=1e-8, just_singles=False):
        ca, cb = _u64(calpha), _u64(cbeta)
        cf = np.ascontiguousarray(coeff, dtype=np.float64)
        o = _SearchOpts(0, h_el_tol, 0.0, int(just_singles), 0)
        n = lib().op_asci_candidates(self.h, C.byref(o), _p(ca), _p(cb), _p(cf), ca.size, E0,
                                     None, None, None, None)
        oa, ob = np.empty(n, dtype=np.uint64), np.empty(n, dtype=np.uint64)
        cm, hd = np.empty(n), np.empty(n)
        lib().op_asci_candidates(self.h, C.byref(o), _p(ca), _p(cb), _p(cf), ca.size, E0,
                                 _p(oa), _p(ob), _p(cm), _p(hd))
        return oa, ob, cm, hd

    def asci_pt2(self, alpha, beta, coeff, E_asci, pt2_tol=1e-16):
        """asci_pt2_constraint (asci/pt2.hpp:61-565) without the constraint partition (which only
        distributes the work: every external determinant belongs to exactly one constraint, so
        the per-determinant sums are the same): contributions with |c*h| >= pt2_tol summed per
        determinant (accumulate_asci_pairs), wavefunction members dropped (inf sentinel,
        :310-311,399-404), EPT2 = sum rv * c_times_matel (determinant_contributions.hpp:49-55).
        Returns (EPT2, NPT2)."""
        _, _, cm, hd = self.asci_candidates(alpha, beta, coeff, E_asci, h_el_tol=pt2_tol)
        fin = np.isfinite(cm)
        return float(np.sum((cm[fin] / hd[fin]) * cm[fin])), int(np.count_nonzero(fin))

    def asci_search(self, calpha, cbeta, coeff, E0, ndets_max, h_el_tol=1e-8,
                    rv_prune_tol=1e-8, just_singles=False):
        """Returns (alpha, beta, stats): selected determinants followed by the core ones."""
        ca, cb = _u64(calpha), _u64(cbeta)
        cf = np.ascontiguousarray(coeff, dtype=np.float64)
        o = _SearchOpts(int(ndets_max), h_el_tol, rv_prune_tol, int(just_singles), 0)
        cap = int(ndets_max) + ca.size + 1024
        stats = np.zeros(8)
        while True:
            oa, ob = np.empty(cap, dtype=np.uint64), np.empty(cap, dtype=np.uint64)
            r = lib().op_asci_search(self.h, C.byref(o), _p(ca), _p(cb), _p(cf), ca.size, E0,
                                     _p(oa), _p(ob), cap, _p(stats))
            if r < 0:
                cap = -r
                continue
            return oa[:r].copy(), ob[:r].copy(), stats


def form_rdms(norb, alpha, beta, C, spin_dep=False, one=True, two=True):
    """SortedDoubleLoopHamiltonianGenerator::form_rdms / form_rdms_spin_dep for bra == ket
    (sorted_double_loop.hpp:512-760; contribution rules util/rdms.hpp). Returns Fortran-ordered
    (n, n) / (n, n, n, n) arrays: (ordm, trdm) or (aa, bb, aaaa, bbbb, aabb)."""
    a, b = _u64(alpha), _u64(beta)
    c = np.ascontiguousarray(C, dtype=np.float64)
    n = int(norb)
    mk1 = lambda: np.zeros(n * n) if one else None
    mk2 = lambda: np.zeros(n ** 4) if two else None
    sh = lambda x, k: None if x is None else x.reshape((n,) * k, order="F")
    if spin_dep:
        o1, o2, t1, t2, t3 = mk1(), mk1(), mk2(), mk2(), mk2()
        lib().op_form_rdms_spin_dep(n, _p(a), _p(b), a.size, _p(c), _p(o1), _p(o2), _p(t1), _p(t2), _p(t3))
        return sh(o1, 2), sh(o2, 2), sh(t1, 4), sh(t2, 4), sh(t3, 4)
    o1, t1 = mk1(), mk2()
    lib().op_form_rdms(n, _p(a), _p(b), a.size, _p(c), _p(o1), _p(t1))
    return sh(o1, 2), sh(t1, 4)


def spmv(rowptr, colind, nzval, x):
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colind, dtype=np.int64)
    nz = np.ascontiguousarray(nzval, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(rp.size - 1)
    lib().op_spmv(rp.size - 1, _p(rp), _p(ci), _p(nz), _p(x), _p(y))
    return y


def extract_diagonal(rowptr, colind, nzval):
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colind, dtype=np.int64)
    nz = np.ascontiguousarray(nzval, dtype=np.float64)
    d = np.empty(rp.size - 1)
    lib().op_extract_diagonal(rp.size - 1, _p(rp), _p(ci), _p(nz), _p(d))
    return d


class DavidsonError(RuntimeError):
    pass


def davidson(rowptr, colind, nzval, max_m: int, tol: float, x0: Optional[np.ndarray] = None,
             guess_policy: bool = True):
    """Returns (E, X, niter, trace[(lambda, rnorm)]). Raises DavidsonError like the
    reference's "Davidson Did Not Converge!" (davidson.hpp:368)."""
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colind, dtype=np.int64)
    nz = np.ascontiguousarray(nzval, dtype=np.float64)
    n = rp.size - 1
    X = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    niter, eig = C.c_int64(0), C.c_double(0.0)
    trace = np.zeros(2 * (max_m + 1))
    if guess_policy:
        rc = lib().op_selected_ci_diag(n, _p(rp), _p(ci), _p(nz), max_m, tol, _p(X),
                                       C.byref(niter), C.byref(eig))
    else:
        if x0 is None:
            X[int(np.argmin(extract_diagonal(rp, ci, nz)))] = 1.0
        rc = lib().op_davidson(n, _p(rp), _p(ci), _p(nz), max_m, tol, _p(X), C.byref(niter),
                               C.byref(eig), _p(trace))
    if rc == 1:
        raise DavidsonError("Davidson Did Not Converge!")
    if rc == 2:
        raise DavidsonError("gram_schmidt: Unable to find orthogonal vector")
    return eig.value, X, niter.value, trace.reshape(-1, 2)


def syev_lower(A: np.ndarray):
    n = A.shape[0]
    M = np.asfortranarray(A, dtype=np.float64).copy(order="F")
    W = np.empty(n)
    lib().op_syev_lower(n, _p(M), n, _p(W))
    return W, M


# ---------------------------------------------------------------------------------------
# ASCI outer loop (numpy restatement)
# ---------------------------------------------------------------------------------------
ASCI_DEFAULTS = dict(
    ntdets_max=100000, ntdets_min=100, ncdets_max=100, core_selection_strategy="percentage",
    core_selection_threshold=0.95, h_el_tol=1e-8, rv_prune_tol=1e-8, just_singles=False,
    grow_factor=8.0, min_grow_factor=1.01, growth_backoff_rate=0.5, growth_recovery_rate=1.1,
    max_refine_iter=6, refine_energy_tol=1e-6, warm_start_davidson=True,
    min_warm_start_overlap=0.5, grow_ci_residual_tolerance=0.0, taper_grow_factor=0.0,
    ci_res_tol=1e-8, ci_max_subspace=200, ci_matel_tol=float(np.finfo(np.float64).eps))


def _selected_ci_diag(ham: Ham, a, b, s, c0, res_tol=None):
    rp, ci, nz = ham.hbuild(a, b, s["ci_matel_tol"])
    E, X, niter, _ = davidson(rp, ci, nz, s["ci_max_subspace"],
                              s["ci_res_tol"] if res_tol is None else res_tol, c0, True)
    return E, X


def asci_iter(ham: Ham, s: dict, ndets_max: int, E0: float, a, b, X, res_tol=None):
    """include/macis/asci/iteration.hpp:50-226. Ties in |c| are broken by current position
    (stable sort); the reference's std::sort leaves them implementation-defined."""
    a, b, X = _u64(a), _u64(b), np.asarray(X, dtype=np.float64)
    if a.size > 1:
        order = np.argsort(-np.abs(X), kind="stable")
        a, b, X = a[order], b[order], X[order]
    if s["core_selection_strategy"] == "fixed":
        nkeep = min(s["ncdets_max"], a.size)
    else:
        w = 0.0
        nkeep = 0
        for i in range(a.size):
            w += X[i] * X[i]
            nkeep += 1
            if w >= s["core_selection_threshold"]:
                break
    if a.size > 1:
        o2 = spin_sort_order(a[:nkeep], b[:nkeep])
        a[:nkeep], b[:nkeep], X[:nkeep] = a[:nkeep][o2], b[:nkeep][o2], X[:nkeep][o2]
    old = {(int(x), int(y)): float(c) for x, y, c in zip(a, b, X)}
    na_, nb_, _ = ham.asci_search(a[:nkeep], b[:nkeep], X[:nkeep], E0, ndets_max,
                                  s["h_el_tol"], s["rv_prune_tol"], s["just_singles"])
    o3 = spin_sort_order(na_, nb_)
    na_, nb_ = na_[o3], nb_[o3]
    c0 = None
    if s["warm_start_davidson"]:
        c0 = np.array([old.get((int(x), int(y)), 0.0) for x, y in zip(na_, nb_)])
        nrm = float(np.sqrt(np.sum(c0 * c0)))
        if nrm < max(s["min_warm_start_overlap"], np.finfo(np.float64).eps):
            c0 = None
        else:
            c0 = c0 * (1.0 / nrm)
    E, Xn = _selected_ci_diag(ham, na_, nb_, s, c0, res_tol)
    return E, na_, nb_, Xn


def asci_grow(ham: Ham, s: dict, E0: float, a, b, X):
    """include/macis/asci/grow.hpp:45-268 (no orbital rotation)."""
    a, b, X = _u64(a), _u64(b), np.asarray(X, dtype=np.float64)
    res_tol = s["grow_ci_residual_tolerance"] if s["grow_ci_residual_tolerance"] > 0 else None
    prev = a.size
    gf = s["grow_factor"]
    while a.size < s["ntdets_max"]:
        eff = gf
        if s["taper_grow_factor"] > 0 and math.ceil(a.size * gf) > s["ntdets_max"]:
            eff = max(s["min_grow_factor"], s["taper_grow_factor"])
        nnew = min(max(s["ntdets_min"], int(math.ceil(a.size * eff))), s["ntdets_max"])
        if nnew <= a.size:
            nnew = min(a.size + 1, s["ntdets_max"])
            if nnew <= a.size:
                break
        E, a, b, X = asci_iter(ham, s, nnew, E0, a, b, X, res_tol)
        if a.size < nnew:
            gf = max(s["min_grow_factor"], gf * s["growth_backoff_rate"])
            if a.size <= prev:
                break
        else:
            gf = min(s["grow_factor"], gf * s["growth_recovery_rate"])
        prev = a.size
        E0 = E
    return E0, a, b, X


def asci_refine(ham: Ham, s: dict, E0: float, a, b, X):
    """include/macis/asci/refine.hpp:44-237 without the oscillation/union branch (raises
    if that branch would be needed so a test never silently diverges)."""
    ndets = len(a)
    for it in range(s["max_refine_iter"]):
        E, a, b, X = asci_iter(ham, s, ndets, E0, a, b, X)
        if len(a) != ndets:
            ndets = len(a)
            if ndets < s["ntdets_min"]:
                break
        dE = E - E0
        if abs(dE) < s["refine_energy_tol"]:
            return E, a, b, X
        E0 = E
    raise RuntimeError("ASCI Refine did not converge")


def asci_run(ham: Ham, na: int, nb: int, refine: bool = True, **settings):
    s = dict(ASCI_DEFAULTS)
    s.update(settings)
    a = np.array([(1 << na) - 1], dtype=np.uint64)
    b = np.array([(1 << nb) - 1], dtype=np.uint64)
    E = ham.matrix_element(a[0], b[0], a[0], b[0])
    X = np.array([1.0])
    E, a, b, X = asci_grow(ham, s, E, a, b, X)
    if refine and s["max_refine_iter"]:
        E, a, b, X = asci_refine(ham, s, E, a, b, X)
    return E, a, b, X
