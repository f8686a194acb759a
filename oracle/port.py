"""TEST INFRASTRUCTURE -- ctypes wrapper for oracle/liboracle_port.so (the plain-C
restatement in oracle/port/oracle_port.c) plus the numpy restatement of the ASCI outer
loop (asci_iter / asci_grow / asci_refine).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module; the product (qdk_chemistry_b200/) never does.

Pinned (tests/test_oracle.py): CSR pattern of the water CISD space against the reference's
golden rowptr (external/macis/tests/csr_hamiltonian.cxx:76-99), Davidson / ASCI energies
against external/macis/tests/{davidson,asci}.cxx, and piecewise against oracle/_ref where
that library is built. Golden data made with oracle/_ref additionally pins: the three generators'
pattern rules, wfn_t<128> determinants (search, H build, RDMs, entropies, a whole grow + refine run),
grow_with_rot, the growth back-off / core-selection scenarios of asci.cxx:577-840 and the refine
loop's oscillation handling (tests/golden/make_golden_*.py, tests/README.md).
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
        L.op_hbuild_rows_gen.restype = i64
        L.op_hbuild_rows_gen.argtypes = [vp, vp, vp, i64, i64, i64, dbl, i32, vp, vp, vp]
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

    def rotate(self, Cm: np.ndarray) -> None:
        """two_index_transform + four_index_transform + generate_integral_intermediates
        (src/macis/transform.cxx:22-96; asci/grow.hpp:196-215): T <- C^T T C, V likewise on all four
        indices, in place like the reference (which rotates the generator's own integrals)."""
        n = self.norb
        Cm = np.asarray(Cm, dtype=np.float64).reshape(n, n, order="F")
        T = self.T.reshape(n, n, order="F")
        V = self.V.reshape(n, n, n, n, order="F")
        T = Cm.T @ T @ Cm
        V = np.einsum("ip,ijkl->pjkl", Cm, V, optimize=True)
        V = np.einsum("jq,pjkl->pqkl", Cm, V, optimize=True)
        V = np.einsum("kr,pqkl->pqrl", Cm, V, optimize=True)
        V = np.einsum("ls,pqrl->pqrs", Cm, V, optimize=True)
        lib().op_ham_destroy(self.h)
        self.T = np.ascontiguousarray(T.reshape(-1, order="F"))
        self.V = np.ascontiguousarray(V.reshape(-1, order="F"))
        self.h = lib().op_ham_create(n, _p(self.T), _p(self.V))

    def intermediates(self):
        n = self.norb
        G, Vr, G2, V2 = np.empty(n ** 3), np.empty(n ** 3), np.empty(n * n), np.empty(n * n)
        lib().op_ham_intermediates(self.h, _p(G), _p(Vr), _p(G2), _p(V2))
        return G, Vr, G2, V2

    def matrix_element(self, bra_a, bra_b, ket_a, ket_b) -> float:
        return lib().op_matrix_element(self.h, int(bra_a), int(bra_b), int(ket_a), int(ket_b))

    def hbuild(self, alpha, beta, thresh: float, rows: Optional[Tuple[int, int]] = None,
               generator: str = "sorted_double_loop"):
        a, b = _u64(alpha), _u64(beta)
        n = a.size
        r0, r1 = rows if rows is not None else (0, n)
        rule = 0 if generator in ("", "sdl", "sorted_double_loop") else 1   # residue_arrays, dynamic_bit_masking
        rp = np.zeros(r1 - r0 + 1, dtype=np.int64)
        nnz = lib().op_hbuild_rows_gen(self.h, _p(a), _p(b), n, r0, r1, thresh, rule, _p(rp), None, None)
        ci = np.empty(nnz, dtype=np.int64)
        nz = np.empty(nnz, dtype=np.float64)
        lib().op_hbuild_rows_gen(self.h, _p(a), _p(b), n, r0, r1, thresh, rule, _p(rp), _p(ci), _p(nz))
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
    ci_res_tol=1e-8, ci_max_subspace=200, ci_matel_tol=float(np.finfo(np.float64).eps),
    grow_with_rot=False, rot_size_start=1000)


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
    """include/macis/asci/grow.hpp:45-268. With grow_with_rot the integrals of ``ham`` are rotated
    in place to natural orbitals after every iteration at or above rot_size_start (:163-258)."""
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
        if s["grow_with_rot"] and a.size >= s["rot_size_start"]:
            ordm, _ = form_rdms(ham.norb, a, b, X, spin_dep=False, one=True, two=False)
            _, vec = np.linalg.eigh(-np.asarray(ordm))      # lapack::syev on -ordm: occupations descending
            ham.rotate(vec)
            _, X = _selected_ci_diag(ham, a, b, s, None)      # diagonal guess, final tolerances
        E0 = E
    return E0, a, b, X


def asci_refine(ham: Ham, s: dict, E0: float, a, b, X, info: Optional[dict] = None):
    """include/macis/asci/refine.hpp:44-237, including the oscillation handling: two consecutive sign
    flips of dE with |dE_prev + dE| < tol diagonalise the union of the last two determinant sets
    (warm-started from the current vector) and extend the iteration budget (:118-205).
    ``info`` (optional) receives {"unions": n, "iterations": n}."""
    a, b, X = _u64(a), _u64(b), np.asarray(X, dtype=np.float64)
    ndets = len(a)
    max_iter = s["max_refine_iter"]
    max_ext, total_ext = s["max_refine_iter"], 0
    prev_dE, osc = 0.0, 0
    prev = None
    unions = 0
    it = 0
    converged = False
    while it < max_iter:
        E, a, b, X = asci_iter(ham, s, ndets, E0, a, b, X)
        if len(a) != ndets:
            ndets = len(a)
            if ndets < s["ntdets_min"]:
                break
        dE = E - E0
        if abs(dE) < s["refine_energy_tol"]:
            E0 = E
            converged = True
            break
        if it > 0 and prev_dE * dE < 0 and abs(prev_dE + dE) < s["refine_energy_tol"]:
            osc += 1
            if osc >= 2 and prev is not None:
                pa, pb = prev
                keys = sorted(set(zip(pa.tolist(), pb.tolist())) | set(zip(a.tolist(), b.tolist())))  # spin order
                ua = np.array([k[0] for k in keys], dtype=np.uint64)
                ub = np.array([k[1] for k in keys], dtype=np.uint64)
                cur = {(int(x), int(y)): float(c) for x, y, c in zip(a, b, X)}
                xu = np.array([cur.get(k, 0.0) for k in keys])
                E_union, Xu = _selected_ci_diag(ham, ua, ub, s, xu)
                ext = min(osc, max_ext - total_ext)
                if ext > 0:
                    max_iter += ext
                    total_ext += ext
                a, b, X = ua, ub, Xu
                ndets = len(a)
                E0 = E_union
                prev, osc, prev_dE = None, 0, 0.0
                unions += 1
                it += 1
                continue
        else:
            osc = 0
        prev_dE = dE
        prev = (a.copy(), b.copy())
        E0 = E
        it += 1
    if info is not None:
        info.update(unions=unions, iterations=it + (1 if converged else 0), extensions=total_ext)
    if not converged:
        msg = "ASCI Refine did not converge"
        if total_ext > 0:   # the reference's message (refine.hpp:222-231)
            msg += f" (oscillation detected, {total_ext} extra iterations granted). "
        raise RuntimeError(msg)
    return E0, a, b, X


def asci_run(ham: Ham, na: int, nb: int, refine: bool = True, **settings):
    s = dict(ASCI_DEFAULTS)
    s.update({k: v for k, v in settings.items() if not k.startswith("_")})
    a = np.array([(1 << na) - 1], dtype=np.uint64)
    b = np.array([(1 << nb) - 1], dtype=np.uint64)
    E = ham.matrix_element(a[0], b[0], a[0], b[0])
    X = np.array([1.0])
    E, a, b, X = asci_grow(ham, s, E, a, b, X)
    if refine and s["max_refine_iter"]:
        E, a, b, X = asci_refine(ham, s, E, a, b, X, settings.get("_info"))
    return E, a, b, X


# ---------------------------------------------------------------------------------------
# Orbital entropies (util/entropies.hpp; SortedDoubleLoopHamiltonianGenerator::form_entropies,
# sorted_double_loop.hpp:760-905). Python loops: small cases only.
# ---------------------------------------------------------------------------------------
ENT_VECS = ("a_ii", "b_ii", "ab_iiii")
ENT_MATS = ("a_ij", "b_ij", "aa_iijj", "bb_iijj", "ab_iijj", "ab_ijjj", "ab_jijj", "ab_jjij", "ab_jjji",
            "ab_ijij", "ab_ijji", "aab_iijjjj", "abb_jjiijj", "aab_iijjii", "abb_iiiijj", "abb_ijiijj",
            "aab_iijjij", "aabb_iijjiijj")


def _occ(x):
    x = int(x)
    return [i for i in range(64) if (x >> i) & 1]


def _sx(bra, ket, ex):
    """single_excitation_sign_indices (sd_operations.hpp:394-402)"""
    bra, ket, ex = int(bra), int(ket), int(ex)
    o1 = ((ket & ex) & -(ket & ex)).bit_length() - 1
    v1 = ((bra & ex) & -(bra & ex)).bit_length() - 1
    lo, hi = min(o1, v1), max(o1, v1)
    between = ket & ((1 << hi) - 1) & ~((1 << (lo + 1)) - 1)
    return o1, v1, -1.0 if bin(between).count("1") & 1 else 1.0


def entropy_intermediates(norb, alpha, beta, C, need_s2=True):
    """OrbitalRDMIntermediates accumulated over the pairs i <= j with |C_i C_j| > 1e-16
    (eval_ordm_intermediates, entropies.hpp:906-962; contributions :212-475). Returns a dict of
    vectors (n) and Fortran-indexed matrices M[i, j]."""
    n = int(norb)
    I = {k: np.zeros(n) for k in ENT_VECS}
    I.update({k: np.zeros((n, n)) for k in ENT_MATS})
    alpha = [int(x) for x in alpha]
    beta = [int(x) for x in beta]
    for i in range(len(alpha)):
        ba, bb = alpha[i], beta[i]
        if ba == 0:
            continue
        oa, ob = _occ(ba), _occ(bb)
        for j in range(i, len(alpha)):
            ka, kb = alpha[j], beta[j]
            if ka == 0:
                continue
            exa, exb = ba ^ ka, bb ^ kb
            ca, cb = bin(exa).count("1"), bin(exb).count("1")
            if ca > (2 if need_s2 else 0) or cb > (2 if need_s2 else 0):
                continue
            val = C[i] * C[j]
            if not abs(val) > 1e-16:
                continue
            if ca == 0 and cb == 0:
                for p in oa:
                    I["a_ii"][p] += val
                for p in ob:
                    I["b_ii"][p] += val
                for q in ob:
                    if q in oa:
                        I["ab_iiii"][q] += val
                if not need_s2:
                    continue
                for q in oa:
                    for p in oa:
                        if p == q:
                            continue
                        I["aa_iijj"][p, q] += val
                        if p in ob:
                            I["aab_iijjii"][p, q] += val
                            if q in ob:
                                I["aabb_iijjiijj"][p, q] += val
                        if q in ob:
                            I["aab_iijjjj"][p, q] += val
                for q in ob:
                    for p in ob:
                        if p == q:
                            continue
                        I["bb_iijj"][p, q] += val
                        if p in oa:
                            I["abb_jjiijj"][q, p] += val
                        if q in oa:
                            I["abb_iiiijj"][q, p] += val
                for q in ob:
                    for p in oa:
                        I["ab_iijj"][p, q] += val
            elif (ca, cb) in ((2, 0), (0, 2)):
                transpose = cb == 2
                bra, ket, ex = (bb, kb, exb) if transpose else (ba, ka, exa)
                occ_os = oa if transpose else ob
                o1, v1, sign = _sx(bra, ket, ex)
                sv = sign * val
                x = I["b_ij"] if transpose else I["a_ij"]
                x[v1, o1] += sv
                x[o1, v1] += sv
                m1, m2 = ("ab_jjij", "ab_jjji") if transpose else ("ab_ijjj", "ab_jijj")
                if o1 in occ_os:
                    I[m1][v1, o1] += sv
                    I[m2][v1, o1] += sv
                if v1 in occ_os:
                    I[m1][o1, v1] += sv
                    I[m2][o1, v1] += sv
                if o1 in occ_os and v1 in occ_os:
                    if transpose:
                        I["aab_iijjij"][o1, v1] += sv
                        I["aab_iijjij"][v1, o1] += sv
                    else:
                        I["abb_ijiijj"][v1, o1] += sv
                        I["abb_ijiijj"][o1, v1] += sv
            elif ca == 2 and cb == 2:
                o2, v2, sb = _sx(ba, ka, exa)
                o1, v1, sa = _sx(bb, kb, exb)
                sv = sa * sb * val
                if o1 == o2 and v1 == v2:
                    I["ab_ijij"][v1, o1] += sv
                    I["ab_ijij"][o1, v1] += sv
                elif o1 == v2 and v1 == o2:
                    I["ab_ijji"][v1, o1] += sv
                    I["ab_ijji"][o1, v1] += sv
    return I


def _eig2(a, b, d):
    hs, hd = 0.5 * (a + d), 0.5 * (a - d)
    w = math.sqrt(hd * hd + b * b)
    return hs - w, hs + w


def entropies_from_intermediates(I, need_s2=True):
    """build_s1_entropy / build_s2_entropy / build_mutual_information (entropies.hpp:476-510,
    552-724, 875-886). The 4x4 block is diagonalised with numpy instead of the reference's Jacobi
    sweeps (same eigenvalues to rounding)."""
    eps = float(np.finfo(np.float64).eps)
    n = len(I["a_ii"])
    a, b, d = I["a_ii"], I["b_ii"], I["ab_iiii"]

    def h(v):
        return -v * math.log(v) if v > eps else 0.0

    s1 = np.array([h(1 - a[i] - b[i] + d[i]) + h(a[i] - d[i]) + h(b[i] - d[i]) + h(d[i]) for i in range(n)])
    if not need_s2:
        return s1, None, None
    g = lambda k: I[k]
    s2 = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            AA, BB, AB = g("aa_iijj")[i, j], g("bb_iijj")[i, j], g("ab_iijj")
            x1, x2, x3, x4, x5 = (g("aab_iijjjj")[i, j], g("abb_jjiijj")[i, j], g("aab_iijjii")[i, j],
                                  g("abb_iiiijj")[i, j], g("aabb_iijjiijj")[i, j])
            e = 0.0
            e += h(1 - a[i] - b[i] - a[j] - b[j] + d[i] + d[j] + AA + AB[i, j] + AB[j, i] + BB - x1 - x2 - x3 - x4 + x5)
            v = a[j] - AB[j, i] - AA - AB[j, j] + x1 + x3 + x2 - x5
            v1 = g("a_ij")[i, j] - g("ab_jijj")[j, i] - g("ab_ijjj")[i, j] + g("abb_ijiijj")[i, j]
            v2 = a[i] - AB[i, j] - AA - AB[i, i] + x1 + x3 + x4 - x5
            e += sum(h(t) for t in _eig2(v, v1, v2))
            v = b[j] - AB[i, j] - BB - AB[j, j] + x4 + x1 + x2 - x5
            v1 = g("b_ij")[i, j] - g("ab_jjij")[j, i] - g("ab_jjji")[i, j] + g("aab_iijjij")[i, j]
            v2 = b[i] - AB[j, i] - BB - AB[i, i] + x2 + x3 + x4 - x5
            e += sum(h(t) for t in _eig2(v, v1, v2))
            e += h(AA - x3 - x1 + x5)
            e += h(BB - x4 - x2 + x5)
            B4 = np.zeros((4, 4))
            B4[0, 0] = AB[j, j] - x1 - x2 + x5
            B4[0, 1] = B4[1, 0] = g("ab_ijjj")[i, j] - g("abb_ijiijj")[i, j]
            B4[0, 2] = B4[2, 0] = -g("ab_jjij")[i, j] + g("aab_iijjij")[i, j]
            B4[0, 3] = B4[3, 0] = g("ab_ijij")[i, j]
            B4[1, 1] = AB[i, j] - x4 - x1 + x5
            B4[1, 2] = B4[2, 1] = -g("ab_ijji")[j, i]
            B4[1, 3] = B4[3, 1] = g("ab_jjji")[j, i] - g("aab_iijjij")[i, j]
            B4[2, 2] = AB[j, i] - x3 - x2 + x5
            B4[2, 3] = B4[3, 2] = -g("ab_jijj")[j, i] + g("abb_ijiijj")[i, j]
            B4[3, 3] = d[i] - x3 - x4 + x5
            e += sum(h(float(t)) for t in np.linalg.eigvalsh(B4))
            e += sum(h(t) for t in _eig2(x1 - x5, -g("aab_iijjij")[i, j], x3 - x5))
            e += sum(h(t) for t in _eig2(x2 - x5, -g("abb_ijiijj")[i, j], x4 - x5))
            e += h(x5)
            s2[i, j] = s2[j, i] = e
    mi = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            mi[i, j] = mi[j, i] = s1[i] + s1[j] - s2[i, j]
    return s1, s2, mi


def form_entropies(norb, alpha, beta, C, need_s2=True):
    return entropies_from_intermediates(entropy_intermediates(norb, alpha, beta, C, need_s2), need_s2)
