"""TEST INFRASTRUCTURE -- ctypes wrapper for oracle/_ref/libmacis_ref.so.

The library is the UNMODIFIED reference (MACIS) compiled by oracle/Makefile from
/root/reference; see oracle/ref_driver.cxx for the reference function behind each
call. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libmacis_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{_LIB_PATH} not built (run `make -C oracle ref` where "
                               "/root/reference exists)")
        L = C.CDLL(_LIB_PATH)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
        L.ref_last_error.restype = C.c_char_p
        L.ref_num_threads.restype = i32
        L.ref_set_num_threads.argtypes = [i32]
        L.ref_set_verbose.argtypes = [i32]
        L.ref_generate_hilbert_space.restype = i64
        L.ref_generate_hilbert_space.argtypes = [i32, i32, i32, i32, vp, i64]
        L.ref_hamgen_create.restype = vp
        L.ref_hamgen_create.argtypes = [i32, i32, vp, vp]
        L.ref_hamgen_destroy.argtypes = [vp]
        L.ref_hamgen_intermediates.argtypes = [vp, vp, vp]
        L.ref_matrix_element.restype = dbl
        L.ref_matrix_element.argtypes = [vp, vp, vp]
        L.ref_form_rdms.argtypes = [vp, i32, vp, i64, vp, i32, vp, vp, vp, vp, vp]
        L.ref_form_entropies.argtypes = [vp, i32, vp, i64, vp, vp, vp, vp]
        L.ref_hbuild.restype = vp
        L.ref_hbuild.argtypes = [vp, i32, vp, i64, vp, i64, dbl, vp]
        L.ref_csr_from_arrays.restype = vp
        L.ref_csr_from_arrays.argtypes = [i64, i64, vp, vp, vp]
        L.ref_csr_nrows.restype = i64
        L.ref_csr_nrows.argtypes = [vp]
        L.ref_csr_nnz.restype = i64
        L.ref_csr_nnz.argtypes = [vp]
        L.ref_csr_copy.argtypes = [vp, vp, vp, vp]
        L.ref_csr_free.argtypes = [vp]
        L.ref_csr_copy_rows.argtypes = [vp, i64, i64, vp, vp, vp]
        L.ref_spmv.restype = dbl
        L.ref_spmv.argtypes = [vp, vp, vp, i32]
        L.ref_csr_diagonal.argtypes = [vp, vp]
        L.ref_davidson.restype = i32
        L.ref_davidson.argtypes = [vp, i64, dbl, vp, vp, vp, i32]
        L.ref_selected_ci_diag.restype = i32
        L.ref_selected_ci_diag.argtypes = [vp, vp, i64, dbl, i64, dbl, vp, vp]
        L.ref_asci_search.restype = i64
        L.ref_asci_search.argtypes = [vp, vp, i64, vp, i64, dbl, vp, vp, i64]
        L.ref_asci_run.restype = vp
        L.ref_asci_run.argtypes = [vp, vp, i32, i32, i32]
        L.ref_asci_result_n.restype = i64
        L.ref_asci_result_n.argtypes = [vp]
        L.ref_asci_result_energy.restype = dbl
        L.ref_asci_result_energy.argtypes = [vp]
        L.ref_asci_result_copy.argtypes = [vp, vp, vp]
        L.ref_asci_result_free.argtypes = [vp]
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class AsciOpts(C.Structure):
    """Mirror of AsciOpts in oracle/ref_driver.cxx (macis::ASCISettings + MCSCFSettings);
    defaults are QDK's (cpp/src/qdk/chemistry/algorithms/microsoft/macis_asci.hpp:34-183)."""

    _fields_ = [
        ("ntdets_max", C.c_int64), ("ntdets_min", C.c_int64), ("ncdets_max", C.c_int64),
        ("core_selection_strategy", C.c_int32), ("just_singles", C.c_int32),
        ("core_selection_threshold", C.c_double), ("h_el_tol", C.c_double),
        ("rv_prune_tol", C.c_double), ("pair_size_max", C.c_int64),
        ("grow_factor", C.c_double), ("min_grow_factor", C.c_double),
        ("growth_backoff_rate", C.c_double), ("growth_recovery_rate", C.c_double),
        ("max_refine_iter", C.c_int64), ("refine_energy_tol", C.c_double),
        ("warm_start_davidson", C.c_int32), ("constraint_level", C.c_int32),
        ("min_warm_start_overlap", C.c_double), ("min_patch_overlap", C.c_double),
        ("grow_ci_residual_tolerance", C.c_double), ("taper_grow_factor", C.c_double),
        ("ci_res_tol", C.c_double), ("ci_max_subspace", C.c_int64),
        ("ci_matel_tol", C.c_double),
        ("grow_with_rot", C.c_int64), ("rot_size_start", C.c_int64),
    ]

    def __init__(self, **kw):
        d = dict(ntdets_max=100000, ntdets_min=100, ncdets_max=100,
                 core_selection_strategy=1, just_singles=0, core_selection_threshold=0.95,
                 h_el_tol=1e-8, rv_prune_tol=1e-8, pair_size_max=500000000,
                 grow_factor=8.0, min_grow_factor=1.01, growth_backoff_rate=0.5,
                 growth_recovery_rate=1.1, max_refine_iter=6, refine_energy_tol=1e-6,
                 warm_start_davidson=1, constraint_level=2, min_warm_start_overlap=0.5,
                 min_patch_overlap=0.3, grow_ci_residual_tolerance=0.0,
                 taper_grow_factor=0.0, ci_res_tol=1e-8, ci_max_subspace=200,
                 ci_matel_tol=float(np.finfo(np.float64).eps), grow_with_rot=0, rot_size_start=1000)
        d.update(kw)
        super().__init__(**d)


# generator selector of oracle/ref_driver.cxx (HamGen::gen)
_GEN = {"sdl": 0, "sorted_double_loop": 0, "double_loop": 1, "residue_arrays": 2, "dynamic_bit_masking": 3}


def words_per_det(nbits: int) -> int:
    return nbits // 64


def nbits_for_norb(norb: int) -> int:
    # dispatch_by_norb, cpp/src/qdk/chemistry/algorithms/microsoft/macis_base.hpp:80-100
    if norb < 32:
        return 64
    if norb < 64:
        return 128
    raise ValueError("oracle/_ref driver instantiates wfn_t<64> and wfn_t<128> only")


def generate_hilbert_space(norb: int, na: int, nb: int, nbits: int = 64) -> np.ndarray:
    from math import comb
    n = comb(norb, na) * comb(norb, nb)
    out = np.empty(n * words_per_det(nbits), dtype=np.uint64)
    r = lib().ref_generate_hilbert_space(nbits, norb, na, nb, _p(out), n)
    if r != n:
        raise RuntimeError(f"ref_generate_hilbert_space returned {r}")
    return out


class Csr:
    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if self.h and _lib is not None:
            _lib.ref_csr_free(self.h)
            self.h = None

    @property
    def n(self) -> int:
        return lib().ref_csr_nrows(self.h)

    @property
    def nnz(self) -> int:
        return lib().ref_csr_nnz(self.h)

    def arrays(self) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        rp = np.empty(self.n + 1, dtype=np.int64)
        ci = np.empty(self.nnz, dtype=np.int64)
        nz = np.empty(self.nnz, dtype=np.float64)
        lib().ref_csr_copy(self.h, _p(rp), _p(ci), _p(nz))
        return rp, ci, nz

    def rowptr(self) -> np.ndarray:
        rp = np.empty(self.n + 1, dtype=np.int64)
        lib().ref_csr_copy_rows(self.h, 0, self.n, _p(rp), None, None)
        return rp

    def rows(self, r0: int, r1: int, rowptr: Optional[np.ndarray] = None):
        """(local rowptr, colind, nzval) of rows [r0, r1) without copying the whole matrix"""
        rp = np.empty(r1 - r0 + 1, dtype=np.int64)
        lib().ref_csr_copy_rows(self.h, r0, r1, _p(rp), None, None)
        ci = np.empty(int(rp[-1]), dtype=np.int64)
        nz = np.empty(int(rp[-1]), dtype=np.float64)
        lib().ref_csr_copy_rows(self.h, r0, r1, None, _p(ci), _p(nz))
        return rp, ci, nz

    def spmv(self, x: np.ndarray, nrep: int = 1) -> Tuple[np.ndarray, float]:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.n, dtype=np.float64)
        t = lib().ref_spmv(self.h, _p(x), _p(y), nrep)
        return y, t

    def diagonal(self) -> np.ndarray:
        d = np.empty(self.n, dtype=np.float64)
        lib().ref_csr_diagonal(self.h, _p(d))
        return d

    def davidson(self, max_m: int, tol: float, x0: Optional[np.ndarray] = None,
                 guess_policy: bool = True):
        """Returns (E, X, niter). guess_policy=True follows serial_selected_ci_diag."""
        n = self.n
        X = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        if not guess_policy and x0 is None:
            X[int(np.argmin(self.diagonal()))] = 1.0
        niter = C.c_int64(0)
        eig = C.c_double(0.0)
        rc = lib().ref_davidson(self.h, max_m, tol, _p(X), C.byref(niter), C.byref(eig),
                                1 if guess_policy else 0)
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())
        return eig.value, X, niter.value


def csr_from_arrays(rowptr, colind, nzval) -> Csr:
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colind, dtype=np.int64)
    nz = np.ascontiguousarray(nzval, dtype=np.float64)
    h = lib().ref_csr_from_arrays(len(rp) - 1, len(ci), _p(rp), _p(ci), _p(nz))
    if not h:
        raise RuntimeError(lib().ref_last_error().decode())
    return Csr(h)


class HamGen:
    """SortedDoubleLoopHamiltonianGenerator (+ DoubleLoop) over caller integrals."""

    def __init__(self, norb: int, T: np.ndarray, V: np.ndarray, nbits: Optional[int] = None):
        self.norb = norb
        self.nbits = nbits or nbits_for_norb(norb)
        self.T = np.ascontiguousarray(T, dtype=np.float64)
        self.V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1)
        assert self.T.size == norb * norb and self.V.size == norb ** 4
        self.h = lib().ref_hamgen_create(self.nbits, norb, _p(self.T), _p(self.V))
        if not self.h:
            raise RuntimeError(lib().ref_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.ref_hamgen_destroy(self.h)
            self.h = None

    def intermediates(self):
        n = self.norb
        G = np.empty(n ** 3)
        Vr = np.empty(n ** 3)
        lib().ref_hamgen_intermediates(self.h, _p(G), _p(Vr))
        return G, Vr

    def matrix_element(self, bra, ket) -> float:
        b = np.atleast_1d(np.asarray(bra, dtype=np.uint64))
        k = np.atleast_1d(np.asarray(ket, dtype=np.uint64))
        return lib().ref_matrix_element(self.h, _p(b), _p(k))

    def hbuild(self, dets: np.ndarray, thresh: float, generator: str = "sdl",
               kets: Optional[np.ndarray] = None) -> Tuple[Csr, float]:
        w = words_per_det(self.nbits)
        d = np.ascontiguousarray(dets, dtype=np.uint64)
        k = None if kets is None else np.ascontiguousarray(kets, dtype=np.uint64)
        sec = C.c_double(0.0)
        h = lib().ref_hbuild(self.h, _GEN[generator], _p(d),
                             d.size // w, _p(k), 0 if k is None else k.size // w,
                             thresh, C.byref(sec))
        if not h:
            raise RuntimeError(lib().ref_last_error().decode())
        return Csr(h), sec.value

    def form_rdms(self, dets: np.ndarray, C: np.ndarray, spin_dep: bool = False,
                  generator: str = "sdl", one: bool = True, two: bool = True):
        """form_rdms -> (ordm, trdm); form_rdms_spin_dep -> (aa, bb, aaaa, bbbb, aabb).
        Matrices are (n, n) / (n, n, n, n) Fortran-ordered views of the reference's buffers."""
        d = np.ascontiguousarray(dets, dtype=np.uint64)
        c = np.ascontiguousarray(C, dtype=np.float64)
        n = self.norb
        mk1 = lambda: np.zeros(n * n) if one else None
        mk2 = lambda: np.zeros(n ** 4) if two else None
        if spin_dep:
            o1, o2, t1, t2, t3 = mk1(), mk1(), mk2(), mk2(), mk2()
        else:
            o1, o2, t1, t2, t3 = mk1(), None, mk2(), None, None
        rc = lib().ref_form_rdms(self.h, _GEN[generator], _p(d), c.size, _p(c),
                                 1 if spin_dep else 0, _p(o1), _p(o2), _p(t1), _p(t2), _p(t3))
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())
        sh = lambda a, k: None if a is None else a.reshape((n,) * k, order="F")
        if spin_dep:
            return sh(o1, 2), sh(o2, 2), sh(t1, 4), sh(t2, 4), sh(t3, 4)
        return sh(o1, 2), sh(t1, 4)

    def form_entropies(self, dets: np.ndarray, C: np.ndarray, s2: bool = True, mi: bool = True,
                       generator: str = "sdl"):
        """form_entropies -> (s1[n], s2[n, n] or None, mutual_information[n, n] or None)"""
        d = np.ascontiguousarray(dets, dtype=np.uint64)
        c = np.ascontiguousarray(C, dtype=np.float64)
        n = self.norb
        s1 = np.zeros(n)
        S2 = np.zeros(n * n) if s2 else None
        MI = np.zeros(n * n) if mi else None
        rc = lib().ref_form_entropies(self.h, _GEN[generator], _p(d), c.size, _p(c),
                                      _p(s1), _p(S2), _p(MI))
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())
        sh = lambda a: None if a is None else a.reshape(n, n, order="F")
        return s1, sh(S2), sh(MI)

    def selected_ci_diag(self, dets: np.ndarray, h_el_tol: float, max_m: int,
                         res_tol: float, c0: Optional[np.ndarray] = None):
        w = words_per_det(self.nbits)
        d = np.ascontiguousarray(dets, dtype=np.uint64)
        n = d.size // w
        Cv = np.zeros(n) if c0 is None else np.array(c0, dtype=np.float64, copy=True)
        E = C.c_double(0.0)
        rc = lib().ref_selected_ci_diag(self.h, _p(d), n, h_el_tol, max_m, res_tol,
                                        _p(Cv), C.byref(E))
        if rc:
            raise RuntimeError(lib().ref_last_error().decode())
        return E.value, Cv

    def asci_search(self, opts: AsciOpts, ndets_max: int, cdets: np.ndarray,
                    coeffs: np.ndarray, E0: float) -> np.ndarray:
        w = words_per_det(self.nbits)
        cd = np.ascontiguousarray(cdets, dtype=np.uint64)
        nc = cd.size // w
        cf = np.ascontiguousarray(coeffs, dtype=np.float64)
        cap = max(4 * ndets_max + nc, 1024)
        while True:
            out = np.empty(cap * w, dtype=np.uint64)
            r = lib().ref_asci_search(self.h, C.byref(opts), ndets_max, _p(cd), nc, E0,
                                      _p(cf), _p(out), cap)
            if r == -(2 ** 63):
                raise RuntimeError(lib().ref_last_error().decode())
            if r < 0:
                cap = -r
                continue
            return out[: r * w].copy()

    def asci_run(self, opts: AsciOpts, na: int, nb: int, refine: bool = True):
        r = lib().ref_asci_run(self.h, C.byref(opts), na, nb, 1 if refine else 0)
        if not r:
            raise RuntimeError(lib().ref_last_error().decode())
        try:
            n = lib().ref_asci_result_n(r)
            E = lib().ref_asci_result_energy(r)
            dets = np.empty(n * words_per_det(self.nbits), dtype=np.uint64)
            Cv = np.empty(n, dtype=np.float64)
            lib().ref_asci_result_copy(r, _p(dets), _p(Cv))
        finally:
            lib().ref_asci_result_free(r)
        return E, dets, Cv


def num_threads() -> int:
    return lib().ref_num_threads()


def set_num_threads(n: int) -> None:
    lib().ref_set_num_threads(n)
