"""Thin object wrappers over the b2ci C ABI (numpy in / numpy out, device-resident handles).

Everything here forwards to libb2ci.so; no arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import AsciSearchOpts, B2ciError, check, lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def words_per_det_for_norb(norb: int) -> int:
    """dispatch_by_norb (cpp/src/qdk/chemistry/algorithms/microsoft/macis_base.hpp:80-100)
    restricted to the widths this build instantiates: wfn_t<64> and wfn_t<128>."""
    if norb < 32:
        return 1
    if norb < 64:
        return 2
    raise ValueError("active spaces with 64 or more orbitals need wfn_t<256+>, not built")


class Context:
    """Device + stream + integrals (+ optional NCCL communicator)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        h = C.c_void_p()
        check(lib().b2ci_ctx_create(device, C.c_void_p(stream or 0), C.byref(h)))
        self.h = h
        self.device = device
        self.norb = 0

    def close(self):
        if getattr(self, "h", None):
            lib().b2ci_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing
    def synchronize(self):
        check(lib().b2ci_ctx_synchronize(self.h))

    def trim(self):
        """Return the library's cached device memory (stream-ordered pool) to the driver."""
        check(lib().b2ci_ctx_trim(self.h))

    @property
    def launch_count(self) -> int:
        return lib().b2ci_ctx_launch_count(self.h)

    def timer_ms(self, name: str) -> float:
        return lib().b2ci_timer_ms(self.h, name.encode())

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_char * 128)()
        check(lib().b2ci_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        assert len(unique_id) == 128
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        check(lib().b2ci_comm_init(self.h, buf, rank, nranks))

    # -- integrals
    def upload_integrals(self, norb: int, T: np.ndarray, V: np.ndarray):
        T = np.ascontiguousarray(T, dtype=np.float64)
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1)
        if T.size != norb * norb or V.size != norb ** 4:
            raise ValueError("T must have norb^2 and V norb^4 entries")
        check(lib().b2ci_integrals_upload(self.h, norb, _p(T), _p(V)))
        self.norb = norb

    def rotate_integrals(self, Cm: np.ndarray):
        """T <- C^T T C and the four-index analogue for V on the device (transform.cxx:22-96), then
        the intermediates again. Returns the rotated (T, V) as flat column-major arrays."""
        n = self.norb
        Cf = np.ascontiguousarray(np.asarray(Cm, dtype=np.float64).reshape(n, n).T).reshape(-1)  # column-major
        T, V = np.empty(n * n), np.empty(n ** 4)
        check(lib().b2ci_integrals_rotate(self.h, _p(Cf), _p(T), _p(V)))
        return T, V

    def download_intermediates(self):
        n = self.norb
        G, Vr, G2, V2 = np.empty(n ** 3), np.empty(n ** 3), np.empty(n * n), np.empty(n * n)
        check(lib().b2ci_integrals_download(self.h, _p(G), _p(Vr), _p(G2), _p(V2)))
        return G, Vr, G2, V2

    # -- determinants
    def upload_dets(self, words: np.ndarray, words_per_det: int = 1) -> "DetList":
        w = np.ascontiguousarray(words, dtype=np.uint64)
        h = C.c_void_p()
        check(lib().b2ci_dets_upload(self.h, _p(w), words_per_det, w.size // words_per_det,
                                     C.byref(h)))
        return DetList(self, h)

    def balanced_partition(self, dets: "DetList", nparts: int, nsamples: int = 1024) -> np.ndarray:
        """Row cuts (nparts + 1 offsets) that balance the estimated connections of a selected-CI list."""
        off = np.zeros(nparts + 1, dtype=np.int64)
        check(lib().b2ci_dets_balanced_partition(self.h, dets.h, nparts, nsamples, _p(off)))
        return off

    def generate_fci(self, norb: int, nalpha: int, nbeta: int) -> "DetList":
        h = C.c_void_p()
        check(lib().b2ci_dets_generate_fci(self.h, norb, nalpha, nbeta, C.byref(h)))
        return DetList(self, h)

    # -- matrices
    def hbuild(self, dets: "DetList", h_thresh: float, rows: Optional[Tuple[int, int]] = None
               ) -> "CsrMatrix":
        r0, r1 = rows if rows is not None else (0, len(dets))
        h = C.c_void_p()
        check(lib().b2ci_hbuild_csr(self.h, dets.h, r0, r1, h_thresh, C.byref(h)))
        return CsrMatrix(self, h)

    GENERATORS = {"": 0, "sorted_double_loop": 0, "residue_arrays": 1, "dynamic_bit_masking": 2}

    def set_hamiltonian_generator(self, name: str) -> None:
        """QDK's hamiltonian_build_algorithm (macis_asci.hpp:174-179): pattern / threshold rules."""
        check(lib().b2ci_set_hamiltonian_generator(self.h, self.GENERATORS[name]))

    def hbuild_patched(self, old_dets: "DetList", old_H: "CsrMatrix", new_dets: "DetList", h_thresh: float,
                       min_overlap: float = 0.3):
        """build_patched_operator (incremental_h_build.hpp:219-356) merged into one CSR of
        ``new_dets``. Returns (matrix or None when the overlap is below ``min_overlap``, n_kept)."""
        h = C.c_void_p()
        nk = C.c_int64(0)
        check(lib().b2ci_hbuild_csr_patched(self.h, old_dets.h, old_H.h, new_dets.h, h_thresh, min_overlap,
                                            C.byref(h), C.byref(nk)))
        return (CsrMatrix(self, h) if h.value else None), nk.value

    def upload_csr(self, rowptr, colind, nzval) -> "CsrMatrix":
        rp = np.ascontiguousarray(rowptr, dtype=np.int64)
        ci = np.ascontiguousarray(colind, dtype=np.int64)
        nz = np.ascontiguousarray(nzval, dtype=np.float64)
        h = C.c_void_p()
        check(lib().b2ci_csr_upload(self.h, rp.size - 1, ci.size, _p(rp), _p(ci), _p(nz),
                                    C.byref(h)))
        return CsrMatrix(self, h)

    # -- ASCI search
    def asci_search(self, core_words: np.ndarray, core_coeffs: np.ndarray, E0: float,
                    ndets_max: int, h_el_tol: float = 1e-8, rv_prune_tol: float = 1e-8,
                    just_singles: bool = False, words_per_det: int = 1, sort_output: bool = False):
        cw = np.ascontiguousarray(core_words, dtype=np.uint64)
        cc = np.ascontiguousarray(core_coeffs, dtype=np.float64)
        nc = cw.size // words_per_det
        o = AsciSearchOpts(int(ndets_max), h_el_tol, rv_prune_tol, int(just_singles), int(sort_output))
        cap = int(ndets_max) + nc + 4096
        stats = np.zeros(8)
        n_out = C.c_int64(0)
        while True:
            out = np.empty(cap * words_per_det, dtype=np.uint64)
            rc = lib().b2ci_asci_search(self.h, C.byref(o), _p(cw), words_per_det, _p(cc), nc, E0,
                                        _p(out), cap, C.byref(n_out), _p(stats))
            if rc == 4 and n_out.value > cap:  # capacity (ties at the cut can exceed ndets_max)
                cap = n_out.value
                continue
            check(rc)
            return out[: n_out.value * words_per_det].copy(), stats

    def asci_candidates(self, core_words, core_coeffs, E0, h_el_tol=1e-8, just_singles=False,
                        words_per_det: int = 1):
        cw = np.ascontiguousarray(core_words, dtype=np.uint64)
        cc = np.ascontiguousarray(core_coeffs, dtype=np.float64)
        nc = cw.size // words_per_det
        o = AsciSearchOpts(0, h_el_tol, 0.0, int(just_singles), 0)
        n_out = C.c_int64(0)
        check(lib().b2ci_asci_candidates(self.h, C.byref(o), _p(cw), words_per_det, _p(cc), nc, E0,
                                         None, None, None, C.byref(n_out)))
        n = n_out.value
        words = np.empty(n * words_per_det, dtype=np.uint64)
        cm, hd = np.empty(n), np.empty(n)
        check(lib().b2ci_asci_candidates(self.h, C.byref(o), _p(cw), words_per_det, _p(cc), nc, E0,
                                         _p(words), _p(cm), _p(hd), C.byref(n_out)))
        return words, cm, hd


    def form_rdms(self, dets: "DetList", C: np.ndarray, spin_dep: bool = False, one: bool = True,
                  two: bool = True):
        """form_rdms -> (ordm, trdm); form_rdms_spin_dep -> (aa, bb, aaaa, bbbb, aabb); Fortran-ordered
        (n, n) / (n, n, n, n) arrays, None where not requested. n = orbitals of the uploaded integrals."""
        c = np.ascontiguousarray(C, dtype=np.float64)
        if c.size != len(dets):
            raise ValueError("one coefficient per determinant")
        n = self.norb
        mk1 = lambda: np.zeros(n * n) if one else None
        mk2 = lambda: np.zeros(n ** 4) if two else None
        sh = lambda x, k: None if x is None else x.reshape((n,) * k, order="F")
        if spin_dep:
            o1, o2, t1, t2, t3 = mk1(), mk1(), mk2(), mk2(), mk2()
            check(lib().b2ci_form_rdms_spin_dep(self.h, dets.h, _p(c), _p(o1), _p(o2), _p(t1), _p(t2), _p(t3)))
            return sh(o1, 2), sh(o2, 2), sh(t1, 4), sh(t2, 4), sh(t3, 4)
        o1, t1 = mk1(), mk2()
        check(lib().b2ci_form_rdms(self.h, dets.h, _p(c), _p(o1), _p(t1)))
        return sh(o1, 2), sh(t1, 4)

    def form_entropies(self, dets: "DetList", C: np.ndarray, two_orbital: bool = True,
                       mutual_information: bool = True):
        """form_entropies -> (s1[n], s2[n, n] or None, mutual information [n, n] or None)"""
        c = np.ascontiguousarray(C, dtype=np.float64)
        if c.size != len(dets):
            raise ValueError("one coefficient per determinant")
        n = self.norb
        s1 = np.zeros(n)
        s2 = np.zeros(n * n) if two_orbital else None
        mi = np.zeros(n * n) if mutual_information else None
        check(lib().b2ci_form_entropies(self.h, dets.h, _p(c), _p(s1), _p(s2), _p(mi)))
        sh = lambda a: None if a is None else a.reshape(n, n, order="F")
        return s1, sh(s2), sh(mi)

    def entropy_intermediates(self, dets: "DetList", C: np.ndarray, need_s2: bool = True) -> np.ndarray:
        c = np.ascontiguousarray(C, dtype=np.float64)
        out = np.zeros(lib().b2ci_entropy_intermediate_count(self.norb, int(need_s2)))
        check(lib().b2ci_entropy_intermediates(self.h, dets.h, _p(c), int(need_s2), _p(out)))
        return out

    def asci_pt2(self, det_words, coeffs, E_asci: float, pt2_tol: float = 1e-16,
                 words_per_det: int = 1):
        """macis::asci_pt2_constraint (asci/pt2.hpp): (EPT2, number of external determinants).
        `det_words` must be spin-sorted, `coeffs` in the same order."""
        dw = np.ascontiguousarray(det_words, dtype=np.uint64)
        cc = np.ascontiguousarray(coeffs, dtype=np.float64)
        e, npt2 = C.c_double(0.0), C.c_int64(0)
        check(lib().b2ci_asci_pt2(self.h, _p(dw), words_per_det, _p(cc), dw.size // words_per_det,
                                  E_asci, pt2_tol, C.byref(e), C.byref(npt2)))
        return e.value, npt2.value


def host_entropies_from_intermediates(norb: int, intermediates: np.ndarray, need_s2: bool = True):
    """Host assembly of (s1, s2, mutual information) from the flat intermediates (no GPU)."""
    I = np.ascontiguousarray(intermediates, dtype=np.float64)
    n = int(norb)
    s1, s2, mi = np.zeros(n), (np.zeros(n * n) if need_s2 else None), (np.zeros(n * n) if need_s2 else None)
    check(lib().b2ci_host_entropies_from_intermediates(n, int(need_s2), _p(I), _p(s1), _p(s2), _p(mi)))
    sh = lambda a: None if a is None else a.reshape(n, n, order="F")
    return s1, sh(s2), sh(mi)


class DetList:
    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    def __len__(self) -> int:
        n = C.c_int64(0)
        check(lib().b2ci_dets_size(self.h, C.byref(n)))
        return n.value

    def download(self, words_per_det: int = 1) -> np.ndarray:
        out = np.empty(len(self) * words_per_det, dtype=np.uint64)
        check(lib().b2ci_dets_download(self.ctx.h, self.h, _p(out), words_per_det))
        return out

    def free(self):
        if self.h and self.ctx.h:
            lib().b2ci_dets_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CsrMatrix:
    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle

    def info(self):
        a, b, c, d = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
        check(lib().b2ci_csr_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    @property
    def nrows(self):
        return self.info()[0]

    @property
    def ncols(self):
        return self.info()[1]

    @property
    def nnz(self):
        return self.info()[2]

    @property
    def row_begin(self):
        return self.info()[3]

    def download(self):
        nrows, _, nnz, _ = self.info()
        rp = np.empty(nrows + 1, dtype=np.int64)
        ci = np.empty(nnz, dtype=np.int64)
        nz = np.empty(nnz, dtype=np.float64)
        check(lib().b2ci_csr_download(self.ctx.h, self.h, _p(rp), _p(ci), _p(nz)))
        return rp, ci, nz

    def download_rowptr(self):
        rp = np.empty(self.nrows + 1, dtype=np.int64)
        check(lib().b2ci_csr_download(self.ctx.h, self.h, _p(rp), None, None))
        return rp

    def device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().b2ci_csr_device_ptrs(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def spmv(self, x: np.ndarray) -> np.ndarray:
        """sigma with HOST buffers (copies inside the call)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != self.ncols:
            raise ValueError("x must have ncols entries")
        y = np.empty(self.nrows, dtype=np.float64)
        check(lib().b2ci_spmv_host(self.ctx.h, self.h, _p(x), _p(y)))
        return y

    def spmv_device(self, x_ptr: int, y_ptr: int):
        """sigma on DEVICE pointers (e.g. torch tensors' data_ptr())."""
        check(lib().b2ci_spmv(self.ctx.h, self.h, C.c_void_p(x_ptr), C.c_void_p(y_ptr)))

    def sigma_sharded(self, x_local_ptr: int, x_full_ptr: int, y_local_ptr: int):
        """all-gather of the row-sharded trial vector + local SpMV (DEVICE pointers)."""
        check(lib().b2ci_sigma_sharded(self.ctx.h, self.h, C.c_void_p(x_local_ptr),
                                       C.c_void_p(x_full_ptr), C.c_void_p(y_local_ptr)))

    def set_row_partition(self, row_offsets):
        """Row offsets of all ranks (nranks + 1 entries): spares the first sharded sigma /
        Davidson call on this block its host-synchronising exchange of block sizes."""
        off = np.ascontiguousarray(row_offsets, dtype=np.int64)
        check(lib().b2ci_csr_set_row_partition(self.ctx.h, self.h, _p(off), off.size - 1))

    def diagonal(self) -> np.ndarray:
        d = np.empty(self.nrows, dtype=np.float64)
        check(lib().b2ci_csr_diagonal(self.ctx.h, self.h, _p(d)))
        return d

    def davidson(self, max_m: int, tol: float, x0: Optional[np.ndarray] = None,
                 guess_policy: bool = True):
        """Returns (E, X, niter, trace). Raises B2ciError("Davidson Did Not Converge!")."""
        n = self.ncols
        X = np.zeros(n) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        if X.size != n:
            raise ValueError("guess must have ncols entries")
        if not guess_policy and x0 is None:
            raise ValueError("Davidson: No Guess Provided")
        niter, eig = C.c_int64(0), C.c_double(0.0)
        trace = np.zeros(2 * (max_m + 1))
        check(lib().b2ci_davidson(self.ctx.h, self.h, max_m, tol, _p(X), 1 if guess_policy else 0,
                                  C.byref(niter), C.byref(eig), _p(trace)))
        return eig.value, X, niter.value, trace.reshape(-1, 2)

    def dense_ground_state(self):
        """Lowest eigenpair by dense diagonalisation (the adapters' small-space branch)."""
        e = C.c_double(0.0)
        x = np.empty(self.nrows, dtype=np.float64)
        check(lib().b2ci_dense_ground_state(self.ctx.h, self.h, C.byref(e), _p(x)))
        return e.value, x

    def free(self):
        if self.h and self.ctx.h:
            lib().b2ci_csr_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def host_matrix_element(norb, T, V, bra_a, bra_b, ket_a, ket_b) -> float:
    T = np.ascontiguousarray(T, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1)
    return lib().b2ci_host_matrix_element(norb, _p(T), _p(V), int(bra_a), int(bra_b), int(ket_a),
                                          int(ket_b))


def host_sym_eig_lower(A: np.ndarray):
    n = A.shape[0]
    M = np.array(A, dtype=np.float64, order="F", copy=True)
    W = np.empty(n)
    check(lib().b2ci_host_sym_eig_lower(n, _p(M), n, _p(W)))
    return W, M


def host_sym_eig_lowest(A: np.ndarray):
    n = A.shape[0]
    M = np.array(A, dtype=np.float64, order="F", copy=True)
    lam = C.c_double(0.0)
    v = np.empty(n)
    check(lib().b2ci_host_sym_eig_lowest(n, _p(M), n, C.byref(lam), _p(v)))
    return lam.value, v
