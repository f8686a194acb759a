"""ctypes binding of libb2ci.so (the C ABI declared in include/b2ci.h).

This is the only place the Python layer touches native code. The library must have been
built in-tree (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C qdk_chemistry_b200/csrc``); there is no fallback of any kind -- a missing library
or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2CI_LIB_PATH") or os.path.join(_HERE, "libb2ci.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "b2ci.h")

_lib = None


class B2ciError(RuntimeError):
    def __init__(self, msg, code=1):
        super().__init__(msg)
        self.code = code


class AsciSearchOpts(C.Structure):
    _fields_ = [("ndets_max", C.c_int64), ("h_el_tol", C.c_double), ("rv_prune_tol", C.c_double),
                ("just_singles", C.c_int32), ("sort_output", C.c_int32)]


NOT_CONVERGED = 3


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2ciError(f"{LIB_PATH} is missing: build the CUDA extension first "
                        "(__graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl, u64 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64
    pp = C.POINTER(vp)
    pi64 = C.POINTER(i64)
    L.b2ci_last_error.restype = C.c_char_p
    L.b2ci_version.restype = C.c_char_p
    L.b2ci_ctx_create.argtypes = [i32, vp, pp]
    L.b2ci_ctx_destroy.argtypes = [vp]
    L.b2ci_ctx_synchronize.argtypes = [vp]
    L.b2ci_ctx_launch_count.restype = i64
    L.b2ci_ctx_launch_count.argtypes = [vp]
    L.b2ci_ctx_trim.argtypes = [vp]
    L.b2ci_comm_unique_id.argtypes = [vp]
    L.b2ci_comm_init.argtypes = [vp, vp, i32, i32]
    L.b2ci_comm_rank.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.b2ci_integrals_upload.argtypes = [vp, i32, vp, vp]
    L.b2ci_integrals_rotate.argtypes = [vp, vp, vp, vp]
    L.b2ci_integrals_download.argtypes = [vp, vp, vp, vp, vp]
    L.b2ci_dets_upload.argtypes = [vp, vp, i32, i64, pp]
    L.b2ci_dets_generate_fci.argtypes = [vp, i32, i32, i32, pp]
    L.b2ci_dets_size.argtypes = [vp, pi64]
    L.b2ci_dets_download.argtypes = [vp, vp, vp, i32]
    L.b2ci_dets_free.argtypes = [vp, vp]
    L.b2ci_hbuild_csr.argtypes = [vp, vp, i64, i64, dbl, pp]
    L.b2ci_set_hamiltonian_generator.argtypes = [vp, i32]
    L.b2ci_hbuild_csr_patched.argtypes = [vp, vp, vp, vp, dbl, dbl, pp, pi64]
    L.b2ci_csr_upload.argtypes = [vp, i64, i64, vp, vp, vp, pp]
    L.b2ci_csr_info.argtypes = [vp, pi64, pi64, pi64, pi64]
    L.b2ci_csr_download.argtypes = [vp, vp, vp, vp, vp]
    L.b2ci_csr_device_ptrs.argtypes = [vp, pp, pp, pp]
    L.b2ci_csr_free.argtypes = [vp, vp]
    L.b2ci_spmv.argtypes = [vp, vp, vp, vp]
    L.b2ci_spmv_host.argtypes = [vp, vp, vp, vp]
    L.b2ci_sigma_sharded.argtypes = [vp, vp, vp, vp, vp]
    L.b2ci_csr_set_row_partition.argtypes = [vp, vp, vp, i32]
    L.b2ci_dets_balanced_partition.argtypes = [vp, vp, i32, i64, vp]
    L.b2ci_csr_diagonal.argtypes = [vp, vp, vp]
    L.b2ci_davidson.argtypes = [vp, vp, i64, dbl, vp, i32, pi64, C.POINTER(dbl), vp]
    L.b2ci_dense_ground_state.argtypes = [vp, vp, C.POINTER(dbl), vp]
    L.b2ci_timer_ms.restype = dbl
    L.b2ci_timer_ms.argtypes = [vp, C.c_char_p]
    L.b2ci_asci_search.argtypes = [vp, vp, vp, i32, vp, i64, dbl, vp, i64, pi64, vp]
    L.b2ci_asci_candidates.argtypes = [vp, vp, vp, i32, vp, i64, dbl, vp, vp, vp, pi64]
    L.b2ci_asci_pt2.argtypes = [vp, vp, i32, vp, i64, dbl, dbl, C.POINTER(dbl), pi64]
    L.b2ci_form_rdms.argtypes = [vp, vp, vp, vp, vp]
    L.b2ci_form_rdms_spin_dep.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.b2ci_form_entropies.argtypes = [vp, vp, vp, vp, vp, vp]
    L.b2ci_entropy_intermediate_count.restype = i64
    L.b2ci_entropy_intermediate_count.argtypes = [i32, i32]
    L.b2ci_entropy_intermediates.argtypes = [vp, vp, vp, i32, vp]
    L.b2ci_host_entropies_from_intermediates.argtypes = [i32, i32, vp, vp, vp, vp]
    L.b2ci_host_matrix_element.restype = dbl
    L.b2ci_host_matrix_element.argtypes = [i32, vp, vp, u64, u64, u64, u64]
    L.b2ci_host_sym_eig_lower.argtypes = [i32, vp, i32, vp]
    L.b2ci_host_sym_eig_lowest.argtypes = [i32, vp, i32, C.POINTER(dbl), vp]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise B2ciError(lib().b2ci_last_error().decode(errors="replace"), rc)


def declared_symbols():
    """Function names declared in include/b2ci.h (used by the CPU-side ABI test)."""
    import re
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2ci_[a-z0-9_]+)\s*\(", text)))
