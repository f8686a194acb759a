#!/usr/bin/env python
"""Piecewise parity of a LARGE selected-CI wavefunction against the unmodified reference (SURVEY.md section 8(d).5:
"CPU oracle for this size may be impractical end-to-end; compare per-iteration pieces -- search on a prefix, H rows on a
sampled row block"). Input: the .npz scripts/asci_scale.py writes with save=PATH (determinant words, coefficients).

  1. H rows: for a few row blocks spread over the list (first rows, last rows, a block in the middle of the longest
     alpha run ...), b2ci_hbuild_csr(row_begin, row_end) on one GPU against the reference's
     make_csr_hamiltonian_block(bra = those determinants, ket = the WHOLE list): rowptr, colind and nzval compared exactly.
  2. search: the NCORE largest-|c| determinants as core, ndets_max = NKEEP: b2ci_asci_search against the reference's
     asci_search -- the selected sets must be identical.

    python scripts/verify_scale.py /tmp/wfn.npz [rows_per_block=64] [nblocks=6] [ncore=2000] [nkeep=100000]
Needs oracle/_ref (travels to the GPU box as a built .so). One JSON line on stdout."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402  (test infrastructure: this script is a checker, not the product)
from qdk_chemistry_b200 import device, workloads as W  # noqa: E402

EPS = float(np.finfo(np.float64).eps)


def main():
    z = np.load(sys.argv[1])
    opt = dict(rows_per_block=64, nblocks=6, ncore=2000, nkeep=100000)
    for a in sys.argv[2:]:
        k, v = a.split("=")
        opt[k] = int(float(v))
    sp = W.config(str(z["workload"]))
    words2 = np.ascontiguousarray(z["words"], dtype=np.uint64)          # (n, 2): alpha, beta
    coeffs = np.asarray(z["coeffs"], dtype=np.float64)
    n = len(coeffs)
    assert sp.norb < 32
    words = (words2[:, 0] & np.uint64(0xFFFFFFFF)) | (words2[:, 1] << np.uint64(32))   # wfn_t<64>
    try:
        ref.set_num_threads(len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    hg = ref.HamGen(sp.norb, sp.T, sp.V)
    ctx = device.Context(0)
    ctx.upload_integrals(sp.norb, sp.T, sp.V)
    dets = ctx.upload_dets(words, 1)
    out = {"workload": str(z["workload"]), "ndets": int(n), "reference_threads": ref.num_threads()}
    # ---- 1. sampled row blocks
    rpb, nblk = opt["rows_per_block"], opt["nblocks"]
    starts = sorted(set(int(x) for x in np.linspace(0, n - rpb, nblk)))
    blocks, all_equal, nnz_checked = [], True, 0
    t_ref = t_gpu = 0.0
    for r0 in starts:
        r1 = r0 + rpb
        t0 = time.perf_counter()
        B = ctx.hbuild(dets, EPS, (r0, r1))
        rp, ci, nz = B.download()
        B.free()
        t_gpu += time.perf_counter() - t0
        Hr, sec = hg.hbuild(words[r0:r1], EPS, kets=words)
        t_ref += sec
        rrp, rci, rnz = Hr.arrays()
        ok = bool(np.array_equal(rp, rrp) and np.array_equal(ci, rci) and np.array_equal(nz, rnz))
        all_equal = all_equal and ok
        nnz_checked += int(rrp[-1])
        blocks.append({"row_begin": r0, "row_end": r1, "nnz": int(rrp[-1]), "equal": ok})
    out["h_rows"] = {"blocks": blocks, "all_equal": all_equal, "nnz_checked": nnz_checked,
                     "reference_seconds": t_ref, "gpu_seconds_incl_download": t_gpu,
                     "what": "b2ci_hbuild_csr row blocks vs make_csr_hamiltonian_block(bra block, all kets): exact"}
    # ---- 2. search from the largest-|c| prefix
    ncore, nkeep = min(opt["ncore"], n), opt["nkeep"]
    top = np.argsort(-np.abs(coeffs), kind="stable")[:ncore]
    core = words[top]
    cc = coeffs[top]
    order = np.lexsort((core >> np.uint64(32), core & np.uint64(0xFFFFFFFF)))   # spin_comparator: alpha-major, then beta
    core, cc = core[order], cc[order]
    E0 = float(z["E"])
    t0 = time.perf_counter()
    sel, stats = ctx.asci_search(core, cc, E0, nkeep)
    t_g = time.perf_counter() - t0
    t0 = time.perf_counter()
    rsel = hg.asci_search(ref.AsciOpts(), nkeep, core, cc, E0)
    t_r = time.perf_counter() - t0
    g, r = set(sel.tolist()), set(rsel.tolist())
    out["search"] = {"ncore": int(ncore), "ndets_max": int(nkeep), "selected_gpu": len(g), "selected_reference": len(r),
                     "identical": bool(g == r), "only_gpu": len(g - r), "only_reference": len(r - g),
                     "contributions": float(stats[0]), "gpu_seconds": t_g, "reference_seconds": t_r,
                     "what": "b2ci_asci_search vs macis::asci_search from the same core determinants"}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
