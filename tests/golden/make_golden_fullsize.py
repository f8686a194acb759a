#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- golden fingerprints of the BENCHMARKED matrices, made with the unmodified reference.

BASELINE configs[1] (2D extended Hubbard 4x3) and configs[2] (Cr2-like CAS(12e,12o)), 853,776 determinants each:
the compiled reference (oracle/_ref) runs its own path end to end --

    generate_hilbert_space -> make_csr_hamiltonian<int64> (ONE symmetric build, sorted_double_loop.hpp:86-451,
    H_thresh = DBL_EPSILON) -> extract_diagonal_elements + davidson (davidson.hpp:259-372, tol 1e-8)

-- and this script records, in tests/golden/fullsize_meta.json,

  * n, nnz, the converged energy and the iteration count,
  * sha256 of the full row pointer (int64), and per block of 33 alpha runs (30,492 rows, 28 blocks) sha256 of the
    block's column indices (int64, little endian) and of its matrix elements (IEEE doubles) -- the same check
    external/macis/tests/csr_hamiltonian.cxx:76-99 makes on water, at the size bench.py quotes its numbers on,
  * wall times and the thread count (evidence only).

Needs /root/reference (build container only; ~26 GB of host memory for configs[2]). Usage:
    python tests/golden/make_golden_fullsize.py [hubbard_4x3] [cr2_cas12]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from qdk_chemistry_b200 import workloads as W  # noqa: E402

EPS = float(np.finfo(np.float64).eps)
RUNS_PER_BLOCK = 33
OUT = os.path.join(ROOT, "tests", "golden", "fullsize_meta.json")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def one(name: str) -> dict:
    sp = W.config(name)
    words = ref.generate_hilbert_space(sp.norb, sp.nalpha, sp.nbeta)
    n = words.size
    nbeta_str = int(np.count_nonzero((words & np.uint64(0xFFFFFFFF)) == (words[0] & np.uint64(0xFFFFFFFF))))
    nruns = n // nbeta_str
    hg = ref.HamGen(sp.norb, sp.T, sp.V)
    t0 = time.perf_counter()
    H, sec = hg.hbuild(words, EPS)          # symmetric: bra == ket, upper triangle + mirror
    print(f"{name}: reference build {sec:.1f} s, nnz {H.nnz}, {ref.num_threads()} threads", flush=True)
    rp = H.rowptr()
    blocks = []
    for r in range(0, nruns, RUNS_PER_BLOCK):
        r0, r1 = r * nbeta_str, min(nruns, r + RUNS_PER_BLOCK) * nbeta_str
        _, ci, nz = H.rows(r0, r1)
        blocks.append({"row_begin": r0, "row_end": r1, "nnz": int(ci.size), "colind_sha256": sha(ci),
                       "nzval_sha256": sha(nz), "colind_sum": int(ci.sum()), "nzval_sum": float(nz.sum())})
    print(f"{name}: fingerprints done ({time.perf_counter() - t0:.0f} s)", flush=True)
    t1 = time.perf_counter()
    E, X, niter = H.davidson(200, 1e-8, guess_policy=False)
    tdav = time.perf_counter() - t1
    y, tsp = H.spmv(X, nrep=3)
    print(f"{name}: E0 = {E!r} after {niter} iterations ({tdav:.0f} s), sigma {tsp * 1e3:.0f} ms", flush=True)
    return {"norb": sp.norb, "nalpha": sp.nalpha, "nbeta": sp.nbeta, "n": int(n), "nnz": int(H.nnz),
            "h_thresh": EPS, "E0_electronic": E, "core_energy": sp.core_energy, "davidson_iterations": int(niter),
            "davidson_tol": 1e-8, "davidson_max_m": 200, "x_norm": float(np.linalg.norm(X)),
            "xHx": float(X @ y), "rowptr_sha256": sha(rp), "runs_per_block": RUNS_PER_BLOCK,
            "rows_per_run": nbeta_str, "blocks": blocks,
            "evidence": {"reference_build_seconds": sec, "reference_davidson_seconds": tdav,
                         "reference_sigma_ms": tsp * 1e3, "threads": ref.num_threads(),
                         "reference_build_nnz_per_s": H.nnz / sec,
                         "note": "full symmetric make_csr_hamiltonian<int64> in the build container"}}


def main():
    names = sys.argv[1:] or ["hubbard_4x3", "cr2_cas12"]
    meta = {}
    if os.path.exists(OUT):
        with open(OUT) as fh:
            meta = json.load(fh)
    for nm in names:
        meta[nm] = one(nm)
        with open(OUT, "w") as fh:
            json.dump(meta, fh, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
