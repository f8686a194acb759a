"""Seeded synthetic active-space integrals + FCIDUMP I/O for the CI hot path.

The five BASELINE.json configs are defined here (shapes and seeds follow
SURVEY.md section 8(d)); both the CUDA path and the oracles consume exactly
these arrays, so parity never depends on how the integrals were made.

Layout handed to the solvers (the reference's own, see
cpp/src/qdk/chemistry/algorithms/microsoft/macis_cas.cpp:58-82 and
external/macis/src/macis/hamiltonian_generator/base.ipp:37-77):

* ``T``: (n, n) float64, symmetric; element (p, q) at ``p + q*n``.
* ``V``: (n**4,) float64, chemists' notation, ``V[p + q*n + r*n*n + s*n**3] = (pq|rs)``
  with the full 8-fold symmetry (so C and Fortran ravel orders coincide).
"""
from __future__ import annotations

import dataclasses
import math
import re
from typing import Dict, Tuple

import numpy as np


@dataclasses.dataclass
class ActiveSpace:
    """Same-shaped stand-in for the slice of ``data::Hamiltonian`` the path reads."""

    name: str
    norb: int
    nalpha: int
    nbeta: int
    T: np.ndarray  # (n, n) float64
    V: np.ndarray  # (n**4,) float64
    core_energy: float = 0.0

    @property
    def fci_dimension(self) -> int:
        return math.comb(self.norb, self.nalpha) * math.comb(self.norb, self.nbeta)


def _sym_outer_sum(n: int, nvec: int, seed: int, scale: float) -> np.ndarray:
    """V(pq|rs) = sum_L B^L_pq B^L_rs, accumulated one L at a time so that every
    element sees the same addition order and the 8-fold symmetry is exact in
    floating point (a Cholesky-shaped, positive semidefinite ERI tensor)."""
    rng = np.random.default_rng(seed)
    V = np.zeros((n, n, n, n))
    for L in range(1, nvec + 1):
        A = rng.normal(0.0, scale / L, size=(n, n))
        B = 0.5 * (A + A.T)
        V += np.multiply.outer(B, B)
    return V


def synthetic_molecular(name: str, norb: int, nalpha: int, nbeta: int, seed_T: int,
                        seed_V: int, eps0: float = -2.0, deps: float = 0.4) -> ActiveSpace:
    """FCIDUMP-shaped dense synthetic integrals (SURVEY.md 8(d) recipes 1, 3, 4, 5)."""
    rng = np.random.default_rng(seed_T)
    S = rng.normal(0.0, 0.1, size=(norb, norb))
    T = np.diag(eps0 + deps * np.arange(norb)) + 0.5 * (S + S.T)
    V = _sym_outer_sum(norb, 3 * norb, seed_V, 0.3)
    return ActiveSpace(name, norb, nalpha, nbeta, np.ascontiguousarray(T),
                       np.ascontiguousarray(V.reshape(-1)), 0.0)


def hubbard_2d(nx: int, ny: int, nalpha: int, nbeta: int, t: float = 1.0, U: float = 4.0,
               V1: float = 1.0, z: float = 1.0, name: str | None = None) -> ActiveSpace:
    """2D extended Hubbard / PPP model on an open nx x ny lattice.

    Follows the reference's PPP builder (cpp/include/qdk/chemistry/utils/
    model_hamiltonians.hpp:197-304): T_ij = -t on bonds, (ii|ii) = U,
    (ii|jj) = V1 on bonds, T_ii -= sum_j V_ij z_j / 1 (two half updates per ordered
    pair), scalar offset sum_{i<j} V_ij z_i z_j returned as the core energy."""
    n = nx * ny
    T = np.zeros((n, n))
    V = np.zeros((n, n, n, n))
    core = 0.0

    def site(ix, iy):
        return ix * ny + iy

    bonds = []
    for ix in range(nx):
        for iy in range(ny):
            if ix + 1 < nx:
                bonds.append((site(ix, iy), site(ix + 1, iy)))
            if iy + 1 < ny:
                bonds.append((site(ix, iy), site(ix, iy + 1)))
    for i in range(n):
        V[i, i, i, i] = U
    for (i, j) in bonds:
        T[i, j] = T[j, i] = -t
        if V1 != 0.0:
            V[i, i, j, j] += V1
            V[j, j, i, i] += V1
            T[i, i] -= V1 * z
            T[j, j] -= V1 * z
            core += V1 * z * z
    return ActiveSpace(name or f"hubbard_{nx}x{ny}", n, nalpha, nbeta, T,
                       np.ascontiguousarray(V.reshape(-1)), core)


# --------------------------------------------------------------------------------------
# FCIDUMP (external/macis/src/macis/fcidump.cxx:340-431 reader semantics: 1-based
# indices, "(pq|rs)" lines fill all 8 permutations, "p q 0 0" one-body symmetric,
# "0 0 0 0" core energy).
# --------------------------------------------------------------------------------------
def read_fcidump(path: str, nalpha: int | None = None, nbeta: int | None = None,
                 name: str | None = None) -> ActiveSpace:
    with open(path, "r") as fh:
        text = fh.read()
    head, _, body = text.partition("&END")
    if not body:
        head, _, body = text.partition("/")
    m = re.search(r"NORB\s*=\s*(\d+)", head)
    if not m:
        raise ValueError(f"{path}: NORB not found in FCIDUMP header")
    n = int(m.group(1))
    nelec = int(re.search(r"NELEC\s*=\s*(\d+)", head).group(1))
    ms2m = re.search(r"MS2\s*=\s*(-?\d+)", head)
    ms2 = int(ms2m.group(1)) if ms2m else 0
    T = np.zeros((n, n))
    V = np.zeros((n, n, n, n))
    core = 0.0
    for line in body.splitlines():
        parts = line.split()
        if len(parts) != 5:
            continue
        try:
            val = float(parts[0])
            p, q, r, s = (int(x) for x in parts[1:])
        except ValueError:
            # integral-last format
            val = float(parts[4])
            p, q, r, s = (int(x) for x in parts[:4])
        if p == 0 and q == 0 and r == 0 and s == 0:
            core = val
        elif r == 0 and s == 0:
            T[p - 1, q - 1] = val
            T[q - 1, p - 1] = val
        else:
            p, q, r, s = p - 1, q - 1, r - 1, s - 1
            for (a, b, c, d) in ((p, q, r, s), (p, q, s, r), (q, p, r, s), (q, p, s, r),
                                 (r, s, p, q), (s, r, p, q), (r, s, q, p), (s, r, q, p)):
                V[a, b, c, d] = val
    na = nalpha if nalpha is not None else (nelec + ms2) // 2
    nb = nbeta if nbeta is not None else (nelec - ms2) // 2
    return ActiveSpace(name or path.rsplit("/", 1)[-1], n, na, nb, T,
                       np.ascontiguousarray(V.reshape(-1)), core)


def write_fcidump(path: str, sp: ActiveSpace, tol: float = 0.0) -> None:
    n = sp.norb
    V = sp.V.reshape(n, n, n, n)
    with open(path, "w") as fh:
        fh.write(f"&FCI NORB={n}, NELEC={sp.nalpha + sp.nbeta}, MS2={sp.nalpha - sp.nbeta},\n")
        fh.write("ORBSYM=" + ",".join("1" for _ in range(n)) + ",\nISYM=1,\n&END\n")
        for p in range(n):
            for q in range(p + 1):
                pq = p * (p + 1) // 2 + q
                for r in range(p + 1):
                    for s in range(r + 1):
                        if r * (r + 1) // 2 + s > pq:
                            continue
                        v = V[p, q, r, s]
                        if abs(v) > tol:
                            fh.write(f"{v:28.20e} {p+1:4d} {q+1:4d} {r+1:4d} {s+1:4d}\n")
        for p in range(n):
            for q in range(p + 1):
                v = sp.T[p, q]
                if abs(v) > tol:
                    fh.write(f"{v:28.20e} {p+1:4d} {q+1:4d}    0    0\n")
        fh.write(f"{sp.core_energy:28.20e}    0    0    0    0\n")


def save_sparse_npz(path: str, sp: ActiveSpace) -> None:
    """Compact fixture format: unique non-zero (pq|rs), p>=q, r>=s, pq>=rs, plus T's lower
    triangle -- i.e. exactly the information content of a FCIDUMP file."""
    n = sp.norb
    V = sp.V.reshape(n, n, n, n)
    vals, idx = [], []
    for p in range(n):
        for q in range(p + 1):
            pq = p * (p + 1) // 2 + q
            for r in range(p + 1):
                for s in range(r + 1):
                    if r * (r + 1) // 2 + s > pq:
                        continue
                    v = V[p, q, r, s]
                    if v != 0.0:
                        vals.append(v)
                        idx.append((p, q, r, s))
    np.savez_compressed(path, norb=n, nalpha=sp.nalpha, nbeta=sp.nbeta, core=sp.core_energy,
                        name=sp.name, T=sp.T, vals=np.array(vals),
                        idx=np.array(idx, dtype=np.int8))


def load_sparse_npz(path: str) -> ActiveSpace:
    z = np.load(path)
    n = int(z["norb"])
    V = np.zeros((n, n, n, n))
    idx = z["idx"].astype(np.int64)
    vals = z["vals"]
    p, q, r, s = idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]
    for (a, b, c, d) in ((p, q, r, s), (p, q, s, r), (q, p, r, s), (q, p, s, r),
                         (r, s, p, q), (s, r, p, q), (r, s, q, p), (s, r, q, p)):
        V[a, b, c, d] = vals
    return ActiveSpace(str(z["name"]), n, int(z["nalpha"]), int(z["nbeta"]),
                       np.ascontiguousarray(z["T"]), np.ascontiguousarray(V.reshape(-1)),
                       float(z["core"]))


def save_npz(path: str, sp: ActiveSpace) -> None:
    np.savez_compressed(path, norb=sp.norb, nalpha=sp.nalpha, nbeta=sp.nbeta, T=sp.T,
                        V=sp.V, core=sp.core_energy, name=sp.name)


def load_npz(path: str) -> ActiveSpace:
    z = np.load(path)
    return ActiveSpace(str(z["name"]), int(z["norb"]), int(z["nalpha"]), int(z["nbeta"]),
                       np.ascontiguousarray(z["T"]), np.ascontiguousarray(z["V"]),
                       float(z["core"]))


# --------------------------------------------------------------------------------------
# The BASELINE.json configs
# --------------------------------------------------------------------------------------
def config(name: str) -> ActiveSpace:
    """Named workloads. ``n2_cas10`` .. ``cr2_asci30`` are BASELINE.json configs[0..4];
    the rest are small parity cases."""
    if name == "n2_cas10":  # configs[0]: N2-like CAS(10e,10o), 63,504 dets
        return synthetic_molecular(name, 10, 5, 5, 1001, 1002)
    if name == "hubbard_4x3":  # configs[1]: 2D extended Hubbard 4x3, 6a6b, 853,776 dets
        return hubbard_2d(4, 3, 6, 6, t=1.0, U=4.0, V1=1.0, name=name)
    if name == "cr2_cas12":  # configs[2]: Cr2-like CAS(12e,12o) dense, 853,776 dets
        return synthetic_molecular(name, 12, 6, 6, 3001, 3002)
    if name == "n2_asci26":  # configs[3]: N2-like ASCI(14e,26o)
        return synthetic_molecular(name, 26, 7, 7, 4001, 4002, deps=0.3)
    if name == "cr2_asci30":  # configs[4]: Cr2-like ASCI(24e,30o)
        return synthetic_molecular(name, 30, 12, 12, 5001, 5002, deps=0.3)
    # small parity cases
    if name == "tiny_cas6":
        return synthetic_molecular(name, 6, 3, 3, 11, 12)
    if name == "small_cas8":
        return synthetic_molecular(name, 8, 4, 3, 21, 22)
    if name == "wide36":  # 32 < norb < 64: wfn_t<128> determinants, 128-bit ASCI keys
        return synthetic_molecular(name, 36, 3, 3, 6001, 6002, deps=0.05)
    if name == "hubbard_3x2":
        return hubbard_2d(3, 2, 3, 3, name=name)
    if name == "hubbard_4x2":
        return hubbard_2d(4, 2, 4, 4, name=name)
    raise KeyError(name)


CONFIG_NAMES = ["n2_cas10", "hubbard_4x3", "cr2_cas12", "n2_asci26", "cr2_asci30"]
