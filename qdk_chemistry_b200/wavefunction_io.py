"""Text wavefunction files of MACIS / pymacis (external/macis/include/macis/wavefunction_io.hpp:22-84).

    <nstates> <norb> <nalpha> <nbeta>
    <coefficient, %30.16e> <canonical string, one of 0 u d 2 per orbital>

`to_canonical_string` / `from_canonical_string` (sd_operations.hpp:478-525): orbital i is '2' when
both spin strings hold it, 'u' alpha only, 'd' beta only, '0' empty. Determinants travel here as the
separate alpha / beta occupation words the C ABI uses (bit p = orbital p).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def to_canonical_string(alpha: int, beta: int, norb: int) -> str:
    out = []
    for i in range(norb):
        a, b = (alpha >> i) & 1, (beta >> i) & 1
        out.append("2" if a and b else "u" if a else "d" if b else "0")
    return "".join(out)


def from_canonical_string(s: str) -> Tuple[int, int]:
    alpha = beta = 0
    for i, ch in enumerate(s[:64]):
        if ch == "2":
            alpha |= 1 << i
            beta |= 1 << i
        elif ch == "u":
            alpha |= 1 << i
        elif ch == "d":
            beta |= 1 << i
    return alpha, beta


def read_wavefunction(path: str):
    """-> (alpha uint64[n], beta uint64[n], coeffs float64[n], (nstates, norb, nalpha, nbeta)).
    Like the reference, the header counts are informational: every following line is read."""
    with open(path) as fh:
        header = fh.readline().split()
        meta = tuple(int(x) for x in header[:4])
        al, be, co = [], [], []
        for line in fh:
            f = line.split()
            if len(f) < 2:
                continue
            a, b = from_canonical_string(f[1])
            al.append(a)
            be.append(b)
            co.append(float(f[0]))
    return np.array(al, dtype=np.uint64), np.array(be, dtype=np.uint64), np.array(co, dtype=np.float64), meta


def write_wavefunction(path: str, norb: int, alpha, beta, coeffs) -> None:
    alpha = np.asarray(alpha, dtype=np.uint64)
    beta = np.asarray(beta, dtype=np.uint64)
    coeffs = np.asarray(coeffs, dtype=np.float64)
    if not (alpha.size == beta.size == coeffs.size):
        raise RuntimeError("Invalid Wave Function Dimensions")
    if alpha.size == 0:
        return  # the reference writes nothing for an empty state list
    na, nb = bin(int(alpha[0])).count("1"), bin(int(beta[0])).count("1")
    with open(path, "w") as fh:
        fh.write(f"{alpha.size} {norb} {na} {nb}\n")
        for a, b, c in zip(alpha, beta, coeffs):
            fh.write(f"{c:30.16e} {to_canonical_string(int(a), int(b), norb)} \n")
