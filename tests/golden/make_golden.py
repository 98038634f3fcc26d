"""Generates tests/golden/getrf_golden.npz.

The reference tree holds no golden vectors (test/runtests.jl is property-based) and Julia is not
available, so the committed fixtures come from an INDEPENDENT implementation of the same pivot rule:
LAPACK {d,s}getrf via scipy (OpenBLAS 0.3.30).  For each seeded input (tests/golden/cases.py) we
store: a checksum of the input, the LAPACK pivots (1-based), info, diag(U), and the full packed LU
for the small cases.  Run from the repo root:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
from scipy.linalg import lapack

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cases import CASES, make_input  # noqa: E402


def main():
    out = {}
    for idx, (name, m, n, dt, special) in enumerate(CASES):
        a = make_input(idx)
        getrf = lapack.dgetrf if dt == "f8" else lapack.sgetrf
        lu, piv, info = getrf(a)
        out[name + "__checksum"] = np.float64(np.asarray(a, dtype=np.float64).sum())
        out[name + "__ipiv"] = (piv + 1).astype(np.int64)
        out[name + "__info"] = np.int64(info)
        out[name + "__diagu"] = np.diag(lu).copy()
        if max(m, n) <= 64:
            out[name + "__lu"] = np.asfortranarray(lu)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "getrf_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
