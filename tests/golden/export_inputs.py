"""Exchange files for tests/golden/make_golden_julia.jl (see its header).

    python tests/golden/export_inputs.py DIR          write the inputs of tests/golden/cases.py as raw column-major files
    python tests/golden/export_inputs.py --pack DIR   collect what the Julia script wrote into tests/golden/julia_golden.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from cases import CASES, make_input  # noqa: E402


def export(d):
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "index.txt"), "w") as idx:
        for i, (name, m, n, dt, _) in enumerate(CASES):
            a = make_input(i)
            a.ravel(order="F").tofile(os.path.join(d, name + ".in"))
            idx.write(f"{name} {m} {n} {dt}\n")


def pack(d):
    out = {}
    for i, (name, m, n, dt, _) in enumerate(CASES):
        a = make_input(i)
        out[f"{name}.input_checksum"] = np.array([float(np.sum(a.astype(np.float64) * np.arange(1, a.size + 1).reshape(a.shape, order="F")))])
        for tag in ("serial", "threaded", "nopiv"):
            f = np.fromfile(os.path.join(d, f"{name}.{tag}.factors"), dtype=dt).reshape((m, n), order="F")
            out[f"{name}.{tag}.factors"] = f
            out[f"{name}.{tag}.info"] = np.fromfile(os.path.join(d, f"{name}.{tag}.info"), dtype=np.int64)
            if tag != "nopiv":
                out[f"{name}.{tag}.ipiv"] = np.fromfile(os.path.join(d, f"{name}.{tag}.ipiv"), dtype=np.int64)
    out["version"] = np.array(open(os.path.join(d, "version.txt")).read().strip())
    np.savez_compressed(os.path.join(HERE, "julia_golden.npz"), **out)


if __name__ == "__main__":
    if sys.argv[1] == "--pack":
        pack(sys.argv[2])
    else:
        export(sys.argv[1])
