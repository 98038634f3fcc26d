"""Seeded inputs of the golden cases (shared by make_golden.py and the tests).

Inputs are regenerated from numpy's PCG64 stream (stable across numpy versions); the golden file
stores a checksum of every input so that stream drift would be detected rather than silently
compared against stale outputs.
"""
import numpy as np

CASES = [  # (name, m, n, dtype, special)
    ("f64_1x1", 1, 1, "f8", None), ("f64_7x7", 7, 7, "f8", None), ("f64_10x12", 10, 12, "f8", None),
    ("f64_50x50", 50, 50, "f8", None), ("f64_52x50", 52, 50, "f8", None), ("f64_64x64", 64, 64, "f8", None),
    ("f64_130x132", 130, 132, "f8", None), ("f64_300x300", 300, 300, "f8", None),
    ("f64_130_zero_col", 130, 130, "f8", "zero_col_77"), ("f64_64_ties", 64, 64, "f8", "ties"),
    ("f64_200_ties", 200, 200, "f8", "ties"),
    ("f32_9x9", 9, 9, "f4", None), ("f32_50x52", 50, 52, "f4", None), ("f32_130x130", 130, 130, "f4", None),
    ("f32_300x302", 300, 302, "f4", None), ("f32_200_zero_col", 200, 200, "f4", "zero_col_5"),
]


def make_input(index: int) -> np.ndarray:
    name, m, n, dt, special = CASES[index]
    rng = np.random.default_rng([12, index])
    a = np.asfortranarray(rng.random((m, n), dtype=np.dtype(dt).type))
    if special and special.startswith("zero_col_"):
        a[:, int(special.split("_")[-1])] = 0
    if special == "ties":   # small integers: many exact ties in |a_ik|; tie-break = first row
        a = np.asfortranarray(rng.integers(-3, 4, size=(m, n)).astype(dt))
    return a
