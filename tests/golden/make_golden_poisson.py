"""Generates tests/golden/ref_poisson.npz: complex spectra before and after the reference's own k-space gradient step
(the lambda "Gradient FFTPeriodicPoissonSolver", src/PoissonSolvers/FFTPeriodicPoissonSolver.hpp:115-150, cut out at
build time and compiled unchanged: oracle/ref_shim/refshim_poisson.cpp).  Run here (needs /root/reference):
    python tests/golden/make_golden_poisson.py"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import refshim  # noqa: E402

CASES = [((8, 6, 10), (0.0, 0.0, 0.0), (4 * math.pi / 8, 4 * math.pi / 6, 4 * math.pi / 10)),
         ((16, 16, 16), (0.5, -1.0, 2.0), (0.1, 0.2, 0.3)),
         ((7, 5, 9), (0.0, 0.0, 0.0), (20.0 / 7, 20.0 / 5, 20.0 / 9))]


def main():
    rng = np.random.default_rng(20261022)
    out = {}
    for i, (ng, origin, h) in enumerate(CASES):
        spec = rng.normal(size=ng[::-1]) + 1j * rng.normal(size=ng[::-1])
        out[f"spec_{i}"] = spec
        for gd in range(3):
            out[f"grad_{i}_{gd}"] = refshim.poisson_grad_kspace(spec, origin, h, gd)
    path = os.path.join(os.path.dirname(__file__), "ref_poisson.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
