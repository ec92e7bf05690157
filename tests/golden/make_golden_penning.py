"""Generates tests/golden/ref_penning.npz: momenta after the reference's own "Kick1" / "Kick2" expressions
(demos/alpine/PenningTrapManager.h:256-272, 313-333, cut out at build time and compiled in place:
oracle/ref_shim/gen_snippets.py + refshim_penning.cpp).  Run here (the container that has /root/reference):
    python tests/golden/make_golden_penning.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402
from oracle import refshim  # noqa: E402


def main():
    rng = np.random.default_rng(20261021)
    n, L = 3000, 20.0
    dt = 0.5 * L / 2048
    pp = oracle.penning_params((0.0, 0.0, 0.0), (L, L, L), dt, 5.0)
    R = [rng.uniform(0, L, n) for _ in range(3)]
    P = [rng.normal(size=n) for _ in range(3)]
    E = [rng.normal(size=n) for _ in range(3)]
    out = {"L": np.array([L]), "dt": np.array([dt]), "R": np.stack(R), "P": np.stack(P), "E": np.stack(E)}
    for which in (1, 2):
        out[f"kick{which}"] = np.stack(refshim.penning_kick(which, R, P, E, (0, 0, 0), (L, L, L), pp.V0, pp.alpha, pp.Bext, pp.DrInv))
    path = os.path.join(os.path.dirname(__file__), "ref_penning.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
