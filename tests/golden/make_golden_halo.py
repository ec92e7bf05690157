"""Generates tests/golden/ref_halo.npz from the REAL in-rank periodic halo code of the reference (oracle/_ref/
libippl_refshim_halo.so: Field/HaloCells.h/.hpp applyPeriodicSerialDim + HaloPeriodicFunctor compiled in place from
/root/reference).  Run here (the container that has /root/reference):  python tests/golden/make_golden_halo.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import refshim  # noqa: E402

SHAPES = [(6, 5, 4), (8, 8, 8), (3, 7, 2), (16, 4, 9)]


def main():
    rng = np.random.default_rng(20261019)
    out = {}
    for i, ng in enumerate(SHAPES):
        n = (ng[0] + 2) * (ng[1] + 2) * (ng[2] + 2)
        f = rng.normal(size=n)
        out[f"in_{i}"] = f
        for mode in ("fill", "accumulate"):
            out[f"{mode}_{i}"] = refshim.halo_periodic(f.copy(), ng, mode)
    path = os.path.join(os.path.dirname(__file__), "ref_halo.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
