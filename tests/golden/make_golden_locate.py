"""Generates tests/golden/ref_locate.npz: destination ranks from the reference's own ownership code (the destRankOf lambda
of ParticleSpatialLayout::locateParticlesPacked and positionInRegion / positionInRegionInclusive, src/Particle/
ParticleSpatialLayout.hpp:316-330, 372-395, cut out at build time: oracle/ref_shim/refshim_locate.cpp) for particles
including every kind of boundary case.  Run here (needs /root/reference):  python tests/golden/make_golden_locate.py"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402
from oracle import refshim  # noqa: E402

CASES = [((16, 16, 16), 2), ((16, 16, 16), 4), ((24, 16, 16), 8), ((17, 9, 33), 6), ((32, 20, 12), 3)]
ORIGIN, H = (0.0, 0.0, 0.0), (4 * math.pi / 16,) * 3


def particles(ng, regs, rng, n=6000):
    Lg = [ng[d] * H[d] for d in range(3)]
    R = [rng.uniform(0, Lg[d], n) for d in range(3)]
    k = 0
    for r in range(len(regs)):     # exactly on every face of every region, and one ulp to either side
        for d in range(3):
            for v in (regs[r][d], regs[r][3 + d], np.nextafter(regs[r][d], -np.inf), np.nextafter(regs[r][3 + d], np.inf)):
                R[d][k] = v
                k += 1
    for d in range(3):
        R[d][k] = 0.0              # the domain's lower corner: only the inclusive fallback finds an owner
    R[0][k + 1] = -1.0             # outside every region: stays on the asking rank
    return R


def main():
    rng = np.random.default_rng(20261024)
    out = {}
    for ci, (ng, nr) in enumerate(CASES):
        regs = oracle.regions(ng, oracle.partition(ng, nr), ORIGIN, H)
        R = particles(ng, regs, rng)
        out[f"R_{ci}"] = np.stack(R)
        for my in range(nr):
            out[f"dest_{ci}_{my}"] = refshim.dest_rank(regs, my, *R)
    path = os.path.join(os.path.dirname(__file__), "ref_locate.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
