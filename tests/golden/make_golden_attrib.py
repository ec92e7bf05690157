"""Generates tests/golden/ref_attrib.npz: fields / attributes produced by the reference's own "ParticleAttrib::scatter" and
"ParticleAttrib::gather" lambda bodies (src/Particle/ParticleAttrib.hpp:167-184, 229-244, cut out at build time and
compiled unchanged on the reference's Interpolation/CIC.h: oracle/ref_shim/refshim_attrib.cpp) -- plain and hash-remapped
sub-range scatter with a per-particle charge, gather with replace and add, on a full box and a sub-box.
Run here (needs /root/reference):  python tests/golden/make_golden_attrib.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402
from oracle import refshim  # noqa: E402

NG, ORIGIN, H = (12, 10, 8), (0.25, -1.0, 3.0), (0.5, 0.125, 1.5)
BOXES = (((0, 0, 0), NG), ((6, 0, 4), (6, 10, 4)))


def main():
    rng = np.random.default_rng(20261023)
    out = {}
    n = 2500
    for bi, (first, nl) in enumerate(BOXES):
        m = oracle.Mesh.make(NG, ORIGIN, H, first=first, nl=nl)
        lo = [ORIGIN[d] + first[d] * H[d] for d in range(3)]
        R = [lo[d] + rng.uniform(0, nl[d] * H[d], n) for d in range(3)]
        for d in range(3):   # lower / upper corner of the box, a cell face, a cell centre
            R[d][0], R[d][1] = lo[d], lo[d] + nl[d] * H[d]
            R[d][2], R[d][3] = lo[d] + 3.0 * H[d], lo[d] + 2.5 * H[d]
        q = rng.normal(size=n)
        hashv = rng.permutation(n).astype(np.int32)[:1800]
        ef = rng.normal(size=m.ext[0] * m.ext[1] * m.ext[2] * 3)
        E0 = [rng.normal(size=n) for _ in range(3)]
        out.update({f"R_{bi}": np.stack(R), f"q_{bi}": q, f"hash_{bi}": hashv, f"ef_{bi}": ef, f"E0_{bi}": np.stack(E0)})
        out[f"scatter_{bi}"] = refshim.attrib_scatter(m, *R, q, oracle.field_zeros(m))
        out[f"scatter_hash_{bi}"] = refshim.attrib_scatter(m, *R, q, oracle.field_zeros(m), begin=100, end=1700, hash=hashv)
        out[f"gather_{bi}"] = np.stack(refshim.attrib_gather(m, *R, ef, [e.copy() for e in E0], add=False))
        out[f"gather_add_{bi}"] = np.stack(refshim.attrib_gather(m, *R, ef, [e.copy() for e in E0], add=True))
    path = os.path.join(os.path.dirname(__file__), "ref_attrib.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
