"""Generates tests/golden/ref_vectors.npz from the REAL reference headers (oracle/_ref shim, built
from /root/reference).  Run here (the container that has /root/reference):
    python tests/golden/make_golden.py
The fixtures pin the oracle (and through it the CUDA path) on boxes where /root/reference and
oracle/_ref do not exist."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402
from oracle import refshim  # noqa: E402


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    # --- CIC scatter / gather on a small anisotropic mesh, including a sub-box with offset first ---
    ng = (12, 10, 8)
    origin = (0.25, -1.0, 3.0)
    h = (0.5, 0.125, 1.5)
    n = 600
    for tag, first, nl in (("full", (0, 0, 0), ng), ("sub", (6, 0, 4), (6, 10, 4))):
        m = oracle.Mesh.make(ng, origin, h, first=first, nl=nl)
        lo = [origin[d] + first[d] * h[d] for d in range(3)]
        x, y, z = [lo[d] + rng.uniform(0, nl[d] * h[d], n) for d in range(3)]
        # edge cases: exactly on lower/upper corner of the box, cell centres, cell faces
        x[0], y[0], z[0] = lo[0], lo[1], lo[2]
        x[1], y[1], z[1] = [lo[d] + nl[d] * h[d] for d in range(3)]
        x[2], y[2], z[2] = [lo[d] + 2.5 * h[d] for d in range(3)]
        x[3], y[3], z[3] = [lo[d] + 3.0 * h[d] for d in range(3)]
        q = rng.normal(size=n)
        rho = oracle.field_zeros(m)
        refshim.scatter(m, x, y, z, q, rho)
        ef = rng.normal(size=rho.size * 3)
        g = refshim.gather(m, x, y, z, ef)
        out.update({f"cic_{tag}_ng": np.array(ng), f"cic_{tag}_first": np.array(first),
                    f"cic_{tag}_nl": np.array(nl), f"cic_{tag}_origin": np.array(origin),
                    f"cic_{tag}_h": np.array(h), f"cic_{tag}_x": x, f"cic_{tag}_y": y, f"cic_{tag}_z": z,
                    f"cic_{tag}_q": q, f"cic_{tag}_rho": rho, f"cic_{tag}_ef": ef,
                    f"cic_{tag}_gx": g[0], f"cic_{tag}_gy": g[1], f"cic_{tag}_gz": g[2]})
    # --- PeriodicBC ---
    lo = [origin[d] for d in range(3)]
    hi = [origin[d] + ng[d] * h[d] for d in range(3)]
    X = [lo[d] + rng.uniform(-0.95, 1.95, 400) * (hi[d] - lo[d]) for d in range(3)]
    for d in range(3):
        X[d][0], X[d][1] = lo[d], hi[d]
        X[d][2] = np.nextafter(lo[d], -np.inf)
        X[d][3] = np.nextafter(hi[d], np.inf)
    B = [a.copy() for a in X]
    refshim.periodic_bc(B[0], B[1], B[2], lo, hi)
    out.update(bc_lo=np.array(lo), bc_hi=np.array(hi), bc_in=np.stack(X), bc_out=np.stack(B))
    # --- Partitioner + FieldLayout neighbour tables ---
    cases = []
    for ngt in ((16, 16, 16), (128, 128, 128), (32, 20, 12), (17, 9, 33)):
        for nr in (2, 3, 4, 6, 8):
            for per in (1, 0):
                boxes = refshim.partition(ngt, nr)
                for my in range(nr):
                    _, nb = refshim.neighbors(ngt, nr, my, periodic=bool(per))
                    key = f"nb_{ngt[0]}_{ngt[1]}_{ngt[2]}_{nr}_{per}_{my}"
                    out[key] = nb
                out[f"boxes_{ngt[0]}_{ngt[1]}_{ngt[2]}_{nr}"] = boxes
                cases.append((*ngt, nr, per))
    out["layout_cases"] = np.array(cases)
    out["matching"] = np.array([refshim.matching_index(i) for i in range(26)])
    path = os.path.join(os.path.dirname(__file__), "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
