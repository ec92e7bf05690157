"""Generates tests/golden/ref_orb.npz from the REAL OrthogonalRecursiveBisection of the reference (oracle/_ref/
libippl_refshim_orb.so: Decomposition/OrthogonalRecursiveBisection.h/.hpp compiled in place from /root/reference on
serial stand-ins).  Run here (the container that has /root/reference):  python tests/golden/make_golden_orb.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import refshim  # noqa: E402

GRIDS = [(32, 32, 32), (64, 16, 24), (20, 33, 17)]
RANKS = (1, 2, 3, 4, 5, 8, 16)
MEDIANS = ([1, 1, 1, 1], [5, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 5], [1, 2, 3, 4, 5, 6, 7, 8], [0] * 8, [3, 3, 3],
           [2, 2, 2, 2, 2, 2], [1, 0, 0, 9, 0, 0, 1], [0.5, 0.25, 0.125, 4.0, 0.125])


def weights(ng, kind, seed):
    z, y, x = np.meshgrid(*[np.arange(n) for n in ng[::-1]], indexing="ij")
    if kind == "uniform":
        return np.ones(ng[::-1])
    if kind == "blob":
        c = [0.3 * ng[0], 0.6 * ng[1], 0.45 * ng[2]]
        s = [0.15 * ng[0], 0.05 * ng[1] + 1, 0.2 * ng[2]]
        return np.exp(-((x - c[0]) / s[0]) ** 2 - ((y - c[1]) / s[1]) ** 2 - ((z - c[2]) / s[2]) ** 2) + 1e-6
    return np.random.default_rng(seed).random(ng[::-1])


def main():
    out = {}
    for gi, ng in enumerate(GRIDS):
        for kind in ("uniform", "blob", "random"):
            w = weights(ng, kind, 100 + gi)
            if kind == "random":
                out[f"w_{gi}_random"] = w
            for nr in RANKS:
                boxes, ok = refshim.orb_repartition(ng, nr, w)
                out[f"boxes_{gi}_{kind}_{nr}"] = boxes
                out[f"ok_{gi}_{kind}_{nr}"] = np.array([int(ok)])
    for i, w in enumerate(MEDIANS):
        out[f"median_w_{i}"] = np.asarray(w, dtype=np.float64)
        out[f"median_{i}"] = np.array([refshim.orb_find_median(w)])
    # neighbour tables of the ORB layouts: FieldLayout::updateLayout(domains) -> findNeighbors of the reference
    for gi, ng in enumerate(GRIDS):
        for kind in ("blob", "random"):
            for nr in (2, 3, 4, 5, 8):
                boxes = out[f"boxes_{gi}_{kind}_{nr}"]
                if not out[f"ok_{gi}_{kind}_{nr}"][0] or (boxes[:, 3:] - boxes[:, :3] + 1).min() < 2:
                    continue
                for my in range(nr):
                    out[f"nb_{gi}_{kind}_{nr}_{my}"] = refshim.neighbors_boxes(ng, boxes, my)
    # scatterR: the reference's own particle loop (index truncation, weights, scatterToField), weight 1
    rng = np.random.default_rng(20261020)
    ng, origin, h, n = (12, 10, 8), (0.25, -1.0, 3.0), (0.5, 0.125, 1.5), 4000
    R = [origin[d] + rng.uniform(0, ng[d] * h[d], n) for d in range(3)]
    for d in range(3):   # on the lower / upper corner of the box, on a cell face, on a cell centre
        R[d][0], R[d][1] = origin[d], origin[d] + ng[d] * h[d]
        R[d][2], R[d][3] = origin[d] + 3.0 * h[d], origin[d] + 2.5 * h[d]
    out.update(sr_ng=np.array(ng), sr_origin=np.array(origin), sr_h=np.array(h), sr_x=R[0], sr_y=R[1], sr_z=R[2],
               sr_field=refshim.orb_scatter_r(ng, origin, h, *R))
    path = os.path.join(os.path.dirname(__file__), "ref_orb.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
