"""Generates tests/golden/ref_region.npz from the REAL RegionLayout of the reference (oracle/_ref/
libippl_refshim_region.so: Region/RegionLayout.h/.hpp + Meshes/UniformCartesian.h/.hpp compiled in place from
/root/reference): the physical region of every rank -- the doubles ParticleSpatialLayout::positionInRegion compares
particle positions with.  Run here (the container that has /root/reference):  python tests/golden/make_golden_region.py"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import refshim  # noqa: E402

GRIDS = [(16, 16, 16), (128, 128, 128), (32, 20, 12), (17, 9, 33), (256, 256, 256), (512, 512, 512)]
RANKS = (1, 2, 3, 4, 6, 8)


def meshes(ng):
    yield (0.0, 0.0, 0.0), tuple(4 * math.pi / n for n in ng)              # LandauDamping
    yield (0.0, 0.0, 0.0), tuple(20.0 / n for n in ng)                     # PenningTrap
    yield (0.0, 0.0, 0.0), tuple(2 * math.pi / 0.21 / n for n in ng)       # BumponTail
    yield (0.5, -1.0, 2.0), (0.1, 0.2, 0.3)


def main():
    out = {}
    for gi, ng in enumerate(GRIDS):
        for nr in RANKS:
            for mi, (origin, h) in enumerate(meshes(ng)):
                out[f"reg_{gi}_{nr}_{mi}"] = refshim.regions(ng, nr, origin, h)
    # after an ORB repartition: unequal boxes
    boxes = np.array([[0, 0, 0, 8, 15, 15], [9, 0, 0, 23, 15, 15]], dtype=np.int32)
    out["orb_boxes"] = boxes
    out["orb_regions"] = refshim.regions((24, 16, 16), 2, (0.0, 0.0, 0.0), (4 * math.pi / 16,) * 3, boxes=boxes)
    path = os.path.join(os.path.dirname(__file__), "ref_region.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
