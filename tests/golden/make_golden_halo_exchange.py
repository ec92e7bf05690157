"""Generates tests/golden/ref_halo_exchange.npz: ghosted fields of EVERY rank before and after the reference's real
BareField::fillHalo / accumulateHalo -- HaloCells::exchangeBoundaries (pack, send / receive with the component tags,
unpack with = or +=) followed by applyPeriodicSerialDim, executed by the reference's own code with one thread per rank
over an in-process mailbox (oracle/ref_shim/refshim_halo.cpp, Communicate/Communicator.h stand-in).  Default partitions
(2, 3, 4, 8 ranks) and one ORB layout with unequal boxes; scalar and Vector<double,3> fields.
Run here (needs /root/reference):  python tests/golden/make_golden_halo_exchange.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402
from oracle import refshim  # noqa: E402

CASES = [((12, 10, 8), 2, None), ((12, 10, 8), 3, None), ((12, 10, 8), 4, None), ((8, 8, 8), 8, None),
         ((12, 8, 8), 2, np.array([[0, 0, 0, 4, 7, 7], [5, 0, 0, 11, 7, 7]], dtype=np.int32)),
         ((10, 8, 8), 4, np.array([[0, 0, 0, 3, 3, 7], [0, 4, 0, 3, 7, 7], [4, 0, 0, 9, 4, 7], [4, 5, 0, 9, 7, 7]],
                                  dtype=np.int32))]


def boxes_of(case):
    ng, nr, b = case
    return oracle.partition(ng, nr) if b is None else b


def main():
    rng = np.random.default_rng(20261025)
    out = {}
    for ci, case in enumerate(CASES):
        ng, nr, b = case
        boxes = boxes_of(case)
        for ncomp in (1, 3):
            fields = [rng.normal(size=int(np.prod(boxes[r, 3:] - boxes[r, :3] + 3)) * ncomp) for r in range(nr)]
            for r in range(nr):
                out[f"in_{ci}_{ncomp}_{r}"] = fields[r]
            for mode in ("fill", "accumulate"):
                got = refshim.halo_exchange(ng, boxes, [f.copy() for f in fields], ncomp, mode, use_boxes=b is not None)
                for r in range(nr):
                    out[f"{mode}_{ci}_{ncomp}_{r}"] = got[r]
    path = os.path.join(os.path.dirname(__file__), "ref_halo_exchange.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
