"""Generates tests/golden/ref_random.npz from the REAL reference headers of the sampling path (Random/Distribution.h,
NormalDistribution.h, Utility.h NewtonRaphson, InverseTransformSampling.h, Randn.h, and the managers' own
CustomDistributionFunctions structs cut out of demos/alpine/*Manager.h at build time) through oracle/_ref/
libippl_refshim_random.so, with the random numbers REPLAYED from the arrays stored next to the results.  Run here (the
container that has /root/reference):
    python tests/golden/make_golden_random.py
The fixtures pin oracle/extras.py (and through it the CUDA sampler) where /root/reference does not exist."""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402
from oracle import refshim  # noqa: E402

CASES = {
    # name: (kind of the shim, kinds of the restatement, par, rmin, rmax)
    "landau": (1, [1, 1, 1], [0.05, 0.5] * 3, [0.0] * 3, [4 * math.pi] * 3),
    "penning": (2, [2, 2, 2], [10.0, 3.0, 10.0, 1.0, 10.0, 4.0], [0.0] * 3, [20.0] * 3),
    "bumpontail": (3, [0, 0, 1], [0.01, 0.21] * 3, [0.0] * 3, [2 * math.pi / 0.21] * 3),
}


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    ng = (32, 32, 32)
    for name, (kind, kinds, par, rmin, rmax) in CASES.items():
        h = [(rmax[d] - rmin[d]) / ng[d] for d in range(3)]
        xs = rng.uniform(rmin[0], rmax[0], 64)
        us = rng.uniform(0.0, 1.0, 64)
        ev = np.array([[[refshim.rand_eval(kind, par, w, d, x, u) for x, u in zip(xs, us)] for d in range(3)] for w in range(5)])
        fp = np.array([refshim.rand_full_pdf(kind, par, (a, b, c)) for a, b, c in zip(xs, xs[::-1], np.roll(xs, 7))])
        out.update({f"{name}_par": np.array(par), f"{name}_kinds": np.array(kinds), f"{name}_xs": xs, f"{name}_us": us,
                    f"{name}_eval": ev, f"{name}_fullpdf": fp})
        for nranks in (1, 2, 4, 8):
            boxes = oracle.partition(ng, nranks)
            regs = oracle.regions(ng, boxes, rmin, h)
            for ntotal in (1 << 20, 10_000_000, 12345):
                nloc, ub, _ = refshim.rand_sampling(kind, par, rmin, rmax, regs, ntotal)
                out[f"{name}_counts_{nranks}_{ntotal}"] = np.array(nloc)
                out[f"{name}_ubounds_{nranks}_{ntotal}"] = ub
        # generate() on rank 3 of 4 with replayed uniforms
        boxes = oracle.partition(ng, 4)
        regs = oracle.regions(ng, boxes, rmin, h)
        nloc, ub, _ = refshim.rand_sampling(kind, par, rmin, rmax, regs, 20000)
        u01 = rng.random((3, nloc[3]))
        _, _, x = refshim.rand_sampling(kind, par, rmin, rmax, regs, 20000, gen_rank=3, u01=u01)
        out.update({f"{name}_gen_regs": regs, f"{name}_gen_u01": u01, f"{name}_gen_x": x, f"{name}_gen_rmin": np.array(rmin),
                    f"{name}_gen_rmax": np.array(rmax)})
    g = rng.normal(size=(500, 3))
    out.update({"randn_g": g, "randn_mu": np.array([0.0, 0.5, 4.0]), "randn_sd": np.array([1.0, 0.25, 1 / math.sqrt(2)]),
                "randn_v": refshim.rand_randn([0.0, 0.5, 4.0], [1.0, 0.25, 1 / math.sqrt(2)], g)})
    path = os.path.join(os.path.dirname(__file__), "ref_random.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
