#!/usr/bin/env python
"""BASELINE.json configs[4]: scatter / gather microbench sweep on one GPU -- 1..64 particles per cell on a 128^3 (or
--grid N) local grid, cell-sorted vs random particle order; scatter (atomic, sorted), gather (E materialised),
fused gather+push and the single-pass fused step reported separately, as particles/s and as achieved algorithmic
GB/s (SURVEY 8d component figures: scatter 24 B with the uniform charge, gather 48 B, gather+push 96 B, step 120 B).

  python scripts/microbench.py [--grid 128] [--ppc 1 2 4 8 16 32 64] [--out profiles/r1_microbench.md]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--ppc", type=int, nargs="+", default=[1, 2, 4, 8, 16, 32, 64])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import ippl_b200 as ib

    ctx = ib.Context(0)
    dev = ctx.device
    ng = (args.grid,) * 3
    L = 4 * np.pi
    h = [L / args.grid] * 3
    mesh = ib.Mesh.make(ng, (0, 0, 0), h)
    ncell = args.grid ** 3
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    dt = 0.5 * h[0]
    push = ib.leapfrog_push(dt)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    ef = ctx.field(mesh, 3)
    ef.normal_(0.0, 0.02, generator=g)
    ctx.halo_fill_periodic(mesh, ef, 3)
    rho = ctx.field(mesh)

    def timed(fn, reset=None):
        ts = []
        for _ in range(args.reps):
            if reset:
                reset()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts[1:])) if len(ts) > 2 else float(np.min(ts))

    rows = []
    for ppc in args.ppc:
        n = ncell * ppc
        cap = int(n * 1.3) + (1 << 16)
        q = -(L ** 3) / n
        base = ib.Particles(cap, dev, q=q)
        for k in "xyz":
            base.arr[k][:n].uniform_(0.0, 1.0, generator=g).mul_(L).clamp_(min=1e-9, max=L)
        for k in ("px", "py", "pz"):
            base.arr[k][:n].normal_(0.0, 1.0, generator=g)
        base.n = n
        srt = ib.Particles(cap, dev, q=q)
        off = ctx.offsets_buffer(mesh)
        ctx.sort_by_cell(mesh, base, srt, off)           # cell-sorted copy (+ offsets for the sorted scatter)
        work = ib.Particles(cap, dev, q=q)
        eout = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
        for order, src in (("sorted", srt), ("random", base)):
            def reset():
                for k in ib.Particles.NAMES:
                    work.arr[k][:n].copy_(src.arr[k][:n])
                work.n = n
            reset()
            x, y, z = (work.arr[k][:n] for k in "xyz")
            r = {"ppc": ppc, "order": order, "n": n}
            r["scatter_atomic"] = timed(lambda: ctx.scatter(mesh, x, y, z, q, rho), reset=lambda: ctx.field_fill(rho, 0.0))
            if order == "sorted":
                r["scatter_sorted"] = timed(lambda: ctx.scatter_sorted(mesh, n, x, y, z, q, off, rho),
                                            reset=lambda: ctx.field_fill(rho, 0.0))
            r["gather"] = timed(lambda: ctx.gather(mesh, x, y, z, ef, eout))
            r["gather_push"] = timed(lambda: ctx.gather_push(mesh, push, work, ef), reset=reset)
            rows.append(r)
        # single-pass fused step on the bucketed store (order is maintained by the step itself)
        bins = ib.Bins(ctx, mesh, cap)
        scratch = ib.Particles(cap, dev, q=q)
        bins.build(base, work)
        for _ in range(2):
            ctx.field_fill(rho, 0.0)
            bins.step(push, work, scratch, ef, rho)
        rows.append({"ppc": ppc, "order": "bucketed", "n": n,
                     "fused_step": timed(lambda: bins.step(push, work, scratch, ef, rho), reset=lambda: ctx.field_fill(rho, 0.0))})
        assert (bins.status()[3] & 7) == 0 and bins.status()[0] == n
        bins.close()
        del base, srt, work, scratch, eout, bins
        torch.cuda.empty_cache()

    bytes_per = {"scatter_atomic": 24, "scatter_sorted": 24, "gather": 48, "gather_push": 96, "fused_step": 120}
    lines = [f"# Round 1: scatter / gather microbench sweep (BASELINE.json configs[4]), {args.grid}^3 grid, 1x B200",
             "",
             f"`python scripts/microbench.py --grid {args.grid}`: uniform random positions, v ~ N(0,1), dt = 0.5 h, CUDA events, median of "
             f"{args.reps - 1} launches after one warm-up.  Each entry: ms | Gparticles/s | algorithmic GB/s (fraction of the measured "
             f"{peak:.0f} GB/s copy peak).  Algorithmic bytes per particle: scatter 24 (uniform charge), gather 48 (R in, E out), "
             "gather+push 96, fused step 120.",
             "",
             "| ppc | order | particles | scatter (atomic) | scatter (sorted) | gather | gather+push | fused step |",
             "|---:|---|---:|---|---|---|---|---|"]

    def cell(r, k):
        if k not in r:
            return "-"
        ms = r[k]
        gps = r["n"] / ms / 1e6
        gbs = bytes_per[k] * r["n"] / ms / 1e6
        return f"{ms:.3f} ms, {gps:.1f} Gp/s, {gbs:.0f} GB/s ({gbs / peak:.2f})"

    for r in rows:
        lines.append(f"| {r['ppc']} | {r['order']} | {r['n']} | " + " | ".join(
            cell(r, k) for k in ("scatter_atomic", "scatter_sorted", "gather", "gather_push", "fused_step")) + " |")
    text = "\n".join(lines) + "\n"
    print(text)
    if args.out:
        with open(args.out, "w") as f:
            f.write(text)
    ctx.close()


if __name__ == "__main__":
    main()
