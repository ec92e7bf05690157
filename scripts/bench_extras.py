#!/usr/bin/env python
"""Secondary measurements that bench.py attaches to its JSON line as `extras.micro` (run as a SEPARATE process after
the headline numbers are final, so that nothing in here can disturb or take down the headline run).

One GPU, 128^3 grid (BASELINE.json configs[4], reduced to the ppc values given):
  * the order-agnostic API kernels (ipplb_scatter_cic, ipplb_gather, ipplb_gather_push) on cell-sorted and on random
    particle order, and the single-pass fused step on the bucketed store, as particles/s;
  * ipplb_bins_build (counting sort into buckets), variant 1 (default) against variant 2 (arrival order, written without
    GPU access): time of each, and whether variant 2 produced the same tables and the same multiset of particles.
Prints ONE JSON line.  Every section is guarded: a failure is reported as {"error": ...} for that section only.

  python scripts/bench_extras.py [--device 0] [--grid 128] [--ppc 8 64] [--reps 4]
"""
import argparse
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--ppc", type=int, nargs="+", default=[8, 64])
    ap.add_argument("--reps", type=int, default=4)
    args = ap.parse_args()
    import torch
    import ippl_b200 as ib

    t_start = time.perf_counter()
    torch.cuda.set_device(args.device)
    ctx = ib.Context(args.device)
    dev = ctx.device
    ng = (args.grid,) * 3
    L = 4 * np.pi
    h = [L / args.grid] * 3
    mesh = ib.Mesh.make(ng, (0, 0, 0), h)
    ncell = args.grid ** 3
    push = ib.leapfrog_push(0.5 * h[0])
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

    out = {"grid": args.grid, "unit": "particles/s", "how": f"CUDA events, median of {max(args.reps - 1, 1)} launches after one warm-up; "
           "uniform random positions, v ~ N(0,1), dt = 0.5 h", "rows": [], "bins_build": []}
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
        work = ib.Particles(cap, dev, q=q)
        # ---- bucket build: variant 1 against variant 2 ------------------------------------------------------------
        try:
            res = {"ppc": ppc, "n": n}
            keep = {}
            for variant in (1, 2):
                bins = ib.Bins(ctx, mesh, cap)
                bins.set_build_variant(variant)
                ms = timed(lambda: bins.build(base, work))
                st = bins.status()
                assert (st[3] & 7) == 0 and st[0] == n, f"variant {variant}: status {st}"
                flat = ib.Particles(n, dev, q=q)
                assert bins.compact(work, flat) == n
                # one fused step on the store just built: its rho depends on every particle sitting in the right bucket
                scratch = ib.Particles(cap, dev, q=q)
                ctx.field_fill(rho, 0.0)
                bins.step(push, work, scratch, ef, rho)
                st = bins.status()
                assert (st[3] & 7) == 0 and st[0] == n, f"variant {variant}: status after one step {st}"
                keep[variant] = ([t.copy() for t in bins.tables()], torch.sort(flat.arr["x"][:n]).values,
                                 torch.sort(flat.arr["pz"][:n]).values, rho.clone())
                del scratch
                res[f"v{variant}_ms"] = ms
                res[f"v{variant}_gpps"] = n / ms / 1e6
                bins.close()
                del flat
            res["v2_same_tables"] = bool(all(np.array_equal(a, b) for a, b in zip(keep[1][0], keep[2][0])))
            res["v2_same_particles"] = bool(torch.equal(keep[1][1], keep[2][1]) and torch.equal(keep[1][2], keep[2][2]))
            res["v2_step_rho_rel_l2"] = float((keep[1][3] - keep[2][3]).norm() / keep[1][3].norm())
            res["speedup"] = res["v1_ms"] / res["v2_ms"]
            del keep
        except Exception as e:  # noqa: BLE001
            res = {"ppc": ppc, "error": f"{type(e).__name__}: {e}"[:400]}
        out["bins_build"].append(res)
        torch.cuda.empty_cache()
        # ---- API kernels, sorted against random order; the fused step ----------------------------------------------
        try:
            srt = ib.Particles(cap, dev, q=q)
            off = ctx.offsets_buffer(mesh)
            ctx.sort_by_cell(mesh, base, srt, off)
            eout = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
            for order, src in (("sorted", srt), ("random", base)):
                def reset():
                    for k in ib.Particles.NAMES:
                        work.arr[k][:n].copy_(src.arr[k][:n])
                    work.n = n
                reset()
                x, y, z = (work.arr[k][:n] for k in "xyz")
                r = {"ppc": ppc, "order": order, "n": n}
                ms = {"scatter_atomic": timed(lambda: ctx.scatter(mesh, x, y, z, q, rho), reset=lambda: ctx.field_fill(rho, 0.0)),
                      "gather": timed(lambda: ctx.gather(mesh, x, y, z, ef, eout)),
                      "gather_push": timed(lambda: ctx.gather_push(mesh, push, work, ef), reset=reset)}
                if order == "sorted":
                    ms["scatter_sorted"] = timed(lambda: ctx.scatter_sorted(mesh, n, x, y, z, q, off, rho),
                                                 reset=lambda: ctx.field_fill(rho, 0.0))
                r.update({k + "_gpps": n / v / 1e6 for k, v in ms.items()})
                out["rows"].append(r)
            del srt, eout
            bins = ib.Bins(ctx, mesh, cap)
            scratch = ib.Particles(cap, dev, q=q)
            bins.build(base, work)
            for _ in range(2):
                ctx.field_fill(rho, 0.0)
                bins.step(push, work, scratch, ef, rho)
            ms = timed(lambda: bins.step(push, work, scratch, ef, rho), reset=lambda: ctx.field_fill(rho, 0.0))
            st = bins.status()
            assert (st[3] & 7) == 0 and st[0] == n
            out["rows"].append({"ppc": ppc, "order": "bucketed", "n": n, "fused_step_gpps": n / ms / 1e6})
            bins.close()
            del scratch
        except Exception as e:  # noqa: BLE001
            out["rows"].append({"ppc": ppc, "error": f"{type(e).__name__}: {e}"[:400], "trace": traceback.format_exc()[-600:]})
        del base, work
        torch.cuda.empty_cache()
    out["seconds"] = time.perf_counter() - t_start
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
