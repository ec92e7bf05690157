#!/usr/bin/env python
"""Secondary measurements that bench.py attaches to its JSON line as `extras.micro*` (run as SEPARATE processes after
the headline numbers are final, so that nothing in here can disturb or take down the headline run -- and one process per
`--part`, so that a fault in a kernel variant that has never run cannot take the measurements of the verified kernels
with it).

One GPU, 128^3 grid (BASELINE.json configs[4], reduced to the ppc values given), uniform random positions:
  --part verified         the order-agnostic API kernels (ipplb_scatter_cic, ipplb_scatter_cic_sorted, ipplb_gather_cic,
                          ipplb_gather_push) on cell-sorted and on random particle order, the single-pass fused step on the
                          bucketed store and ipplb_bins_build, as particles/s;
  --part gather_variants  the gather kernels with variant 2 of the field loads (ipplb_ctx_set_gather_variant: 16-byte loads
                          per x-pair of stencil nodes) next to variant 1, and whether the results are the same bits;
  --part build_variants   ipplb_bins_build variant 2 (ipplb_bins_set_build_variant: arrival order, warp-aggregated tile
                          cursors) next to variant 1: time of each, same tables, same multiset of particles, and the rho of
                          one fused step on either store.
Both variants were written without GPU access; variant 1 is the default everywhere.  Prints ONE JSON line.

  python scripts/bench_extras.py --part verified [--device 0] [--grid 128] [--ppc 8 64] [--reps 4]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--part", default="verified", choices=["verified", "gather_variants", "build_variants"])
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
    L = 4 * np.pi
    h = [L / args.grid] * 3
    mesh = ib.Mesh.make((args.grid,) * 3, (0, 0, 0), h)
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

    def gpps(n, ms):
        return n / ms / 1e6

    def zero_rho():
        ctx.field_fill(rho, 0.0)

    out = {"part": args.part, "grid": args.grid, "unit": "G particles/s (keys *_gpps), ms (keys *_ms)",
           "how": f"CUDA events, median of {max(args.reps - 1, 1)} launches after one warm-up; uniform random positions, v ~ N(0,1), "
                  "dt = 0.5 h", "rows": []}
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

        def sorted_copy():
            srt, off = ib.Particles(cap, dev, q=q), ctx.offsets_buffer(mesh)
            ctx.sort_by_cell(mesh, base, srt, off)
            return srt, off

        def loader(src):
            def reset():
                for k in ib.Particles.NAMES:
                    work.arr[k][:n].copy_(src.arr[k][:n])
                work.n = n
            return reset

        if args.part == "verified":
            srt, off = sorted_copy()
            eout = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
            for order, src in (("sorted", srt), ("random", base)):
                reset = loader(src)
                reset()
                x, y, z = (work.arr[k][:n] for k in "xyz")
                r = {"ppc": ppc, "order": order, "n": n,
                     "scatter_atomic_gpps": gpps(n, timed(lambda: ctx.scatter(mesh, x, y, z, q, rho), reset=zero_rho)),
                     "gather_gpps": gpps(n, timed(lambda: ctx.gather(mesh, x, y, z, ef, eout))),
                     "gather_push_gpps": gpps(n, timed(lambda: ctx.gather_push(mesh, push, work, ef), reset=reset))}
                if order == "sorted":
                    reset()     # back to the cell-sorted positions the offsets describe
                    r["scatter_sorted_gpps"] = gpps(n, timed(lambda: ctx.scatter_sorted(mesh, n, x, y, z, q, off, rho), reset=zero_rho))
                out["rows"].append(r)
            del srt, eout
            bins = ib.Bins(ctx, mesh, cap)
            scratch = ib.Particles(cap, dev, q=q)
            build_ms = timed(lambda: bins.build(base, work))
            for _ in range(2):
                zero_rho()
                bins.step(push, work, scratch, ef, rho)
            ms = timed(lambda: bins.step(push, work, scratch, ef, rho), reset=zero_rho)
            st = bins.status()
            assert (st[3] & 7) == 0 and st[0] == n, f"fused store: status {st}"
            out["rows"].append({"ppc": ppc, "order": "bucketed", "n": n, "fused_step_gpps": gpps(n, ms),
                                "bins_build_ms": build_ms, "bins_build_gpps": gpps(n, build_ms)})
            bins.close()
            del scratch

        elif args.part == "gather_variants":
            srt, off = sorted_copy()
            eout = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
            for order, src in (("sorted", srt), ("random", base)):
                reset = loader(src)
                reset()
                x, y, z = (work.arr[k][:n] for k in "xyz")
                r = {"ppc": ppc, "order": order, "n": n}
                res = {}
                for variant in (1, 2):
                    ctx.set_gather_variant(variant)
                    r[f"gather_v{variant}_gpps"] = gpps(n, timed(lambda: ctx.gather(mesh, x, y, z, ef, eout)))
                    r[f"gather_push_v{variant}_gpps"] = gpps(n, timed(lambda: ctx.gather_push(mesh, push, work, ef), reset=reset))
                    # eout: this variant's gather of src; work: ONE push of src by this variant (reset precedes every call)
                    res[variant] = [e.clone() for e in eout] + [work.arr[k][:n].clone() for k in ib.Particles.NAMES]
                    reset()
                ctx.set_gather_variant(1)
                r["v2_same_bits"] = bool(all(torch.equal(a, b) for a, b in zip(res[1], res[2])))
                r["gather_speedup"] = r["gather_v2_gpps"] / r["gather_v1_gpps"]
                r["gather_push_speedup"] = r["gather_push_v2_gpps"] / r["gather_push_v1_gpps"]
                del res
                out["rows"].append(r)
            del srt, eout

        else:   # build_variants
            r = {"ppc": ppc, "order": "random", "n": n}
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
                zero_rho()
                bins.step(push, work, scratch, ef, rho)
                st = bins.status()
                assert (st[3] & 7) == 0 and st[0] == n, f"variant {variant}: status after one step {st}"
                keep[variant] = ([t.copy() for t in bins.tables()], torch.sort(flat.arr["x"][:n]).values,
                                 torch.sort(flat.arr["pz"][:n]).values, rho.clone())
                r[f"build_v{variant}_ms"] = ms
                r[f"build_v{variant}_gpps"] = gpps(n, ms)
                bins.close()
                del flat, scratch
            r["v2_same_tables"] = bool(all(np.array_equal(a, b) for a, b in zip(keep[1][0], keep[2][0])))
            r["v2_same_particles"] = bool(torch.equal(keep[1][1], keep[2][1]) and torch.equal(keep[1][2], keep[2][2]))
            r["v2_step_rho_rel_l2"] = float((keep[1][3] - keep[2][3]).norm() / keep[1][3].norm())
            r["build_speedup"] = r["build_v1_ms"] / r["build_v2_ms"]
            del keep
            out["rows"].append(r)
        del base, work
        torch.cuda.empty_cache()
    out["seconds"] = time.perf_counter() - t_start
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
