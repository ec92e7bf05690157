"""A/B timing of the fused step's kernel variants (IPPLB_FUSED_VAR) inside ONE process on the C2 workload
(LandauDamping 128^3, 2^27 particles): the variants are interleaved round-robin so that drift of the box (clocks,
memory placement) hits all of them alike.  usage: python scripts/ab_fused.py "4 6 12" [rounds] [steps_per_round]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ippl_b200 as ib  # noqa: E402

variants = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0 4").split()]
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
log2n = int(os.environ.get("AB_LOG2N", "27"))
grid = int(os.environ.get("AB_GRID", "128"))

ctx = ib.Context(0)
dev = ctx.device
n = 1 << log2n
L = 4 * np.pi
h = [L / grid] * 3
mesh = ib.Mesh.make((grid,) * 3, (0, 0, 0), h)
dt = min(0.05, 0.5 * h[0])
q = -(L ** 3) / n
cap = int(n * 1.25)
parts, scratch = ib.Particles(cap, dev, q=q), ib.Particles(cap, dev)
landau = ib.Dist.make([1, 1, 1], [0.05, 0.5] * 3)
regs = ib.Layout((grid,) * 3, 1).regions((0, 0, 0), h)
counts, ub = ib.sample_counts(landau, [0.0] * 3, [L] * 3, regs, n)
ctx.sample_positions(landau, ub[0][:3], ub[0][3:], 42, 0, n, parts)
ctx.sample_normal([0.0] * 3, [1.0] * 3, 42, 0, n, parts)
for d, k in enumerate("xyz"):
    parts.arr[k][:n].clamp_(min=float(np.nextafter(0.0, 1.0)), max=L)
parts.n = n
rho, ef = ctx.field(mesh), ctx.field(mesh, 3)
ctx.scatter(mesh, parts.arr["x"], parts.arr["y"], parts.arr["z"], q, rho, end=n)
ctx.halo_accumulate_periodic(mesh, rho)
ctx.field_density(mesh, rho, h[0] ** 3, q * n / L ** 3)
sol = ib.Poisson(ctx, mesh)
sol.solve(rho, ef)
ctx.halo_fill_periodic(mesh, ef, 3)
bins = ib.Bins(ctx, mesh, cap)
bins.build(parts, scratch)
parts.arr, scratch.arr = scratch.arr, parts.arr
push = ib.leapfrog_push(dt)


def step():
    ctx.pic_step(mesh, push, parts, scratch, None, ef, rho, do_sort=2, bins=bins)


for _ in range(3):
    step()
torch.cuda.synchronize()
res = {v: [] for v in variants}
for r in range(rounds):
    for v in variants:
        os.environ["IPPLB_FUSED_VAR"] = str(v)
        step()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        torch.cuda.synchronize()
        res[v].append(a.elapsed_time(b) / steps)
        st = bins.status()
        if st[0] != n or (st[3] & 7):
            print(f"var {v} round {r}: status {st}")
st = bins.status()
assert st[0] == n and (st[3] & 7) == 0, st
base = np.median(res[variants[0]])
for v in variants:
    x = np.array(res[v])
    print(f"var {v:3d}: median {np.median(x):.4f} ms/step  min {x.min():.4f}  max {x.max():.4f}  vs first {np.median(x) / base:.4f}")
