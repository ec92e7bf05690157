"""Debug: run the fused step on C2 for many steps with an exit buffer; report the first leaver on a periodic single rank."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ippl_b200 as ib
var = sys.argv[1] if len(sys.argv) > 1 else "0"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
os.environ["IPPLB_FUSED_VAR"] = var
ctx = ib.Context(0); dev = ctx.device
n = 1 << 27; grid = 128; L = 4 * np.pi; h = [L / grid] * 3
mesh = ib.Mesh.make((grid,) * 3, (0, 0, 0), h)
dt = min(0.05, 0.5 * h[0]); q = -(L ** 3) / n; cap = int(n * 1.25)
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
sol = ib.Poisson(ctx, mesh); sol.solve(rho, ef); ctx.halo_fill_periodic(mesh, ef, 3)
print("E absmax", float(ef.abs().max()), "P absmax", max(float(parts.arr[k][:n].abs().max()) for k in ("px", "py", "pz")))
bins = ib.Bins(ctx, mesh, cap); bins.build(parts, scratch)
parts.arr, scratch.arr = scratch.arr, parts.arr
push = ib.leapfrog_push(dt)
exit_buf = torch.zeros(6 * 4096, dtype=torch.float64, device=dev)
for it in range(nsteps):
    rho.zero_()
    bins.step(push, parts, scratch, ef, rho, exit_buf=exit_buf)
    st = bins.status()
    if st[2] or st[0] != n or (st[3] & 7):
        print("step", it, "status", st)
        print(exit_buf.view(-1, 6)[:max(1, st[2])].cpu().numpy())
        break
else:
    print("no leak in", nsteps, "steps; P absmax", max(float(parts.arr[k].abs().max()) for k in ("px", "py", "pz")))
