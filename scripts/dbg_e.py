import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ippl_b200 as ib
ctx = ib.Context(0); dev = ctx.device
n = 1 << 27; grid = 128; L = 4 * np.pi; h = [L / grid] * 3
mesh = ib.Mesh.make((grid,) * 3, (0, 0, 0), h)
q = -(L ** 3) / n; cap = n
parts = ib.Particles(cap, dev, q=q)
landau = ib.Dist.make([1, 1, 1], [0.05, 0.5] * 3)
regs = ib.Layout((grid,) * 3, 1).regions((0, 0, 0), h)
counts, ub = ib.sample_counts(landau, [0.0] * 3, [L] * 3, regs, n)
print("ubounds", ub)
ctx.sample_positions(landau, ub[0][:3], ub[0][3:], 42, 0, n, parts)
for k in "xyz":
    x = parts.arr[k][:n]
    print(k, "min", float(x.min()), "max", float(x.max()), "n<=0", int((x <= 0).sum()), "n>L", int((x > L).sum()), "nan", int(torch.isnan(x).sum()))
    hist = torch.histc(x, bins=16, min=0, max=L)
    print("   hist", (hist / n * 16).cpu().numpy().round(3))
for d, k in enumerate("xyz"):
    parts.arr[k][:n].clamp_(min=float(np.nextafter(0.0, 1.0)), max=L)
rho, ef = ctx.field(mesh), ctx.field(mesh, 3)
ctx.scatter(mesh, parts.arr["x"], parts.arr["y"], parts.arr["z"], q, rho)
ctx.halo_accumulate_periodic(mesh, rho)
r3 = rho.view(130, 130, 130)[1:-1, 1:-1, 1:-1]
print("rho raw interior min/max/mean", float(r3.min()), float(r3.max()), float(r3.mean()), "expected mean", q * 64)
ctx.field_density(mesh, rho, h[0] ** 3, q * n / L ** 3)
print("rho dens interior min/max/mean", float(r3.min()), float(r3.max()), float(r3.mean()))
am = int(r3.abs().argmax()); print("argmax rho (z,y,x)", np.unravel_index(am, (128, 128, 128)))
sol = ib.Poisson(ctx, mesh); sol.solve(rho, ef)
e4 = ef.view(130, 130, 130, 3)[1:-1, 1:-1, 1:-1]
print("E interior absmax per comp", [float(e4[..., c].abs().max()) for c in range(3)], "rms", [float(e4[..., c].pow(2).mean().sqrt()) for c in range(3)])
am = int(e4.abs().amax(dim=3).argmax()); print("argmax E (z,y,x)", np.unravel_index(am, (128, 128, 128)))
