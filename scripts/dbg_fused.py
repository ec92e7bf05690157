import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import ippl_b200 as ib
from util import normal_velocities
ctx = ib.Context(0)
ppc, vscale = 40, 3.0
nr = (20, 16, 12)
n = nr[0] * nr[1] * nr[2] * ppc
h = [4 * np.pi / 16] * 3
L = [nr[d] * h[d] for d in range(3)]
mg = ib.Mesh.make(nr, (0, 0, 0), h)
rng = np.random.default_rng(100 + ppc)
R = [rng.uniform(0, L[d], n) for d in range(3)]
P = [vscale * p for p in normal_velocities(n, seed=7)]
dt = 0.5 * h[0]
ef = 0.2 * rng.normal(size=mg.cells * 3)
q = -0.37
push = ib.leapfrog_push(dt)
pa = ib.Particles.from_host(R, P, ctx.device, q=q)
cap = int(1.6 * n) + 4096
src = ib.Particles.from_host(R, P, ctx.device, q=q)
pb, sc = ib.Particles(cap, ctx.device, q=q), ib.Particles(cap, ctx.device, q=q)
bins = ib.Bins(ctx, mg, cap)
bins.build(src, pb)
print("build", bins.status())
rho = ctx.field(mg)
efd = torch.from_numpy(ef).to(ctx.device)
exit_buf = torch.zeros(6 * 1000, dtype=torch.float64, device=ctx.device)
for it in range(3):
    ctx.gather_push(mg, push, pa, efd)
    Ro = pa.host()
    print("ref range", [(a.min(), a.max(), Ld) for a, Ld in zip(Ro[:3], L)])
    rho.zero_()
    bins.step(push, pb, sc, efd, rho, exit_buf=exit_buf)
    st = bins.status()
    print("it", it, st)
    if st[2]:
        print(exit_buf.view(6, 1000)[:, :st[2]].cpu().numpy().T)
