"""Experiment: cost of gather_push / atomic scatter as particles drift away from cell-sorted order."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ippl_b200 as ib

ctx = ib.Context(0)
dev = ctx.device
n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 27
nr = (128, 128, 128)
L = 4 * np.pi
h = [L / 128] * 3
mesh = ib.Mesh.make(nr, (0, 0, 0), h)
g = torch.Generator(device=dev); g.manual_seed(1)
parts, scratch = ib.Particles(n, dev, q=-1.0), ib.Particles(n, dev)
for k in "xyz":
    parts.arr[k].uniform_(0, L, generator=g).clamp_(max=float(np.nextafter(L, 0)))
for k in ("px", "py", "pz"):
    parts.arr[k].normal_(0, 1, generator=g)
parts.n = n
off = ctx.offsets_buffer(mesh)
rho, ef = ctx.field(mesh), ctx.field(mesh, 3)
ef.normal_(0, 0.02, generator=g)
dt = 0.5 * h[0]

def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)

ctx.sort_by_cell(mesh, parts, scratch, off); parts.arr, scratch.arr = scratch.arr, parts.arr
push = ib.leapfrog_push(dt)
print("steps_since_sort gather_push_ms scatter_atomic_ms")
for s in range(0, 7):
    tp = timed(lambda: ctx.gather_push(mesh, push, parts, ef))
    rho.zero_()
    ts = timed(lambda: ctx.scatter(mesh, parts.arr["x"], parts.arr["y"], parts.arr["z"], -1.0, rho))
    print(s + 1, round(tp, 3), round(ts, 3))
t_keys = timed(lambda: ctx.sort_by_cell(mesh, parts, scratch, off))
print("sort_ms", round(t_keys, 3))
