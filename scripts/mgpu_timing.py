"""Per-phase timing of the multi-GPU step (torchrun): fused step / migrate / halo exchanges."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import ippl_b200 as ib
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
ctx = ib.Context(local); dev = ctx.device
dist.init_process_group("nccl", device_id=dev)
uid = [ib.nccl_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0); ctx.comm_init(rank, world, uid[0])
dims = [128, 128, 128]; v, d = world, 0
while v > 1:
    dims[d] *= 2; v //= 2; d = (d + 1) % 3
n_local = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 27
h = [4 * np.pi / 128] * 3; origin = (0.0, 0.0, 0.0)
layout = ib.Layout(tuple(dims), world); mesh = layout.mesh(rank, origin, h); ctx.set_layout(layout, origin, h)
reg = layout.regions(origin, h)[rank]
cap = int(n_local * 1.25)
g = torch.Generator(device=dev); g.manual_seed(42 + 100 * rank)
parts, scratch = ib.Particles(cap, dev, q=-1.0), ib.Particles(cap, dev)
for d_, k in enumerate("xyz"):
    parts.arr[k][:n_local].uniform_(0.0, 1.0, generator=g).mul_(reg[3 + d_] - reg[d_]).add_(reg[d_])
    parts.arr[k][:n_local].clamp_(min=float(np.nextafter(reg[d_], np.inf)), max=float(reg[3 + d_]))
for k in ("px", "py", "pz"):
    parts.arr[k][:n_local].normal_(0.0, 1.0, generator=g)
parts.n = n_local
bins = ib.Bins(ctx, mesh, cap)
rho, ef = ctx.field(mesh), ctx.field(mesh, 3)
ef.normal_(0.0, 0.02, generator=g)
exit_cap = max(n_local // 16, 1 << 16)
exit_buf = torch.zeros(6 * exit_cap, dtype=torch.float64, device=dev)
bins.build(parts, scratch); parts.arr, scratch.arr = scratch.arr, parts.arr
push = ib.leapfrog_push(0.5 * h[0])
def ev(): return torch.cuda.Event(enable_timing=True)
acc = {}
for it in range(8):
    e = [ev() for _ in range(6)]
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e[0].record(); ctx.halo_exchange(ef, 3, "fill")
    e[1].record(); ctx.field_fill(rho, 0.0); bins.step(push, parts, scratch, ef, rho, exit_buf=exit_buf, region=list(reg))
    e[2].record(); t0 = time.perf_counter(); st = bins.status(); t1 = time.perf_counter()
    e[3].record(); sent, recv = bins.migrate(parts, exit_buf, rho)
    e[4].record(); ctx.halo_exchange(rho, 1, "accumulate")
    e[5].record(); torch.cuda.synchronize()
    if it >= 3:
        for name, a, b in (("fillE", 0, 1), ("fused", 1, 2), ("status_sync", 2, 3), ("migrate", 3, 4), ("accum", 4, 5), ("total", 0, 5)):
            acc.setdefault(name, []).append(e[a].elapsed_time(e[b]))
if rank == 0:
    print({k: round(float(np.mean(v)), 3) for k, v in acc.items()}, "sent", sum(sent), "status", bins.status())
dist.destroy_process_group(); ctx.close()
