"""One small oracle-checked multi-rank step over NCCL (the peer-memory migration path), callable from inside a running
job: bench.py --gpus N runs it before its timed region and prints the result as "parity".  The oracle is the checker
here, never the thing measured.  Every rank evaluates the oracle for ALL ranks on the same seeded input and compares its
own share (same recipe as tests/test_gpu_loop.py, which runs it on in-process ranks)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def canon(cols):
    a = np.stack(cols, axis=1)
    return a[np.lexsort(a.T[::-1])]


def multi_rank_step(ctx, dist, rank, world, steps=2):
    import torch
    import ippl_b200 as ib
    import oracle
    from util import normal_velocities, rel_l2
    dev = ctx.device
    ng = (24, 16, 16)
    h = [4 * np.pi / 16] * 3
    origin = (0.0, 0.0, 0.0)
    Lg = [ng[d] * h[d] for d in range(3)]
    layout = ib.Layout(ng, world)
    boxes = layout.boxes()
    mesh = layout.mesh(rank, origin, h)
    ctx.set_layout(layout, origin, h)
    regs = layout.regions(origin, h)
    n = 60000
    rng = np.random.default_rng(2024)
    R = [rng.uniform(0, Lg[d], n) for d in range(3)]
    R[0][:4] = [0.0, Lg[0], regs[0][3], np.nextafter(regs[0][3], np.inf)]
    P = [2.0 * p for p in normal_velocities(n, seed=5)]
    dt, q = 0.5 * h[0], -0.01
    names = ("x", "y", "z", "px", "py", "pz")
    meshes_o = [oracle.Mesh.make(ng, origin, h, first=tuple(boxes[r, :3]), nl=tuple(boxes[r, 3:] - boxes[r, :3] + 1)) for r in range(world)]
    ef_o = [0.1 * np.random.default_rng(100 + r).normal(size=m.ext[0] * m.ext[1] * m.ext[2] * 3) for r, m in enumerate(meshes_o)]
    own = oracle.locate(regs, 0, R[0], R[1], R[2])
    parts_o = [{k: a[own == r].copy() for k, a in zip(names, R + P)} for r in range(world)]
    mine = parts_o[rank]
    cap = 4 * n // world + 4096
    src = ib.Particles.from_host([mine[k] for k in "xyz"], [mine[k] for k in ("px", "py", "pz")], dev, q=q)
    cur, nxt = ib.Particles(cap, dev, q=q), ib.Particles(cap, dev, q=q)
    bins = ib.Bins(ctx, mesh, cap)
    bins.build(src, cur)
    ef = torch.from_numpy(ef_o[rank].copy()).to(dev)
    rho = ctx.field(mesh)
    ctx.migrate_connect(max(n // 2, 1024))
    worst_rho, counts_exact, particles_exact, e_exact = 0.0, True, True, True
    for it in range(steps):
        oracle.halo_full(ng, boxes, ef_o, 3, "fill")
        for r in range(world):
            p = parts_o[r]
            nn = len(p["x"])
            E = [np.zeros(nn) for _ in range(3)]
            oracle.gather_cic(meshes_o[r], p["x"], p["y"], p["z"], ef_o[r], E)
            for d, k in enumerate(("px", "py", "pz")):
                oracle.kick(p[k], E[d], 0.5 * dt)
                oracle.kick(p[k], E[d], 0.5 * dt)
            for kx, kp in zip("xyz", ("px", "py", "pz")):
                oracle.drift(p[kx], p[kp], dt)
        wrapped = []
        for p in parts_o:
            w = {k: p[k].copy() for k in "xyz"}
            for d, k in enumerate("xyz"):
                oracle.periodic_bc(w[k], 0.0, Lg[d])
            wrapped.append(w)
        dests = [oracle.locate(regs, r, w["x"], w["y"], w["z"]) for r, w in enumerate(wrapped)]
        sent_o = [[int((dests[r] == t).sum()) if t != r else 0 for t in range(world)] for r in range(world)]
        parts_o = oracle.update(ng, boxes, origin, h, parts_o)
        rho_o = [oracle.field_zeros(m) for m in meshes_o]
        for r in range(world):
            p = parts_o[r]
            oracle.scatter_cic(meshes_o[r], p["x"], p["y"], p["z"], q, rho_o[r])
        oracle.halo_full(ng, boxes, rho_o, 1, "accumulate")

        ctx.halo_exchange(ef, 3, "fill")
        e_exact &= bool(np.array_equal(ef.cpu().numpy(), ef_o[rank]))
        rho.zero_()
        bins.step(ib.leapfrog_push(dt), cur, nxt, ef, rho, region=list(regs[rank]))
        bins.migrate_async(cur, rho)
        ctx.halo_exchange(rho, 1, "accumulate")
        sent, recv = ctx.migrate_counts()
        counts_exact &= sent == sent_o[rank] and recv == [sent_o[t][rank] for t in range(world)]
        nloc, ntail, nexit, flags = bins.status()
        counts_exact &= (flags & 7) == 0 and nloc == len(parts_o[rank]["x"])
        out = ib.Particles(max(nloc, 1), dev)
        bins.compact(cur, out)
        particles_exact &= bool(np.array_equal(canon(out.host()), canon([parts_o[rank][k] for k in names])))
        worst_rho = max(worst_rho, float(rel_l2(rho.cpu().numpy(), rho_o[rank])))
    bins.close()
    layout.close()
    t = torch.tensor([worst_rho, 0.0 if counts_exact else 1.0, 0.0 if particles_exact else 1.0, 0.0 if e_exact else 1.0],
                     device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res = {"rho_rel_l2": float(t[0]), "counts_exact": float(t[1]) == 0.0, "particles_bit_exact": float(t[2]) == 0.0,
           "e_halo_bit_exact": float(t[3]) == 0.0, "steps": steps, "ranks": world,
           "what": "24x16x16 mesh, 60000 particles, fused step + peer-memory migration + halo exchanges over NCCL against the "
                   "oracle's all-ranks step; max over ranks"}
    assert res["counts_exact"] and res["particles_bit_exact"] and res["e_halo_bit_exact"] and res["rho_rel_l2"] <= 1e-12, res
    return res


if __name__ == "__main__":   # torchrun --nproc-per-node N tests/mgpu_parity.py
    import torch
    import torch.distributed as dist
    import ippl_b200 as ib
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    ctx = ib.Context(local)
    dist.init_process_group("nccl", device_id=ctx.device)
    uid = [ib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])
    r = multi_rank_step(ctx, dist, rank, world, steps=4)
    if rank == 0:
        print("MGPU_PARITY_OK", r)
    dist.destroy_process_group()
    ctx.close()
