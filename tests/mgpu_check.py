"""Multi-GPU parity of the fused path against the CPU oracle (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py

Every rank evaluates the oracle for ALL ranks on the same seeded input (halo fill of E, gather + push + BC,
ParticleSpatialLayout::update, scatter, halo accumulate -- the reference order) and compares its own share
with what the CUDA path produced over NCCL:
  * ownership and migration counts: bit-exact (sent / received per rank pair, particles per rank),
  * the particles a rank holds after the step: the same multiset, bit for bit,
  * rho after accumulateHalo and E after fillHalo: relative L2 <= 1e-12 (north_star tolerance),
  * then one orthogonal-recursive-bisection repartition (scatterR weights, plane sums reduced over
    NCCL, findMedian cuts) whose boxes equal the oracle's binaryRepartition, followed by LoadBalancer::updateLayout
    (new FieldLayout, re-laid fields, ParticleSpatialLayout::update to the new owners, re-bucketing) and two more
    steps on the new decomposition held to the same checks.
Used by tests/test_gpu_parity.py::test_multi_gpu_fused_step (skipped with fewer than 2 GPUs)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOL = 1e-12


def canon(cols):
    a = np.stack(cols, axis=1)
    return a[np.lexsort(a.T[::-1])]


def main():
    import torch
    import torch.distributed as dist
    import ippl_b200 as ib
    import oracle
    from util import normal_velocities, rel_l2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    ctx = ib.Context(local)
    dev = ctx.device
    dist.init_process_group("nccl", device_id=dev)
    uid = [ib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(rank, world, uid[0])

    ng = (24, 16, 16)
    L0 = 4 * np.pi
    h = [L0 / 16] * 3
    origin = (0.0, 0.0, 0.0)
    Lg = [ng[d] * h[d] for d in range(3)]
    layout = ib.Layout(ng, world)
    boxes = layout.boxes()
    assert np.array_equal(boxes, oracle.partition(ng, world)), "FieldLayout boxes differ from the oracle"
    mesh = layout.mesh(rank, origin, h)
    ctx.set_layout(layout, origin, h)
    regs = layout.regions(origin, h)
    assert np.array_equal(regs, oracle.regions(ng, boxes, origin, h))

    # ---- identical global input on every rank ---------------------------------------------------------
    n = 60000
    rng = np.random.default_rng(2024)
    R = [rng.uniform(0, Lg[d], n) for d in range(3)]
    R[0] = np.clip(np.mod(rng.normal(0.3 * Lg[0], 0.2 * Lg[0], n), Lg[0]), 1e-9, Lg[0])  # off-centre blob: ORB has work to do
    R[0][:4] = [0.0, Lg[0], regs[0][3], np.nextafter(regs[0][3], np.inf)]  # on region / domain boundaries
    P = [2.0 * p for p in normal_velocities(n, seed=5)]
    dt, q = 0.5 * h[0], -0.01
    meshes_o = [oracle.Mesh.make(ng, origin, h, first=tuple(boxes[r, :3]), nl=tuple(boxes[r, 3:] - boxes[r, :3] + 1))
                for r in range(world)]
    ef_o = [0.1 * np.random.default_rng(100 + r).normal(size=m.ext[0] * m.ext[1] * m.ext[2] * 3) for r, m in enumerate(meshes_o)]
    own = oracle.locate(regs, 0, R[0], R[1], R[2])
    parts_o = [{k: a[own == r].copy() for k, a in zip(("x", "y", "z", "px", "py", "pz"), R + P)} for r in range(world)]

    # ---- CUDA side ------------------------------------------------------------------------------------------
    mine = parts_o[rank]
    n0 = len(mine["x"])
    cap = 4 * n // world + 4096
    src = ib.Particles.from_host([mine[k] for k in "xyz"], [mine[k] for k in ("px", "py", "pz")], dev, q=q)
    cur, nxt = ib.Particles(cap, dev, q=q), ib.Particles(cap, dev, q=q)
    bins = ib.Bins(ctx, mesh, cap)
    bins.build(src, cur)
    ef = torch.from_numpy(ef_o[rank].copy()).to(dev)
    rho = ctx.field(mesh)
    exit_cap = n
    exit_buf = torch.zeros(6 * exit_cap, dtype=torch.float64, device=dev)
    region = list(regs[rank])

    def step(it):
        nonlocal parts_o
        # oracle, all ranks: fillHalo(E); gather; kick, kick, drift; update (BC + migrate); scatter; accumulateHalo
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
        before = [len(p["x"]) for p in parts_o]
        # migration matrix of the oracle (after the BC that update applies)
        lo = [0 * h[d] + origin[d] for d in range(3)]
        hi = [ng[d] * h[d] + origin[d] for d in range(3)]
        wrapped = []
        for p in parts_o:
            w = {k: p[k].copy() for k in "xyz"}
            for d, k in enumerate("xyz"):
                oracle.periodic_bc(w[k], lo[d], hi[d])
            wrapped.append(w)
        dests = [oracle.locate(regs, r, w["x"], w["y"], w["z"]) for r, w in enumerate(wrapped)]
        sent_o = [[int((dests[r] == t).sum()) if t != r else 0 for t in range(world)] for r in range(world)]
        parts_o = oracle.update(ng, boxes, origin, h, parts_o)
        rho_o = [oracle.field_zeros(m) for m in meshes_o]
        for r in range(world):
            p = parts_o[r]
            oracle.scatter_cic(meshes_o[r], p["x"], p["y"], p["z"], q, rho_o[r])
        oracle.halo_full(ng, boxes, rho_o, 1, "accumulate")

        # CUDA: same step over NCCL
        ctx.halo_exchange(ef, 3, "fill")
        assert rel_l2(ef.cpu().numpy(), ef_o[rank]) <= TOL, "fillHalo(E) differs"
        rho.zero_()
        bins.step(ib.leapfrog_push(dt), cur, nxt, ef, rho, exit_buf=exit_buf, region=region)
        sent, recv = bins.migrate(cur, exit_buf, rho)
        ctx.halo_exchange(rho, 1, "accumulate")
        assert sent == sent_o[rank], f"rank {rank} step {it}: sent {sent} != oracle {sent_o[rank]}"
        assert recv == [sent_o[t][rank] for t in range(world)], f"rank {rank} step {it}: recv {recv}"
        nloc, ntail, nexit, flags = bins.status()
        assert (flags & 7) == 0 and nloc == len(parts_o[rank]["x"]) == cur.n, (nloc, len(parts_o[rank]["x"]), cur.n, flags)
        out = ib.Particles(max(nloc, 1), dev)
        assert bins.compact(cur, out) == nloc
        want = [parts_o[rank][k] for k in ("x", "y", "z", "px", "py", "pz")]
        assert np.array_equal(canon(out.host()), canon(want)), f"rank {rank} step {it}: particles differ from the oracle"
        err = rel_l2(rho.cpu().numpy(), rho_o[rank])
        assert err <= TOL, f"rank {rank} step {it}: rho rel L2 {err}"
        tot = torch.tensor([nloc], device=dev, dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot[0]) == n
        if rank == 0:
            print(f"step {it}: ok  sent={sent} recv={recv} n_local={nloc} tail={ntail} rho_rel_l2={err:.2e}", flush=True)

    for it in range(3):
        step(it)

    # ---- ORB repartition (OrthogonalRecursiveBisection::binaryRepartition + LoadBalancer::updateLayout) ----------
    from oracle import extras as ox
    nloc = bins.status()[0]
    flat = ib.Particles(cap, dev, q=q)
    assert bins.compact(cur, flat) == nloc
    w = ctx.field(mesh)
    ctx.scatter(mesh, flat.arr["x"][:nloc], flat.arr["y"][:nloc], flat.arr["z"][:nloc], 1.0, w)   # scatterR: weight 1
    ctx.halo_exchange(w, 1, "accumulate")
    new_boxes, ok = ctx.orb_repartition(mesh, world, w)
    w_o = [oracle.field_zeros(m) for m in meshes_o]
    for r in range(world):
        p = parts_o[r]
        oracle.scatter_cic(meshes_o[r], p["x"], p["y"], p["z"], 1.0, w_o[r])
    oracle.halo_full(ng, boxes, w_o, 1, "accumulate")
    W = np.zeros(ng[::-1])
    for r in range(world):
        b = boxes[r]
        W[b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1] = oracle.interior(w_o[r], meshes_o[r])
    want_boxes, want_ok = ox.orb_repartition(ng, world, W)
    assert ok and want_ok, "ORB could not repartition"
    assert np.array_equal(new_boxes, np.asarray(want_boxes, dtype=np.int32)), f"ORB boxes {new_boxes.tolist()} != oracle {want_boxes}"
    assert not np.array_equal(new_boxes, boxes), "the blob should have moved the cut"
    # LoadBalancer::updateLayout: fields and particle layout follow the new FieldLayout, then pc->update()
    layout.set_boxes(new_boxes)
    boxes = layout.boxes()
    mesh = layout.mesh(rank, origin, h)
    ctx.set_layout(layout, origin, h)
    regs = layout.regions(origin, h)
    assert np.array_equal(regs, oracle.regions(ng, boxes, origin, h))
    region = list(regs[rank])
    meshes_o = [oracle.Mesh.make(ng, origin, h, first=tuple(boxes[r, :3]), nl=tuple(boxes[r, 3:] - boxes[r, :3] + 1))
                for r in range(world)]
    ef_o = [0.1 * np.random.default_rng(200 + r).normal(size=m.ext[0] * m.ext[1] * m.ext[2] * 3) for r, m in enumerate(meshes_o)]
    before = [len(p["x"]) for p in parts_o]
    parts_o = oracle.update(ng, boxes, origin, h, parts_o)
    flat.n = nloc
    sent, recv = ctx.update(flat)
    assert flat.n == len(parts_o[rank]["x"]), (flat.n, len(parts_o[rank]["x"]))
    want = [parts_o[rank][k] for k in ("x", "y", "z", "px", "py", "pz")]
    assert np.array_equal(canon(flat.host()), canon(want)), f"rank {rank}: particles after the ORB update differ from the oracle"
    imb_old = max(before) / (n / world)
    imb_new = max(len(p["x"]) for p in parts_o) / (n / world)
    assert imb_new < imb_old, (imb_old, imb_new)
    bins.close()
    bins = ib.Bins(ctx, mesh, cap)
    bins.build(flat, cur)
    ef = torch.from_numpy(ef_o[rank].copy()).to(dev)
    rho = ctx.field(mesh)
    if rank == 0:
        print(f"ORB: boxes {boxes.tolist()} imbalance {imb_old:.3f} -> {imb_new:.3f}, moved {sum(sent)} particles from rank 0", flush=True)
    for it in range(3, 5):
        step(it)

    dist.barrier()
    if rank == 0:
        print(f"MGPU_CHECK_OK world={world}")
    bins.close()
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
