"""World-size-2 (and 4) checks of the N > 1 host-side logic on CPU over a real process group (torch.distributed, gloo):
every rank runs the product's host code (FieldLayout mirror, sampling counts, ORB state machine of the C-ABI) on ITS
share only and the collectives are real -- per-rank plane sums all-reduced like perpendicularReduction + allreduce
(OrthogonalRecursiveBisection.hpp:44-62), the per-destination count exchange of ParticleSpatialLayout::update
(ParticleSpatialLayout.hpp:150-170) as an all-gather.  Results must equal the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ippl_b200 as ib
    import oracle
    from oracle import extras as ox
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ng = (24, 16, 20)
        h = (0.5, 0.75, 0.4)
        origin = (0.0, -1.0, 2.0)
        layout = ib.Layout(ng, world)
        boxes = layout.boxes()
        regs = layout.regions(origin, h)
        # ---- sampling counts: each rank takes its entry, the sum over ranks is the total ------------------------------
        Lg = [ng[d] * h[d] for d in range(3)]
        dist_b = ib.Dist.make([2, 1, 0], [origin[0] + 0.4 * Lg[0], 0.2 * Lg[0], 0.05, 2 * np.pi / Lg[1], 0.0, 1.0])
        rmin, rmax = list(origin), [origin[d] + Lg[d] for d in range(3)]
        ntotal = 1_000_003
        nloc, _ = ib.sample_counts(dist_b, rmin, rmax, regs, ntotal)
        mine = torch.tensor([nloc[rank]], dtype=torch.int64)
        dist.all_reduce(mine)
        assert int(mine[0]) == ntotal
        # ---- particles: same seeded global set everywhere, each rank keeps what it owns -----------------------------------
        rng = np.random.default_rng(99)
        n = 50_000
        R = [np.clip(np.mod(rng.normal(0.35 * Lg[d], 0.2 * Lg[d], n), Lg[d]), 1e-9, Lg[d]) + origin[d] for d in range(3)]
        own = oracle.locate(regs, 0, *R)
        my = [r[own == rank] for r in R]
        # ---- ORB: scatterR weights of my particles on my box (+ the halo contributions the neighbours would send: the
        #      global field is assembled with an all-reduce, which is what accumulateHalo + plane allreduce amount to) ---------
        mo = oracle.Mesh.make(ng, origin, h)
        wloc = oracle.field_zeros(mo)
        oracle.scatter_cic(mo, *my, 1.0, wloc)
        oracle.halo_periodic(wloc, mo.ext, 1, 1, (1, 1, 1), "accumulate")
        W = torch.from_numpy(np.ascontiguousarray(oracle.interior(wloc, mo)))
        dist.all_reduce(W)                       # global weight field (every rank's deposit summed)
        W = W.numpy()
        b = boxes[rank]
        orb = ib.Orb(ng, world)
        while True:
            nxt = orb.next()
            if nxt is None:
                break
            lo, hi, axis = nxt
            # perpendicularReduction: my box clipped to the domain, zeros elsewhere
            red = np.zeros(hi[axis] - lo[axis] + 1)
            clo = [max(lo[d], b[d]) for d in range(3)]
            chi = [min(hi[d], b[3 + d]) for d in range(3)]
            if all(chi[d] >= clo[d] for d in range(3)):
                sub = W[clo[2]:chi[2] + 1, clo[1]:chi[1] + 1, clo[0]:chi[0] + 1]
                part = sub.sum(axis=tuple(a for a in range(3) if a != 2 - axis))
                red[clo[axis] - lo[axis]: chi[axis] - lo[axis] + 1] = part
            t = torch.from_numpy(red)
            dist.all_reduce(t)                   # comm.allreduce(reducedRank, reduced, ..., std::plus)
            orb.cut(t.numpy())
        new_boxes, ok = orb.finish()
        want, want_ok = ox.orb_repartition(ng, world, W)
        assert ok == want_ok and np.array_equal(new_boxes, np.asarray(want, dtype=np.int32)), (new_boxes, want)
        gathered = [torch.zeros(world * 6, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(new_boxes.reshape(-1).copy()))
        assert all(torch.equal(g, gathered[0]) for g in gathered)      # every rank derived the same layout
        # ---- updateLayout + update(): per-destination counts, exchanged as in ParticleSpatialLayout::update -----------------
        layout.set_boxes(new_boxes)
        regs2 = layout.regions(origin, h)
        assert np.array_equal(regs2, oracle.regions(ng, new_boxes, origin, h))
        dest = oracle.locate(regs2, rank, *my)
        sent = torch.tensor([int((dest == r).sum()) for r in range(world)], dtype=torch.int64)
        matrix = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(matrix, sent)
        M = torch.stack(matrix).numpy()          # M[s][d]: particles rank s sends to rank d (diagonal: stays)
        assert M.sum() == n and M[rank].sum() == len(my[0])
        new_counts = M.sum(axis=0)
        want_counts = np.bincount(oracle.locate(regs2, 0, *R), minlength=world)
        assert np.array_equal(new_counts, want_counts)
        old_counts = np.bincount(own, minlength=world)
        assert new_counts.max() < old_counts.max()   # the repartition balanced the blob
        layout.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_orb_sampling_and_count_exchange_over_gloo(world, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def _slab_worker(rank, world, port, out_dir):
    """one rank of the slab-decomposed FFT solve: MY plan (ipplb_slabplan_*), MY box, real messages (gloo send / recv), numpy
    transforms on my slabs only; the result is compared with the whole-domain solve computed redundantly from the seed"""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import ippl_b200 as ib
    import oracle
    from test_slabplan_cpu import Rank, orb_like, transforms
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        for ng, kind in (((16, 12, 10), "default"), ((24, 16, 16), "orb")):
            origin, h = (0.0, 0.5, -1.0), (0.3, 0.25, 0.4)
            layout = ib.Layout(ng, world)
            if kind == "orb":
                layout.set_boxes(orb_like(ng, world))
            rng = np.random.default_rng(7)
            rho_g = rng.normal(size=(ng[2], ng[1], ng[0]))
            rho_g -= rho_g.mean()
            N, nxh = ng[0] * ng[1] * ng[2], ng[0] // 2 + 1
            rhat = np.fft.rfftn(rho_g) / N
            want = np.stack([np.fft.irfftn(rhat * np.broadcast_to(M, rho_g.shape)[:, :, :nxh], s=rho_g.shape, axes=(0, 1, 2)) * N
                             for M in oracle.poisson_kspace_multipliers(ng, origin, h)], axis=-1)
            me = Rank(layout, rank, origin, h)
            g, f = me.plan.nghost, me.first
            rho = me.buf["rho"].reshape(me.ext[2], me.ext[1], me.ext[0])
            rho[g:-g, g:-g, g:-g] = rho_g[f[2]:f[2] + me.nl[2], f[1]:f[1] + me.nl[1], f[0]:f[0] + me.nl[0]]
            for phase in range(4):
                me.copies(phase, 0)
                reqs, keep = [], []
                for m in me.plan.rows(phase, 1):
                    if m["peer"] == rank:    # a message to myself is a local copy
                        me.buf["recv"][m["roff"]:m["roff"] + m["rcount"]] = me.buf["send"][m["soff"]:m["soff"] + m["scount"]]
                        continue
                    if m["scount"]:
                        t = torch.from_numpy(me.buf["send"][m["soff"]:m["soff"] + m["scount"]].copy())
                        keep.append(t)
                        reqs.append(dist.isend(t, m["peer"]))
                    if m["rcount"]:
                        t = torch.empty(m["rcount"], dtype=torch.float64)
                        keep.append((t, m))
                        reqs.append(dist.irecv(t, m["peer"]))
                for r in reqs:
                    r.wait()
                for item in keep:
                    if isinstance(item, tuple):
                        t, m = item
                        me.buf["recv"][m["roff"]:m["roff"] + m["rcount"]] = t.numpy()
                me.copies(phase, 2)
                if phase < 3:
                    transforms(me, phase, origin, h)
            ef = me.buf["ef"].reshape(me.ext[2], me.ext[1], me.ext[0], 3)[g:-g, g:-g, g:-g]
            ref = want[f[2]:f[2] + me.nl[2], f[1]:f[1] + me.nl[1], f[0]:f[0] + me.nl[0]]
            assert np.isfinite(ef).all() and np.max(np.abs(ef - ref)) <= 1e-12 * np.abs(want).max()
            me.plan.close()
            layout.close()
        open(os.path.join(out_dir, f"slab_ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_slab_fft_plan_over_gloo(world, tmp_path):
    """the four exchanges of the slab-decomposed solve as real messages between processes"""
    import torch.multiprocessing as mp
    mp.spawn(_slab_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"slab_ok{r}").exists() for r in range(world))
