"""Multi-rank parity on ONE GPU: every rank of a 2 / 4 / 8-rank job is a context of this process (ipplb_loop_*), device
copies are the transport, and the kernels, tables and host logic are the ones the NCCL path runs.  Held to the oracle's
all-ranks simulation of the reference:
  * HaloCells::exchangeBoundaries + applyPeriodicSerialDim (src/Field/HaloCells.hpp:109-336): fillHalo bit-exact,
    accumulateHalo <= 1e-12 (atomic order), scalar and vector fields;
  * ParticleSpatialLayout::update (src/Particle/ParticleSpatialLayout.hpp:115-464, ParticleBase.hpp:175-393): ownership
    and counts exact (incl. particles on region / domain faces), every rank's particles the same multiset bit for bit;
  * the fused step with the ownership test + the peer-memory migration + both halo exchanges, several steps: counts
    exact, particles bit-exact, rho <= 1e-12;
on the default FieldLayout decomposition and on an ORB-style layout with unequal boxes."""
import numpy as np
import pytest

import oracle
from util import normal_velocities, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12


def canon(cols):
    a = np.stack(cols, axis=1)
    return a[np.lexsort(a.T[::-1])]


def orb_like_boxes(ng, world, shifts=((0, 2), (1, -1))):
    """unequal boxes that tile the domain (what an ORB repartition produces): cuts moved off the middle"""
    boxes = oracle.partition(ng, world).copy()
    # move every interior cut along x by +2 cells and along y by -1 cell where the layout has such cuts
    for d, shift in shifts:
        cuts = sorted(set(int(b[d]) for b in boxes) - {0})
        for c in cuts:
            for b in boxes:
                if b[d] == c:
                    b[d] = c + shift
                if b[3 + d] == c - 1:
                    b[3 + d] = c - 1 + shift
    return boxes


class Job:
    """`world` in-process ranks on the current device, bound to a layout"""

    def __init__(self, world, ng, boxes=None):
        import ippl_b200 as ib
        self.ib, self.world, self.ng = ib, world, ng
        self.h = [4 * np.pi / 16] * 3
        self.origin = (0.0, 0.0, 0.0)
        self.ctxs = [ib.Context(0) for _ in range(world)]
        self.loop = ib.Loop(self.ctxs)
        self.layout = ib.Layout(ng, world)
        if boxes is not None:
            self.layout.set_boxes(boxes)
        self.boxes = self.layout.boxes()
        for c in self.ctxs:
            c.set_layout(self.layout, self.origin, self.h)
        self.regs = self.layout.regions(self.origin, self.h)
        assert np.array_equal(self.regs, oracle.regions(ng, self.boxes, self.origin, self.h))
        self.meshes = [self.layout.mesh(r, self.origin, self.h) for r in range(world)]
        self.meshes_o = [oracle.Mesh.make(ng, self.origin, self.h, first=tuple(self.boxes[r, :3]),
                                          nl=tuple(self.boxes[r, 3:] - self.boxes[r, :3] + 1)) for r in range(world)]
        self.Lg = [ng[d] * self.h[d] for d in range(3)]

    def close(self):
        self.loop.close()
        self.layout.close()
        for c in self.ctxs:
            c.close()


LAYOUTS = [(2, "default"), (4, "default"), (8, "default"), (4, "orb"), (8, "orb")]


def make_job(world, kind, ng=(24, 16, 16)):
    if kind == "orb_z":   # also the z cut off the middle (tests/loop_orb_z_check.py: the 127 / 129 split of the PenningTrap blob)
        return Job(world, ng, orb_like_boxes(ng, world, shifts=((0, 2), (1, -1), (2, -1))))
    return Job(world, ng, orb_like_boxes(ng, world) if kind == "orb" else None)


@pytest.mark.parametrize("world,kind", LAYOUTS)
@pytest.mark.parametrize("ncomp", [1, 3])
def test_loop_halo_exchange_vs_oracle(world, kind, ncomp):
    import torch
    job = make_job(world, kind)
    try:
        dev = job.ctxs[0].device
        for mode in ("fill", "accumulate"):
            f_o = [np.random.default_rng(10 * r + ncomp).normal(size=m.ext[0] * m.ext[1] * m.ext[2] * ncomp) for r, m in enumerate(job.meshes_o)]
            f_g = [torch.from_numpy(a.copy()).to(dev) for a in f_o]
            oracle.halo_full(job.ng, job.boxes, f_o, ncomp, mode)
            job.loop.halo_exchange(f_g, ncomp, mode)
            for r in range(world):
                got = f_g[r].cpu().numpy()
                if mode == "fill":
                    assert np.array_equal(got, f_o[r]), f"fillHalo rank {r}"
                else:
                    assert rel_l2(got, f_o[r]) <= TOL, f"accumulateHalo rank {r}"
    finally:
        job.close()


def _global_particles(job, n, seed):
    rng = np.random.default_rng(seed)
    Lg, regs = job.Lg, job.regs
    R = [rng.uniform(0, Lg[d], n) for d in range(3)]
    R[0] = np.clip(np.mod(rng.normal(0.3 * Lg[0], 0.2 * Lg[0], n), Lg[0]), 1e-9, Lg[0])
    # on region / domain faces (the strict and the inclusive test of positionInRegion)
    R[0][:4] = [0.0, Lg[0], regs[0][3], np.nextafter(regs[0][3], np.inf)]
    R[1][4:6] = [regs[-1][1], np.nextafter(regs[-1][1], -np.inf)]
    P = [2.0 * p for p in normal_velocities(n, seed=seed + 1)]
    return R, P


@pytest.mark.parametrize("world,kind", LAYOUTS)
def test_loop_update_vs_oracle(world, kind):
    """ipplb_update's phases (locate_kernel, count exchange, pack_leavers_kernel, transport, unpack_arrivals_kernel, hole
    filling) through the in-process transport == the oracle's ParticleSpatialLayout::update"""
    job = make_job(world, kind)
    ib = job.ib
    try:
        dev = job.ctxs[0].device
        n = 40000
        R, P = _global_particles(job, n, 77)
        own = oracle.locate(job.regs, 0, R[0], R[1], R[2])
        names = ("x", "y", "z", "px", "py", "pz")
        parts_o = [{k: a[own == r].copy() for k, a in zip(names, R + P)} for r in range(world)]
        # move them (some across several ranks), no BC yet: update applies it
        for p in parts_o:
            for kx, kp in zip("xyz", ("px", "py", "pz")):
                p[kx] += 0.9 * p[kp]
        cap = n
        parts_g = []
        lo, hi = [0.0] * 3, job.Lg
        for r in range(world):
            p = ib.Particles.from_host([parts_o[r][k] for k in "xyz"], [parts_o[r][k] for k in ("px", "py", "pz")], dev,
                                       capacity=cap)
            job.ctxs[r].apply_periodic_bc(p.arr["x"], p.arr["y"], p.arr["z"], lo, hi, n=p.n)   # ParticleLayout::applyBC
            parts_g.append(p)
        wrapped = []
        for p in parts_o:
            w = {k: p[k].copy() for k in "xyz"}
            for d, k in enumerate("xyz"):
                oracle.periodic_bc(w[k], lo[d], hi[d])
            wrapped.append(w)
        dests = [oracle.locate(job.regs, r, w["x"], w["y"], w["z"]) for r, w in enumerate(wrapped)]
        sent_o = [[int((dests[r] == t).sum()) if t != r else 0 for t in range(world)] for r in range(world)]
        parts_o = oracle.update(job.ng, job.boxes, job.origin, job.h, parts_o)
        sent, recv = job.loop.update(parts_g)
        assert sent == sent_o
        assert recv == [[sent_o[t][r] for t in range(world)] for r in range(world)]
        assert sum(p.n for p in parts_g) == n
        for r in range(world):
            want = [parts_o[r][k] for k in names]
            assert parts_g[r].n == len(want[0])
            assert np.array_equal(canon(parts_g[r].host()), canon(want)), f"rank {r}: particles differ from the oracle"
    finally:
        job.close()


@pytest.mark.parametrize("world,kind", LAYOUTS)
def test_loop_fused_step_and_peer_migration_vs_oracle(world, kind, push_kind="leapfrog"):
    """The multi-GPU step of bench.py / the facade (fillHalo(E), fused step with ownership test, peer-memory migration,
    accumulateHalo(rho)) on in-process ranks against the oracle's all-ranks step, four steps.  (push_kind "penning": the
    PenningTrap kicks instead of the leapfrog ones, tests/loop_orb_z_check.py.)"""
    import torch
    job = make_job(world, kind)
    ib = job.ib
    try:
        dev = job.ctxs[0].device
        ng, h, origin, boxes, regs = job.ng, job.h, job.origin, job.boxes, job.regs
        n = 60000
        R, P = _global_particles(job, n, 2024)
        dt, q = 0.5 * h[0], -0.01
        names = ("x", "y", "z", "px", "py", "pz")
        own = oracle.locate(regs, 0, R[0], R[1], R[2])
        parts_o = [{k: a[own == r].copy() for k, a in zip(names, R + P)} for r in range(world)]
        ef_o = [0.1 * np.random.default_rng(100 + r).normal(size=m.ext[0] * m.ext[1] * m.ext[2] * 3) for r, m in enumerate(job.meshes_o)]
        cap = 4 * n // world + 4096
        cur, nxt, bins, ef, rho = [], [], [], [], []
        for r in range(world):
            mine = parts_o[r]
            src = ib.Particles.from_host([mine[k] for k in "xyz"], [mine[k] for k in ("px", "py", "pz")], dev, q=q)
            c, x = ib.Particles(cap, dev, q=q), ib.Particles(cap, dev, q=q)
            b = ib.Bins(job.ctxs[r], job.meshes[r], cap)
            b.build(src, c)
            cur.append(c); nxt.append(x); bins.append(b)
            ef.append(torch.from_numpy(ef_o[r].copy()).to(dev))
            rho.append(job.ctxs[r].field(job.meshes[r]))
        job.loop.migrate_connect(max(n // 2, 1024))
        push = ib.leapfrog_push(dt) if push_kind == "leapfrog" else ib.penning_push(dt, origin, tuple(job.Lg))
        pp = oracle.penning_params(origin, tuple(job.Lg), dt) if push_kind == "penning" else None
        for it in range(4):
            # ---- oracle, all ranks: fillHalo(E); gather; kick, kick, drift; update (BC + migrate); scatter; accumulateHalo
            oracle.halo_full(ng, boxes, ef_o, 3, "fill")
            for r in range(world):
                p = parts_o[r]
                nn = len(p["x"])
                E = [np.zeros(nn) for _ in range(3)]
                oracle.gather_cic(job.meshes_o[r], p["x"], p["y"], p["z"], ef_o[r], E)
                if pp is not None:      # PenningTrapManager.h:313-333 (closing kick of the last step), :256-272
                    Rl, Pl = [p[k] for k in "xyz"], [p[k] for k in ("px", "py", "pz")]
                    oracle.penning_kick(2, pp, Rl, Pl, E)
                    oracle.penning_kick(1, pp, Rl, Pl, E)
                else:
                    for d, k in enumerate(("px", "py", "pz")):
                        oracle.kick(p[k], E[d], 0.5 * dt)
                        oracle.kick(p[k], E[d], 0.5 * dt)
                for kx, kp in zip("xyz", ("px", "py", "pz")):
                    oracle.drift(p[kx], p[kp], dt)
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
            rho_o = [oracle.field_zeros(m) for m in job.meshes_o]
            for r in range(world):
                p = parts_o[r]
                oracle.scatter_cic(job.meshes_o[r], p["x"], p["y"], p["z"], q, rho_o[r])
            oracle.halo_full(ng, boxes, rho_o, 1, "accumulate")
            # ---- CUDA, all ranks in this process
            job.loop.halo_exchange(ef, 3, "fill")
            for r in range(world):
                assert np.array_equal(ef[r].cpu().numpy(), ef_o[r]), "fillHalo(E) differs"
                rho[r].zero_()
                bins[r].step(push, cur[r], nxt[r], ef[r], rho[r], region=list(regs[r]))   # exit_buf None: peer mode
            job.loop.bins_migrate(bins, cur, rho)
            job.loop.halo_exchange(rho, 1, "accumulate")
            total = 0
            for r in range(world):
                sent, recv = job.ctxs[r].migrate_counts()
                assert sent == sent_o[r], f"rank {r} step {it}: sent {sent} != oracle {sent_o[r]}"
                assert recv == [sent_o[t][r] for t in range(world)], f"rank {r} step {it}: recv {recv}"
                nloc, ntail, nexit, flags = bins[r].status()
                assert (flags & 7) == 0 and nloc == len(parts_o[r]["x"]), (r, it, nloc, len(parts_o[r]["x"]), flags)
                total += nloc
                out = ib.Particles(max(nloc, 1), dev)
                assert bins[r].compact(cur[r], out) == nloc
                want = [parts_o[r][k] for k in names]
                assert np.array_equal(canon(out.host()), canon(want)), f"rank {r} step {it}: particles differ from the oracle"
                assert rel_l2(rho[r].cpu().numpy(), rho_o[r]) <= TOL, f"rank {r} step {it}: rho"
            assert total == n
        for b in bins:
            b.close()
    finally:
        job.close()
