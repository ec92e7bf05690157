"""ipplb_ctx_set_gather_variant(2): ipplb_gather_cic (3 components) and ipplb_gather_push with 16-byte field loads per x-pair of
stencil nodes (ippl_b200/csrc/push.cuh, gather_point3_vec) -- bit for bit the oracle's gather (the reference's
ParticleAttrib::gather, src/Particle/ParticleAttrib.hpp:193-246 with src/Interpolation/CIC.hpp:47-66) and push.  The kernels
were written after this round's GPU budget was spent and have not run on a GPU yet.  Not collected by name:
tests/test_zz_variants_gpu.py runs this file in its own pytest process behind an xfail mark.  Variant 1 stays the default.

Cases: whole domain and sub-domain meshes, particles on the lower / upper corners, faces and cell centres (tests/
test_gpu_parity._case); ghosted extents even and odd (a stencil row starts on a 16-byte boundary or 8 bytes behind one, both
along one row of nodes when the x extent is odd); all three extents odd, where the upper-corner particle's last row is the
last six doubles of the field."""
import numpy as np
import pytest

import oracle
from test_gpu_parity import _dev
from util import normal_velocities

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import ippl_b200 as ib
    c = ib.Context(0)
    c.set_gather_variant(2)
    yield c
    c.close()


def _mesh_case(seed, ng, first=None, nl=None, n=30000):
    import ippl_b200 as ib
    rng = np.random.default_rng(seed)
    origin, h = (0.25, -1.0, 3.0), (0.5, 0.125, 1.5)
    first, nl = (first, nl) if first is not None else ((0, 0, 0), ng)
    lo = [origin[d] + first[d] * h[d] for d in range(3)]
    R = [lo[d] + rng.uniform(0, nl[d] * h[d], n) for d in range(3)]
    for d in range(3):
        R[d][0] = lo[d]                          # lower corner
        R[d][1] = lo[d] + nl[d] * h[d]           # upper corner: the last node of the ghosted box is part of its stencil
        R[d][2] = lo[d] + 2.5 * h[d]
        R[d][3] = lo[d] + 3.0 * h[d]
    R[0][4], R[1][4], R[2][4] = lo[0] + nl[0] * h[0], lo[1], lo[2] + nl[2] * h[2]
    mo = oracle.Mesh.make(ng, origin, h, first=first, nl=nl)
    mg = ib.Mesh.make(ng, origin, h, first=first, nl=nl)
    return rng, mo, mg, R


MESHES = [((12, 10, 8), None, None), ((12, 10, 8), (6, 0, 4), (6, 10, 4)), ((11, 10, 8), None, None), ((11, 9, 7), None, None),
          ((13, 9, 7), (2, 1, 0), (9, 7, 7))]


@pytest.mark.parametrize("ng,first,nl", MESHES)
def test_gather_variant2_bit_exact(ctx, ng, first, nl):
    rng, mo, mg, R = _mesh_case(31, ng, first, nl)
    n = len(R[0])
    ef = rng.normal(size=mg.cells * 3)
    want = [np.zeros(n) for _ in range(3)]
    oracle.gather_cic(mo, *R, ef, want)
    out = [ctx.zeros(n) for _ in range(3)]
    Rd, efd = [_dev(ctx, r) for r in R], _dev(ctx, ef)
    ctx.gather(mg, *Rd, efd, out)
    for d in range(3):
        assert np.array_equal(out[d].cpu().numpy(), want[d])
    oracle.gather_cic(mo, *R, ef, want, add=True)     # addToAttribute
    ctx.gather(mg, *Rd, efd, out, add=True)
    for d in range(3):
        assert np.array_equal(out[d].cpu().numpy(), want[d])
    # a field that does not start on a 16-byte boundary takes variant 1 (same bits)
    shifted = ctx.zeros(mg.cells * 3 + 1)
    shifted[1:].copy_(efd)
    out2 = [ctx.zeros(n) for _ in range(3)]
    ctx.gather(mg, *Rd, shifted[1:], out2)
    oracle.gather_cic(mo, *R, ef, want)
    for d in range(3):
        assert np.array_equal(out2[d].cpu().numpy(), want[d])


@pytest.mark.parametrize("kind", ["leapfrog", "penning"])
@pytest.mark.parametrize("ng", [(16, 16, 16), (15, 13, 11)])
def test_gather_push_variant2_bit_exact(ctx, kind, ng):
    import ippl_b200 as ib
    n = 40000
    Ld = 20.0
    h = [Ld / 16] * 3
    L = [ng[d] * h[d] for d in range(3)]
    mo = oracle.Mesh.make(ng, (0, 0, 0), h)
    mg = ib.Mesh.make(ng, (0, 0, 0), h)
    rng = np.random.default_rng(41)
    R = [rng.uniform(0, np.nextafter(L[d], 0), n) for d in range(3)]
    for d in range(3):
        R[d][0], R[d][1] = 0.0, np.nextafter(L[d], 0)
    P = normal_velocities(n, seed=43)
    dt = 0.5 * Ld / 2048 if kind == "penning" else 0.05
    ef = rng.normal(size=mg.cells * 3)
    Ro, Po = [r.copy() for r in R], [p.copy() for p in P]
    E = [np.zeros(n) for _ in range(3)]
    oracle.gather_cic(mo, *Ro, ef, E)
    if kind == "penning":
        pp = oracle.penning_params((0, 0, 0), tuple(L), dt)
        oracle.penning_kick(2, pp, Ro, Po, E)
        oracle.penning_kick(1, pp, Ro, Po, E)
        push = ib.penning_push(dt, (0, 0, 0), tuple(L))
    else:
        for _ in range(2):
            for d in range(3):
                oracle.kick(Po[d], E[d], 0.5 * dt)
        push = ib.leapfrog_push(dt)
    for d in range(3):
        oracle.drift(Ro[d], Po[d], dt)
    for d in range(3):
        oracle.periodic_bc(Ro[d], 0.0, ng[d] * h[d] + 0.0)
    parts = ib.Particles.from_host(R, P, ctx.device, q=-1.0)
    ctx.gather_push(mg, push, parts, _dev(ctx, ef))
    for a, b in zip(parts.host(), Ro + Po):
        assert np.array_equal(a, b)


def test_gather_variant_argument_check(ctx):
    import ippl_b200 as ib
    with pytest.raises(ib.IpplbError):
        ctx.set_gather_variant(0)
    ctx.set_gather_variant(2)
