"""The invariants the reference's own unit tests hold for this path, checked on
the oracle: unit_tests/PIC/PIC.cpp, Particle/ParticleBC.cpp, Field/Halo.cpp, Particle/ParticleUpdate.cpp."""
import numpy as np
import pytest

import oracle


def _mesh(ng=(8, 8, 8), L=1.0):
    return oracle.Mesh.make(ng, (0, 0, 0), tuple(L / n for n in ng))


def test_pic_scatter_conserves_charge_and_gather_constant():
    # unit_tests/PIC/PIC.cpp:109-143: 32 particles, mt19937_64 seed 42; sum rho == N*q (10 eps rel),
    # gather of constant 1 gives exactly N
    rng = np.random.default_rng(42)
    m = _mesh()
    n = 32
    x, y, z = [rng.uniform(0, 1, n) for _ in range(3)]
    rho = oracle.field_zeros(m)
    oracle.scatter_cic(m, x, y, z, 0.5, rho)
    oracle.halo_periodic(rho, m.ext, 1, 1, (1, 1, 1), "accumulate")
    tot = oracle.field_sum(rho, m.ext)
    assert abs(tot - n * 0.5) / (n * 0.5) < 10 * np.finfo(float).eps
    ef = np.ones(rho.size)
    out = [np.zeros(n)]
    oracle.gather_cic(m, x, y, z, ef, out)
    assert float(np.sum(out[0])) == float(n)


def test_adjointness_scatter_gather():
    # KernelGatherScatterTest.cpp: <scatter q, f> == <q, gather f>
    rng = np.random.default_rng(3)
    m = _mesh((6, 7, 5))
    n = 500
    x, y, z = [rng.uniform(0, 1, n) for _ in range(3)]
    q = rng.normal(size=n)
    f = rng.normal(size=m.ext[0] * m.ext[1] * m.ext[2])
    rho = oracle.field_zeros(m)
    oracle.scatter_cic(m, x, y, z, q, rho)
    g = [np.zeros(n)]
    oracle.gather_cic(m, x, y, z, f, g)
    assert abs(np.dot(rho, f) - np.dot(q, g[0])) < 1e-10


def test_periodic_bc_closed_form():
    # unit_tests/Particle/ParticleBC.cpp:79-218: shift by one period lands inside
    lo, hi = 0.0, 1.0
    # (the formula x - L*(int)((x-mid)*2/L) is only a wrap for overshoots below half a period)
    x = np.array([-0.25, 1.25, 0.5, -0.499, 1.499])
    want = np.array([0.75, 0.25, 0.5, 0.501, 0.499])
    oracle.periodic_bc(x, lo, hi)
    assert np.allclose(x, want, atol=1e-15)


def test_halo_fill_then_accumulate_identity():
    # unit_tests/Field/Halo.cpp: fill of a constant field leaves ghosts == constant; accumulate of a
    # field of ones adds the ghost layers back (interior boundary cells get 1 + number of images)
    m = _mesh((4, 5, 6))
    f = oracle.field_zeros(m)
    oracle.interior(f, m)[...] = 1.0
    oracle.halo_periodic(f, m.ext, 1, 1, (1, 1, 1), "fill")
    assert np.all(f == 1.0)
    oracle.halo_periodic(f, m.ext, 1, 1, (1, 1, 1), "accumulate")
    a = oracle.interior(f, m)
    assert a[2, 2, 2] == 1.0 and a[0, 2, 2] == 2.0 and a[0, 0, 2] == 4.0 and a[0, 0, 0] == 8.0


@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
def test_update_conserves_and_places(nranks):
    # unit_tests/Particle/ParticleUpdate.cpp:222-815: counts, charge, tags conserved; 0 misplaced
    rng = np.random.default_rng(11)
    ng, origin, h = (16, 12, 8), (0.0, 0.0, 0.0), (0.5, 0.25, 1.0)
    L = [ng[d] * h[d] for d in range(3)]
    boxes = oracle.partition(ng, nranks)
    regs = oracle.regions(ng, boxes, origin, h)
    parts, tag0 = [], 0
    for r in range(nranks):
        n = 1000 + 100 * r
        p = {k: rng.uniform(-0.5, 1.5, n) * L[d] for d, k in enumerate("xyz")}  # anywhere, incl. outside
        p["tag"] = np.arange(tag0, tag0 + n, dtype=np.float64)
        tag0 += n
        parts.append(p)
    out = oracle.update(ng, boxes, origin, h, parts)
    assert sum(len(p["x"]) for p in out) == tag0
    tags = np.sort(np.concatenate([p["tag"] for p in out]))
    assert np.array_equal(tags, np.arange(tag0))
    for r, p in enumerate(out):
        d = oracle.locate(regs, r, p["x"], p["y"], p["z"])
        assert np.all(d == r)  # zero misplaced particles
        for dd, k in enumerate("xyz"):
            assert np.all(p[k] >= regs[r, dd]) and np.all(p[k] <= regs[r, 3 + dd])


def test_update_all_on_one_rank_and_empty_ranks():
    ng, origin, h = (8, 8, 8), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    boxes = oracle.partition(ng, 4)
    rng = np.random.default_rng(5)
    parts = [{k: rng.uniform(0, 8, 2000) for k in "xyz"}] + [{k: np.zeros(0) for k in "xyz"} for _ in range(3)]
    out = oracle.update(ng, boxes, origin, h, parts)
    assert sum(len(p["x"]) for p in out) == 2000
    assert all(len(p["x"]) > 0 for p in out)
