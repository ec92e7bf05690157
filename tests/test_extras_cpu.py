"""CPU checks of the SURVEY 8f rows that have host-only logic: the sampling oracle's Philox against the Random123
known-answer vectors, the rank counts of InverseTransformSampling (C-ABI host function vs the oracle restatement),
and the ORB state machine of the C-ABI against the oracle's binaryRepartition + the invariants of the reference's
unit_tests/PIC/ORB.cpp (every rank keeps a box; the boxes tile the domain)."""
import math

import numpy as np
import pytest

import ippl_b200 as ib
import oracle
from oracle import extras as ox


def test_philox_known_answer_vectors():
    """Random123 kat_vectors, philox4x32 10 rounds"""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        got = ox.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want
    u0, u1 = ox.philox_uniform2(42, np.arange(1000), 0)
    assert u0.min() >= 0.0 and u0.max() < 1.0 and abs(u0.mean() - 0.5) < 0.05 and not np.array_equal(u0, u1)


def _dists():
    L = 4 * math.pi
    kb = 0.21
    return {
        "landau": (ox.Dist([1, 1, 1], [0.05, 0.5] * 3), ib.Dist.make([1, 1, 1], [0.05, 0.5] * 3), [0.0] * 3, [L] * 3),
        "bumpontail": (ox.Dist([0, 0, 1], [0.01, kb] * 3), ib.Dist.make([0, 0, 1], [0.01, kb] * 3), [0.0] * 3,
                       [2 * math.pi / kb] * 3),
        "penning": (ox.Dist([2, 2, 2], [10.0, 3.0, 10.0, 1.0, 10.0, 4.0]),
                    ib.Dist.make([2, 2, 2], [10.0, 3.0, 10.0, 1.0, 10.0, 4.0]), [0.0] * 3, [20.0] * 3),
    }


@pytest.mark.parametrize("name", ["landau", "bumpontail", "penning"])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 8])
def test_sample_counts_match_oracle(name, nranks):
    od, bd, rmin, rmax = _dists()[name]
    ng = (32, 32, 32)
    h = [(rmax[d] - rmin[d]) / ng[d] for d in range(3)]
    boxes = oracle.partition(ng, nranks)
    regs = oracle.regions(ng, boxes, rmin, h)
    for ntotal in (1 << 20, 10_000_000, 12345):
        want_n, want_u = ox.sample_counts(od, rmin, rmax, regs, ntotal)
        got_n, got_u = ib.sample_counts(bd, rmin, rmax, regs, ntotal)
        assert sum(got_n) == ntotal == sum(want_n)
        assert got_n == want_n
        assert np.array_equal(np.asarray(want_u), got_u)   # both sides call libm on the same doubles


def _global_weight(ng, kind, seed=0):
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(*[np.arange(n) for n in ng[::-1]], indexing="ij")
    if kind == "uniform":
        return np.ones(ng[::-1])
    if kind == "blob":   # PenningTrap-like Gaussian blob off centre
        c = [0.3 * ng[0], 0.6 * ng[1], 0.45 * ng[2]]
        s = [0.15 * ng[0], 0.05 * ng[1] + 1, 0.2 * ng[2]]
        return np.exp(-((x - c[0]) / s[0]) ** 2 - ((y - c[1]) / s[1]) ** 2 - ((z - c[2]) / s[2]) ** 2) + 1e-6
    return rng.random(ng[::-1])


def _drive_state_machine(ng, nranks, w):
    orb = ib.Orb(ng, nranks)
    ncuts = 0
    while True:
        nxt = orb.next()
        if nxt is None:
            break
        lo, hi, axis = nxt
        sub = w[lo[2]:hi[2] + 1, lo[1]:hi[1] + 1, lo[0]:hi[0] + 1]
        other = tuple(a for a in range(3) if a != 2 - axis)
        orb.cut(sub.sum(axis=other))
        ncuts += 1
    assert ncuts == nranks - 1
    return orb.finish()


@pytest.mark.parametrize("nranks", [1, 2, 3, 4, 5, 8, 16])
@pytest.mark.parametrize("kind", ["uniform", "blob", "random"])
def test_orb_state_machine_matches_oracle_and_tiles_the_domain(nranks, kind):
    for ng in [(32, 32, 32), (64, 16, 24), (20, 33, 17)]:
        w = _global_weight(ng, kind, seed=nranks)
        boxes, ok = _drive_state_machine(ng, nranks, w)
        want, want_ok = ox.orb_repartition(ng, nranks, w)
        assert ok == want_ok
        assert np.array_equal(boxes, np.asarray(want, dtype=np.int32))
        # unit_tests/PIC/ORB.cpp invariants: one box per rank, disjoint, covering the domain
        assert boxes.shape == (nranks, 6)
        cover = np.zeros(ng[::-1], dtype=np.int32)
        for b in boxes:
            assert (b[3:] >= b[:3]).all()
            cover[b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1] += 1
        assert (cover == 1).all()
        if ok and kind == "blob" and nranks in (2, 4, 8):
            # the point of ORB: weights are balanced far better than the equal-volume partition
            def imbalance(bx):
                loads = [w[b[2]:b[5] + 1, b[1]:b[4] + 1, b[0]:b[3] + 1].sum() for b in bx]
                return max(loads) / (sum(loads) / len(loads))
            assert imbalance(boxes) <= imbalance(oracle.partition(ng, nranks)) + 1e-12
        if ok:   # the new boxes are a valid layout for the product's FieldLayout mirror
            L = ib.Layout(ng, nranks)
            L.set_boxes(boxes)
            assert np.array_equal(L.boxes(), boxes)
            if (boxes[:, 3:] - boxes[:, :3] + 1).min() >= 2:
                for my in range(nranks):
                    assert np.array_equal(L.neighbors(my), oracle.neighbors(ng, boxes, my))
            L.close()


def test_orb_find_median_edge_cases():
    """findMedian's special cases (OrthogonalRecursiveBisection.hpp:185-216) through the C-ABI state machine"""
    for w in ([1, 1, 1, 1], [5, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 5], [1, 2, 3, 4, 5, 6, 7, 8], [0] * 8, [3, 3, 3]):
        ng = (len(w), 2, 2)
        orb = ib.Orb(ng, 2)
        lo, hi, axis = orb.next()
        assert axis == 0
        orb.cut(np.asarray(w, dtype=np.float64))
        boxes, ok = orb.finish()
        m = ox.orb_find_median([float(x) for x in w])
        assert boxes[0][3] == m and boxes[1][0] == m + 1


def test_alpine_oracle_loops_invariants():
    """oracle.extras.AlpineOracle (the PenningTrap / BumponTail managers restated on the CPU): the invariants the
    reference itself enforces or implies -- charge conservation below AlpineManager's 1e-10 abort threshold at every
    scatter, the imposed BumponTail perturbation's E_z energy from linear theory at t = 0, particle count fixed."""
    nr, n = (16, 16, 16), 2_000_000      # shot noise of the fundamental mode: sqrt(2 / n) = 0.001 against delta = 0.01
    kb = 0.21
    L = 2 * math.pi / kb
    d = ox.Dist([0, 0, 1], [0.01, kb] * 3)
    R, _ = ox.sample_positions(d, [0.0, 0.0, d.cdf(0.0, 2)], [L, L, d.cdf(L, 2)], 42, 0, n)
    R = [np.clip(r, 1e-12, L) for r in R]
    P = ox.sample_normal([0.0] * 3, [1 / math.sqrt(2)] * 3, 42, 0, n)
    s = ox.AlpineOracle("bumpontail", nr, R, P, parallel=False)
    s.pre_run()
    assert s.rel_err < 1e-10
    # the imposed mode: rho = rho0 (1 + delta cos(k z)), rho0 = Q / V = -1  ->  E_z = -(delta / k) sin(k z); project the
    # computed field on sin(k z) (the total E_z energy is dominated by shot noise at 49 particles per cell)
    ez = oracle.interior(s.Ef, s.mesh, 3)[..., 2]
    zc = (np.arange(nr[2]) + 0.5) * s.hr[2]
    amp = 2.0 * float(np.mean(ez * np.sin(kb * zc)[:, None, None]))
    assert abs(amp / (-(0.01 / kb)) - 1.0) < 0.35, amp
    assert s.history[0][1] > 0.5 * (0.5 * (0.01 / kb) ** 2 * L ** 3)
    for _ in range(2):
        s.step()
        assert s.rel_err < 1e-10 and len(s.R[0]) == n
        assert all((r >= 0).all() and (r <= L).all() for r in s.R)
    # PenningTrap: blob stays inside the box; Kick1 + Kick2 with E = 0 rotate P about z without changing |P_xy| beyond
    # the quadrupole's work, so compare against the closed form of one rotation step on a particle at the trap centre
    d = ox.Dist([2, 2, 2], [10.0, 3.0, 10.0, 1.0, 10.0, 4.0])
    ub = [d.cdf(0.0, k) for k in range(3)] + [d.cdf(20.0, k) for k in range(3)]
    n = 200_000
    R, _ = ox.sample_positions(d, ub[:3], ub[3:], 7, 0, n)
    R = [np.clip(r, 1e-12, 20.0) for r in R]
    P = ox.sample_normal([0.0] * 3, [1.0] * 3, 7, 0, n)
    s = ox.AlpineOracle("penning", nr, R, P, parallel=False)
    s.pre_run()
    assert s.rel_err < 1e-10 and s.history[0][1] > 0 and abs(s.history[0][2] / (1.5 * n) - 1) < 0.02
    s.step()
    assert s.rel_err < 1e-10
    pp = s.pp
    Rc = [np.array([10.0]), np.array([10.0]), np.array([10.0])]      # trap centre: E_ext = 0
    Pc = [np.array([1.0]), np.array([0.0]), np.array([0.5])]
    E0 = [np.zeros(1) for _ in range(3)]
    oracle.penning_kick(1, pp, Rc, Pc, E0)
    oracle.penning_kick(2, pp, Rc, Pc, E0)
    # Boris-like rotation: |P_xy| is conserved to O((alpha B)^3) per step (alpha B = 0.012), P_z untouched
    assert abs(math.hypot(Pc[0][0], Pc[1][0]) - 1.0) < 1e-5 and Pc[2][0] == 0.5
    ang = math.atan2(Pc[1][0], Pc[0][0])
    assert abs(ang - 2 * math.atan(-pp.alpha * pp.Bext)) < 1e-5    # rotation by 2 atan(|alpha| B) + O((alpha B)^3) per step
