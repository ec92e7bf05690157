"""Pins the sampling restatement (oracle/extras.py) and the C-ABI's host-side count function against the REAL reference
headers of the sampling path: live through oracle/_ref/libippl_refshim_random.so where /root/reference exists, and
everywhere against tests/golden/ref_random.npz (the committed outputs of those headers, tests/golden/
make_golden_random.py).  Random numbers are replayed, so what is compared is the reference's arithmetic: distribution
functions (the managers' own CustomDistributionFunctions structs, cut out of demos/alpine/*Manager.h at build time, and
NormalDistribution), getFullPdf, NewtonRaphson::solve, the InverseTransformSampling constructor (rank counts + CDF bounds),
generate() / fill_random, randn."""
import math
import os

import numpy as np
import pytest

import ippl_b200 as ib
import oracle
from oracle import extras as ox
from oracle import refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("landau", "penning", "bumpontail")
SHIM_KIND = {"landau": 1, "penning": 2, "bumpontail": 3}
NG = (32, 32, 32)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_random.npz"))


def _dist(gold, name):
    kinds, par = [int(k) for k in gold[f"{name}_kinds"]], [float(p) for p in gold[f"{name}_par"]]
    return ox.Dist(kinds, par), ib.Dist.make(kinds, par), par


@pytest.mark.parametrize("name", NAMES)
def test_distribution_functions_vs_reference(gold, name):
    od, _, par = _dist(gold, name)
    xs, us, ev = gold[f"{name}_xs"], gold[f"{name}_us"], gold[f"{name}_eval"]
    for d in range(3):
        # scalars go through libm on both sides: bit-exact
        assert np.array_equal(np.array([od.cdf(float(x), d) for x in xs]), ev[0][d])
        assert np.allclose(od.pdf(xs, d), ev[1][d], rtol=4e-16, atol=0)      # numpy's vector cos / exp: last ulp
        assert np.array_equal(np.array([float(od.estimate(float(u), d)) for u in xs]), ev[2][d])
        assert np.array_equal(np.array([od.cdf(float(x), d) - float(u) for x, u in zip(xs, us)]), ev[3][d])
        assert np.allclose(od.pdf(xs, d), ev[4][d], rtol=4e-16, atol=0)
    fp = np.array([float(od.pdf(np.array([a]), 0)[0] * od.pdf(np.array([b]), 1)[0] * od.pdf(np.array([c]), 2)[0])
                   for a, b, c in zip(xs, xs[::-1], np.roll(xs, 7))])
    assert np.allclose(fp, gold[f"{name}_fullpdf"], rtol=1e-15, atol=0)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_rank_counts_and_bounds_vs_reference(gold, name, nranks):
    """InverseTransformSampling's constructor: the restatement AND the product's ipplb_sample_counts, exactly"""
    od, bd, par = _dist(gold, name)
    rmin = [0.0] * 3
    rmax = {"landau": [4 * math.pi] * 3, "penning": [20.0] * 3, "bumpontail": [2 * math.pi / 0.21] * 3}[name]
    h = [(rmax[d] - rmin[d]) / NG[d] for d in range(3)]
    regs = oracle.regions(NG, oracle.partition(NG, nranks), rmin, h)
    for ntotal in (1 << 20, 10_000_000, 12345):
        want_n, want_u = gold[f"{name}_counts_{nranks}_{ntotal}"], gold[f"{name}_ubounds_{nranks}_{ntotal}"]
        on, ou = ox.sample_counts(od, rmin, rmax, regs, ntotal)
        bn, bu = ib.sample_counts(bd, rmin, rmax, regs, ntotal)
        assert on == list(want_n) == bn
        assert np.array_equal(np.asarray(ou), want_u) and np.array_equal(bu, want_u)
        if refshim.random_available():
            ln, lu, _ = refshim.rand_sampling(SHIM_KIND[name], par, rmin, rmax, regs, ntotal)
            assert ln == bn and np.array_equal(lu, bu)


@pytest.mark.parametrize("name", NAMES)
def test_generate_vs_reference(gold, name):
    """generate() / fill_random / NewtonRaphson with the reference's uniforms replayed.  Tolerance: the restatement
    evaluates sin / cos / erf / exp with numpy's vector kernels, the reference with libm: a last-ulp difference in the
    residual can move Newton's stopping iteration, so the bound is the stopping tolerance over the pdf (1e-12 / pdf),
    stated as 1e-10 absolute; typical agreement is 1e-15."""
    od, _, par = _dist(gold, name)
    regs, u01, want = gold[f"{name}_gen_regs"], gold[f"{name}_gen_u01"], gold[f"{name}_gen_x"]
    rmin, rmax = gold[f"{name}_gen_rmin"], gold[f"{name}_gen_rmax"]
    _, ub = ox.sample_counts(od, rmin, rmax, regs, 20000)
    got = ox.newton_positions(od, ub[3][:3], ub[3][3:], [u01[d] for d in range(3)])
    for d in range(3):
        assert np.max(np.abs(got[d] - want[d])) <= 1e-10
        assert np.median(np.abs(got[d] - want[d])) <= 1e-14
    if refshim.random_available():
        _, _, live = refshim.rand_sampling(SHIM_KIND[name], par, rmin, rmax, regs, 20000, gen_rank=3, u01=u01)
        assert np.array_equal(live, want)      # the fixture is what the headers produce today
        for k in range(0, 40, 7):              # NewtonRaphson::solve alone, scalar path: bit-exact against libm
            u = float(ub[3][0] + (ub[3][3] - ub[3][0]) * u01[0][k])
            x0 = float(od.estimate(u, 0))
            want1 = refshim.rand_newton(SHIM_KIND[name], par, 0, x0, u)
            x, it = x0, 0
            while it < 20 and abs(od.cdf(x, 0) - u) > 1e-12:
                x = x - ((od.cdf(x, 0) - u) / float(od.pdf(np.array([x]), 0)[0]))
                it += 1
            assert abs(x - want1) <= 1e-13


def test_randn_vs_reference(gold):
    g, mu, sd = gold["randn_g"], gold["randn_mu"], gold["randn_sd"]
    want = gold["randn_v"]
    got = np.stack([mu[d] + sd[d] * g[:, d] for d in range(3)], axis=1)
    assert np.array_equal(got, want)
    if refshim.random_available():
        assert np.array_equal(refshim.rand_randn(mu, sd, g), want)
