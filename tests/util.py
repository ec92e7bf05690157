"""Shared helpers for the tests: seeded initial conditions of the alpine mini-apps; the time budget of the tests that run
code for the first time."""
import os
import time

import numpy as np

# The round-end `pytest -m gpu` run has a fixed time limit for the WHOLE suite.  Tests that execute code which has never
# met a GPU (xfail-marked, each in its own process) must not be able to use it up between them if several of them hang:
# they share one deadline, counted from the start of the pytest session (tests/conftest.py sets IPPLB_TEST_SESSION_T0).
FIRST_RUN_DEADLINE_S = float(os.environ.get("IPPLB_FIRST_RUN_DEADLINE_S", "840"))


def first_run_timeout(wanted_s):
    """timeout for a first-execution subprocess: `wanted_s`, cut to what is left before the shared deadline; gives up
    (pytest.xfail) when less than 20 s are left"""
    import pytest
    t0 = float(os.environ.get("IPPLB_TEST_SESSION_T0", time.time()))
    left = FIRST_RUN_DEADLINE_S - (time.time() - t0)
    if left < 20.0:
        pytest.xfail(f"the suite's time budget for first executions is spent ({FIRST_RUN_DEADLINE_S:.0f} s from the session start)")
    return min(float(wanted_s), left)


def landau_positions(n, L, alpha=0.05, kw=0.5, seed=42):
    """Inverse-transform sample of f(x) = 1 + alpha*cos(kw*x) on [0, L) per dimension (the
    distribution of demos/alpine/LandauDampingManager.h:216-230; Newton on the CDF like
    src/Random/InverseTransformSampling.h:172-256, atol 1e-12, <= 20 iterations).  The RNG stream is
    our own (numpy PCG64): the reference's Kokkos pool is backend dependent."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(3):
        u = rng.random(n)
        target = u * (L + alpha / kw * np.sin(kw * L))
        x = u * L
        for _ in range(20):
            f = x + alpha / kw * np.sin(kw * x) - target
            x = x - f / (1.0 + alpha * np.cos(kw * x))
            if np.max(np.abs(f)) < 1e-12:
                break
        x = np.clip(x, 0.0, np.nextafter(L, 0.0))
        out.append(np.ascontiguousarray(x))
    return out


def normal_velocities(n, seed=43, mu=(0, 0, 0), sd=(1, 1, 1)):
    rng = np.random.default_rng(seed)
    return [np.ascontiguousarray(rng.normal(mu[d], sd[d], n)) for d in range(3)]


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
