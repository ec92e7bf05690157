"""Shared helpers for the tests: seeded initial conditions of the alpine mini-apps."""
import numpy as np


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
