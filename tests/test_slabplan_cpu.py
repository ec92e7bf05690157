"""The host-side plan of the slab-decomposed FFT Poisson solve (ippl_b200/csrc/slabplan.cpp, C-ABI ipplb_slabplan_*),
executed here with numpy for ALL ranks of a job: the same copy tables and messages the device executor
(ippl_b200/csrc/fftdist.cu) runs, with numpy FFTs in place of cuFFT.  Held to the oracle's whole-domain solve
(FFTPeriodicPoissonSolver::solve, GRAD output): <= 1e-12 relative.  Also checked: every message has its mirror image on the
peer, the buffers the plan declares are large enough, the boxes' interiors are fully written and the ghost layers are
left alone.  Runs without a GPU."""
import numpy as np
import pytest

import ippl_b200 as ib
import oracle


def strided(buf, off, strides, n, elem):
    """view of a sub-box inside a flat buffer (elements of `elem` doubles; complex buffers are complex arrays here)"""
    item = buf.itemsize
    return np.lib.stride_tricks.as_strided(buf[off:], shape=(n[2], n[1], n[0]), strides=(strides[2] * item, strides[1] * item, strides[0] * item))


class Rank:
    def __init__(self, layout, r, origin, h):
        self.plan = ib.SlabPlan(layout, r)
        p = self.plan
        b = layout.boxes()[r]
        g = p.nghost
        self.nl = tuple(int(b[3 + d] - b[d] + 1) for d in range(3))
        self.ext = tuple(n + 2 * g for n in self.nl)
        self.first = tuple(int(x) for x in b[:3])
        cells = self.ext[0] * self.ext[1] * self.ext[2]
        self.buf = {
            "rho": np.full(cells, np.nan), "ef": np.full(3 * cells, np.nan),
            "real": np.full(max(p.size["real"], 1), np.nan),
            "spec2d": np.full(max(p.size["spec2d"] // 2, 1), np.nan + 0j),
            "specz": np.full(max(p.size["specz"] // 2, 1), np.nan + 0j),
            "send": np.full(max(p.size["send"], 2), np.nan), "recv": np.full(max(p.size["recv"], 2), np.nan),
        }

    def view(self, name, elem):
        b = self.buf[name]
        if elem == 2 and b.dtype != np.complex128:   # send / recv hold doubles; complex phases see them as complex pairs
            return b.view(np.complex128)
        return b

    def copies(self, phase, which):
        for c in self.plan.rows(phase, which):
            src, dst = self.view(c["src"], c["elem"]), self.view(c["dst"], c["elem"])
            n = c["n"]
            last_s = c["src_off"] + sum((n[a] - 1) * c["ss"][a] for a in range(3))
            last_d = c["dst_off"] + sum((n[a] - 1) * c["ds"][a] for a in range(3))
            assert 0 <= c["src_off"] and last_s < len(src), ("source out of range", phase, which, c, len(src))
            assert 0 <= c["dst_off"] and last_d < len(dst), ("destination out of range", phase, which, c, len(dst))
            strided(dst, c["dst_off"], c["ds"], n, c["elem"])[...] = strided(src, c["src_off"], c["ss"], n, c["elem"])


def exchange(ranks, phase):
    msgs = [{m["peer"]: m for m in rk.plan.rows(phase, 1)} for rk in ranks]
    for a, rk in enumerate(ranks):
        for peer, m in msgs[a].items():
            mirror = msgs[peer].get(a)
            assert mirror is not None and mirror["rcount"] == m["scount"] and mirror["scount"] == m["rcount"], (phase, a, peer, m, mirror)
            assert m["soff"] + m["scount"] <= len(rk.buf["send"]) and m["roff"] + m["rcount"] <= len(rk.buf["recv"])
    for a, rk in enumerate(ranks):
        for peer, m in msgs[a].items():
            if m["scount"]:
                r = msgs[peer][a]
                ranks[peer].buf["recv"][r["roff"]:r["roff"] + r["rcount"]] = rk.buf["send"][m["soff"]:m["soff"] + m["scount"]]


def transforms(rk, step, origin, h):
    p = rk.plan
    nx, ny, nz = p.ng
    nxh, nzl, nyl = p.nxh, p.ze - p.zs, p.ye - p.ys
    S2, SZ, SR = nzl * ny * nxh, nz * nyl * nxh, nzl * ny * nx
    if step == 0 and nzl:
        real = rk.buf["real"][:SR].reshape(nzl, ny, nx)
        rk.buf["spec2d"][:S2] = np.fft.rfft2(real, axes=(1, 2)).ravel()
    elif step == 1 and nyl:
        sz = rk.buf["specz"]
        rh = np.fft.fft(sz[:SZ].reshape(nz, nyl, nxh), axis=0) / (nx * ny * nz)
        mult = oracle.poisson_kspace_multipliers((nx, ny, nz), origin, h)     # [nz][ny][nx] broadcastable pieces
        for c in range(3):
            M = np.broadcast_to(mult[c], (nz, ny, nx))[:, p.ys:p.ye, :nxh]
            # unnormalised inverse, like cuFFT
            sz[(1 + c) * SZ:(2 + c) * SZ] = (np.fft.ifft(rh * M, axis=0) * nz).ravel()
    elif step == 2 and nzl:
        for c in range(3):
            spec = rk.buf["spec2d"][c * S2:(c + 1) * S2].reshape(nzl, ny, nxh)
            rk.buf["real"][c * SR:(c + 1) * SR] = (np.fft.irfft2(spec, s=(ny, nx), axes=(1, 2)) * (nx * ny)).ravel()


def orb_like(ng, world):
    boxes = oracle.partition(ng, world).copy()
    for d, shift in ((0, 1), (1, -1), (2, 3)):
        cuts = sorted(set(int(b[d]) for b in boxes) - {0})
        for c in cuts:
            for b in boxes:
                if b[d] == c:
                    b[d] = c + shift
                if b[3 + d] == c - 1:
                    b[3 + d] = c - 1 + shift
    return boxes


CASES = [((16, 12, 10), 1, "default"), ((16, 12, 10), 2, "default"), ((16, 12, 10), 3, "default"), ((12, 16, 20), 4, "default"),
         ((16, 16, 16), 8, "default"), ((24, 16, 16), 8, "orb"), ((16, 12, 10), 4, "orb"), ((10, 6, 5), 8, "default"),
         ((9, 7, 11), 5, "default")]


@pytest.mark.parametrize("ng,world,kind", CASES)
def test_slab_plan_executed_with_numpy_matches_whole_domain_solve(ng, world, kind):
    origin, h = (0.0, 0.5, -1.0), (0.3, 0.25, 0.4)
    layout = ib.Layout(ng, world)
    if kind == "orb":
        layout.set_boxes(orb_like(ng, world))
    boxes = layout.boxes()
    rng = np.random.default_rng(7)
    rho_g = rng.normal(size=(ng[2], ng[1], ng[0]))
    rho_g -= rho_g.mean()
    # the reference transforms real-to-complex and back (heFFTe r2c / c2r, src/FFT/FFT.hpp:118-193): the half spectrum in x.
    # For even sizes this equals oracle.poisson_grad (full complex transforms); for odd sizes the reference's "notMid" rule
    # zeroes index N/2 but not its mirror image, the spectrum is not Hermitian, and only the half-spectrum form is what a
    # r2c / c2r solver -- the reference's and cuFFT's -- computes.
    N = ng[0] * ng[1] * ng[2]
    nxh = ng[0] // 2 + 1
    rhat = np.fft.rfftn(rho_g) / N
    want = np.stack([np.fft.irfftn(rhat * np.broadcast_to(M, rho_g.shape)[:, :, :nxh], s=rho_g.shape, axes=(0, 1, 2)) * N
                     for M in oracle.poisson_kspace_multipliers(ng, origin, h)], axis=-1)   # [nz][ny][nx][3]
    if all(n % 2 == 0 for n in ng):
        assert np.max(np.abs(want - oracle.poisson_grad(rho_g, origin, h))) <= 1e-12 * np.abs(want).max()
    ranks = [Rank(layout, r, origin, h) for r in range(world)]
    g = ranks[0].plan.nghost
    for rk in ranks:   # interior <- my box of rho; ghosts stay NaN and must never be read
        f = rk.first
        rho = rk.buf["rho"].reshape(rk.ext[2], rk.ext[1], rk.ext[0])
        rho[g:-g, g:-g, g:-g] = rho_g[f[2]:f[2] + rk.nl[2], f[1]:f[1] + rk.nl[1], f[0]:f[0] + rk.nl[0]]
    for phase in range(4):
        for rk in ranks:
            rk.copies(phase, 0)
        exchange(ranks, phase)
        for rk in ranks:
            rk.copies(phase, 2)
        if phase < 3:
            for rk in ranks:
                transforms(rk, phase, origin, h)
    scale = np.abs(want).max()
    for r, rk in enumerate(ranks):
        f = rk.first
        ef = rk.buf["ef"].reshape(rk.ext[2], rk.ext[1], rk.ext[0], 3)
        got = ef[g:-g, g:-g, g:-g]
        ref = want[f[2]:f[2] + rk.nl[2], f[1]:f[1] + rk.nl[1], f[0]:f[0] + rk.nl[0]]
        assert np.isfinite(got).all(), f"rank {r}: E interior not fully written"
        assert np.max(np.abs(got - ref)) <= 1e-12 * scale, (r, np.max(np.abs(got - ref)) / scale)
        halo = ef.copy()
        halo[g:-g, g:-g, g:-g] = np.nan
        assert np.isnan(halo).all(), f"rank {r}: the solve wrote into E's ghost layers"
        # the reference's inverse transform lands in rho's storage: rho interior <- last gradient component
        rho = rk.buf["rho"].reshape(rk.ext[2], rk.ext[1], rk.ext[0])
        assert np.array_equal(rho[g:-g, g:-g, g:-g], got[..., 2])
    for rk in ranks:
        rk.plan.close()
    layout.close()


def test_slab_ranges_tile_the_axes():
    layout = ib.Layout((20, 13, 7), 5)
    plans = [ib.SlabPlan(layout, r) for r in range(5)]
    assert [(p.zs, p.ze) for p in plans] == [(0, 2), (2, 4), (4, 5), (5, 6), (6, 7)]
    assert plans[0].ys == 0 and plans[-1].ye == 13 and all(a.ye == b.ys for a, b in zip(plans, plans[1:]))
    for p in plans:
        p.close()
    layout.close()
