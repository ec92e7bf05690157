"""CPU oracle for the IPPL particle-mesh hot path -- TEST INFRASTRUCTURE ONLY.

ctypes/numpy front-end of oracle/ippl_oracle.cpp (line-by-line restatement of the reference
algorithm; every C function cites the reference file:line it follows) plus the numpy restatement
of the periodic FFT Poisson solve (non-owned stage).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package; the product path
(ippl_b200) never does and has no CPU fallback.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libippl_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ippl_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "_build/libippl_oracle.so"])
    return _LIB_PATH


class Mesh(C.Structure):
    _fields_ = [("ng", C.c_int * 3), ("first", C.c_int * 3), ("nl", C.c_int * 3),
                ("nghost", C.c_int), ("origin", C.c_double * 3), ("h", C.c_double * 3)]

    @staticmethod
    def make(ng, origin, h, first=(0, 0, 0), nl=None, nghost=1):
        m = Mesh()
        nl = ng if nl is None else nl
        for d in range(3):
            m.ng[d], m.first[d], m.nl[d] = int(ng[d]), int(first[d]), int(nl[d])
            m.origin[d], m.h[d] = float(origin[d]), float(h[d])
        m.nghost = nghost
        return m

    @property
    def ext(self):
        return tuple(self.nl[d] + 2 * self.nghost for d in range(3))


class Penning(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("length", C.c_double * 3), ("V0", C.c_double),
                ("alpha", C.c_double), ("Bext", C.c_double), ("DrInv", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_field_sum.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _i3(v):
    return (C.c_int * 3)(*[int(t) for t in v])


def _d3(v):
    return (C.c_double * 3)(*[float(t) for t in v])


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def sample_landau(n, alpha, k, lo, hi, seed):
    """positions of LandauDampingManager::initializeParticles along one dimension (bench.py's CPU baseline)"""
    x = np.empty(n)
    lib().orc_sample_landau(C.c_long(n), C.c_double(alpha), C.c_double(k), C.c_double(lo), C.c_double(hi),
                            C.c_ulonglong(seed), _p(x))
    return x


def sample_normal(n, seed):
    p = np.empty(n)
    lib().orc_sample_normal(C.c_long(n), C.c_ulonglong(seed), _p(p))
    return p


def field_zeros(mesh, ncomp=1):
    ex, ey, ez = mesh.ext
    return np.zeros(ex * ey * ez * ncomp, dtype=np.float64)


def scatter_cic(mesh, x, y, z, q, rho, begin=0, end=None, hash=None, parallel=False):
    """ParticleAttrib::scatter kernel (no halo).  q: ndarray or python float (uniform charge)."""
    n = len(x)
    end = n if end is None else end
    qa = q if isinstance(q, np.ndarray) else None
    qs = 0.0 if qa is not None else float(q)
    lib().orc_scatter_cic(C.byref(mesh), C.c_long(begin), C.c_long(end), _p(x), _p(y), _p(z), _p(qa),
                          C.c_double(qs), _p(hash), _p(rho), int(parallel))
    return rho


def gather_cic(mesh, x, y, z, efield, out, add=False, parallel=False):
    ncomp = len(out)
    arr = (C.c_void_p * ncomp)(*[o.ctypes.data for o in out])
    lib().orc_gather_cic(C.byref(mesh), C.c_long(len(x)), _p(x), _p(y), _p(z), _p(efield), ncomp, arr,
                         int(add), int(parallel))
    return out


def periodic_bc(x, lo, hi, parallel=False):
    lib().orc_periodic_bc(C.c_long(len(x)), _p(x), C.c_double(lo), C.c_double(hi), int(parallel))
    return x


def kick(p, e, c, parallel=False):
    lib().orc_kick(C.c_long(len(p)), _p(p), _p(e), C.c_double(c), int(parallel))


def drift(r, p, dt, parallel=False):
    lib().orc_drift(C.c_long(len(r)), _p(r), _p(p), C.c_double(dt), int(parallel))


def penning_params(origin, length, dt, Bext=5.0):
    pp = Penning()
    for d in range(3):
        pp.origin[d], pp.length[d] = origin[d], length[d]
    pp.V0 = 30 * length[2]
    pp.alpha = -0.5 * dt
    pp.Bext = Bext
    pp.DrInv = 1.0 / (1 + (pp.alpha * Bext) ** 2)
    return pp


def penning_kick(which, pp, R, P, E):
    fn = lib().orc_penning_kick1 if which == 1 else lib().orc_penning_kick2
    fn(C.byref(pp), C.c_long(len(R[0])), _p(R[0]), _p(R[1]), _p(R[2]), _p(P[0]), _p(P[1]), _p(P[2]),
       _p(E[0]), _p(E[1]), _p(E[2]))


def halo_periodic(v, ext, ncomp, nghost, serial, mode):
    """mode: 'fill' or 'accumulate' (HaloCells::applyPeriodicSerialDim)."""
    lib().orc_halo_periodic(_p(v), _i3(ext), ncomp, nghost, _i3(serial), 0 if mode == "fill" else 1)
    return v


def partition(ng, nranks, parallel=(1, 1, 1)):
    boxes = np.zeros((nranks, 6), dtype=np.int32)
    rc = lib().orc_partition(_i3(ng), _i3(parallel), nranks, _p(boxes))
    if rc != 0:
        raise RuntimeError("partition failed")
    return boxes


def neighbors(ng, boxes, my, nghost=1, periodic=True):
    out = np.zeros((512, 14), dtype=np.int32)
    n = lib().orc_neighbors(_i3(ng), len(boxes), _p(np.ascontiguousarray(boxes, dtype=np.int32)), nghost,
                            int(periodic), my, _p(out), 512)
    assert n <= 512
    return out[:n].copy()


def matching_index(i):
    return lib().orc_matching_index(int(i))


def halo_exchange(ng, boxes, fields, ncomp, mode, nghost=1, periodic=True):
    arr = (C.c_void_p * len(fields))(*[f.ctypes.data for f in fields])
    lib().orc_halo_exchange(_i3(ng), len(boxes), _p(np.ascontiguousarray(boxes, dtype=np.int32)), nghost,
                            int(periodic), arr, ncomp, 0 if mode == "fill" else 1)


def regions(ng, boxes, origin, h):
    out = np.zeros((len(boxes), 6), dtype=np.float64)
    lib().orc_regions(_i3(ng), len(boxes), _p(np.ascontiguousarray(boxes, dtype=np.int32)), _d3(origin),
                      _d3(h), _p(out))
    return out


def locate(regs, my, x, y, z):
    dest = np.zeros(len(x), dtype=np.int32)
    lib().orc_locate(len(regs), _p(regs), my, C.c_long(len(x)), _p(x), _p(y), _p(z), _p(dest))
    return dest


def field_sum(v, ext, nghost=1):
    return lib().orc_field_sum(_p(v), _i3(ext), nghost)


def density(v, ext, nghost, cell_volume, q_over_size):
    lib().orc_density(_p(v), _i3(ext), nghost, C.c_double(cell_volume), C.c_double(q_over_size))


def pic_step_nosolve(mesh, R, P, E, q, dt, efield, rho):
    lib().orc_pic_step_nosolve(C.byref(mesh), C.c_long(len(R[0])), _p(R[0]), _p(R[1]), _p(R[2]),
                               _p(P[0]), _p(P[1]), _p(P[2]), _p(E[0]), _p(E[1]), _p(E[2]),
                               C.c_double(q), C.c_double(dt), _p(efield), _p(rho))


# ----------------------------------------------------------------------------------------------
# Multi-rank accumulate/fill exactly as BareField::accumulateHalo / fillHalo order them
# (src/Field/BareField.hpp:152-172): inter-rank exchange first, then the in-rank periodic wrap for
# the dims whose local extent equals the global one.
# ----------------------------------------------------------------------------------------------
def halo_full(ng, boxes, fields, ncomp, mode, nghost=1, periodic=True):
    if len(boxes) > 1:
        halo_exchange(ng, boxes, fields, ncomp, mode, nghost, periodic)
    if periodic:
        for r, f in enumerate(fields):
            nl = boxes[r, 3:6] - boxes[r, 0:3] + 1
            serial = [int(nl[d] == ng[d]) for d in range(3)]
            halo_periodic(f, tuple(int(n) + 2 * nghost for n in nl), ncomp, nghost, serial, mode)


# ----------------------------------------------------------------------------------------------
# ParticleSpatialLayout::update for all ranks in one process (src/Particle/ParticleSpatialLayout.hpp:
# 115-314): periodic BC, locate, remove leavers with the reference's hole-filling order
# (ParticleBase::internalDestroy, src/Particle/ParticleBase.hpp:175-276: i-th hole among the first
# N-d slots <- i-th survivor of the tail), append arrivals in ascending source rank
# (ParticleSpatialLayout.hpp:221-247, 306-308).  parts[r] = dict of equally long 1-D arrays with
# keys 'x','y','z' + any other attributes.  The order of arrivals inside one source is an atomic
# cursor race in the reference; the oracle uses ascending particle index.
# ----------------------------------------------------------------------------------------------
def update(ng, boxes, origin, h, parts):
    nranks = len(boxes)
    regs = regions(ng, boxes, origin, h)
    lo = [0 * h[d] + origin[d] for d in range(3)]
    hi = [ng[d] * h[d] + origin[d] for d in range(3)]
    for p in parts:
        for d, k in enumerate("xyz"):
            periodic_bc(p[k], lo[d], hi[d])
    if nranks < 2:
        return parts
    dests = [locate(regs, r, p["x"], p["y"], p["z"]) for r, p in enumerate(parts)]
    out = []
    for r, p in enumerate(parts):
        leaving = dests[r] != r
        n, nd = len(leaving), int(leaving.sum())
        keep = {}
        if nd == n:
            keep = {k: v[:0].copy() for k, v in p.items()}
        else:
            holes = np.nonzero(leaving[: n - nd])[0]
            fill = np.nonzero(~leaving[n - nd:])[0] + (n - nd)
            for k, v in p.items():
                w = v[: n - nd].copy()
                w[holes] = v[fill[: len(holes)]]
                keep[k] = w
        out.append(keep)
    for r in range(nranks):
        for src in range(nranks):
            if src == r:
                continue
            sel = np.nonzero(dests[src] == r)[0]
            if len(sel):
                for k in out[r]:
                    out[r][k] = np.concatenate([out[r][k], parts[src][k][sel]])
    return out


# ----------------------------------------------------------------------------------------------
# FFTPeriodicPoissonSolver::solve, GRAD output (src/PoissonSolvers/FFTPeriodicPoissonSolver.hpp:
# 53-169): forward r2c scaled by 1/N (heFFTe scale::full, src/FFT/Backend/Heffte.h:282-284),
# per gd: rho_hat * -(i * k_gd * factor), k_d = notMid * 2*pi/Len * (i_d - shift*N_d) with the
# Nyquist mode zeroed (:134-137) and factor = 1/|k|^2 (0 at DC, :147-148), unscaled c2r
# (Heffte.h:287-288) copied into E[gd] interior (:155-160).  NON-OWNED stage; numpy pocketfft is
# the stand-in for heFFTe/FFTW here.  rho_int: (nz, ny, nx) interior array (x fastest).
# Returns E interior as (nz, ny, nx, 3).
# ----------------------------------------------------------------------------------------------
def poisson_kspace_multipliers(N, origin, h):
    """The three k-space multipliers -(i k_gd / |k|^2) of FFTPeriodicPoissonSolver::solve, GRAD output
    (src/PoissonSolvers/FFTPeriodicPoissonSolver.hpp:115-150: shift, Nyquist ("notMid") rule, k = 0 guard), arrays
    [nz][ny][nx] for N = (nx, ny, nz)."""
    k = []
    for d in range(3):
        rmax = origin[d] + N[d] * h[d]
        Len = rmax - origin[d]
        i = np.arange(N[d])
        shift = i > (N[d] // 2)
        notmid = i != (N[d] // 2)
        k.append(notmid * 2 * np.pi / Len * (i - shift * N[d]))
    KX = k[0][None, None, :]
    KY = k[1][None, :, None]
    KZ = k[2][:, None, None]
    Dr = KX * KX + KY * KY + KZ * KZ
    nz_ = Dr != 0.0
    factor = nz_ * (1.0 / (Dr + (~nz_) * 1.0))
    return [-(1j * K * factor) for K in (KX, KY, KZ)]


def poisson_grad(rho_int, origin, h):
    nz, ny, nx = rho_int.shape
    rhat = np.fft.fftn(rho_int) / (nx * ny * nz)
    E = np.empty((nz, ny, nx, 3))
    for gd, M in enumerate(poisson_kspace_multipliers((nx, ny, nz), origin, h)):
        tmp = rhat * M
        E[..., gd] = np.real(np.fft.ifftn(tmp)) * (nx * ny * nz)
    return E


def interior(v, mesh, ncomp=1):
    ex, ey, ez = mesh.ext
    g = mesh.nghost
    a = v.reshape(ez, ey, ex, ncomp) if ncomp > 1 else v.reshape(ez, ey, ex)
    return a[g:ez - g, g:ey - g, g:ex - g]


class LandauOracle:
    """Single-rank alpine LandauDamping loop (demos/alpine/LandauDampingManager.h:66-157, 265-320,
    339-366 and AlpineManager.h:157-245) on the oracle kernels.  Particle initial conditions are
    INPUTS (the reference's Kokkos RNG stream is backend/thread dependent)."""

    def __init__(self, nr, R, P, parallel=True):
        self.nr = tuple(nr)
        kw, self.alpha = 0.5, 0.05
        self.rmax = 2 * np.pi / kw
        self.hr = [self.rmax / n for n in nr]
        self.origin = [0.0, 0.0, 0.0]
        self.Q = -1.0 * self.rmax * self.rmax * self.rmax
        self.dt = min(0.05, 0.5 * min(self.hr))
        self.mesh = Mesh.make(nr, self.origin, self.hr)
        self.R = [np.ascontiguousarray(r, dtype=np.float64).copy() for r in R]
        self.P = [np.ascontiguousarray(p, dtype=np.float64).copy() for p in P]
        self.n = len(self.R[0])
        self.q = self.Q / self.n
        self.E = [np.zeros(self.n) for _ in range(3)]
        self.rho = field_zeros(self.mesh)
        self.Ef = field_zeros(self.mesh, 3)
        self.time = 0.0
        self.par = parallel
        self.history = []
        self.rel_err = 0.0

    def scatter(self):
        self.rho[:] = 0.0
        scatter_cic(self.mesh, *self.R, self.q, self.rho, parallel=self.par)
        halo_periodic(self.rho, self.mesh.ext, 1, 1, (1, 1, 1), "accumulate")
        self.rel_err = abs((self.Q - field_sum(self.rho, self.mesh.ext)) / self.Q)
        cell = self.hr[0] * self.hr[1] * self.hr[2]
        size = 1.0
        for d in range(3):
            size *= self.rmax - 0.0
        density(self.rho, self.mesh.ext, 1, cell, self.Q / size)

    def solve(self):
        E = poisson_grad(interior(self.rho, self.mesh), self.origin, self.hr)
        interior(self.Ef, self.mesh, 3)[...] = E

    def gather(self):
        halo_periodic(self.Ef, self.mesh.ext, 3, 1, (1, 1, 1), "fill")
        gather_cic(self.mesh, *self.R, self.Ef, self.E, parallel=self.par)

    def dump(self):
        ex = interior(self.Ef, self.mesh, 3)[..., 0]
        e2 = float(np.sum(ex ** 2))
        energy = e2 * self.hr[0] * self.hr[1] * self.hr[2]
        self.history.append((self.time, energy, float(np.max(np.abs(ex)))))

    def pre_run(self):
        self.scatter()
        self.solve()
        self.gather()
        self.dump()

    def step(self):
        dt = self.dt
        for d in range(3):
            kick(self.P[d], self.E[d], 0.5 * dt, self.par)
        for d in range(3):
            drift(self.R[d], self.P[d], dt, self.par)
        for d in range(3):
            periodic_bc(self.R[d], 0.0 * self.hr[d] + 0.0, self.nr[d] * self.hr[d] + 0.0, self.par)
        self.scatter()
        self.solve()
        self.gather()
        for d in range(3):
            kick(self.P[d], self.E[d], 0.5 * dt, self.par)
        self.time += dt
        self.dump()
