"""oracle/extras.py -- CPU restatement (numpy) of the rows SURVEY.md 8f marks "next": particle sampling, orthogonal
recursive bisection and the dump reductions.  TEST INFRASTRUCTURE ONLY (same rules as ippl_oracle.cpp: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import it; the product never does).

PARITY STATUS
  * sampling: PINNED against the reference's own headers for everything but the random stream -- Random/Distribution.h,
    NormalDistribution.h, Utility.h (NewtonRaphson), InverseTransformSampling.h (constructor = rank counts + CDF bounds,
    generate / fill_random), Randn.h and the managers' own CustomDistributionFunctions structs (cut out of
    demos/alpine/*Manager.h at build time) are compiled in place from /root/reference with replayed random numbers
    (oracle/ref_shim/refshim_random.cpp -> oracle/_ref/libippl_refshim_random.so) and compared live and through the
    committed fixture tests/golden/ref_random.npz (tests/test_oracle_random_pinned.py).  The STREAM itself is "parity
    unpinned" by nature: the reference draws from Kokkos::Random_XorShift64_Pool, whose stream assignment is per thread /
    per lock and backend dependent; ours is the published Philox4x32-10, implemented here independently of
    the product and pinned by the Random123 known-answer vectors.
  * ORB: findCutAxis / findMedian / cutDomain / binaryRepartition restated from
    src/Decomposition/OrthogonalRecursiveBisection.hpp:14-232 and PINNED against that very code: the reference's
    OrthogonalRecursiveBisection.h/.hpp (+ FieldLayout::updateLayout) are compiled in place on serial stand-ins
    (oracle/ref_shim/refshim_orb.cpp -> oracle/_ref/libippl_refshim_orb.so) and compared live and through the committed
    tests/golden/ref_orb.npz (tests/test_oracle_orb_pinned.py: 63 weight fields x rank counts, exact boxes); plus the
    invariants of the reference's own test unit_tests/PIC/ORB.cpp (every rank keeps a box, boxes tile the domain).
  * dumps: PenningTrapManager.h:346-389, LandauDampingManager.h:339-366, BumponTailInstabilityManager.h:448-480.
  * AlpineOracle (PenningTrap / BumponTail loops): built from the kernels of ippl_oracle.cpp, which ARE pinned against
    the reference's headers; the loops themselves restate the managers line by line.  The reference holds no known-answer
    file for these two apps (only FieldLandau_valid_result.csv exists): "parity unpinned" at app level beyond the
    invariants in tests/test_extras_cpu.py (charge conservation at the reference's 1e-10 abort threshold, the imposed
    perturbation's field energy from linear theory, the B-field rotation conserving |P| when E = 0).
"""
import math

import numpy as np

U32 = np.uint64(0xFFFFFFFF)


# ---- Philox4x32-10 (Salmon et al., SC'11), vectorised over counters ------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c = [np.asarray(x, dtype=np.uint64) & U32 for x in (c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0) & U32, np.uint64(k1) & U32
    m0, m1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    s32 = np.uint64(32)
    for _ in range(10):
        p0, p1 = m0 * c[0], m1 * c[2]
        c = [(p1 >> s32) ^ c[1] ^ k0, p1 & U32, (p0 >> s32) ^ c[3] ^ k1, p0 & U32]
        k0 = (k0 + np.uint64(0x9E3779B9)) & U32
        k1 = (k1 + np.uint64(0xBB67AE85)) & U32
    return c


def philox_uniform2(seed, ids, dim, block=0):
    """two uniforms in [0,1) per id: 53 bits of (c0:c1) and of (c2:c3); counter = (id lo, id hi, dim, block)"""
    ids = np.asarray(ids, dtype=np.uint64)
    seed = int(seed)
    c = philox4x32_10(ids & U32, ids >> np.uint64(32), np.full(ids.shape, dim, np.uint64),
                      np.full(ids.shape, block, np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    a = (c[0] << np.uint64(32)) | c[1]
    b = (c[2] << np.uint64(32)) | c[3]
    scale = 2.0 ** -53
    return (a >> np.uint64(11)).astype(np.float64) * scale, (b >> np.uint64(11)).astype(np.float64) * scale


# ---- distributions of the alpine managers -------------------------------------------------------------------
UNIFORM, COSINE, NORMAL = 0, 1, 2
_erf = np.vectorize(math.erf, otypes=[np.float64])


class Dist:
    """ippl::random::Distribution<double, 3, 6, Functions> with one kind per dimension (par[2d], par[2d+1])."""

    def __init__(self, kind, par):
        self.kind, self.par = list(kind), [float(p) for p in par]

    def cdf(self, x, d):
        a, b = self.par[2 * d], self.par[2 * d + 1]
        if np.ndim(x) == 0:             # scalars through libm (what the reference's host code calls)
            x = float(x)
            if self.kind[d] == COSINE:
                return x + (a / b) * math.sin(b * x)
            if self.kind[d] == NORMAL:
                return 0.5 * (1 + math.erf((x - a) / (b * math.sqrt(2.0))))
            return x
        if self.kind[d] == COSINE:      # LandauDampingManager.h:23-28
            return x + (a / b) * np.sin(b * x)
        if self.kind[d] == NORMAL:      # NormalDistribution.h:11-13
            return 0.5 * (1 + _erf((x - a) / (b * math.sqrt(2.0))))
        return x + 0.0                  # UniformDistribution.h:17-20

    def pdf(self, x, d):
        a, b = self.par[2 * d], self.par[2 * d + 1]
        if self.kind[d] == COSINE:      # LandauDampingManager.h:31-35
            return 1.0 + a * np.cos(b * x)
        if self.kind[d] == NORMAL:      # NormalDistribution.h:16-20
            return (1.0 / (b * math.sqrt(2 * math.pi))) * np.exp(-(x - a) * (x - a) / (2 * b * b))
        return np.ones_like(np.asarray(x, dtype=np.float64))

    def estimate(self, u, d):
        if self.kind[d] == NORMAL:      # NormalDistribution.h:23-25
            return self.par[2 * d] + 0.0 * u * self.par[2 * d + 1]
        return u + self.par[d] * 0.0    # LandauDampingManager.h:38-42


def sample_counts(dist, rmin, rmax, regions, ntotal):
    """InverseTransformSampling::updateBounds for every rank, InverseTransformSampling.h:106-131.
    regions[r] = (min[3], max[3]).  Returns (nlocal[r], ubounds[r] = umin[3] + umax[3])."""
    nloc, ub = [], []
    for reg in regions:
        pnr = pdr = 1.0
        um, uM = [], []
        for d in range(3):
            lo, hi = float(reg[d]), float(reg[3 + d])
            nr = float(dist.cdf(hi, d)) - float(dist.cdf(lo, d))
            dr = float(dist.cdf(float(rmax[d]), d)) - float(dist.cdf(float(rmin[d]), d))
            pnr *= nr
            pdr *= dr
            um.append(float(dist.cdf(lo, d)))
            uM.append(float(dist.cdf(hi, d)))
        nloc.append(int((pnr / pdr) * ntotal))
        ub.append(um + uM)
    rest = int(ntotal - sum(nloc))
    for r in range(len(nloc)):
        if r < rest:
            nloc[r] += 1
    return nloc, ub


def newton_positions(dist, umin, umax, u01):
    """fill_random::operator() + NewtonRaphson::solve per dimension given the uniforms u01[d] in [0,1)."""
    out = []
    for d in range(3):
        u = umin[d] + (umax[d] - umin[d]) * u01[d]
        x = np.array(dist.estimate(u, d), dtype=np.float64)
        it = np.zeros(x.shape, dtype=np.int64)
        while True:
            act = (it < 20) & (np.abs(dist.cdf(x, d) - u) > 1e-12)
            if not act.any():
                break
            xn = x - ((dist.cdf(x, d) - u) / dist.pdf(x, d))
            x = np.where(act, xn, x)
            it += act
        out.append(x)
    return out


def sample_positions(dist, umin, umax, seed, first_id, n):
    ids = np.arange(first_id, first_id + n, dtype=np.uint64)
    u01 = [philox_uniform2(seed, ids, d)[0] for d in range(3)]
    return newton_positions(dist, umin, umax, u01), u01


def sample_normal(mu, sd, seed, first_id, n):
    """randn::operator(), Randn.h:82-94, normals by Box-Muller on counter dimensions 3 and 4"""
    ids = np.arange(first_id, first_id + n, dtype=np.uint64)
    a0, a1 = philox_uniform2(seed, ids, 3)
    b0, b1 = philox_uniform2(seed, ids, 4)
    two_pi = 6.283185307179586476925286766559
    r0, r1 = np.sqrt(-2.0 * np.log(1.0 - a0)), np.sqrt(-2.0 * np.log(1.0 - b0))
    g = [r0 * np.cos(two_pi * a1), r0 * np.sin(two_pi * a1), r1 * np.cos(two_pi * b1)]
    return [mu[d] + sd[d] * g[d] for d in range(3)]


def full_pdf_field(dist, nl, first, origin, h):
    """rho(cell) = getFullPdf((global index + 0.5) * h + origin): interior array [nz][ny][nx]"""
    ax = [(np.arange(nl[d]) + first[d] + 0.5) * h[d] + origin[d] for d in range(3)]
    px, py, pz = (dist.pdf(ax[d], d) for d in range(3))
    return ((1.0 * px[None, None, :]) * py[None, :, None]) * pz[:, None, None]


# ---- orthogonal recursive bisection -----------------------------------------------------------------------------
def orb_find_median(w):
    """OrthogonalRecursiveBisection::findMedian, .hpp:185-216"""
    n = len(w)
    if n == 4:
        return 1
    tot = 0.0
    for x in w:
        tot += float(x)
    half = 0.5 * tot
    curr = 0.0
    for i in range(n - 1):
        curr += float(w[i])
        if curr >= half:
            if i == 0:
                return 1
            previous = curr - float(w[i])
            if (curr + previous) <= tot and curr != half:
                return i - 1 if i == n - 2 else i
            return i - 1 if i > 1 else 1
    return n - 3


def orb_repartition(ng, nranks, weight):
    """binaryRepartition, .hpp:14-105, on the GLOBAL interior weight array weight[z][y][x] (the sum over ranks of
    perpendicularReduction's per-rank plane sums is the plane sum of the global field).  Returns (boxes, ok),
    boxes[r] = lo[3] + hi[3] inclusive."""
    domains = [[0, 0, 0, ng[0] - 1, ng[1] - 1, ng[2] - 1]]
    procs = [nranks]
    it, maxprocs = 0, nranks
    while maxprocs > 1:
        d = domains[it]
        lens = [d[3 + k] - d[k] + 1 for k in range(3)]
        axis = 0
        for k in range(1, 3):          # std::max_element: first maximum
            if lens[axis] < lens[k]:
                axis = k
        sub = weight[d[2]:d[5] + 1, d[1]:d[4] + 1, d[0]:d[3] + 1]
        other = tuple(a for a in range(3) if a != 2 - axis)   # array axes are (z, y, x)
        reduced = sub.sum(axis=other)
        median = orb_find_median(reduced)
        mid = median + d[axis]
        left, right = list(d), list(d)
        left[3 + axis] = mid
        right[axis] = mid + 1
        domains[it] = left
        domains.insert(it + 1, right)
        temp = procs[it]
        procs[it] = temp // 2
        procs.insert(it + 1, temp - procs[it])
        maxprocs = 0
        for i, p in enumerate(procs):
            if p > maxprocs:
                maxprocs, it = p, i
    ok = all(dm[3 + k] - dm[k] + 1 != 1 for dm in domains for k in range(3))
    return domains, ok


# ---- dump reductions --------------------------------------------------------------------------------------------
def energy_stats(E_int):
    """E_int[z][y][x][3] interior.  (sum E_d^2 [3], max|E_d| [3], sum dot(E,E))"""
    s2 = [float(np.sum(E_int[..., d] ** 2)) for d in range(3)]
    mx = [float(np.max(np.abs(E_int[..., d]))) for d in range(3)]
    dot = float(np.sum((E_int[..., 0] ** 2 + E_int[..., 1] ** 2) + E_int[..., 2] ** 2))
    return s2, mx, dot


def kinetic(P):
    """sum_i dot(P_i, P_i), PenningTrapManager.h:354-362"""
    return float(np.sum((P[0] * P[0] + P[1] * P[1]) + P[2] * P[2]))


# ---- the other two alpine mini-apps as single-rank CPU loops ------------------------------------------------------
def _alpine_base():
    import oracle as _o
    return _o


class AlpineOracle:
    """PenningTrap / BumponTailInstability loops on the oracle kernels, same structure as oracle.LandauOracle
    (AlpineManager.h:157-245 for scatter + getDensity; the managers' pre_run / LeapFrogStep / dump).  Initial particles
    are INPUTS.  kind = "bumpontail": leapfrog push, dump = (Ez energy, Ez max norm)
    (BumponTailInstabilityManager.h:315-369, 448-500); kind = "penning": Kick1 / drift / Kick2 in the external fields,
    dump = (potential energy 0.5 h^3 sum dot(E,E), kinetic 0.5 sum dot(P,P), |Ex|, |Ey|, |Ez|)
    (PenningTrapManager.h:242-336, 346-389)."""

    def __init__(self, kind, nr, R, P, parallel=True):
        o = _alpine_base()
        self.o, self.kind, self.nr = o, kind, tuple(nr)
        if kind == "bumpontail":
            kw = 0.21
            self.rmax = 2 * math.pi / kw
            self.Q = -1.0 * self.rmax * self.rmax * self.rmax
            self.hr = [self.rmax / n for n in nr]
            self.dt = min(0.05, 0.5 * min(self.hr))
        elif kind == "penning":
            self.rmax = 20.0
            self.Q = -1562.5
            self.hr = [self.rmax / n for n in nr]
            self.dt = 0.5 * (self.rmax / 2048)
            self.pp = o.penning_params((0.0, 0.0, 0.0), (self.rmax,) * 3, self.dt, 5.0)
        else:
            raise ValueError(kind)
        self.origin = [0.0, 0.0, 0.0]
        self.mesh = o.Mesh.make(nr, self.origin, self.hr)
        self.R = [np.ascontiguousarray(r, dtype=np.float64).copy() for r in R]
        self.P = [np.ascontiguousarray(p, dtype=np.float64).copy() for p in P]
        self.n = len(self.R[0])
        self.q = self.Q / self.n
        self.E = [np.zeros(self.n) for _ in range(3)]
        self.rho = o.field_zeros(self.mesh)
        self.Ef = o.field_zeros(self.mesh, 3)
        self.time, self.par, self.history = 0.0, parallel, []

    def scatter(self):
        o = self.o
        self.rho[:] = 0.0
        o.scatter_cic(self.mesh, *self.R, self.q, self.rho, parallel=self.par)
        o.halo_periodic(self.rho, self.mesh.ext, 1, 1, (1, 1, 1), "accumulate")
        self.rel_err = abs((self.Q - o.field_sum(self.rho, self.mesh.ext)) / self.Q)
        cell = self.hr[0] * self.hr[1] * self.hr[2]
        size = 1.0
        for _ in range(3):
            size *= self.rmax - 0.0
        o.density(self.rho, self.mesh.ext, 1, cell, self.Q / size)

    def solve(self):
        o = self.o
        E = o.poisson_grad(o.interior(self.rho, self.mesh), self.origin, self.hr)
        o.interior(self.Ef, self.mesh, 3)[...] = E

    def gather(self):
        o = self.o
        o.halo_periodic(self.Ef, self.mesh.ext, 3, 1, (1, 1, 1), "fill")
        o.gather_cic(self.mesh, *self.R, self.Ef, self.E, parallel=self.par)

    def dump(self):
        Ei = self.o.interior(self.Ef, self.mesh, 3)
        cell = self.hr[0] * self.hr[1] * self.hr[2]
        if self.kind == "bumpontail":
            ez = Ei[..., 2]
            self.history.append((self.time, float(np.sum(ez ** 2)) * cell, float(np.max(np.abs(ez)))))
        else:
            s2, _, dot = energy_stats(Ei)
            self.history.append((self.time, 0.5 * cell * dot, 0.5 * kinetic(self.P), math.sqrt(s2[0]), math.sqrt(s2[1]),
                                 math.sqrt(s2[2])))

    def pre_run(self):
        self.scatter()
        self.solve()
        self.gather()
        self.dump()

    def step(self):
        o, dt = self.o, self.dt
        if self.kind == "penning":
            o.penning_kick(1, self.pp, self.R, self.P, self.E)
        else:
            for d in range(3):
                o.kick(self.P[d], self.E[d], 0.5 * dt, self.par)
        for d in range(3):
            o.drift(self.R[d], self.P[d], dt, self.par)
        for d in range(3):
            o.periodic_bc(self.R[d], 0.0 * self.hr[d] + 0.0, self.nr[d] * self.hr[d] + 0.0, self.par)
        self.scatter()
        self.solve()
        self.gather()
        if self.kind == "penning":
            o.penning_kick(2, self.pp, self.R, self.P, self.E)
        else:
            for d in range(3):
                o.kick(self.P[d], self.E[d], 0.5 * dt, self.par)
        self.time += dt
        self.dump()
