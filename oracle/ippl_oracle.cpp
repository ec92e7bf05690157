// ============================================================================
// ippl_oracle.cpp -- CPU restatement of IPPL's particle-mesh hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library, and only as the checker or
// the timed CPU baseline.  The product path (ippl_b200/csrc) never links,
// imports or falls back to it.
//
// PARITY STATUS: the reference cannot be compiled here as a whole (Kokkos
// 5.2, heFFTe and MPI are absent), so every function below is
// a line-by-line restatement of the cited reference code (paths relative to
// /root/reference).  It is pinned three ways (see tests/test_oracle_*.py):
//   (1) against the reference's OWN headers for the pure-arithmetic pieces
//       (CIC.hpp weights/fold order, ParticleBC.h PeriodicBC, Index/NDIndex/
//       Partitioner/FieldLayout neighbour tables, HaloCells.hpp's periodic wrap
//       and multi-rank exchange, RegionLayout's rank regions, ownership, the
//       scatter / gather and PenningTrap kick lambdas), compiled from /root/reference with a minimal Kokkos
//       stand-in into oracle/_ref (oracle/ref_shim/, Makefile target `ref`);
//   (2) against every invariant the reference's unit tests hold for the
//       path;
//   (3) against the one known-answer file of the reference,
//       demos/alpine/validation/FieldLandau_valid_result.csv, at the
//       reference's own tolerance.
//
// Compile with -ffp-contract=off: the restatement performs exactly the IEEE
// operations the C++ source of the reference spells out, in that order.
// Field storage: ghosted, x fastest: idx = i + ex*(j + ey*k), ex = nl[0]+2*ng.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

struct orc_mesh {
    int ng[3];         // global cells per dim
    int first[3];      // first global cell index of the local box (lDom.first())
    int nl[3];         // local cells per dim
    int nghost;        // ghost layers (reference: 1, BareField.hpp:100)
    double origin[3];  // mesh origin
    double h[3];       // mesh spacing
};

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// bench.py's reference arm runs under torchrun, which exports OMP_NUM_THREADS=1: the arm sets the thread count itself
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// ---------------------------------------------------------------------------
// Synthetic Landau initial condition for the timed CPU baseline (bench.py only; the parity tests feed identical arrays
// to both sides instead).  Same distributions as LandauDampingManager::initializeParticles
// (demos/alpine/LandauDampingManager.h:159-254): per dimension x by inverse transform of
// cdf(x) = x + (alpha / k) sin(k x) on [lo, hi] with the reference's Newton iteration (Random/Utility.h:27-60:
// while iter < 20 && |f| > 1e-12: x -= f / pdf), velocities N(0, 1).  The uniform stream is a counter hash
// (splitmix64), not Kokkos' pool: only the distribution matters for a throughput baseline.
// ---------------------------------------------------------------------------
static inline double orc_u01(unsigned long long i, unsigned long long seed) {
    unsigned long long z = i * 0x9E3779B97F4A7C15ULL + seed * 0xD1B54A32D192ED03ULL + 0x632BE59BD9B4E019ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

void orc_sample_landau(long n, double alpha, double k, double lo, double hi, unsigned long long seed, double* x) {
    const double c0 = lo + (alpha / k) * std::sin(k * lo), c1 = hi + (alpha / k) * std::sin(k * hi);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const double u = c0 + (c1 - c0) * orc_u01((unsigned long long)i, seed);
        double v       = u;  // estimate(u) = u (LandauDampingManager.h:36-38)
        for (int it = 0; it < 20; ++it) {
            const double f = v + (alpha / k) * std::sin(k * v) - u;
            if (std::fabs(f) <= 1e-12) break;
            v -= f / (1.0 + alpha * std::cos(k * v));
        }
        x[i] = v < lo ? lo : (v > hi ? hi : v);
    }
}

void orc_sample_normal(long n, unsigned long long seed, double* p) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const double u1 = orc_u01((unsigned long long)(2 * i), seed), u2 = orc_u01((unsigned long long)(2 * i + 1), seed);
        p[i] = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
}

// ---------------------------------------------------------------------------
// CIC index/weights: src/Particle/ParticleAttrib.hpp:174-179 (scatter) and
// :229-234 (gather):  l = (x - origin) * invdx + 0.5; index = (int) l;
// whi = l - index; wlo = 1.0 - whi; args = index - lDom.first() + nghost.
// invdx = 1.0 / dx is formed once on the host (:153, :217).
// ---------------------------------------------------------------------------
struct Cic {
    double whi[3], wlo[3];
    long args[3];
};

static inline void cic_setup(const orc_mesh* m, const double invdx[3], double x, double y, double z,
                             Cic& c) {
    const double pos[3] = {x, y, z};
    for (int d = 0; d < 3; ++d) {
        double l  = (pos[d] - m->origin[d]) * invdx[d] + 0.5;
        int index = (int)l;
        c.whi[d]  = l - index;
        c.wlo[d]  = 1.0 - c.whi[d];
        c.args[d] = (long)(index - m->first[d] + m->nghost);
    }
}

// src/Interpolation/CIC.hpp:6-24: bit d of the point id set -> (args[d]-1, wlo[d]),
// clear -> (args[d], whi[d]).  The weight product is the unary RIGHT fold
// (w0 * (w1 * w2)) of CIC.hpp:32-33 / :55.
static inline double cic_weight(const Cic& c, int p) {
    double w0 = (p & 1) ? c.wlo[0] : c.whi[0];
    double w1 = (p & 2) ? c.wlo[1] : c.whi[1];
    double w2 = (p & 4) ? c.wlo[2] : c.whi[2];
    return w0 * (w1 * w2);
}
static inline long cic_index(const Cic& c, int p, long ex, long ey) {
    long i = (p & 1) ? c.args[0] - 1 : c.args[0];
    long j = (p & 2) ? c.args[1] - 1 : c.args[1];
    long k = (p & 4) ? c.args[2] - 1 : c.args[2];
    return i + ex * (j + ey * k);
}

// ParticleAttrib<T>::scatter kernel body, src/Particle/ParticleAttrib.hpp:167-184
// + detail::scatterToField, CIC.hpp:26-45 (points visited in order 0..7, each
// `view(idx) += val * w`).  q == NULL means every particle carries q_scalar.
// `parallel` != 0 uses OpenMP + atomics exactly like the reference's
// Kokkos::atomic_add on the OpenMP backend (the timed CPU baseline);
// parallel == 0 is the deterministic serial order used by the checker.
// hash (may be NULL) is the optional index remap of :170-171.
void orc_scatter_cic(const orc_mesh* m, long begin, long end, const double* x, const double* y,
                     const double* z, const double* q, double q_scalar, const int* hash,
                     double* rho, int parallel) {
    const double invdx[3] = {1.0 / m->h[0], 1.0 / m->h[1], 1.0 / m->h[2]};
    const long ex = m->nl[0] + 2 * m->nghost, ey = m->nl[1] + 2 * m->nghost;
    if (!parallel) {
        for (long idx = begin; idx < end; ++idx) {
            long i = hash ? hash[idx] : idx;
            Cic c;
            cic_setup(m, invdx, x[i], y[i], z[i], c);
            const double val = q ? q[i] : q_scalar;
            for (int p = 0; p < 8; ++p) rho[cic_index(c, p, ex, ey)] += val * cic_weight(c, p);
        }
    } else {
#pragma omp parallel for schedule(static)
        for (long idx = begin; idx < end; ++idx) {
            long i = hash ? hash[idx] : idx;
            Cic c;
            cic_setup(m, invdx, x[i], y[i], z[i], c);
            const double val = q ? q[i] : q_scalar;
            for (int p = 0; p < 8; ++p) {
                double add = val * cic_weight(c, p);
                double* t  = &rho[cic_index(c, p, ex, ey)];
#pragma omp atomic
                *t += add;
            }
        }
    }
}

// ParticleAttrib<T>::gather kernel body, src/Particle/ParticleAttrib.hpp:226-244 +
// detail::gatherFromField, CIC.hpp:47-66: sum over points is the RIGHT fold
// g0 + (g1 + (... + (g6 + g7))), each g_p = w_p * view(idx_p) per component.
// E field is AoS Vector<double,3> per cell (24 B), ghosted.
void orc_gather_cic(const orc_mesh* m, long n, const double* x, const double* y, const double* z,
                    const double* efield, int ncomp, double** out, int add_to_attribute,
                    int parallel) {
    const double invdx[3] = {1.0 / m->h[0], 1.0 / m->h[1], 1.0 / m->h[2]};
    const long ex = m->nl[0] + 2 * m->nghost, ey = m->nl[1] + 2 * m->nghost;
#pragma omp parallel for schedule(static) if (parallel)
    for (long i = 0; i < n; ++i) {
        Cic c;
        cic_setup(m, invdx, x[i], y[i], z[i], c);
        double w[8];
        long id[8];
        for (int p = 0; p < 8; ++p) {
            w[p]  = cic_weight(c, p);
            id[p] = cic_index(c, p, ex, ey);
        }
        for (int d = 0; d < ncomp; ++d) {
            double acc = w[7] * efield[id[7] * ncomp + d];
            for (int p = 6; p >= 0; --p) acc = w[p] * efield[id[p] * ncomp + d] + acc;
            if (add_to_attribute)
                out[d][i] += acc;
            else
                out[d][i] = acc;
        }
    }
}

// PeriodicBC::operator(), src/Particle/ParticleBC.h:73-76, applied per dimension
// by ParticleLayout::applyBC, src/Particle/ParticleLayout.hpp:34-74 (lower faces
// only).  extent = max - min, middle = (min + max) / 2 (ParticleBC.h:50-51).
void orc_periodic_bc(long n, double* x, double lo, double hi, int parallel) {
    const double extent = hi - lo;
    const double middle = (lo + hi) / 2;
#pragma omp parallel for schedule(static) if (parallel)
    for (long i = 0; i < n; ++i) {
        double value = x[i];
        x[i]         = value - extent * (int)((value - middle) * 2 / extent);
    }
}

// ParticleAttrib::operator=(Expression), src/Particle/ParticleAttrib.hpp:118-130 for
// the two alpine expressions (demos/alpine/LandauDampingManager.h:281,286,318):
//   kick : P = P - (0.5*dt) * E   -> p[i] = p[i] - c * e[i]      (c = 0.5*dt)
//   drift: R = R + dt * P         -> r[i] = r[i] + dt * p[i]
void orc_kick(long n, double* p, const double* e, double c, int parallel) {
#pragma omp parallel for schedule(static) if (parallel)
    for (long i = 0; i < n; ++i) p[i] = p[i] - c * e[i];
}
void orc_drift(long n, double* r, const double* p, double dt, int parallel) {
#pragma omp parallel for schedule(static) if (parallel)
    for (long i = 0; i < n; ++i) r[i] = r[i] + dt * p[i];
}

// PenningTrap Kick1 / Kick2, demos/alpine/PenningTrapManager.h:256-272, 313-333.
// length/origin per dim, V0 = 30*length[2], alpha = -0.5*dt, DrInv = 1/(1+(alpha*B)^2).
struct orc_penning {
    double origin[3], length[3], V0, alpha, Bext, DrInv;
};
static inline void penning_eext(const orc_penning* pp, double x, double y, double z, double ex,
                                double ey, double ez, double& Ex, double& Ey, double& Ez) {
    const double l2 = std::pow(pp->length[2], 2);
    Ex = -(x - pp->origin[0] - 0.5 * pp->length[0]) * (pp->V0 / (2 * l2));
    Ey = -(y - pp->origin[1] - 0.5 * pp->length[1]) * (pp->V0 / (2 * l2));
    Ez = (z - pp->origin[2] - 0.5 * pp->length[2]) * (pp->V0 / (l2));
    Ex += ex;
    Ey += ey;
    Ez += ez;
}
void orc_penning_kick1(const orc_penning* pp, long n, const double* x, const double* y,
                       const double* z, double* px, double* py, double* pz, const double* ex,
                       const double* ey, const double* ez) {
#pragma omp parallel for schedule(static)
    for (long j = 0; j < n; ++j) {
        double Ex, Ey, Ez;
        penning_eext(pp, x[j], y[j], z[j], ex[j], ey[j], ez[j], Ex, Ey, Ez);
        px[j] += pp->alpha * (Ex + py[j] * pp->Bext);
        py[j] += pp->alpha * (Ey - px[j] * pp->Bext);
        pz[j] += pp->alpha * Ez;
    }
}
void orc_penning_kick2(const orc_penning* pp, long n, const double* x, const double* y,
                       const double* z, double* px, double* py, double* pz, const double* ex,
                       const double* ey, const double* ez) {
#pragma omp parallel for schedule(static)
    for (long j = 0; j < n; ++j) {
        double Ex, Ey, Ez;
        penning_eext(pp, x[j], y[j], z[j], ex[j], ey[j], ez[j], Ex, Ey, Ez);
        const double a = pp->alpha, B = pp->Bext;
        px[j] = pp->DrInv * (px[j] + a * (Ex + py[j] * B + a * B * Ey));
        py[j] = pp->DrInv * (py[j] + a * (Ey - px[j] * B - a * B * Ex));
        pz[j] += a * Ez;
    }
}

// ---------------------------------------------------------------------------
// In-rank periodic wrap: HaloCells::applyPeriodicSerialDim + HaloPeriodicFunctor,
// src/Field/HaloCells.hpp:59-87, 297-336.  For d = 0,1,2 in order, if the local
// extent equals the global one: for i in [0,nghost) and ALL indices (ghosts
// included) of the other dims, with N = extent(d)-1:
//   left = v[nghost+i], right = v[N-nghost-i], glow = v[nghost-1-i], gup = v[N-nghost+1+i]
//   fill (assign):            glow = right;  gup = left
//   accumulate (rhs_plus_assign, HaloCells.h:119-121): right += glow; left += gup
// mode: 0 = fill, 1 = accumulate.  serial[d] != 0 <=> dim d is un-split.
// ---------------------------------------------------------------------------
void orc_halo_periodic(double* v, const int ext[3], int ncomp, int nghost, const int serial[3],
                       int mode) {
    const long e0 = ext[0], e1 = ext[1], e2 = ext[2];
    for (int d = 0; d < 3; ++d) {
        if (!serial[d]) continue;
        const int N = ext[d] - 1;
        long end[3] = {e0, e1, e2};
        end[d]      = nghost;
        for (long c2 = 0; c2 < end[2]; ++c2)
            for (long c1 = 0; c1 < end[1]; ++c1)
                for (long c0 = 0; c0 < end[0]; ++c0) {
                    long co[3] = {c0, c1, c2};
                    const long i = co[d];
                    auto at = [&](long cd) {
                        long t[3] = {co[0], co[1], co[2]};
                        t[d]      = cd;
                        return (t[0] + e0 * (t[1] + e1 * t[2])) * ncomp;
                    };
                    long left = at(nghost + i), right = at(N - nghost - i);
                    long glow = at(nghost - 1 - i), gup = at(N - nghost + 1 + i);
                    for (int c = 0; c < ncomp; ++c) {
                        if (mode == 0) {
                            v[glow + c] = v[right + c];
                            v[gup + c]  = v[left + c];
                        } else {
                            v[right + c] += v[glow + c];
                            v[left + c] += v[gup + c];
                        }
                    }
                }
    }
}

// ---------------------------------------------------------------------------
// Index / NDIndex / Partitioner restatement (host value types).
// ---------------------------------------------------------------------------
struct Box {
    int lo[3], hi[3];  // inclusive, stride 1
    int len(int d) const { return hi[d] - lo[d] + 1; }
};
static Box grow(const Box& b, int n) {
    Box r = b;
    for (int d = 0; d < 3; ++d) {
        r.lo[d] -= n;
        r.hi[d] += n;
    }
    return r;
}
// Index::touches, src/Index/Index.hpp:153-155
static bool touches(const Box& a, const Box& b) {
    for (int d = 0; d < 3; ++d)
        if (!(a.lo[d] <= b.hi[d] && a.hi[d] >= b.lo[d])) return false;
    return true;
}
static Box intersect(const Box& a, const Box& b) {
    Box r;
    for (int d = 0; d < 3; ++d) {
        r.lo[d] = std::max(a.lo[d], b.lo[d]);
        r.hi[d] = std::min(a.hi[d], b.hi[d]);
    }
    return r;
}

// Partitioner::split, src/Partition/Partitioner.hpp:15-123, with Index::split
// (mid = first + length/2 - 1, src/Index/Index.hpp:162-169) and the ratio split
// (mid = first + (int)(length*a + 0.5) - 1, :183-191).
// boxes out: [nranks][6] = lo0,lo1,lo2,hi0,hi1,hi2 (inclusive global indices).
int orc_partition(const int ng[3], const int is_parallel[3], int nsplits, int* boxes_out) {
    std::vector<Box> dom(nsplits);
    for (int d = 0; d < 3; ++d) {
        dom[0].lo[d] = 0;
        dom[0].hi[d] = ng[d] - 1;
    }
    auto split_half = [](const Box& b, Box& l, Box& r, int d) {
        int mid = b.lo[d] + b.len(d) / 2 - 1;
        l = b; r = b;
        l.hi[d] = mid;
        r.lo[d] = mid + 1;
    };
    auto split_ratio = [](const Box& b, Box& l, Box& r, int d, double a) {
        int mid = b.lo[d] + static_cast<int>(b.len(d) * a + 0.5) - 1;
        l = b; r = b;
        l.hi[d] = mid;
        r.lo[d] = mid + 1;
    };
    int v, rm;
    unsigned d = 0;
    for (v = nsplits, rm = 0; v > 1; v /= 2) rm += (v % 2);
    if (rm == 0) {
        std::vector<Box> copy(nsplits);
        for (v = 1; v < nsplits; v *= 2) {
            while (!is_parallel[d])
                if (++d == 3) d = 0;
            for (int i = 0, j = 0; i < v; ++i, j += 2) split_half(dom[i], copy[j], copy[j + 1], d);
            std::copy(copy.begin(), copy.begin() + v * 2, dom.begin());
            if (++d == 3) d = 0;
        }
    } else {
        int vtot = 1;
        for (v = 1; v < 2 * nsplits; ++v) {
            int v1, v2;
            for (v2 = v, v1 = 1; v2 > 1; v2 /= 2) v1 = 2 * v1 + (v2 % 2);
            int vl = 0, vr = nsplits;
            while (v1 > 1) {
                if ((v1 % 2) == 1)
                    vl = vl + (vr - vl) / 2;
                else
                    vr = vl + (vr - vl) / 2;
                v1 /= 2;
            }
            v2 = vl + (vr - vl) / 2;
            if (v2 > vl) {
                double a = v2 - vl;
                a /= vr - vl;
                vr          = v2;
                Box left    = dom[vl];
                double lmax = 0, len;
                int dd_sel  = -1;
                for (int dd = 0; dd < 3; ++dd)
                    if (is_parallel[dd])
                        if ((len = left.len(dd)) > lmax) {
                            lmax   = len;
                            dd_sel = dd;
                        }
                Box temp;
                split_ratio(dom[vl], temp, dom[vr], dd_sel, a);
                dom[vl] = temp;
                ++vtot;
            }
        }
        if (vtot != nsplits) return -1;
    }
    for (int r = 0; r < nsplits; ++r)
        for (int k = 0; k < 3; ++k) {
            boxes_out[r * 6 + k]     = dom[r].lo[k];
            boxes_out[r * 6 + 3 + k] = dom[r].hi[k];
        }
    return 0;
}

// ---------------------------------------------------------------------------
// FieldLayout::findNeighbors / findPeriodicNeighbors / addNeighbors / getBounds,
// src/FieldLayout/FieldLayout.hpp:203-341.  One entry per (component, neighbour):
//   comp  : base-3 component index (digit d: 0 lower, 1 upper, 2 parallel)
//   rank  : neighbour rank
//   send  : lo[3],hi[3) local ghosted indices of MY interior strip (INTERNAL_TO_HALO send)
//   recv  : lo[3],hi[3) local ghosted indices of MY ghost strip    (INTERNAL_TO_HALO recv)
// ---------------------------------------------------------------------------
struct NbrEntry {
    int comp, rank;
    int send_lo[3], send_hi[3], recv_lo[3], recv_hi[3];
};

struct LayoutCtx {
    Box gdom;
    std::vector<Box> boxes;
    int nghost;
    bool periodic;
};

static void get_bounds(const Box& nd1, const Box& nd2, const Box& offset, int nghost, int lo[3],
                       int hi[3]) {
    Box gnd     = grow(nd2, nghost);
    Box overlap = intersect(gnd, nd1);
    for (int i = 0; i < 3; ++i) {
        lo[i] = (overlap.lo[i] - offset.lo[i]) + nghost;
        hi[i] = (overlap.hi[i] - offset.lo[i]) + nghost + 1;
    }
}

static void add_neighbors(const Box& gnd, const Box& nd, const Box& ndNeighbor, const Box& isect,
                          int nghost, int rank, std::vector<NbrEntry>& out) {
    NbrEntry e;
    get_bounds(nd, ndNeighbor, nd, nghost, e.send_lo, e.send_hi);
    get_bounds(ndNeighbor, nd, nd, nghost, e.recv_lo, e.recv_hi);
    int index = 0;
    for (int d = 0, digit = 1; d < 3; ++d, digit *= 3) {
        if (isect.len(d) == nghost) {
            if (gnd.lo[d] != isect.lo[d]) index += digit;
        } else {
            index += 2 * digit;
        }
    }
    e.comp = index;
    e.rank = rank;
    out.push_back(e);
}

static int periodic_offset(const LayoutCtx& L, const Box& nd, int d, int k) {
    const int period = L.gdom.len(d);
    if (k == 0) {
        if (nd.hi[d] == L.gdom.hi[d]) return -period;
    } else {
        if (nd.lo[d] == L.gdom.lo[d]) return period;
    }
    return 0;
}

static void find_periodic(const LayoutCtx& L, const Box& local, Box& grown, Box& nbr, int rank,
                          std::map<unsigned, int>& offsets, unsigned d0, unsigned codim,
                          std::vector<NbrEntry>& out) {
    for (unsigned d = d0; d < 3; ++d) {
        for (int k = 0; k < 2; ++k) {
            int offset = offsets[d] = periodic_offset(L, local, d, k);
            if (offset == 0) continue;
            grown.lo[d] += offset;
            grown.hi[d] += offset;
            if (touches(grown, nbr)) {
                Box isect = intersect(grown, nbr);
                for (auto& [dd, off] : offsets) {
                    nbr.lo[dd] -= off;
                    nbr.hi[dd] -= off;
                }
                add_neighbors(grown, local, nbr, isect, L.nghost, rank, out);
                for (auto& [dd, off] : offsets) {
                    nbr.lo[dd] += off;
                    nbr.hi[dd] += off;
                }
            }
            if (codim + 1 < 3) find_periodic(L, local, grown, nbr, rank, offsets, d + 1, codim + 1, out);
            grown.lo[d] -= offset;
            grown.hi[d] -= offset;
            offsets.erase(d);
        }
    }
}

static std::vector<NbrEntry> find_neighbors(const LayoutCtx& L, int my) {
    std::vector<NbrEntry> out;
    const Box& nd = L.boxes[my];
    Box gnd       = grow(nd, L.nghost);
    for (int rank = 0; rank < (int)L.boxes.size(); ++rank) {
        if (rank == my) continue;
        Box nbr = L.boxes[rank];
        if (touches(gnd, nbr)) {
            Box isect = intersect(gnd, nbr);
            add_neighbors(gnd, nd, nbr, isect, L.nghost, rank, out);
        }
        if (L.periodic) {
            std::map<unsigned, int> offsets;
            find_periodic(L, nd, gnd, nbr, rank, offsets, 0, 0, out);
        }
    }
    // neighbors_m[index] lists are filled in rank-ascending discovery order; a stable
    // sort by component reproduces the per-component vectors concatenated by index.
    std::stable_sort(out.begin(), out.end(),
                     [](const NbrEntry& a, const NbrEntry& b) { return a.comp < b.comp; });
    return out;
}

static LayoutCtx make_layout(const int ng[3], int nranks, const int* boxes, int nghost,
                             int periodic) {
    LayoutCtx L;
    for (int d = 0; d < 3; ++d) {
        L.gdom.lo[d] = 0;
        L.gdom.hi[d] = ng[d] - 1;
    }
    L.boxes.resize(nranks);
    for (int r = 0; r < nranks; ++r)
        for (int k = 0; k < 3; ++k) {
            L.boxes[r].lo[k] = boxes[r * 6 + k];
            L.boxes[r].hi[k] = boxes[r * 6 + 3 + k];
        }
    L.nghost   = nghost;
    L.periodic = periodic != 0;
    return L;
}

// Returns the number of entries; fills out[i*14 + ...] = comp, rank, send_lo[3], send_hi[3],
// recv_lo[3], recv_hi[3] (up to max_entries).
int orc_neighbors(const int ng[3], int nranks, const int* boxes, int nghost, int periodic, int my,
                  int* out, int max_entries) {
    LayoutCtx L = make_layout(ng, nranks, boxes, nghost, periodic);
    auto v      = find_neighbors(L, my);
    for (int i = 0; i < (int)v.size() && i < max_entries; ++i) {
        int* o = out + i * 14;
        o[0]   = v[i].comp;
        o[1]   = v[i].rank;
        for (int d = 0; d < 3; ++d) {
            o[2 + d]  = v[i].send_lo[d];
            o[5 + d]  = v[i].send_hi[d];
            o[8 + d]  = v[i].recv_lo[d];
            o[11 + d] = v[i].recv_hi[d];
        }
    }
    return (int)v.size();
}

// FieldLayout::getMatchingIndex, src/FieldLayout/FieldLayout.hpp:25-38
int orc_matching_index(int index) {
    static const int digit_swap[3] = {1, 0, 2};
    int match                      = 0;
    for (unsigned d = 1; d < 27; d *= 3) {
        match += digit_swap[index % 3] * d;
        index /= 3;
    }
    return match;
}

// ---------------------------------------------------------------------------
// HaloCells::exchangeBoundaries simulated for ALL ranks inside one process,
// src/Field/HaloCells.hpp:109-242 (+ pack :244-265, unpack :267-285).  Every rank
// first packs + "sends" (tag = HALO + component), then receives in component
// order matching tag HALO + getMatchingIndex(component) from that source, FIFO
// per (source, dest, tag) as MPI guarantees.
// mode 0 = fillHalo (INTERNAL_TO_HALO, assign): send my `send` strip, unpack into `recv`.
// mode 1 = accumulateHalo (HALO_TO_INTERNAL, +=): send my `recv` strip, add into `send`.
// fields[r] is rank r's ghosted field (ncomp doubles per cell).
// ---------------------------------------------------------------------------
void orc_halo_exchange(const int ng[3], int nranks, const int* boxes, int nghost, int periodic,
                       double** fields, int ncomp, int mode) {
    LayoutCtx L = make_layout(ng, nranks, boxes, nghost, periodic);
    std::vector<std::vector<NbrEntry>> nb(nranks);
    for (int r = 0; r < nranks; ++r) nb[r] = find_neighbors(L, r);
    using Key = std::tuple<int, int, int>;  // src, dst, tag
    std::map<Key, std::vector<std::vector<double>>> mail;
    auto ext = [&](int r, int d) { return L.boxes[r].len(d) + 2 * nghost; };
    for (int r = 0; r < nranks; ++r) {
        const long e0 = ext(r, 0), e1 = ext(r, 1);
        for (auto& e : nb[r]) {
            const int* lo = mode == 0 ? e.send_lo : e.recv_lo;
            const int* hi = mode == 0 ? e.send_hi : e.recv_hi;
            std::vector<double> buf;
            // pack order (HaloPackFunctor, :18-35): l = a0 + s0*(a1 + s1*a2), i.e. dim 0 fastest
            for (int k = lo[2]; k < hi[2]; ++k)
                for (int j = lo[1]; j < hi[1]; ++j)
                    for (int i = lo[0]; i < hi[0]; ++i)
                        for (int c = 0; c < ncomp; ++c)
                            buf.push_back(fields[r][(i + e0 * (j + e1 * k)) * ncomp + c]);
            mail[Key(r, e.rank, e.comp)].push_back(std::move(buf));
        }
    }
    std::map<Key, size_t> cursor;
    for (int r = 0; r < nranks; ++r) {
        const long e0 = ext(r, 0), e1 = ext(r, 1);
        for (auto& e : nb[r]) {
            const int* lo = mode == 0 ? e.recv_lo : e.send_lo;
            const int* hi = mode == 0 ? e.recv_hi : e.send_hi;
            Key key(e.rank, r, orc_matching_index(e.comp));
            auto& q = mail[key];
            size_t& cur = cursor[key];
            if (cur >= q.size()) {
                std::fprintf(stderr, "orc_halo_exchange: unmatched recv %d<-%d comp %d\n", r, e.rank,
                             e.comp);
                std::abort();
            }
            const std::vector<double>& buf = q[cur++];
            size_t expect = (size_t)(hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]) * ncomp;
            if (buf.size() != expect) {
                std::fprintf(stderr, "orc_halo_exchange: size mismatch %zu vs %zu\n", buf.size(),
                             expect);
                std::abort();
            }
            size_t l = 0;
            for (int k = lo[2]; k < hi[2]; ++k)
                for (int j = lo[1]; j < hi[1]; ++j)
                    for (int i = lo[0]; i < hi[0]; ++i)
                        for (int c = 0; c < ncomp; ++c) {
                            double& t = fields[r][(i + e0 * (j + e1 * k)) * ncomp + c];
                            if (mode == 0)
                                t = buf[l++];
                            else
                                t += buf[l++];
                        }
        }
    }
}

// ---------------------------------------------------------------------------
// Ownership: ParticleSpatialLayout::locateParticlesPacked destRankOf lambda,
// src/Particle/ParticleSpatialLayout.hpp:372-395 with positionInRegion (:324-330:
// pos > min && pos <= max) and positionInRegionInclusive (:316-322).  Regions are
// RegionLayout::convertNDIndex boxes: min = first*h + origin, max = (last+1)*h +
// origin (src/Region/RegionLayout.hpp:68-98, src/Meshes/UniformCartesian.h:45-53).
// Search order own -> (neighbours, all ranks ascending) -> inclusive ascending ->
// own.  The neighbour list only reorders a search over DISJOINT strict regions,
// so "own, then all ranks ascending" visits the same unique match.
// ---------------------------------------------------------------------------
void orc_regions(const int ng[3], int nranks, const int* boxes, const double origin[3],
                 const double h[3], double* regions /*[nranks][6]: min0..2,max0..2*/) {
    (void)ng;
    for (int r = 0; r < nranks; ++r)
        for (int d = 0; d < 3; ++d) {
            int first            = boxes[r * 6 + d];
            int last             = boxes[r * 6 + 3 + d] + 1;
            regions[r * 6 + d]     = first * h[d] + origin[d];
            regions[r * 6 + 3 + d] = last * h[d] + origin[d];
        }
}

void orc_locate(int nranks, const double* regions, int my, long n, const double* x,
                const double* y, const double* z, int* dest) {
    auto in_strict = [&](int r, const double p[3]) {
        for (int d = 0; d < 3; ++d)
            if (!(p[d] > regions[r * 6 + d])) return false;
        for (int d = 0; d < 3; ++d)
            if (!(p[d] <= regions[r * 6 + 3 + d])) return false;
        return true;
    };
    auto in_incl = [&](int r, const double p[3]) {
        for (int d = 0; d < 3; ++d)
            if (!(p[d] >= regions[r * 6 + d])) return false;
        for (int d = 0; d < 3; ++d)
            if (!(p[d] <= regions[r * 6 + 3 + d])) return false;
        return true;
    };
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const double p[3] = {x[i], y[i], z[i]};
        int dst           = -1;
        if (in_strict(my, p)) dst = my;
        for (int r = 0; dst < 0 && r < nranks; ++r)
            if (in_strict(r, p)) dst = r;
        for (int r = 0; dst < 0 && r < nranks; ++r)
            if (in_incl(r, p)) dst = r;
        if (dst < 0) dst = my;
        dest[i] = dst;
    }
}

// BareField::sum over interior cells, src/Field/BareField.hpp:224-240 (serial order).
double orc_field_sum(const double* v, const int ext[3], int nghost) {
    double s = 0.0;
    for (int k = nghost; k < ext[2] - nghost; ++k)
        for (int j = nghost; j < ext[1] - nghost; ++j)
            for (int i = nghost; i < ext[0] - nghost; ++i)
                s += v[i + (long)ext[0] * (j + (long)ext[1] * k)];
    return s;
}

// AlpineManager::getDensity, demos/alpine/AlpineManager.h:225-245:
//   rho = rho / cellVolume;  rho = rho - (Q / size)   on interior cells
// (BareField::operator=(Expression) runs over getRangePolicy(view, nghost), BareField.hpp:195-203)
void orc_density(double* v, const int ext[3], int nghost, double cell_volume, double q_over_size) {
    for (int k = nghost; k < ext[2] - nghost; ++k)
        for (int j = nghost; j < ext[1] - nghost; ++j)
            for (int i = nghost; i < ext[0] - nghost; ++i) {
                long l = i + (long)ext[0] * (j + (long)ext[1] * k);
                v[l]   = v[l] / cell_volume;
                v[l]   = v[l] - q_over_size;
            }
}

// One fused reference-order PIC push on the CPU, used ONLY as the timed CPU baseline of
// bench.py (the reference performs these as separate passes in exactly this order,
// demos/alpine/LandauDampingManager.h:265-320: kick, drift, applyBC(x3), scatter, [solve],
// gather, kick).  Returns nothing; the solve is not part of the metric.
void orc_pic_step_nosolve(const orc_mesh* m, long n, double* x, double* y, double* z, double* px,
                          double* py, double* pz, double* ex, double* ey, double* ez,
                          double q_scalar, double dt, const double* efield, double* rho) {
    double* P[3] = {px, py, pz};
    double* R[3] = {x, y, z};
    double* E[3] = {ex, ey, ez};
    for (int d = 0; d < 3; ++d) orc_kick(n, P[d], E[d], 0.5 * dt, 1);
    for (int d = 0; d < 3; ++d) orc_drift(n, R[d], P[d], dt, 1);
    for (int d = 0; d < 3; ++d)
        orc_periodic_bc(n, R[d], m->origin[d], m->ng[d] * m->h[d] + m->origin[d], 1);
    const int ext[3] = {m->nl[0] + 2 * m->nghost, m->nl[1] + 2 * m->nghost,
                        m->nl[2] + 2 * m->nghost};
    const long cells = (long)ext[0] * ext[1] * ext[2];
    std::memset(rho, 0, sizeof(double) * cells);
    orc_scatter_cic(m, 0, n, x, y, z, nullptr, q_scalar, nullptr, rho, 1);
    const int serial[3] = {1, 1, 1};
    orc_halo_periodic(rho, ext, 1, m->nghost, serial, 1);
    orc_gather_cic(m, n, x, y, z, efield, 3, E, 0, 1);
    for (int d = 0; d < 3; ++d) orc_kick(n, P[d], E[d], 0.5 * dt, 1);
}

}  // extern "C"
