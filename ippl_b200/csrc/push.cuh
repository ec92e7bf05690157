// push.cuh -- gather stencil + particle push device functions shared by the kernels.
#pragma once
#include <cmath>

#include "cic.cuh"

namespace ipplb {

// ------------------------------------------------------------------------------------------------
// Gather (API-faithful): E_p = sum_p w_p * F(node_p), right fold like CIC.hpp:63-65.
// ------------------------------------------------------------------------------------------------
// node(p) -> linear ghosted node index of stencil point p
template <int NCOMP, typename NodeFn>
__device__ __forceinline__ void gather_point_at(const double whi[3], NodeFn node, const double* __restrict__ f,
                                                double out[NCOMP]) {
    double w[8];
    long id[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        w[p]  = cic_weight(whi, p);
        id[p] = node(p) * NCOMP;
    }
#pragma unroll
    for (int d = 0; d < NCOMP; ++d) {
        double acc = dmul(w[7], __ldg(&f[id[7] + d]));
#pragma unroll
        for (int p = 6; p >= 0; --p) acc = dadd(dmul(w[p], __ldg(&f[id[p] + d])), acc);
        out[d] = acc;
    }
}
template <int NCOMP>
__device__ __forceinline__ void gather_point(const MeshDev& m, const Cic& c,
                                             const double* __restrict__ f, double out[NCOMP]) {
    gather_point_at<NCOMP>(c.whi, [&](int p) { return cic_node(m, c.a, p); }, f, out);
}

// Gather, variant 2 (three components, AoS-3 field): the two x-neighbours of a stencil row are 6 contiguous doubles that
// start on a 16-byte boundary or 8 bytes behind one, so a row is three 16-byte loads (+ one 8-byte load in the odd
// case) instead of six 8-byte loads: 14 load instructions per particle instead of 24.  On unordered particles every lane
// of a load touches its own sector, and the number of such load wavefronts is what the kernel pays for.  Same weights,
// same fold order (CIC.hpp:63-65): bit-identical to gather_point<3>.  `f` must be 16-byte aligned.
__device__ __forceinline__ void load_pair3(const double* __restrict__ f, long node_lo, double v[6]) {
    const long e  = node_lo * 3;   // first element of the pair
    const long e0 = e & ~1L;       // the 16-byte aligned element at or one before it
    const double2* q = reinterpret_cast<const double2*>(f + e0);
    const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    if (e & 1) {
        v[0] = a.y; v[1] = b.x; v[2] = b.y; v[3] = c.x; v[4] = c.y;
        v[5] = __ldg(f + e + 5);   // (not a fourth 16-byte load: its upper half may lie behind the end of the field)
    } else {
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
    }
}
__device__ __forceinline__ void gather_point3_vec(const MeshDev& m, const Cic& c, const double* __restrict__ f, double out[3]) {
    double w[8], v[4][6];
#pragma unroll
    for (int p = 0; p < 8; ++p) w[p] = cic_weight(c.whi, p);
#pragma unroll
    for (int r = 0; r < 4; ++r) load_pair3(f, cic_node(m, c.a, 2 * r + 1), v[r]);   // stencil point 2r+1 = (a0-1, ., .), 2r = its +x neighbour
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double acc = dmul(w[7], v[3][d]);
#pragma unroll
        for (int p = 6; p >= 0; --p) acc = dadd(dmul(w[p], (p & 1) ? v[p >> 1][d] : v[p >> 1][3 + d]), acc);
        out[d] = acc;
    }
}

struct PushDev {
    int kind, do_kick2, do_kick1, do_drift, do_bc;
    double dt, c;          // c = 0.5*dt
    double lo[3], ext[3], mid[3], hext[3];  // periodic BC constants (ParticleBC.h:43-52); hext = ext / 2
    double p_origin[3], p_half_len[3], cxy, cz, alpha, Bext, DrInv, aB;  // penning
};

__device__ __forceinline__ void penning_field(const PushDev& P, double x, double y, double z,
                                              const double E[3], double Ee[3]) {
    // Eext_x = -(x - origin - 0.5*length) * (V0 / (2 * length_z^2)), etc.; then += E
    Ee[0] = dadd(dmul(-dsub(dsub(x, P.p_origin[0]), P.p_half_len[0]), P.cxy), E[0]);
    Ee[1] = dadd(dmul(-dsub(dsub(y, P.p_origin[1]), P.p_half_len[1]), P.cxy), E[1]);
    Ee[2] = dadd(dmul(dsub(dsub(z, P.p_origin[2]), P.p_half_len[2]), P.cz), E[2]);
}

__device__ __forceinline__ void push_particle(const PushDev& P, double r[3], double p[3],
                                              const double E[3]) {
    if (P.kind == IPPLB_PUSH_LEAPFROG) {
        if (P.do_kick2) {
#pragma unroll
            for (int d = 0; d < 3; ++d) p[d] = dsub(p[d], dmul(P.c, E[d]));
        }
        if (P.do_kick1) {
#pragma unroll
            for (int d = 0; d < 3; ++d) p[d] = dsub(p[d], dmul(P.c, E[d]));
        }
    } else {
        double Ee[3];
        penning_field(P, r[0], r[1], r[2], E, Ee);
        const double a = P.alpha, B = P.Bext;
        if (P.do_kick2) {
            // P0 = DrInv * (P0 + a * (Ex + P1*B + a*B*Ey));  a*B*Ey parses as (a*B)*Ey
            p[0] = dmul(P.DrInv,
                        dadd(p[0], dmul(a, dadd(dadd(Ee[0], dmul(p[1], B)), dmul(P.aB, Ee[1])))));
            p[1] = dmul(P.DrInv,
                        dadd(p[1], dmul(a, dsub(dsub(Ee[1], dmul(p[0], B)), dmul(P.aB, Ee[0])))));
            p[2] = dadd(p[2], dmul(a, Ee[2]));
        }
        if (P.do_kick1) {
            p[0] = dadd(p[0], dmul(a, dadd(Ee[0], dmul(p[1], B))));
            p[1] = dadd(p[1], dmul(a, dsub(Ee[1], dmul(p[0], B))));
            p[2] = dadd(p[2], dmul(a, Ee[2]));
        }
    }
    if (P.do_drift) {
#pragma unroll
        for (int d = 0; d < 3; ++d) r[d] = dadd(r[d], dmul(P.dt, p[d]));
    }
    if (P.do_bc) {
#pragma unroll
        for (int d = 0; d < 3; ++d) r[d] = periodic_wrap(r[d], P.ext[d], P.mid[d], P.hext[d]);
    }
}

// the steady-state leapfrog step with every sub-step on (kick2, kick1, drift, periodic BC): the same operations in the
// same order as push_particle, without the flag tests
__device__ __forceinline__ void push_leapfrog_full(const PushDev& P, double r[3], double p[3], const double E[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) p[d] = dsub(p[d], dmul(P.c, E[d]));
#pragma unroll
    for (int d = 0; d < 3; ++d) p[d] = dsub(p[d], dmul(P.c, E[d]));
#pragma unroll
    for (int d = 0; d < 3; ++d) r[d] = dadd(r[d], dmul(P.dt, p[d]));
#pragma unroll
    for (int d = 0; d < 3; ++d) r[d] = periodic_wrap(r[d], P.ext[d], P.mid[d], P.hext[d]);
}

inline PushDev make_push_dev(const ipplb_mesh* mesh, const ipplb_push* push) {
    PushDev P;
    std::memset(&P, 0, sizeof(P));
    P.kind     = push->kind;
    P.do_kick2 = push->do_kick2;
    P.do_kick1 = push->do_kick1;
    P.do_drift = push->do_drift;
    P.do_bc    = push->do_bc;
    P.dt       = push->dt;
    P.c        = 0.5 * push->dt;
    for (int d = 0; d < 3; ++d) {
        // region of the global domain: RegionLayout::convertNDIndex (min = 0*h + origin, max = N*h + origin)
        const double lo = 0 * mesh->h[d] + mesh->origin[d];
        const double hi = mesh->ng[d] * mesh->h[d] + mesh->origin[d];
        P.lo[d]  = lo;
        P.ext[d] = hi - lo;
        P.mid[d] = (lo + hi) / 2;
        P.hext[d] = P.ext[d] * 0.5;
    }
    if (push->kind == IPPLB_PUSH_PENNING) {
        const double l2 = std::pow(push->length[2], 2);
        for (int d = 0; d < 3; ++d) {
            P.p_origin[d]   = push->origin[d];
            P.p_half_len[d] = 0.5 * push->length[d];
        }
        P.cxy   = push->V0 / (2 * l2);
        P.cz    = push->V0 / (l2);
        P.alpha = push->alpha;
        P.Bext  = push->Bext;
        P.DrInv = push->DrInv;
        P.aB    = push->alpha * push->Bext;
    }
    return P;
}

}  // namespace ipplb
