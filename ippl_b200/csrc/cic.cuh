// cic.cuh -- device helpers for the cloud-in-cell arithmetic, bit-faithful to the reference.
//
// Every floating-point operation goes through the __d*_rn intrinsics so nvcc can neither
// contract (FMA) nor reassociate them: the sequence of IEEE operations is exactly the one the
// reference's C++ spells out (src/Particle/ParticleAttrib.hpp:174-179, src/Interpolation/CIC.hpp:6-66),
// which makes gather / push / keys bit-comparable with the CPU oracle.
#pragma once
#include "common.cuh"

namespace ipplb {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// l = (x - origin) * invdx + 0.5; index = (int) l; whi = l - index; wlo = 1.0 - whi
__device__ __forceinline__ void cic_axis(double pos, double origin, double invdx, int& index,
                                         double& whi) {
    double l = dadd(dmul(dsub(pos, origin), invdx), 0.5);
    index    = __double2int_rz(l);
    whi      = dsub(l, (double)index);
}

struct Cic {
    double whi[3];
    int a[3];  // args = index - lDom.first() + nghost  (ghosted local index of the upper node)
};

__device__ __forceinline__ void cic_setup(const MeshDev& m, double x, double y, double z, Cic& c) {
    int idx;
    cic_axis(x, m.origin[0], m.invdx[0], idx, c.whi[0]);
    c.a[0] = idx - m.first[0] + m.nghost;
    cic_axis(y, m.origin[1], m.invdx[1], idx, c.whi[1]);
    c.a[1] = idx - m.first[1] + m.nghost;
    cic_axis(z, m.origin[2], m.invdx[2], idx, c.whi[2]);
    c.a[2] = idx - m.first[2] + m.nghost;
}

// weight of stencil point p: bit d set -> wlo[d] (node args[d]-1), clear -> whi[d] (node args[d]);
// product is the right fold w0 * (w1 * w2)  (CIC.hpp:6-14, 32-33)
__device__ __forceinline__ double cic_weight(const double whi[3], int p) {
    double w0 = (p & 1) ? dsub(1.0, whi[0]) : whi[0];
    double w1 = (p & 2) ? dsub(1.0, whi[1]) : whi[1];
    double w2 = (p & 4) ? dsub(1.0, whi[2]) : whi[2];
    return dmul(w0, dmul(w1, w2));
}

__device__ __forceinline__ long cic_node(const MeshDev& m, const int a[3], int p) {
    long i = a[0] - (p & 1);
    long j = a[1] - ((p >> 1) & 1);
    long k = a[2] - ((p >> 2) & 1);
    return i + (long)m.ex * (j + (long)m.ey * k);
}

// cell key used by the sort: (index - first) per dim in [0, nl[d]] (a particle sitting exactly on
// the upper region boundary has index == first + nl), x fastest.
__device__ __forceinline__ int cell_key(const MeshDev& m, const int a[3]) {
    int cx = a[0] - m.nghost, cy = a[1] - m.nghost, cz = a[2] - m.nghost;
    return cx + (m.nl[0] + 1) * (cy + (m.nl[1] + 1) * cz);
}

// PeriodicBC::operator(), src/Particle/ParticleBC.h:73-76
__device__ __forceinline__ double periodic_wrap(double v, double extent, double middle) {
    double t = __ddiv_rn(dmul(dsub(v, middle), 2.0), extent);
    return dsub(v, dmul(extent, (double)__double2int_rz(t)));
}

}  // namespace ipplb
