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

// A rank that owns a whole periodic axis: ghost layer g aliases the opposite interior layer (what
// HaloCells::applyPeriodicSerialDim copies / adds, src/Field/HaloCells.hpp:297-336)
__device__ __forceinline__ int wrap_axis(int g, int nl, int nghost) {
    return g < nghost ? g + nl : (g >= nl + nghost ? g - nl : g);
}
__device__ __forceinline__ long cic_node_wrapped(const MeshDev& m, const int a[3], int p) {
    const long i = wrap_axis(a[0] - (p & 1), m.nl[0], m.nghost);
    const long j = wrap_axis(a[1] - ((p >> 1) & 1), m.nl[1], m.nghost);
    const long k = wrap_axis(a[2] - ((p >> 2) & 1), m.nl[2], m.nghost);
    return i + (long)m.ex * (j + (long)m.ey * k);
}

// Cell key used by the sort: c[d] = index[d] - first[d] in [0, nl[d]] (a particle sitting exactly on
// the upper region boundary has index == first + nl).  Keys are TILE-MAJOR: the key space is cut into
// 4x4x4-cell tiles, key = tile_id * 64 + cell_in_tile (x fastest inside the tile and across tiles), so
// any run of consecutive sorted particles is spatially compact in 3-D (what the fused push/move/deposit
// kernel needs for its shared-memory window).
constexpr int TILE = 4;
constexpr int TILE_CELLS = TILE * TILE * TILE;

__host__ __device__ __forceinline__ int tiles_along(int nl) { return (nl + 1 + TILE - 1) / TILE; }

__device__ __forceinline__ int cell_key_c(const MeshDev& m, int cx, int cy, int cz) {
    const int ntx = tiles_along(m.nl[0]), nty = tiles_along(m.nl[1]);
    const int tile = (cx >> 2) + ntx * ((cy >> 2) + nty * (cz >> 2));
    return tile * TILE_CELLS + ((cz & 3) << 4) + ((cy & 3) << 2) + (cx & 3);
}

__device__ __forceinline__ int cell_key(const MeshDev& m, const int a[3]) {
    return cell_key_c(m, a[0] - m.nghost, a[1] - m.nghost, a[2] - m.nghost);
}

// inverse: key -> args (ghosted local index of the cell's upper node)
__device__ __forceinline__ void key_to_args(const MeshDev& m, int key, int a[3]) {
    const int ntx = tiles_along(m.nl[0]), nty = tiles_along(m.nl[1]);
    const int tile = key >> 6, in = key & 63;
    const int tx = tile % ntx, ty = (tile / ntx) % nty, tz = tile / (ntx * nty);
    a[0] = tx * TILE + (in & 3) + m.nghost;
    a[1] = ty * TILE + ((in >> 2) & 3) + m.nghost;
    a[2] = tz * TILE + (in >> 4) + m.nghost;
}

// PeriodicBC::operator(), src/Particle/ParticleBC.h:73-76.  half_extent = extent * 0.5 (exact), formed on the host.
// The wrap itself (an fp64 division) is out of line: ~1 % of the particles per step cross the domain boundary.
#ifndef IPPLB_FAR_ATTR
#define IPPLB_FAR_ATTR __forceinline__
#endif
static __device__ IPPLB_FAR_ATTR double periodic_wrap_far(double v, double extent, double off) {
    const double num = dmul(off, 2.0);
    const double t   = __ddiv_rn(num, extent);
    return dsub(v, dmul(extent, (double)__double2int_rz(t)));
}
__device__ __forceinline__ double periodic_wrap(double v, double extent, double middle, double half_extent) {
    const double off = dsub(v, middle);
    // num = off * 2 (exact).  |num| < extent  =>  |num/extent| < 1 after correct rounding  =>  (int) = 0  =>
    // v - extent*0 == v: skip the fp64 division for particles that stay inside (bit-identical result); the test is
    // made on off against extent / 2 (both scalings by 2 are exact), which saves a multiply on the fast path
    if (fabs(off) < half_extent) return v;
    return periodic_wrap_far(v, extent, off);
}
__device__ __forceinline__ double periodic_wrap(double v, double extent, double middle) {
    return periodic_wrap(v, extent, middle, dmul(extent, 0.5));
}

}  // namespace ipplb
