// philox.h -- Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) as a
// counter-based uniform stream: key = 64-bit seed, counter = (particle id lo, hi, dimension, block).  Integer
// arithmetic only up to the final exact conversion to double, so any independent implementation of the published
// algorithm (the test oracle has its own, in numpy) draws bit-identical uniforms; pinned by the Random123
// known-answer vectors in tests/.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PHILOX_HD __host__ __device__ inline
#else
#define PHILOX_HD inline
#endif

PHILOX_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// two uniforms in [0, 1) with 53 random bits each: exact in double
PHILOX_HD void philox_uniform2(uint64_t seed, uint64_t id, uint32_t dim, uint32_t block, double u[2]) {
    uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), dim, block};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)c[0] << 32) | c[1], b = ((uint64_t)c[2] << 32) | c[3];
    u[0] = (double)(a >> 11) * 0x1.0p-53;
    u[1] = (double)(b >> 11) * 0x1.0p-53;
}
PHILOX_HD double philox_uniform(uint64_t seed, uint64_t id, uint32_t dim, uint32_t block) {
    double u[2];
    philox_uniform2(seed, id, dim, block, u);
    return u[0];
}

// three standard normals by Box-Muller on the uniforms of dimensions 3 and 4 of the particle's counter
PHILOX_HD void philox_normal3(uint64_t seed, uint64_t id, double g[3]) {
    double a[2], b[2];
    philox_uniform2(seed, id, 3u, 0u, a);
    philox_uniform2(seed, id, 4u, 0u, b);
    const double two_pi = 6.283185307179586476925286766559;
    const double r0 = sqrt(-2.0 * log(1.0 - a[0])), r1 = sqrt(-2.0 * log(1.0 - b[0]));  // 1 - u in (0, 1]
    g[0] = r0 * cos(two_pi * a[1]);
    g[1] = r0 * sin(two_pi * a[1]);
    g[2] = r1 * cos(two_pi * b[1]);
}
