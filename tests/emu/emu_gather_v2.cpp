// emu_gather_v2.cpp -- TEST INFRASTRUCTURE: ippl_b200/csrc/push.cuh (the device functions themselves, not a restatement)
// compiled for the host with the handful of intrinsics they use spelled as plain C++ (IEEE fp64 without contraction:
// -ffp-contract=off), and gather_point3_vec (variant 2: 16-byte loads per x-pair of stencil nodes) compared bit for bit
// with gather_point<3> (variant 1, the kernel that is pinned to the oracle on the GPU) for every particle of several
// meshes -- even / odd ghosted extents, a sub-domain, particles on the lower / upper corners and faces.  Every load is checked against the field's
// byte range (and the field ends at a PROT_NONE page).  Not a product path.
#include <sys/mman.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include <cuda_runtime.h>
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int __double2int_rz(double a) { return (int)a; }
static long loads16 = 0, loads8 = 0;
static const char *field_lo = nullptr, *field_hi = nullptr;   // every load must lie inside the field
static void check_range(const void* p, size_t bytes) {
    if ((const char*)p < field_lo || (const char*)p + bytes > field_hi) { std::printf("load outside the field\n"); std::exit(2); }
}
static inline double2 __ldg(const double2* p) {
    if ((uintptr_t)p & 15) { std::printf("misaligned 16-byte load\n"); std::exit(2); }
    check_range(p, 16);
    ++loads16;
    return *p;
}
static inline double __ldg(const double* p) { check_range(p, 8); ++loads8; return *p; }

#include "ippl_b200/csrc/push.cuh"

using namespace ipplb;

// a field of `n` doubles whose last element is the last 8 bytes before an inaccessible page
struct GuardedField {
    char* base = nullptr;
    size_t bytes = 0;
    double* f = nullptr;
    explicit GuardedField(size_t n) {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        const size_t need = ((n * sizeof(double) + page - 1) / page) * page;
        bytes = need + page;
        base  = (char*)mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        mprotect(base + need, page, PROT_NONE);
        f = (double*)(base + need - n * sizeof(double));
    }
    ~GuardedField() { munmap(base, bytes); }
};

static int run_mesh(const int ng[3], const int first[3], const int nl[3]) {
    ipplb_mesh M{};
    const double origin[3] = {0.25, -1.0, 3.0}, h[3] = {0.5, 0.125, 1.5};
    for (int d = 0; d < 3; ++d) { M.ng[d] = ng[d]; M.first[d] = first[d]; M.nl[d] = nl[d]; M.origin[d] = origin[d]; M.h[d] = h[d]; }
    M.nghost = 1;
    const MeshDev m = make_mesh_dev(&M);
    const size_t cells = (size_t)m.ex * m.ey * m.ez;
    // The field starts on a 16-byte boundary (device allocations do; the product falls back to variant 1 otherwise).  With
    // an odd number of doubles it then cannot also end at the guard page: 8 spare bytes sit in between, and the range
    // check in the load wrappers covers them.
    GuardedField G(cells * 3 + ((cells * 3) & 1));
    double* f = G.f;
    field_lo  = (const char*)f;
    field_hi  = (const char*)(f + cells * 3);
    std::mt19937_64 rng(17);
    std::normal_distribution<double> N01(0.0, 1.0);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    for (size_t i = 0; i < cells * 3; ++i) f[i] = N01(rng);
    double lo[3];
    for (int d = 0; d < 3; ++d) lo[d] = origin[d] + first[d] * h[d];
    const int n = 20000;
    int bad = 0;
    for (int i = 0; i < n && !bad; ++i) {
        double r[3];
        for (int d = 0; d < 3; ++d) r[d] = lo[d] + U(rng) * nl[d] * h[d];
        if (i < 8)          // the 8 corners of the box (the upper ones reach the last node of the ghosted box)
            for (int d = 0; d < 3; ++d) r[d] = (i >> d) & 1 ? lo[d] + nl[d] * h[d] : lo[d];
        else if (i < 11)    // on a node, on a cell centre, on a face
            for (int d = 0; d < 3; ++d) r[d] = lo[d] + (i == 8 ? 3.0 : i == 9 ? 2.5 : (d == 1 ? 0.0 : 1.75)) * h[d];
        Cic c;
        cic_setup(m, r[0], r[1], r[2], c);
        double g1[3], g2[3];
        gather_point<3>(m, c, f, g1);
        gather_point3_vec(m, c, f, g2);
        if (std::memcmp(g1, g2, sizeof(g1)) != 0) {
            std::printf("particle %d (%.17g %.17g %.17g): variant 1 (%.17g %.17g %.17g) != variant 2 (%.17g %.17g %.17g)\n", i, r[0], r[1],
                        r[2], g1[0], g1[1], g1[2], g2[0], g2[1], g2[2]);
            bad = 1;
        }
    }
    std::printf("mesh %dx%dx%d box first (%d %d %d) n (%d %d %d), ghosted %dx%dx%d (%zu cells): %s\n", ng[0], ng[1], ng[2], first[0], first[1],
                first[2], nl[0], nl[1], nl[2], m.ex, m.ey, m.ez, cells, bad ? "FAILED" : "ok");
    return bad;
}

int main() {
    const int cases[5][9] = {{12, 10, 8, 0, 0, 0, 12, 10, 8}, {12, 10, 8, 6, 0, 4, 6, 10, 4}, {11, 10, 8, 0, 0, 0, 11, 10, 8},
                             {11, 9, 7, 0, 0, 0, 11, 9, 7},   {13, 9, 7, 2, 1, 0, 9, 7, 7}};
    int bad = 0;
    for (auto& c : cases) bad |= run_mesh(c, c + 3, c + 6);
    // per particle: variant 1 issues 24 8-byte loads, variant 2 12 16-byte loads and an 8-byte one per row that starts 8 bytes
    // behind a 16-byte boundary
    std::printf("loads: %ld 16-byte, %ld 8-byte over %d particles\n", loads16, loads8, 5 * 20000);
    if (loads16 != 12L * 5 * 20000) { std::printf("unexpected number of 16-byte loads\n"); bad = 1; }
    std::printf(bad ? "EMU_GATHER_V2_FAILED\n" : "EMU_GATHER_V2_OK\n");
    return bad;
}
