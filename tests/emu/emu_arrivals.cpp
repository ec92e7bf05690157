// emu_arrivals.cpp -- TEST INFRASTRUCTURE: the source text of arrivals_p2p_kernel (ippl_b200/csrc/comm.cu: the kernel that
// drops a step's arrivals from the peer-memory inbox into their buckets on every rank of a multi-GPU run; cut out between
// its markers by tests/test_kernel_text_cpu.py and passed as KERNEL_TEXT) compiled for the host together with the product's
// own cic.cuh / bins.h and executed by a lock-step block emulator (256 host threads per block meeting at __syncthreads,
// real atomics).  The kernel ran on GPUs in round 2; its check that refuses records outside the rank's box was added
// afterwards without GPU access.  Checked here: every record inside the box lands exactly once, in the bucket of its tile
// or -- where the bucket is full -- in the tail, with all six values; records on the upper faces of the box (cell index ==
// nl) are accepted; records outside the box (zeroed memory, a non-finite position, the next rank's cell) are refused and
// flagged, and touch nothing; a full tail is flagged; the status words add up; the deposited charge is q per stored record.
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int __double2int_rz(double a) {   // cvt.rzi.s32.f64: NaN -> 0, saturating
    if (a != a) return 0;
    if (a >= 2147483647.0) return 2147483647;
    if (a <= -2147483648.0) return -2147483647 - 1;
    return (int)a;
}
struct Dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;
#define __launch_bounds__(n)
#undef __global__
#define __global__
#undef __shared__
#define __shared__ static

static std::barrier<>* block_bar = nullptr;
static void __syncthreads() { block_bar->arrive_and_wait(); }
static int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static int atomicSub(int* p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_RELAXED); }
static int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static double atomicAdd(double* p, double v) {
    uint64_t* a = reinterpret_cast<uint64_t*>(p);
    uint64_t old = __atomic_load_n(a, __ATOMIC_RELAXED), want;
    double cur;
    do {
        std::memcpy(&cur, &old, 8);
        cur += v;
        std::memcpy(&want, &cur, 8);
    } while (!__atomic_compare_exchange_n(a, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return cur - v;
}

#include "ippl_b200/csrc/bins.h"
#include "ippl_b200/csrc/cic.cuh"

namespace ipplb {
#include KERNEL_TEXT
}
using namespace ipplb;

static void launch(int grid, const ArriveArgs& a) {
    blockDim.x = 256;
    gridDim.x  = (unsigned)grid;
    for (int b = 0; b < grid; ++b) {
        std::barrier<> bar(256);
        block_bar = &bar;
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < 256; ++t)
            ts.emplace_back([&, t] { threadIdx.x = t; blockIdx.x = (unsigned)b; arrivals_p2p_kernel(a); });
        for (auto& t : ts) t.join();
    }
}

int main() {
    // rank 1 of 3: a box that does not start at cell 0 and whose sizes are not multiples of the tile size
    ipplb_mesh M{};
    const int ng[3] = {24, 18, 10}, first[3] = {9, 0, 3}, nl[3] = {9, 18, 7};
    for (int d = 0; d < 3; ++d) { M.ng[d] = ng[d]; M.first[d] = first[d]; M.nl[d] = nl[d]; M.origin[d] = 0.5 * d; M.h[d] = 0.25 + 0.125 * d; }
    M.nghost = 1;
    const MeshDev m = make_mesh_dev(&M);
    const int ntx = tiles_along(nl[0]), nty = tiles_along(nl[1]), ntz = tiles_along(nl[2]), nt = ntx * nty * ntz;
    const int nranks = 3, me = 1;
    const long seg_cap = 700;
    const double q = -0.75;
    int bad = 0;
    for (int scenario = 0; scenario < 3; ++scenario) {   // 0: clean arrivals; 1: + records outside the box; 2: tiny buckets and a tail that fills up
        std::mt19937_64 rng(3 + scenario);
        std::uniform_real_distribution<double> U(0.0, 1.0);
        std::vector<double> inbox((size_t)nranks * seg_cap * 6, 0.0);
        std::vector<int> matrix(nranks * nranks, 0);
        const int from[3] = {650, 0, 500};   // rank 1 sends nothing to itself here
        struct Rec { double v[6]; int tile; bool inside; };
        std::vector<Rec> recs;
        for (int s = 0; s < nranks; ++s) {
            matrix[s * nranks + me] = from[s];
            matrix[s * nranks + (me + 1) % nranks] = 123;   // other destinations: not this rank's business
            for (int j = 0; j < from[s]; ++j) {
                Rec r;
                for (int d = 0; d < 3; ++d) r.v[d] = M.origin[d] + (first[d] + U(rng) * nl[d]) * M.h[d];
                if (j % 97 == 5) r.v[0] = M.origin[0] + (first[0] + nl[0]) * M.h[0];     // on the upper x face: cell index == nl
                if (j % 89 == 7) for (int d = 0; d < 3; ++d) r.v[d] = M.origin[d] + (first[d] + nl[d]) * M.h[d];   // upper corner
                r.inside = true;
                if (scenario >= 1) {
                    if (j % 50 == 1) { r.v[0] = r.v[1] = r.v[2] = 0.0; r.inside = false; }                                  // unwritten memory
                    if (j % 50 == 2) { r.v[1] = std::numeric_limits<double>::quiet_NaN(); r.inside = false; }               // non-finite
                    if (j % 50 == 3) { r.v[0] = M.origin[0] + (first[0] + nl[0] + 1.0) * M.h[0]; r.inside = false; }         // the next rank's cell
                    if (j % 50 == 4) { r.v[2] = M.origin[2] + (first[2] - 2.0) * M.h[2]; r.inside = false; }                 // below the box in z
                }
                for (int d = 3; d < 6; ++d) r.v[d] = 1000.0 * s + j + 0.125 * d;
                int c[3];
                for (int d = 0; d < 3; ++d) c[d] = (int)((r.v[d] - M.origin[d]) * (1.0 / M.h[d]) + 0.5) - first[d];
                r.tile = r.inside ? (c[0] >> 2) + ntx * ((c[1] >> 2) + nty * (c[2] >> 2)) : -1;
                std::memcpy(&inbox[((size_t)s * seg_cap + j) * 6], r.v, sizeof(r.v));
                recs.push_back(r);
            }
        }
        // tables: bucket t at start[t] with room cap[t]; the tail behind the buckets
        std::vector<int> start(nt), cap(nt), count(nt, 0), state(BS_WORDS, 0), misc(BM_WORDS, 0);
        long run = 0;
        for (int t = 0; t < nt; ++t) { start[t] = (int)run; cap[t] = scenario == 2 ? 3 : 200; run += cap[t]; }
        const int tail_room = scenario == 2 ? 400 : 5000;
        const int capacity  = (int)run + tail_room;
        state[BS_TAIL_START] = (int)run;
        std::vector<double> out[6];
        for (auto& o : out) o.assign((size_t)capacity + 16, -7.0);
        std::vector<double> rho((size_t)m.ex * m.ey * m.ez, 0.0);
        ArriveArgs a;
        a.m = m; a.matrix = matrix.data(); a.nranks = nranks; a.me = me; a.inbox = inbox.data(); a.seg_cap = seg_cap;
        a.start = start.data(); a.cap = cap.data(); a.count = count.data(); a.state = state.data(); a.misc = misc.data();
        a.capacity = capacity; a.ntx = ntx; a.nty = nty;
        for (int k = 0; k < 6; ++k) a.out[k] = out[k].data();
        a.q = q; a.rho = rho.data();
        launch(3, a);
        // ---- checks
        long inside = 0, refused = 0;
        for (auto& r : recs) (r.inside ? inside : refused)++;
        long stored = 0, in_buckets = 0, in_tail = 0;
        std::vector<char> hit(recs.size(), 0);
        auto find = [&](long g, int want_tile) {   // which record sits in slot g?
            for (size_t i = 0; i < recs.size(); ++i) {
                if (hit[i] || !recs[i].inside) continue;
                bool same = true;
                for (int k = 0; k < 6 && same; ++k) same = out[k][g] == recs[i].v[k];
                if (same) {
                    if (want_tile >= 0 && recs[i].tile != want_tile) { std::printf("slot %ld: record of tile %d in bucket %d\n", g, recs[i].tile, want_tile); bad = 1; }
                    hit[i] = 1;
                    return true;
                }
            }
            return false;
        };
        for (int t = 0; t < nt; ++t) {
            if (count[t] > cap[t]) { std::printf("scenario %d: tile %d count %d > cap %d\n", scenario, t, count[t], cap[t]); bad = 1; }
            for (int j = 0; j < cap[t]; ++j) {
                const long g = (long)start[t] + j;
                if (j < count[t]) { if (!find(g, t)) { std::printf("scenario %d: bucket %d slot %d holds no arrival of that tile\n", scenario, t, j); bad = 1; } else { ++stored; ++in_buckets; } }
                else if (out[0][g] != -7.0) { std::printf("scenario %d: slot %ld behind bucket %d's count was written\n", scenario, g, t); bad = 1; }
            }
        }
        const int tc = state[BS_TAIL_COUNT];
        for (long g = state[BS_TAIL_START]; g < (long)capacity + 16; ++g) {
            const long j = g - state[BS_TAIL_START];
            if (j < tc) { if (!find(g, -1)) { std::printf("scenario %d: tail slot %ld holds no arrival\n", scenario, j); bad = 1; } else { ++stored; ++in_tail; } }
            else if (out[0][g] != -7.0) { std::printf("scenario %d: slot %ld behind the tail was written\n", scenario, g); bad = 1; }
        }
        const bool tail_full = scenario == 2;
        const int flags = misc[BM_ST_FLAGS];
        if (!tail_full && stored != inside) { std::printf("scenario %d: %ld stored, %ld inside the box\n", scenario, stored, inside); bad = 1; }
        if (tail_full && (in_tail != tail_room || !(flags & IPPLB_FLAG_CAPACITY))) { std::printf("scenario %d: tail %ld of %d, flags %d\n", scenario, in_tail, tail_room, flags); bad = 1; }
        if (!tail_full && (flags & IPPLB_FLAG_CAPACITY)) { std::printf("scenario %d: capacity flag without a full tail\n", scenario); bad = 1; }
        if (((flags & IPPLB_FLAG_INTERNAL) != 0) != (refused > 0)) { std::printf("scenario %d: %ld records outside the box, flags %d\n", scenario, refused, flags); bad = 1; }
        if (misc[BM_ST_BUCKETED] != in_buckets || misc[BM_ST_TAIL] != in_tail || misc[BM_ST_TOTAL] != stored) {
            std::printf("scenario %d: status words %d / %d / %d, found %ld / %ld / %ld\n", scenario, misc[BM_ST_BUCKETED], misc[BM_ST_TAIL], misc[BM_ST_TOTAL], in_buckets, in_tail, stored);
            bad = 1;
        }
        double sum = 0.0;
        for (double v : rho) sum += v;
        if (std::fabs(sum - q * (double)stored) > 1e-9 * (1.0 + std::fabs(q * stored))) { std::printf("scenario %d: deposited charge %.15g, want %.15g\n", scenario, sum, q * stored); bad = 1; }
        std::printf("scenario %d: %zu records, %ld inside the box, %ld refused, %ld in buckets, %ld in the tail, flags %d: %s\n", scenario, recs.size(),
                    inside, refused, in_buckets, in_tail, flags, bad ? "FAILED" : "ok");
    }
    std::printf(bad ? "EMU_ARRIVALS_FAILED\n" : "EMU_ARRIVALS_OK\n");
    return bad;
}
