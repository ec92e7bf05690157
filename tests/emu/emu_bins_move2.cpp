// emu_bins_move2.cpp -- TEST INFRASTRUCTURE: the source text of bins_move2_kernel (ippl_b200/csrc/bins.cu, cut out between
// its markers by tests/test_kernel_text_cpu.py and passed as KERNEL_TEXT) compiled for the host and executed by a
// lock-step warp emulator: the 32 lanes of a warp are 32 host threads that meet at every warp intrinsic
// (__match_any_sync, __shfl_sync), warps and blocks run one after the other, atomics are real atomics.  Checks what the
// fused step needs from the bucket build: every particle lands once, inside the bucket of its tile, with all six
// attributes; the tile cursors end at the tiles' totals; what exceeds a bucket's capacity is not written anywhere.
// Not a performance model and not a product path.
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

struct Dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;
#define __global__
#define __launch_bounds__(n)

struct Warp {
    std::barrier<> bar{32};
    long long slot[32];
};
static thread_local Warp* warp_ = nullptr;
static thread_local unsigned lane_ = 0;

static unsigned __match_any_sync(unsigned, int v) {
    warp_->slot[lane_] = v;
    warp_->bar.arrive_and_wait();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l)
        if (warp_->slot[l] == v) m |= 1u << l;
    warp_->bar.arrive_and_wait();
    return m;
}
static int __shfl_sync(unsigned, int v, int src) {
    warp_->slot[lane_] = v;
    warp_->bar.arrive_and_wait();
    const int r = (int)warp_->slot[src & 31];
    warp_->bar.arrive_and_wait();
    return r;
}
static int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static int __popc(unsigned v) { return __builtin_popcount(v); }
static int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <class T>
static T __ldcs(const T* p) { return *p; }

struct SoA6 {
    const double* in[6];
    double* out[6];
};

#include KERNEL_TEXT

static void launch(int grid, long n, const int* keys, int* cursor, const int* start, const int* cap, SoA6 P) {
    blockDim.x = 256;
    gridDim.x  = (unsigned)grid;
    for (int b = 0; b < grid; ++b)
        for (int w = 0; w < 8; ++w) {
            Warp W;
            std::vector<std::thread> lanes;
            for (unsigned l = 0; l < 32; ++l)
                lanes.emplace_back([&, l] {
                    warp_ = &W; lane_ = l;
                    threadIdx.x = (unsigned)w * 32 + l; blockIdx.x = (unsigned)b;
                    bins_move2_kernel(n, keys, cursor, start, cap, P);
                });
            for (auto& t : lanes) t.join();
        }
}

static double attr(int a, long id) { return (double)id * 8.0 + a + 0.25; }

static int run_case(const char* name, long n, int ntiles, const std::vector<int>& tile_of, int shrink_tile) {
    std::vector<int> keys(n), total(ntiles, 0), start(ntiles), cap(ntiles), cursor(ntiles, 0);
    std::mt19937 rng(5);
    for (long i = 0; i < n; ++i) {
        keys[i] = tile_of[i] * 64 + (int)(rng() % 64);   // tile-major keys: key >> 6 is the tile
        ++total[tile_of[i]];
    }
    long run = 0;
    for (int t = 0; t < ntiles; ++t) {
        start[t] = (int)run;
        cap[t]   = total[t] + 7;
        if (t == shrink_tile) cap[t] = total[t] / 2;    // a bucket that is too small: the excess must not be written
        run += cap[t] + 3;                              // gaps between the buckets: nothing may land there
    }
    const long slots = run + 8;
    std::vector<double> in[6], out[6];
    SoA6 P;
    for (int a = 0; a < 6; ++a) {
        in[a].resize(n);
        out[a].assign(slots, -1.0);
        for (long i = 0; i < n; ++i) in[a][i] = attr(a, i);
        P.in[a]  = in[a].data();
        P.out[a] = out[a].data();
    }
    launch(3, n, keys.data(), cursor.data(), start.data(), cap.data(), P);
    int bad = 0;
    std::vector<char> seen(n, 0);
    long written = 0;
    for (int t = 0; t < ntiles && !bad; ++t) {
        if (cursor[t] != total[t]) { std::printf("%s: tile %d cursor %d != total %d\n", name, t, cursor[t], total[t]); bad = 1; }
        const int stored = std::min(total[t], cap[t]);
        for (int j = 0; j < cap[t] + 3 && !bad; ++j) {
            const long g = (long)start[t] + j;
            if (j >= stored) {
                for (int a = 0; a < 6; ++a)
                    if (out[a][g] != -1.0) { std::printf("%s: slot %ld behind the stored part of tile %d was written\n", name, g, t); bad = 1; }
                continue;
            }
            const long id = (long)((out[0][g] - 0.25) / 8.0);
            if (id < 0 || id >= n || seen[id] || tile_of[id] != t) { std::printf("%s: tile %d slot %d holds particle %ld\n", name, t, j, id); bad = 1; break; }
            seen[id] = 1;
            ++written;
            for (int a = 0; a < 6; ++a)
                if (out[a][g] != attr(a, id)) { std::printf("%s: attribute %d of particle %ld differs\n", name, a, id); bad = 1; }
        }
    }
    long expect = 0;
    for (int t = 0; t < ntiles; ++t) expect += std::min(total[t], cap[t]);
    if (!bad && written != expect) { std::printf("%s: %ld particles stored, expected %ld\n", name, written, expect); bad = 1; }
    std::printf("%s: n=%ld tiles=%d stored=%ld %s\n", name, n, ntiles, written, bad ? "FAILED" : "ok");
    return bad;
}

int main() {
    int bad = 0;
    std::mt19937 rng(11);
    {   // random tiles, n not a multiple of the warp size
        const long n = 20011; const int nt = 97;
        std::vector<int> t(n);
        for (auto& v : t) v = (int)(rng() % nt);
        bad |= run_case("random", n, nt, t, -1);
        bad |= run_case("random, one bucket too small", n, nt, t, 13);
    }
    {   // tile-ordered input: long runs of lanes with the same tile
        const long n = 15000; const int nt = 40;
        std::vector<int> t(n);
        for (long i = 0; i < n; ++i) t[i] = (int)(i * nt / n);
        bad |= run_case("sorted", n, nt, t, -1);
    }
    {   // every particle in one tile; fewer particles than one warp
        std::vector<int> t(5000, 3);
        bad |= run_case("one tile", 5000, 8, t, -1);
        std::vector<int> s(7, 1);
        bad |= run_case("seven particles", 7, 4, s, -1);
    }
    std::printf(bad ? "EMU_BINS_MOVE2_FAILED\n" : "EMU_BINS_MOVE2_OK\n");
    return bad;
}
