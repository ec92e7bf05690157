// emu_shim_reduce.cpp -- TEST INFRASTRUCTURE: the device code behind Kokkos::parallel_for / parallel_reduce of
// include/ippl/KokkosShim.cuh (for_kernel, reduce1_kernel, reduce2_kernel, block_join, atomic_join and the reducers; cut out
// between their markers by tests/test_kernel_text_cpu.py and passed as REDUCER_TEXT / KERNEL_TEXT) compiled for the host and
// executed by a lock-step BLOCK emulator: the 256 threads of a block are 256 host threads; the lanes of a warp meet at every
// warp shuffle, the whole block at every __syncthreads; __shared__ storage is one array per block run; atomics are real.
// The shim's own host-emulation mode replaces these kernels by host loops, so this is the only place their text runs
// without a GPU.  Checks sums, maxima and minima over ranges shorter than a warp, not a multiple of the block, and longer
// than one sweep of the grid, with one and two reducers; and that for_kernel visits every index exactly once.
#include <atomic>
#include <barrier>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

struct Dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim, gridDim;
#define __global__
#define __device__
#define __host__
#define __launch_bounds__(n)
#define __shared__ static

struct Block {
    std::barrier<> all{256};
    std::barrier<> warp[8] = {std::barrier<>(32), std::barrier<>(32), std::barrier<>(32), std::barrier<>(32),
                              std::barrier<>(32), std::barrier<>(32), std::barrier<>(32), std::barrier<>(32)};
    double slot[256];
};
static thread_local Block* blk_ = nullptr;

static void __syncthreads() { blk_->all.arrive_and_wait(); }
static double __shfl_xor_sync(unsigned, double v, int o) {
    const unsigned t = threadIdx.x, w = t >> 5;
    blk_->slot[t] = v;
    blk_->warp[w].arrive_and_wait();
    const double r = blk_->slot[(t & ~31u) | ((t ^ (unsigned)o) & 31u)];
    blk_->warp[w].arrive_and_wait();
    return r;
}
static unsigned long long atomicCAS(unsigned long long* a, unsigned long long expected, unsigned long long desired) {
    __atomic_compare_exchange_n(a, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;   // the value found, like the device function
}
static double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }

#include REDUCER_TEXT
#include KERNEL_TEXT

template <class K>
static void launch(int grid, K kernel_for_thread) {
    blockDim.x = 256;
    gridDim.x  = (unsigned)grid;
    for (int b = 0; b < grid; ++b) {
        Block B;
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < 256; ++t)
            ts.emplace_back([&, t] {
                blk_ = &B; threadIdx.x = t; blockIdx.x = (unsigned)b;
                kernel_for_thread();
            });
        for (auto& t : ts) t.join();
    }
}
static int grid_for(long n) {   // shim::grid_for(n, 256)
    long g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

static double value(long i) { return std::sin(0.37 * (double)i) * (1.0 + (double)(i % 7)); }

int main() {
    int bad = 0;
    const long sizes[] = {1, 5, 31, 32, 33, 255, 256, 257, 1000, 4099};
    for (long n : sizes) {
        const long b = 3, e = 3 + n;   // a range that does not start at 0
        double sum = 0.0, mx = -DBL_MAX, mn = DBL_MAX;
        for (long i = b; i < e; ++i) { const double v = value(i); sum += v; mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
        auto f1 = [](std::size_t i, double& a) { a += value((long)i); };
        auto fm = [](std::size_t i, double& a) { const double v = value((long)i); a = v > a ? v : a; };
        auto f2 = [](std::size_t i, double& a, double& m) { const double v = value((long)i); a += v; m = v < m ? v : m; };
        for (int grid : {grid_for(n), 2}) {   // the shim's grid, and a small one that forces the grid-stride loop
            double s1 = Sum<double>::identity(), m1 = Max<double>::identity(), s2[2] = {Sum<double>::identity(), Min<double>::identity()};
            launch(grid, [&] { reduce1_kernel<decltype(f1), Sum<double>>(b, e, f1, &s1); });
            launch(grid, [&] { reduce1_kernel<decltype(fm), Max<double>>(b, e, fm, &m1); });
            launch(grid, [&] { reduce2_kernel<decltype(f2), Sum<double>, Min<double>>(b, e, f2, s2); });
            const double tol = 1e-12 * (1.0 + std::fabs(sum)) + 1e-13 * (double)n;
            if (std::fabs(s1 - sum) > tol || m1 != mx || std::fabs(s2[0] - sum) > tol || s2[1] != mn) {
                std::printf("n=%ld grid=%d: sum %.17g / %.17g (want %.17g), max %.17g (want %.17g), min %.17g (want %.17g)\n", n, grid, s1, s2[0],
                            sum, m1, mx, s2[1], mn);
                bad = 1;
            }
        }
        std::vector<int> hits(e + 4, 0);
        int* h = hits.data();
        auto ff = [h](std::size_t i) { __atomic_fetch_add(&h[i], 1, __ATOMIC_RELAXED); };
        launch(grid_for(n), [&] { for_kernel<decltype(ff)>(b, e, ff); });
        for (long i = 0; i < e + 4; ++i)
            if (hits[i] != (i >= b && i < e ? 1 : 0)) { std::printf("for_kernel n=%ld: index %ld visited %d times\n", n, i, hits[i]); bad = 1; break; }
    }
    std::printf(bad ? "EMU_SHIM_REDUCE_FAILED\n" : "EMU_SHIM_REDUCE_OK\n");
    return bad;
}
