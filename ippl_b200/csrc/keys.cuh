// keys.cuh -- cell keys + histogram kernel shared by the counting sort (particles.cu) and the bucket build (bins.cu).
#pragma once
#include "cic.cuh"

namespace ipplb {

// ------------------------------------------------------------------------------------------------
// Counting sort by cell (integer keys).  Pass 1: keys + histogram (run-aggregated atomics);
// exclusive scan (CUB); pass 2: claim a slot from the per-cell cursor and move all attributes.
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
sort_keys_kernel(MeshDev m, long n, const double* __restrict__ x, const double* __restrict__ y,
                 const double* __restrict__ z, int* __restrict__ keys, int* __restrict__ counts) {
    const unsigned lane = threadIdx.x & 31u;
    const long stride   = (long)gridDim.x * blockDim.x;
    long base = (long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    for (; base < n; base += stride) {
        const long i = base + lane;
        int key      = -1 - (int)lane;
        if (i < n) {
            Cic c;
            cic_setup(m, x[i], y[i], z[i], c);
            key     = cell_key(m, c.a);
            keys[i] = key;
        }
        const int prev    = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head   = (lane == 0) || (prev != key);
        const unsigned hm = __ballot_sync(0xffffffffu, head);
        if (head && i < n) {
            // run length = distance to the next head (or to lane 32)
            const unsigned above = hm & ~((2u << lane) - 1u);  // heads strictly above this lane
            const int next       = above ? __ffs(above) - 1 : 32;
            int len              = next - (int)lane;
            const long rem       = n - i;
            if (len > rem) len = (int)rem;
            atomicAdd(&counts[key], len);
        }
    }
}

}  // namespace ipplb
