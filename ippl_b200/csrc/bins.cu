// bins.cu -- the cell-ordered particle store behind the fused step: bucket tables, device-side planning,
// initial build (counting sort into buckets), tail append, compaction to the contiguous API view.
#include <cub/device/device_scan.cuh>

#include "bins.h"
#include "keys.cuh"

namespace ipplb {

__device__ __forceinline__ int round16(int v) { return (v + 15) & ~15; }

// ---- planning (PLAN_CTAS co-resident CTAs with two grid barriers; any number of tiles) ------------------------
// o = buffer that was just written: its cursor counts every particle that WANTED the tile, including those
// that overflowed into the tail.  i = the other buffer, planned here as the next output: every bucket gets
// room for the tile's current total plus slack (the flux through a tile's faces is a few per cent per step).
constexpr int PLAN_CTAS = 64, PLAN_NT = 256;

struct PlanArgs {
    int nt;
    const int* cap_o;
    int *count_o, *state_o, *start_i, *cap_i, *count_i, *state_i, *misc;
    long long* scratch;  // [PLAN_CTAS][4] partial sums, then [4] barrier words (as long long)
    int* exit_cnt;  // [2][MAX_RANKS]: live counters of the step, snapshot of the last step (for the migration)
    int exit_ranks, seg_cap;
    int capacity, slack_div, slack_sqrt, slack_const, tail_reserve;
};

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (atomicAdd(bar, 0u) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

template <int N>
__device__ __forceinline__ void block_sum(long long (&v)[N], long long (*red)[PLAN_NT / 32]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        long long x = v[k];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[k][warp] = x;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        long long x = 0;
        for (int w = 0; w < PLAN_NT / 32; ++w) x += red[k][w];
        v[k] = x;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(PLAN_NT) bins_plan_kernel(const PlanArgs a) {
    __shared__ long long red[3][PLAN_NT / 32];
    __shared__ long long wsum[PLAN_NT / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5, b = blockIdx.x, G = gridDim.x;
    unsigned int* bar = reinterpret_cast<unsigned int*>(a.scratch + (size_t)PLAN_CTAS * 4);
    const int per = (a.nt + G - 1) / G;
    const int j0 = min(a.nt, b * per), j1 = min(a.nt, j0 + per);
    // phase 1: clamp the counts of the buffer just written, sums for the capacity budget
    long long v[3] = {0, 0, 0};
    for (int j = j0 + t; j < j1; j += PLAN_NT) {
        const int total = a.count_o[j];
        const int c     = min(total, a.cap_o[j]);
        a.count_o[j]    = c;      // particles really stored in the bucket
        a.cap_i[j]      = total;  // scratch for phase 2
        v[0] += round16(total);
        v[1] += total / a.slack_div + a.slack_sqrt * (int)sqrtf((float)total) + a.slack_const;
        v[2] += c;
    }
    block_sum(v, red);
    if (t < 3) a.scratch[b * 4 + t] = v[t];
    grid_barrier(bar, G);
    long long tot[3] = {0, 0, 0};
    for (int g = 0; g < G; ++g) {
        tot[0] += __ldcg(&a.scratch[g * 4 + 0]);
        tot[1] += __ldcg(&a.scratch[g * 4 + 1]);
        tot[2] += __ldcg(&a.scratch[g * 4 + 2]);
    }
    int flags = 0;
    const long long avail = (long long)a.capacity - a.tail_reserve - tot[0];
    double scale = 1.0;
    if (tot[1] > avail) {
        scale = avail > 0 ? (double)avail / (double)tot[1] : 0.0;
        flags |= IPPLB_FLAG_SLACK_SCALED;
    }
    // phase 2: wanted capacities, per-CTA sums
    long long mine[1] = {0};
    for (int j = j0 + t; j < j1; j += PLAN_NT) {
        const int total = a.cap_i[j];
        const int slack = total / a.slack_div + a.slack_sqrt * (int)sqrtf((float)total) + a.slack_const;
        const int want  = round16(total + (int)(slack * scale));
        a.cap_i[j]      = want;
        mine[0] += want;
    }
    block_sum(mine, red);
    if (t == 0) a.scratch[b * 4 + 3] = mine[0];
    grid_barrier(bar + 1, G);
    long long run = 0, all = 0;
    for (int g = 0; g < G; ++g) {
        const long long sgm = __ldcg(&a.scratch[g * 4 + 3]);
        if (g < b) run += sgm;
        all += sgm;
    }
    // phase 3: exclusive scan inside my segment, batches of PLAN_NT tiles
    for (int base = j0; base < j1; base += PLAN_NT) {
        const int j    = base + t;
        const int want = j < j1 ? a.cap_i[j] : 0;
        long long inc  = want;
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        long long off = 0, bt = 0;
        for (int w = 0; w < PLAN_NT / 32; ++w) {
            if (w < warp) off += wsum[w];
            bt += wsum[w];
        }
        if (j < j1) {
            // never plan a bucket beyond the arrays: what does not fit overflows into the tail / is flagged
            long long s = run + off + inc - want;
            if (s > a.capacity) s = a.capacity;
            a.start_i[j] = (int)s;
            a.cap_i[j]   = (int)(s + want <= a.capacity ? want : a.capacity - s);
            a.count_i[j] = 0;
        }
        run += bt;
        __syncthreads();
    }
    if (b == 0 && t == 0) {
        if (all > a.capacity) flags |= IPPLB_FLAG_CAPACITY;
        a.state_i[BS_TAIL_START] = (int)(all < a.capacity ? all : a.capacity);
        a.state_i[BS_TAIL_COUNT] = 0;
        // tail of the buffer just written: clamp the cursor to what fitted
        int tc         = a.state_o[BS_TAIL_COUNT];
        const int room = a.capacity - a.state_o[BS_TAIL_START];
        if (tc > room) tc = room > 0 ? room : 0;
        a.state_o[BS_TAIL_COUNT] = tc;
        a.misc[BM_ST_TOTAL]      = (int)(tot[2] + tc);
        a.misc[BM_ST_BUCKETED]   = (int)tot[2];
        a.misc[BM_ST_TAIL]       = tc;
        // leavers per destination rank: snapshot for ipplb_bins_migrate, live counters re-armed for the next step
        int nexit = 0;
        for (int r = 0; r < a.exit_ranks; ++r) {
            const int c = a.exit_cnt[r];
            nexit += c;
            if (c > a.seg_cap) flags |= IPPLB_FLAG_EXIT_OVERFLOW;
            a.exit_cnt[MAX_RANKS + r] = c;
            a.exit_cnt[r]             = 0;
        }
        a.misc[BM_ST_EXIT]       = nexit;
        a.misc[BM_ST_TAIL_START] = a.state_o[BS_TAIL_START];
        // error bits are sticky until the next ipplb_bins_build: a status read after K steps sees a drop in any of them
        a.misc[BM_ST_FLAGS]      = (a.misc[BM_ST_FLAGS] & 7) | a.misc[BM_FLAGS] | flags;
        // the step's scratch words (work counter, flags) are re-armed here: no memset in front of the next step
        a.misc[BM_WORK] = 0; a.misc[BM_EXIT] = 0; a.misc[BM_FLAGS] = 0; a.misc[BM_SPARE] = 0;
    }
    // the last CTA to get here resets the barrier words for the next launch
    if (t == 0) {
        __threadfence();
        if (atomicAdd(bar + 2, 1u) == (unsigned)G - 1) {
            bar[0] = 0;
            bar[1] = 0;
            bar[2] = 0;
        }
    }
}

// ---- build ---------------------------------------------------------------------------------------------------
// totals per tile from the per-cell offsets; written as the "cursor" of the buffer that bins_plan treats as
// just-written, with caps equal to the totals (nothing overflowed)
__global__ void tile_totals_kernel(int nt, const int* __restrict__ cell_off, int* __restrict__ cap_o,
                                   int* __restrict__ count_o) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
        const int tot = cell_off[(t + 1) * TILE_CELLS] - cell_off[t * TILE_CELLS];
        cap_o[t]      = tot;
        count_o[t]    = tot;
    }
}

__global__ void set_counts_kernel(int nt, const int* __restrict__ tot, const int* __restrict__ cap,
                                  int* __restrict__ count, int* __restrict__ misc) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
        count[t] = min(tot[t], cap[t]);
        if (tot[t] > cap[t]) atomicOr(&misc[BM_FLAGS], IPPLB_FLAG_CAPACITY);
    }
}

struct SoA6 {
    const double* in[6];
    double* out[6];
};

__global__ void __launch_bounds__(256)
bins_move_kernel(long n, const int* __restrict__ keys, const int* __restrict__ cell_off,
                 int* __restrict__ cell_cursor, const int* __restrict__ start, const int* __restrict__ cap,
                 SoA6 P) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int key  = keys[i];
        const int tile = key >> 6;
        const int in_t = cell_off[key] - cell_off[tile << 6] + atomicAdd(&cell_cursor[key], 1);
        if (in_t < cap[tile]) {
            const long g = (long)start[tile] + in_t;
#pragma unroll
            for (int a = 0; a < 6; ++a) P.out[a][g] = P.in[a][i];
        }
    }
}

// Build variant 2: the buckets are filled in ARRIVAL order (one cursor per tile, not per cell -- the order inside a bucket
// is free: the fused step appends arrivals and overflow the same way).  What it changes against bins_move_kernel:
//   * the write position of a tile advances sequentially over the kernel, so the 8-byte stores of consecutive arrivals
//     land in the same 32-byte sectors and merge in L2 before they reach HBM (per-cell positions keep 64x more
//     partially written sectors open than L2 holds: every store becomes a read-modify-write of a sector);
//   * lanes of a warp that go to the same tile take their slots with ONE atomic (match.any + popc): cell-ordered or
//     nearly ordered input (a re-bucketing after ipplb_bins_compact / an ORB repartition) costs one atomic per run;
//   * no per-cell offset look-ups (two random 4-byte loads per particle), inputs read with evict-first loads so that
//     the streaming side does not push the half-written sectors out of L2.
// Same tables, same guarantees as variant 1 (count <= cap, nothing dropped: cap >= the tile's total at build time).
// [host-emulation begin: bins_move2_kernel]  (tests/test_kernel_text_cpu.py compiles the text between the markers for the
// host with a lock-step warp emulator, tests/emu/emu_bins_move2.cpp, and runs it)
__global__ void __launch_bounds__(256)
bins_move2_kernel(long n, const int* __restrict__ keys, int* __restrict__ tile_cursor, const int* __restrict__ start,
                  const int* __restrict__ cap, SoA6 P) {
    const unsigned lane = threadIdx.x & 31u;
    const long stride   = (long)gridDim.x * blockDim.x;
    // all lanes of a warp iterate together (match / shuffle need the full warp)
    for (long base = (long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const long i     = base + lane;
        const bool valid = i < n;
        const int tile   = valid ? (__ldcs(&keys[i]) >> 6) : (-1 - (int)lane);   // idle lanes: distinct groups of one
        double v[6];
        if (valid) {
#pragma unroll
            for (int a = 0; a < 6; ++a) v[a] = __ldcs(&P.in[a][i]);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, tile);
        const int leader     = __ffs(peers) - 1;
        const int rank       = __popc(peers & ((1u << lane) - 1u));
        int first = 0;
        if (valid && (int)lane == leader) first = atomicAdd(&tile_cursor[tile], __popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        if (valid) {
            const int in_t = first + rank;
            if (in_t < cap[tile]) {
                const long g = (long)start[tile] + in_t;
#pragma unroll
                for (int a = 0; a < 6; ++a) P.out[a][g] = v[a];
            }
        }
    }
}
// [host-emulation end: bins_move2_kernel]

// ---- append / compact -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bins_append_kernel(long count, const int* __restrict__ state, int capacity, SoA6 P, int* __restrict__ misc) {
    const long base   = (long)state[BS_TAIL_START] + state[BS_TAIL_COUNT];
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const long g = base + i;
        if (g < capacity) {
#pragma unroll
            for (int a = 0; a < 6; ++a) P.out[a][g] = P.in[a][i];
        } else if (i == count - 1) {
            atomicOr(&misc[BM_ST_FLAGS], IPPLB_FLAG_CAPACITY);
        }
    }
}

__global__ void bins_append_commit_kernel(long count, int* __restrict__ state, int capacity,
                                          int* __restrict__ misc) {
    const long room = (long)capacity - state[BS_TAIL_START] - state[BS_TAIL_COUNT];
    const int add   = (int)(count < room ? count : (room > 0 ? room : 0));
    state[BS_TAIL_COUNT] += add;
    misc[BM_ST_TOTAL] += add;
    misc[BM_ST_TAIL] += add;
}

// one CTA per tile: bucket -> contiguous; the last CTAs copy the tail
__global__ void __launch_bounds__(256)
bins_compact_kernel(int nt, const int* __restrict__ start, const int* __restrict__ count,
                    const int* __restrict__ off, const int* __restrict__ state, SoA6 P) {
    for (int b = blockIdx.x; b < nt + 64; b += gridDim.x) {
        long src, dst;
        int cnt;
        if (b < nt) {
            src = start[b];
            dst = off[b];
            cnt = count[b];
            for (int j = threadIdx.x; j < cnt; j += 256) {
#pragma unroll
                for (int a = 0; a < 6; ++a) P.out[a][dst + j] = P.in[a][src + j];
            }
        } else {
            const int part = b - nt;  // 64 CTAs share the tail
            src = state[BS_TAIL_START];
            dst = off[nt];
            cnt = state[BS_TAIL_COUNT];
            for (long j = (long)part * 256 + threadIdx.x; j < cnt; j += 64 * 256) {
#pragma unroll
                for (int a = 0; a < 6; ++a) P.out[a][dst + j] = P.in[a][src + j];
            }
        }
    }
}

int bins_plan(ipplb_ctx* ctx, ipplb_bins* b, int o, int seg_cap) {
    const int i = 1 - o;
    PlanArgs a;
    a.nt = b->ntiles;
    a.cap_o = b->cap(o); a.count_o = b->count(o); a.state_o = b->state(o);
    a.start_i = b->start(i); a.cap_i = b->cap(i); a.count_i = b->count(i); a.state_i = b->state(i);
    a.misc = b->misc();
    a.scratch = b->d_plan;
    a.exit_cnt = b->d_exit_cnt; a.exit_ranks = b->exit_ranks; a.seg_cap = seg_cap;
    a.capacity = (int)b->capacity;
    a.slack_div = b->slack_div; a.slack_sqrt = b->slack_sqrt; a.slack_const = b->slack_const;
    a.tail_reserve = (int)(b->capacity / 50) + 1024;
    // all PLAN_CTAS CTAs are co-resident (64 <= SM count, launched after the previous kernel of the stream
    // has drained), which is what the in-kernel grid barriers rely on
    bins_plan_kernel<<<PLAN_CTAS, PLAN_NT, 0, ctx->stream>>>(a);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

static int grid_for(const ipplb_ctx* ctx, long n, int block, int per_sm) {
    long want = (n + block - 1) / block;
    long cap  = (long)ctx->num_sms * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_bins_create(ipplb_ctx* ctx, const ipplb_mesh* mesh, long capacity, ipplb_bins** out) {
    IPPLB_REQUIRE(ctx && mesh && out && capacity > 0, "bins_create: bad arguments");
    IPPLB_REQUIRE(capacity < (1L << 31) - 64, "bins_create: capacity must stay below 2^31 elements per rank");
    ipplb_bins* b = new ipplb_bins();
    b->mesh     = *mesh;
    b->ntx      = tiles_along(mesh->nl[0]);
    b->nty      = tiles_along(mesh->nl[1]);
    b->ntz      = tiles_along(mesh->nl[2]);
    b->ntiles   = b->ntx * b->nty * b->ntz;
    b->ncells   = (long)b->ntiles * TILE_CELLS;
    b->capacity = capacity & ~15L;
    cudaError_t e = cudaMalloc(&b->d_tab, sizeof(int) * b->tab_words());
    if (e == cudaSuccess) e = cudaMalloc(&b->d_cell, sizeof(int) * (size_t)(b->ncells + 1));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_plan, sizeof(long long) * (PLAN_CTAS * 4 + 4));
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_plan, 0, sizeof(long long) * (PLAN_CTAS * 4 + 4), ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_exit_cnt, sizeof(int) * 2 * MAX_RANKS);
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_exit_cnt, 0, sizeof(int) * 2 * MAX_RANKS, ctx->stream);
    if (e == cudaSuccess) e = cudaMallocHost(&b->h_status, sizeof(int) * (BM_WORDS + MAX_RANKS));
    if (e == cudaSuccess) e = cudaMemsetAsync(b->d_tab, 0, sizeof(int) * b->tab_words(), ctx->stream);
    if (e != cudaSuccess) {
        set_error("bins_create: %s", cudaGetErrorString(e));
        ipplb_bins_destroy(b);
        return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? IPPLB_ERR_NO_DEVICE : IPPLB_ERR_CUDA;
    }
    *out = b;
    return IPPLB_OK;
}

int ipplb_bins_destroy(ipplb_bins* b) {
    if (!b) return IPPLB_OK;
    if (b->d_tab) cudaFree(b->d_tab);
    if (b->d_cell) cudaFree(b->d_cell);
    if (b->d_plan) cudaFree(b->d_plan);
    if (b->d_exit_cnt) cudaFree(b->d_exit_cnt);
    if (b->h_status) cudaFreeHost(b->h_status);
    if (b->ev) {
        for (int i = 0; i < 2 * ipplb_bins::NEV; ++i) cudaEventDestroy(b->ev[i]);
        delete[] b->ev;
    }
    delete b;
    return IPPLB_OK;
}

int ipplb_bins_ntiles(const ipplb_bins* b) { return b ? b->ntiles : -1; }

int ipplb_bins_set_timing(ipplb_bins* b, int on) {
    IPPLB_REQUIRE(b, "bins_set_timing: bad arguments");
    if (on && !b->ev) {
        b->ev = new cudaEvent_t[2 * ipplb_bins::NEV];
        for (int i = 0; i < 2 * ipplb_bins::NEV; ++i) IPPLB_CUDA(cudaEventCreate(&b->ev[i]));
    }
    b->timing = on ? 1 : 0;
    b->ev_n   = 0;
    return IPPLB_OK;
}

int ipplb_bins_set_build_variant(ipplb_bins* b, int variant) {
    IPPLB_REQUIRE(b && (variant == 1 || variant == 2), "bins_set_build_variant: variant is 1 or 2");
    b->build_variant = variant;
    return IPPLB_OK;
}

int ipplb_bins_kernel_ms(ipplb_ctx* ctx, ipplb_bins* b, double* ms_out, int max_out, int* n_out) {
    IPPLB_REQUIRE(ctx && b && n_out, "bins_kernel_ms: bad arguments");
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n = b->ev_n < ipplb_bins::NEV ? b->ev_n : ipplb_bins::NEV;
    *n_out      = n;
    for (int i = 0; i < n && i < max_out; ++i) {
        float ms = 0.f;
        IPPLB_CUDA(cudaEventElapsedTime(&ms, b->ev[2 * i], b->ev[2 * i + 1]));
        ms_out[i] = ms;
    }
    return IPPLB_OK;
}

int ipplb_bins_build(ipplb_ctx* ctx, ipplb_bins* b, const ipplb_particles* in, ipplb_particles* out) {
    IPPLB_REQUIRE(ctx && b && in && out, "bins_build: bad arguments");
    IPPLB_REQUIRE(in->q == nullptr, "bins_build: per-particle charge arrays take the unfused path");
    IPPLB_REQUIRE(out->capacity >= b->capacity, "bins_build: output bundle smaller than the bins capacity");
    IPPLB_REQUIRE(in->n <= b->capacity, "bins_build: more particles than capacity");
    const long n = in->n;
    MeshDev m    = make_mesh_dev(&b->mesh);
    int rc;
    if ((rc = ensure(ctx, ctx->keys, sizeof(int) * (size_t)(n > 0 ? n : 1)))) return rc;
    if ((rc = ensure(ctx, ctx->counts, sizeof(int) * (size_t)(b->ncells + 1)))) return rc;
    int* keys   = (int*)ctx->keys.ptr;
    int* counts = (int*)ctx->counts.ptr;
    IPPLB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (b->ncells + 1), ctx->stream));
    IPPLB_CUDA(cudaMemsetAsync(b->misc(), 0, sizeof(int) * BM_WORDS, ctx->stream));
    if (n > 0) {
        sort_keys_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(m, n, in->x, in->y, in->z, keys,
                                                                             counts);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, b->d_cell, (int)(b->ncells + 1), ctx->stream);
    if ((rc = ensure(ctx, ctx->cub_tmp, tmp_bytes))) return rc;
    IPPLB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.ptr, tmp_bytes, counts, b->d_cell,
                                             (int)(b->ncells + 1), ctx->stream));
    ctx->launches += 2;
    const int c = b->cur, o = 1 - b->cur;
    const int tg = (b->ntiles + 255) / 256;
    // totals -> "just written" pseudo buffer o; plan c from them; place the particles; plan o for the first step
    tile_totals_kernel<<<tg, 256, 0, ctx->stream>>>(b->ntiles, b->d_cell, b->cap(o), b->count(o));
    IPPLB_CHECK_LAUNCH(ctx);
    IPPLB_CUDA(cudaMemsetAsync(b->state(o), 0, sizeof(int) * BS_WORDS, ctx->stream));
    if ((rc = bins_plan(ctx, b, o))) return rc;
    set_counts_kernel<<<tg, 256, 0, ctx->stream>>>(b->ntiles, b->count(o), b->cap(c), b->count(c), b->misc());
    IPPLB_CHECK_LAUNCH(ctx);
    IPPLB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (b->ncells + 1), ctx->stream));
    if (n > 0) {
        SoA6 P;
        const double* pi[6] = {in->x, in->y, in->z, in->px, in->py, in->pz};
        double* po[6]       = {out->x, out->y, out->z, out->px, out->py, out->pz};
        for (int a = 0; a < 6; ++a) {
            IPPLB_REQUIRE(pi[a] && po[a] && pi[a] != po[a], "bins_build: null or aliased particle arrays");
            P.in[a]  = pi[a];
            P.out[a] = po[a];
        }
        if (b->build_variant == 2)
            bins_move2_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(n, keys, counts, b->start(c), b->cap(c), P);
        else
            bins_move_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(n, keys, b->d_cell, counts,
                                                                                 b->start(c), b->cap(c), P);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    if ((rc = bins_plan(ctx, b, c))) return rc;
    b->built      = true;
    out->n        = n;
    out->q        = nullptr;
    out->q_scalar = in->q_scalar;
    return IPPLB_OK;
}

int ipplb_bins_status(ipplb_ctx* ctx, ipplb_bins* b, long* n_local, long* n_tail, long* n_exit, int* flags) {
    IPPLB_REQUIRE(ctx && b, "bins_status: bad arguments");
    IPPLB_CUDA(cudaMemcpyAsync(b->h_status, b->misc(), sizeof(int) * BM_WORDS, cudaMemcpyDeviceToHost,
                               ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_local) *n_local = b->h_status[BM_ST_TOTAL];
    if (n_tail) *n_tail = b->h_status[BM_ST_TAIL];
    if (n_exit) *n_exit = b->h_status[BM_ST_EXIT];
    if (flags) *flags = b->h_status[BM_ST_FLAGS];
    return IPPLB_OK;
}

int ipplb_bins_append(ipplb_ctx* ctx, ipplb_bins* b, ipplb_particles* cur, const double* const src[6],
                      long count) {
    IPPLB_REQUIRE(ctx && b && cur && src && count >= 0, "bins_append: bad arguments");
    IPPLB_REQUIRE(b->built, "bins_append: call ipplb_bins_build first");
    if (count == 0) return IPPLB_OK;
    SoA6 P;
    double* po[6] = {cur->x, cur->y, cur->z, cur->px, cur->py, cur->pz};
    for (int a = 0; a < 6; ++a) {
        P.in[a]  = src[a];
        P.out[a] = po[a];
    }
    bins_append_kernel<<<grid_for(ctx, count, 256, 8), 256, 0, ctx->stream>>>(count, b->state(b->cur),
                                                                              (int)b->capacity, P, b->misc());
    IPPLB_CHECK_LAUNCH(ctx);
    bins_append_commit_kernel<<<1, 1, 0, ctx->stream>>>(count, b->state(b->cur), (int)b->capacity, b->misc());
    IPPLB_CHECK_LAUNCH(ctx);
    cur->n += count;
    return IPPLB_OK;
}

int ipplb_bins_compact(ipplb_ctx* ctx, ipplb_bins* b, const ipplb_particles* cur, ipplb_particles* out) {
    IPPLB_REQUIRE(ctx && b && cur && out, "bins_compact: bad arguments");
    IPPLB_REQUIRE(b->built, "bins_compact: call ipplb_bins_build first");
    long n = 0;
    int flags = 0, rc;
    if ((rc = ipplb_bins_status(ctx, b, &n, nullptr, nullptr, &flags))) return rc;
    IPPLB_REQUIRE(out->capacity >= n, "bins_compact: output capacity too small");
    size_t tmp_bytes = 0;
    int* off = b->d_cell;  // scratch, ntiles + 1 <= ncells + 1
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, b->count(b->cur), off, b->ntiles + 1, ctx->stream);
    if ((rc = ensure(ctx, ctx->cub_tmp, tmp_bytes))) return rc;
    // ntiles + 1 items so that off[ntiles] = sum of the counts (the extra input word is another table word
    // inside d_tab and does not influence any output that is read)
    IPPLB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.ptr, tmp_bytes, b->count(b->cur), off, b->ntiles + 1,
                                             ctx->stream));
    ctx->launches += 2;
    SoA6 P;
    const double* pi[6] = {cur->x, cur->y, cur->z, cur->px, cur->py, cur->pz};
    double* po[6]       = {out->x, out->y, out->z, out->px, out->py, out->pz};
    for (int a = 0; a < 6; ++a) {
        IPPLB_REQUIRE(pi[a] && po[a] && pi[a] != po[a], "bins_compact: null or aliased particle arrays");
        P.in[a]  = pi[a];
        P.out[a] = po[a];
    }
    const int grid = b->ntiles + 64 < ctx->num_sms * 16 ? b->ntiles + 64 : ctx->num_sms * 16;
    bins_compact_kernel<<<grid, 256, 0, ctx->stream>>>(b->ntiles, b->start(b->cur), b->count(b->cur), off,
                                                       b->state(b->cur), P);
    IPPLB_CHECK_LAUNCH(ctx);
    out->n        = n;
    out->q        = nullptr;
    out->q_scalar = cur->q_scalar;
    return IPPLB_OK;
}

int ipplb_bins_tables(ipplb_ctx* ctx, ipplb_bins* b, int* start_host, int* cap_host, int* count_host) {
    IPPLB_REQUIRE(ctx && b, "bins_tables: bad arguments");
    const size_t bytes = sizeof(int) * (size_t)b->ntiles;
    if (start_host) IPPLB_CUDA(cudaMemcpyAsync(start_host, b->start(b->cur), bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (cap_host) IPPLB_CUDA(cudaMemcpyAsync(cap_host, b->cap(b->cur), bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (count_host) IPPLB_CUDA(cudaMemcpyAsync(count_host, b->count(b->cur), bytes, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return IPPLB_OK;
}

}  // extern "C"
