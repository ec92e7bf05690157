// fused.cu -- the B200-first PIC step: two passes over the particles, no separate sort or scatter.
//
//   pass A  push_count : read R,P; gather E (L1/L2); kick/kick/drift/BC in registers; histogram of the
//                        NEW cell keys.  Nothing but the histogram is written.               (48 B/particle)
//   scan               : exclusive scan of the histogram -> new cell offsets
//   pass B  push_move_deposit : one CTA per 4x4x4-cell tile of the OLD order (dynamic tile scheduler).
//                        Re-reads R,P, recomputes the identical push, bins the particles by NEW cell in
//                        shared memory (integer atomics on a 12^3-cell window around the tile), reserves
//                        one contiguous run per (CTA, destination cell) from the global cell cursors,
//                        writes R,P out run by run (coalesced), and deposits the charge from the
//                        shared-memory-sorted particles with one thread per (cell, stencil node) --
//                        register accumulation, one RED.F64 per (cell,node).                  (96 B/particle)
//
// The particles leave pass B exactly cell-sorted (tile-major keys) with cell_offsets valid, so the next
// step's gather reads are warp-coherent.  E never round-trips through HBM and rho is produced without
// re-reading the particles: 144 B/particle-step of DRAM traffic instead of ~250 for push + sort + scatter.
// Positions and momenta are bit-identical to the unfused path (same device functions, same order).
#include <cub/device/device_scan.cuh>

#include "push.cuh"

namespace ipplb {

constexpr int WIN_H     = 4;                 // halo cells around the home tile inside the window
constexpr int WIN       = TILE + 2 * WIN_H;  // 12
constexpr int WIN_CELLS = WIN * WIN * WIN;   // 1728
constexpr unsigned short NOSLOT = 0xFFFFu;

struct FusedArgs {
    MeshDev m;
    PushDev P;
    const double *x, *y, *z, *px, *py, *pz;  // input order
    double *ox, *oy, *oz, *opx, *opy, *opz;  // output (cell-sorted) order
    const int* offsets_old;  // [ncells+1] offsets of the input order (nullptr: input is unsorted)
    long n_sorted;           // particles covered by offsets_old
    long n;                  // all input particles (the rest is an unsorted tail, e.g. migration arrivals)
    int ncells, ntiles;
    int tile_items;  // ntiles when the input carries cell offsets, else 0 (everything is 'tail')
    int* counts;   // pass A: histogram; pass B: running cursors (initialised with the new offsets)
    const double* ef;
    double* rho;
    double q_scalar;
    int* work;  // [0] dynamic work counter, [1] number of leavers, [2] exit overflow flag
    double* exit_buf;  // [6][exit_cap] leavers (particles whose new cell is outside the local key space)
    int exit_cap;
};

// new position/momentum + cell of one particle; returns false when the particle left the local box
struct Pushed {
    double r[3], p[3], whi[3];
    int c[3];  // new cell coords in [0, nl] if inside
    bool inside;
};

__device__ __forceinline__ void push_one(const FusedArgs& A, long i, Pushed& o) {
    o.r[0] = A.x[i]; o.r[1] = A.y[i]; o.r[2] = A.z[i];
    o.p[0] = A.px[i]; o.p[1] = A.py[i]; o.p[2] = A.pz[i];
    Cic c;
    cic_setup(A.m, o.r[0], o.r[1], o.r[2], c);
    double E[3];
    gather_point<3>(A.m, c, A.ef, E);
    push_particle(A.P, o.r, o.p, E);
    Cic cn;
    cic_setup(A.m, o.r[0], o.r[1], o.r[2], cn);
    o.inside = true;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        o.whi[d] = cn.whi[d];
        o.c[d]   = cn.a[d] - A.m.nghost;
        if (o.c[d] < 0 || o.c[d] > A.m.nl[d]) o.inside = false;
    }
}

// ---- pass A ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
push_count_kernel(FusedArgs A) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += stride) {
        Pushed o;
        push_one(A, i, o);
        if (o.inside) atomicAdd(&A.counts[cell_key_c(A.m, o.c[0], o.c[1], o.c[2])], 1);
    }
}

// ---- pass B ----------------------------------------------------------------------------------------
template <int NT, int K>
struct FusedSmem {
    static constexpr int CAP = NT * K;
    double dat[6][CAP];            // arrival-order R',P'; later reused for the cell-sorted weights
    unsigned short local[CAP];     // window cell of the particle in arrival slot s (NOSLOT: not binned)
    unsigned short rank[CAP];      // rank inside (chunk, cell)
    unsigned short perm[CAP];      // sorted position -> arrival slot
    int hist[WIN_CELLS];
    int prefix[WIN_CELLS + 1];
    int gbase[WIN_CELLS];
    unsigned short list[WIN_CELLS];  // non-empty window cells
    int warp_sums[32];
    int nne, item, total;
};

// direct path for a particle whose destination is outside the chunk's window: claim one slot from the
// cell cursor, scattered write, 8 reductions
__device__ __forceinline__ void place_direct(const FusedArgs& A, const Pushed& o) {
    const int key = cell_key_c(A.m, o.c[0], o.c[1], o.c[2]);
    const int g   = atomicAdd(&A.counts[key], 1);
    A.ox[g] = o.r[0]; A.oy[g] = o.r[1]; A.oz[g] = o.r[2];
    A.opx[g] = o.p[0]; A.opy[g] = o.p[1]; A.opz[g] = o.p[2];
    const int a[3] = {o.c[0] + A.m.nghost, o.c[1] + A.m.nghost, o.c[2] + A.m.nghost};
#pragma unroll
    for (int p = 0; p < 8; ++p)
        atomicAdd(&A.rho[cic_node(A.m, a, p)], dmul(A.q_scalar, cic_weight(o.whi, p)));
}

__device__ __forceinline__ void place_exit(const FusedArgs& A, const Pushed& o) {
    const int e = atomicAdd(&A.work[1], 1);
    if (e < A.exit_cap) {
        for (int d = 0; d < 3; ++d) {
            A.exit_buf[(long)d * A.exit_cap + e]       = o.r[d];
            A.exit_buf[(long)(3 + d) * A.exit_cap + e] = o.p[d];
        }
    } else {
        A.work[2] = 1;
    }
}

template <int NT, int K, int MINB>
__global__ void __launch_bounds__(NT, MINB)
push_move_deposit_kernel(FusedArgs A) {
    using S = FusedSmem<NT, K>;
    constexpr int CAP = S::CAP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S& s = *reinterpret_cast<S*>(smem_raw);
    const int t    = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int ntx = tiles_along(A.m.nl[0]), nty = tiles_along(A.m.nl[1]);
    const long tail       = A.n - A.n_sorted;
    const int tail_items  = (int)((tail + CAP - 1) / CAP);
    const int nitems      = A.tile_items + tail_items;

    for (int c = t; c < WIN_CELLS; c += NT) s.hist[c] = 0;
    if (t == 0) s.nne = 0;
    __syncthreads();

    for (;;) {
        __syncthreads();  // everybody is done with the previous work item (and has read s.item)
        if (t == 0) s.item = atomicAdd(&A.work[0], 1);
        __syncthreads();
        const int item = s.item;
        if (item >= nitems) break;
        long pbeg, pend;
        int hx, hy, hz;  // home tile coords
        if (item < A.tile_items) {
            pbeg = A.offsets_old[item * TILE_CELLS];
            pend = A.offsets_old[(item + 1) * TILE_CELLS];
            hx = item % ntx; hy = (item / ntx) % nty; hz = item / (ntx * nty);
        } else {
            pbeg = A.n_sorted + (long)(item - A.tile_items) * CAP;
            pend = min(pbeg + (long)CAP, A.n);
            // home = tile of the first particle's OLD cell (arrivals cluster near faces; far ones take
            // the direct path)
            Cic c0;
            cic_setup(A.m, A.x[pbeg], A.y[pbeg], A.z[pbeg], c0);
            hx = min(max(c0.a[0] - A.m.nghost, 0), A.m.nl[0]) >> 2;
            hy = min(max(c0.a[1] - A.m.nghost, 0), A.m.nl[1]) >> 2;
            hz = min(max(c0.a[2] - A.m.nghost, 0), A.m.nl[2]) >> 2;
        }
        const int wox = hx * TILE - WIN_H, woy = hy * TILE - WIN_H, woz = hz * TILE - WIN_H;

        for (long cb = pbeg; cb < pend; cb += CAP) {
            // ---- P1: push, stage in arrival order, bin by new cell ---------------------------------
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                const int slot = k * NT + t;
                const long i   = cb + slot;
                unsigned short loc = NOSLOT, rk = 0;
                if (i < pend) {
                    Pushed o;
                    push_one(A, i, o);
                    if (!o.inside) {
                        place_exit(A, o);
                    } else {
                        const int wx = o.c[0] - wox, wy = o.c[1] - woy, wz = o.c[2] - woz;
                        if ((unsigned)wx < (unsigned)WIN && (unsigned)wy < (unsigned)WIN &&
                            (unsigned)wz < (unsigned)WIN) {
                            loc = (unsigned short)((wz * WIN + wy) * WIN + wx);
                            rk  = (unsigned short)atomicAdd(&s.hist[loc], 1);
#pragma unroll
                            for (int d = 0; d < 3; ++d) {
                                s.dat[d][slot]     = o.r[d];
                                s.dat[3 + d][slot] = o.p[d];
                            }
                        } else {
                            place_direct(A, o);
                        }
                    }
                }
                s.local[slot] = loc;
                s.rank[slot]  = rk;
            }
            __syncthreads();
            // ---- P2: exclusive scan of the window histogram, reserve one run per non-empty cell ------
            {
                constexpr int IPT = (WIN_CELLS + NT - 1) / NT;
                int v[IPT], sum = 0;
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    const int c = t * IPT + j;
                    v[j] = c < WIN_CELLS ? s.hist[c] : 0;
                    sum += v[j];
                }
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += y;
                }
                if (lane == 31) s.warp_sums[warp] = inc;
                __syncthreads();
                if (warp == 0) {
                    int w = lane < NT / 32 ? s.warp_sums[lane] : 0;
                    int wi = w;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int y = __shfl_up_sync(0xffffffffu, wi, o);
                        if (lane >= o) wi += y;
                    }
                    s.warp_sums[lane] = wi - w;  // exclusive
                    if (lane == 31) s.total = wi;
                }
                __syncthreads();
                int run = s.warp_sums[warp] + inc - sum;
#pragma unroll
                for (int j = 0; j < IPT; ++j) {
                    const int c = t * IPT + j;
                    if (c < WIN_CELLS) {
                        s.prefix[c] = run;
                        if (v[j] > 0) {
                            const int cx = c % WIN + wox, cy = (c / WIN) % WIN + woy, cz = c / (WIN * WIN) + woz;
                            s.gbase[c] = atomicAdd(&A.counts[cell_key_c(A.m, cx, cy, cz)], v[j]);
                            s.list[atomicAdd(&s.nne, 1)] = (unsigned short)c;
                        }
                        run += v[j];
                    }
                }
                if (t == 0) s.prefix[WIN_CELLS] = s.total;
            }
            __syncthreads();
            // ---- P3: sorted position -> arrival slot ------------------------------------------------
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int slot = k * NT + t;
                const unsigned short loc = s.local[slot];
                if (loc != NOSLOT) s.perm[s.prefix[loc] + s.rank[slot]] = (unsigned short)slot;
            }
            __syncthreads();
            // ---- P4: coalesced write-out in cell-sorted order; weights of the new position ----------
            const int ntot = s.total;
            double w[K][3];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int p = k * NT + t;
                if (p < ntot) {
                    const int slot = s.perm[p];
                    const int c    = s.local[slot];
                    const long g   = (long)s.gbase[c] + s.rank[slot];
                    const double r0 = s.dat[0][slot], r1 = s.dat[1][slot], r2 = s.dat[2][slot];
                    A.ox[g] = r0; A.oy[g] = r1; A.oz[g] = r2;
                    A.opx[g] = s.dat[3][slot]; A.opy[g] = s.dat[4][slot]; A.opz[g] = s.dat[5][slot];
                    int idx;
                    cic_axis(r0, A.m.origin[0], A.m.invdx[0], idx, w[k][0]);
                    cic_axis(r1, A.m.origin[1], A.m.invdx[1], idx, w[k][1]);
                    cic_axis(r2, A.m.origin[2], A.m.invdx[2], idx, w[k][2]);
                }
            }
            __syncthreads();
            // ---- P5: weights into the (now free) staging buffer, in sorted order ---------------------
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int p = k * NT + t;
                if (p < ntot) {
                    s.dat[0][p] = w[k][0];
                    s.dat[1][p] = w[k][1];
                    s.dat[2][p] = w[k][2];
                }
            }
            __syncthreads();
            // ---- P6: deposit, one thread per (non-empty cell, stencil node) --------------------------
            const int nitems_dep = s.nne * 8;
            for (int j = t; j < nitems_dep; j += NT) {
                const int c    = s.list[j >> 3];
                const int node = j & 7;
                const int b = s.prefix[c], e = s.prefix[c + 1];
                double acc = 0.0;
                for (int p = b; p < e; ++p) {
                    double w0 = s.dat[0][p], w1 = s.dat[1][p], w2 = s.dat[2][p];
                    if (node & 1) w0 = 1.0 - w0;
                    if (node & 2) w1 = 1.0 - w1;
                    if (node & 4) w2 = 1.0 - w2;
                    acc += w0 * (w1 * w2);
                }
                const int a[3] = {c % WIN + wox + A.m.nghost, (c / WIN) % WIN + woy + A.m.nghost,
                                  c / (WIN * WIN) + woz + A.m.nghost};
                atomicAdd(&A.rho[cic_node(A.m, a, node)], A.q_scalar * acc);
            }
            __syncthreads();
            // reset the window tables for the next chunk
            for (int j = t; j < s.nne; j += NT) s.hist[s.list[j]] = 0;
            __syncthreads();
            if (t == 0) s.nne = 0;
            __syncthreads();
        }
    }
}

static int grid_for(const ipplb_ctx* ctx, long n, int block, int per_sm) {
    long want = (n + block - 1) / block;
    long cap  = (long)ctx->num_sms * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

constexpr int F_NT = 384, F_K = 4, F_MINB = 2;

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_step_fused(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push,
                     ipplb_particles* p, ipplb_particles* scratch, int* cell_offsets, long n_sorted,
                     const double* efield, double* rho, double* exit_buf, int exit_cap,
                     int* n_exit_host) {
    IPPLB_REQUIRE(ctx && mesh && push && p && scratch && cell_offsets && efield && rho,
                  "step_fused: bad arguments");
    IPPLB_REQUIRE(p->q == nullptr, "step_fused: per-particle charge arrays take the unfused path");
    IPPLB_REQUIRE(scratch->capacity >= p->n, "step_fused: scratch capacity too small");
    IPPLB_REQUIRE(n_sorted >= 0 && n_sorted <= p->n, "step_fused: bad n_sorted");
    const long ncells = ipplb_sort_ncells(mesh);
    int rc;
    if ((rc = ensure(ctx, ctx->counts, sizeof(int) * (size_t)(ncells + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->misc, sizeof(int) * 64))) return rc;
    if ((rc = ensure(ctx, ctx->keys, sizeof(int) * (size_t)(ncells + 1)))) return rc;
    int* counts  = (int*)ctx->counts.ptr;
    int* work    = (int*)ctx->misc.ptr;
    int* newoffs = (int*)ctx->keys.ptr;
    FusedArgs A;
    A.m = make_mesh_dev(mesh);
    A.P = make_push_dev(mesh, push);
    A.x = p->x; A.y = p->y; A.z = p->z; A.px = p->px; A.py = p->py; A.pz = p->pz;
    A.ox = scratch->x; A.oy = scratch->y; A.oz = scratch->z;
    A.opx = scratch->px; A.opy = scratch->py; A.opz = scratch->pz;
    A.offsets_old = cell_offsets;
    A.n_sorted = n_sorted;
    A.n = p->n;
    A.ncells = (int)ncells;
    A.ntiles = (int)(ncells / TILE_CELLS);
    A.tile_items = n_sorted > 0 ? A.ntiles : 0;
    A.counts = counts;
    A.ef = efield; A.rho = rho; A.q_scalar = p->q_scalar;
    A.work = work;
    A.exit_buf = exit_buf; A.exit_cap = exit_buf ? exit_cap : 0;
    IPPLB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (ncells + 1), ctx->stream));
    IPPLB_CUDA(cudaMemsetAsync(work, 0, sizeof(int) * 16, ctx->stream));
    if (p->n > 0) {
        push_count_kernel<<<grid_for(ctx, p->n, 256, 16), 256, 0, ctx->stream>>>(A);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, newoffs, (int)(ncells + 1), ctx->stream);
    if ((rc = ensure(ctx, ctx->cub_tmp, tmp_bytes))) return rc;
    IPPLB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.ptr, tmp_bytes, counts, newoffs,
                                             (int)(ncells + 1), ctx->stream));
    ctx->launches += 2;
    IPPLB_CUDA(cudaMemcpyAsync(counts, newoffs, sizeof(int) * (ncells + 1), cudaMemcpyDeviceToDevice,
                               ctx->stream));
    if (p->n > 0) {
        using S = FusedSmem<F_NT, F_K>;
        auto kern = push_move_deposit_kernel<F_NT, F_K, F_MINB>;
        static bool attr_set = false;
        if (!attr_set) {
            IPPLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(S)));
            attr_set = true;
        }
        kern<<<ctx->num_sms * F_MINB, F_NT, sizeof(S), ctx->stream>>>(A);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    // new offsets become the caller's cell_offsets; the number of stayers is offsets[ncells]
    IPPLB_CUDA(cudaMemcpyAsync(cell_offsets, newoffs, sizeof(int) * (ncells + 1),
                               cudaMemcpyDeviceToDevice, ctx->stream));
    int* h = (int*)ctx->reduce_host;
    IPPLB_CUDA(cudaMemcpyAsync(h, work, sizeof(int) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaMemcpyAsync(h + 4, newoffs + ncells, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int n_exit = h[1], n_stay = h[4];
    if (h[2]) {
        set_error("step_fused: %d particles left the local box but the exit buffer holds %d", n_exit, A.exit_cap);
        return IPPLB_ERR_CAPACITY;
    }
    if ((long)n_stay + n_exit != p->n) {
        set_error("step_fused: particle count mismatch (%d stay + %d exit != %ld)", n_stay, n_exit, p->n);
        return IPPLB_ERR_CUDA;
    }
    if (n_exit_host) *n_exit_host = n_exit;
    ipplb_particles tswap = *p;
    *p             = *scratch;
    *scratch       = tswap;
    p->n           = n_stay;
    p->q           = nullptr;
    p->q_scalar    = scratch->q_scalar;
    scratch->n     = 0;
    return IPPLB_OK;
}

}  // extern "C"
