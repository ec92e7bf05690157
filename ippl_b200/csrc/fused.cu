// fused.cu -- the B200-first PIC step: ONE pass over bucketed particles (see include/ippl_b200.h, ipplb_bins).
//
// Storage: per-tile buckets (tile = 4x4x4 key cells) with slack, double-buffered; device-resident tables
// start/cap/count per tile, re-planned on the device after every step (bins.cu).
//
// Kernel (persistent, 2 CTAs per SM, dynamic tile scheduler):
//   producer warp : walks (tile, chunk) work items, streams the chunk's six SoA runs into a 2-stage shared
//                   memory ring with bulk async copies (cp.async.bulk + mbarrier complete_tx, i.e. TMA 1-D)
//                   and stages the tile's 5x5x5-node E window as x-pairs so the gather is 12 LDS.128.
//   consumer warps: P1 gather E from shared memory, kick/kick/drift/BC in registers, bin by NEW cell with
//                      integer shared-memory atomics on an 8x8x8-cell window (tile-major ids);
//                   P2 scan of the window histogram, one global reservation per destination TILE (<= 27);
//                   P3 sorted position -> arrival slot permutation;
//                   P4 coalesced store of R,P into next step's buckets (overflow -> tail), CIC weights of
//                      the new positions into the freed staging columns, in sorted order;
//                   P5 charge deposit: 2 lanes per non-empty cell walk the cell's weights, register
//                      accumulation of the 8 nodes, 4 RED.F64 per lane.
// HBM traffic: 48 B read + 48 B written per particle.  E never round-trips through HBM, rho is produced
// without re-reading the particles, and there is no sort pass.  Positions and momenta are bit-identical to
// the unfused path (same device functions, same operation order).
#include <climits>
#include <cstdint>

#include "bins.h"
#include "push.cuh"

namespace ipplb {

constexpr int WH        = 2;              // window halo (cells) around the home tile
constexpr int WIN       = TILE + 2 * WH;  // 8
constexpr int WIN_CELLS = WIN * WIN * WIN;
constexpr int NSLOT     = 27;             // destination tiles touched by the window
constexpr unsigned short NOSLOT = 0xFFFFu;
constexpr int EP_N = 3 * 5 * 5 * 4;       // E x-pairs per tile window

enum { CH_TILE = 0, CH_TAIL = 1, CH_STOP = 2 };

struct ChunkDesc {
    int kind, cnt;
    int hx, hy, hz;  // home tile coords
    int pad[3];
};

struct StepArgs {
    MeshDev m;
    PushDev P;
    const double* in[6];
    double* out[6];
    const int *start_in, *count_in, *state_in;
    const int *start_out, *cap_out;
    int *cursor_out, *state_out;
    int* misc;
    const double* ef;
    double* rho;
    double q;
    double* exit_buf;
    int exit_cap;
    int capacity;
    int ntx, nty, ntz, ntiles;
    int check_owner;
    double rmin[3], rmax[3];
};

template <int NT, int K>
struct StepSmem {
    static constexpr int CAP = NT * K;
    struct Stage {
        double dat[6][CAP];
        double2 ep[EP_N];
    } st[2];
    ChunkDesc desc[2];
    unsigned long long full[2], empty[2];
    int hist[WIN_CELLS];
    int prefix[WIN_CELLS + 1];
    unsigned short local[CAP], rank[CAP], perm[CAP];
    unsigned short list[WIN_CELLS];
    unsigned short cellxyz[WIN_CELLS];
    unsigned char tsof[WIN_CELLS];
    int tsbase[NSLOT + 1];
    int adj[NSLOT], lim[NSLOT], tadj[NSLOT];
    int warp_sums[32];
    int nne, total;
};

// ---- PTX helpers: mbarrier + bulk async copy (TMA 1-D) -----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
template <int NT>
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// window coordinate -> (segment, start, width): cells [0,2) belong to the lower neighbour tile, [2,6) to the
// home tile, [6,8) to the upper neighbour
__device__ __forceinline__ void win_seg(int w, int& s, int& l, int& wd) {
    s  = (w + 2) >> 2;
    l  = w - (s == 0 ? 0 : 4 * s - 2);
    wd = (s == 1) ? 4 : 2;
}

// ---- slow paths ------------------------------------------------------------------------------------------
__device__ __forceinline__ void place_exit(const StepArgs& A, const double r[3], const double p[3]) {
    const int e = atomicAdd(&A.misc[BM_EXIT], 1);
    if (e < A.exit_cap) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            A.exit_buf[(long)d * A.exit_cap + e]       = r[d];
            A.exit_buf[(long)(3 + d) * A.exit_cap + e] = p[d];
        }
    } else {
        atomicOr(&A.misc[BM_FLAGS], IPPLB_FLAG_EXIT_OVERFLOW);
    }
}

// a particle whose destination is outside the chunk's window (or that sits in the unsorted tail): claim one
// slot of the destination bucket, scattered write, 8 reductions
__device__ __forceinline__ void place_direct(const StepArgs& A, const double r[3], const double p[3],
                                             const int c[3], const double whi[3]) {
    const int tile = (c[0] >> 2) + A.ntx * ((c[1] >> 2) + A.nty * (c[2] >> 2));
    int slot       = atomicAdd(&A.cursor_out[tile], 1);
    long g;
    if (slot < A.cap_out[tile]) {
        g = (long)A.start_out[tile] + slot;
    } else {
        g = (long)A.state_out[BS_TAIL_START] + atomicAdd(&A.state_out[BS_TAIL_COUNT], 1);
        if (g >= A.capacity) {
            atomicOr(&A.misc[BM_FLAGS], IPPLB_FLAG_CAPACITY);
            g = -1;
        }
    }
    if (g >= 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            A.out[d][g]     = r[d];
            A.out[3 + d][g] = p[d];
        }
    }
    const int a[3] = {c[0] + A.m.nghost, c[1] + A.m.nghost, c[2] + A.m.nghost};
#pragma unroll
    for (int n = 0; n < 8; ++n) atomicAdd(&A.rho[cic_node(A.m, a, n)], dmul(A.q, cic_weight(whi, n)));
}

__device__ __forceinline__ bool owned_by_me(const StepArgs& A, const double r[3], const int c[3]) {
    if (A.check_owner) {
        // ParticleSpatialLayout::positionInRegion (ParticleSpatialLayout.hpp:316-330): pos > min && pos <= max
        return r[0] > A.rmin[0] && r[0] <= A.rmax[0] && r[1] > A.rmin[1] && r[1] <= A.rmax[1] &&
               r[2] > A.rmin[2] && r[2] <= A.rmax[2];
    }
    return c[0] >= 0 && c[0] <= A.m.nl[0] && c[1] >= 0 && c[1] <= A.m.nl[1] && c[2] >= 0 && c[2] <= A.m.nl[2];
}

// gather from the staged x-pair window: identical operation order to gather_point<3> (push.cuh)
__device__ __forceinline__ void gather_pairs(const double2* __restrict__ ep, int lx, int ly, int lz,
                                             const double whi[3], double E[3]) {
    double w[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) w[n] = cic_weight(whi, n);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // node n: bit0 -> lower x (pair.x) else upper x (pair.y); bit1 -> row ly else ly+1; bit2 -> lz else lz+1
        const double2 p11 = ep[((c * 5 + lz) * 5 + ly) * 4 + lx];          // bits 1,2 set   (n = 6,7)
        const double2 p01 = ep[((c * 5 + lz) * 5 + ly + 1) * 4 + lx];      // bit 2 set      (n = 4,5)
        const double2 p10 = ep[((c * 5 + lz + 1) * 5 + ly) * 4 + lx];      // bit 1 set      (n = 2,3)
        const double2 p00 = ep[((c * 5 + lz + 1) * 5 + ly + 1) * 4 + lx];  // none           (n = 0,1)
        double acc = dmul(w[7], p11.x);
        acc        = dadd(dmul(w[6], p11.y), acc);
        acc        = dadd(dmul(w[5], p01.x), acc);
        acc        = dadd(dmul(w[4], p01.y), acc);
        acc        = dadd(dmul(w[3], p10.x), acc);
        acc        = dadd(dmul(w[2], p10.y), acc);
        acc        = dadd(dmul(w[1], p00.x), acc);
        acc        = dadd(dmul(w[0], p00.y), acc);
        E[c]       = acc;
    }
}

// ---- the kernel --------------------------------------------------------------------------------------------
template <int NT, int K, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) fused_step_kernel(const StepArgs A) {
    using S           = StepSmem<NT, K>;
    constexpr int CAP = S::CAP;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    S& s           = *reinterpret_cast<S*>(smem_raw);
    const int t    = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;

    // ---- one-time tables ---------------------------------------------------------------------------------
    if (t == 0) {
        int run = 0;
        for (int ts = 0; ts < NSLOT; ++ts) {
            s.tsbase[ts] = run;
            const int wx = (ts % 3 == 1) ? 4 : 2, wy = ((ts / 3) % 3 == 1) ? 4 : 2, wz = (ts / 9 == 1) ? 4 : 2;
            run += wx * wy * wz;
        }
        s.tsbase[NSLOT] = run;
        s.nne           = 0;
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    for (int c = t; c < WIN_CELLS; c += NT + 32) {
        const int wx = c & 7, wy = (c >> 3) & 7, wz = c >> 6;
        int sx, lx, dx, sy, ly, dy, sz, lz, dz;
        win_seg(wx, sx, lx, dx);
        win_seg(wy, sy, ly, dy);
        win_seg(wz, sz, lz, dz);
        const int ts = sx + 3 * (sy + 3 * sz);
        const int id = s.tsbase[ts] + (lz * dy + ly) * dx + lx;
        s.cellxyz[id] = (unsigned short)(wx | (wy << 4) | (wz << 8));
        s.tsof[id]    = (unsigned char)ts;
        s.hist[c]     = 0;
    }
    __syncthreads();

    // ======================================= producer warp ==================================================
    if (warp == NT / 32) {
        const int tail_start = A.state_in[BS_TAIL_START];
        const int tail_count = A.state_in[BS_TAIL_COUNT];
        int rem = 0, kind = CH_STOP, hx = 0, hy = 0, hz = 0;
        long pbeg = 0;
        for (unsigned seq = 0;; ++seq) {
            const int st = seq & 1;
            mbar_wait(&s.empty[st], ((seq >> 1) & 1) ^ 1);
            while (rem == 0) {
                int it = 0;
                if (lane == 0) it = atomicAdd(&A.misc[BM_WORK], 1);
                it = __shfl_sync(0xffffffffu, it, 0);
                if (it < A.ntiles) {
                    rem  = A.count_in[it];
                    pbeg = A.start_in[it];
                    kind = CH_TILE;
                    hx   = it % A.ntx;
                    hy   = (it / A.ntx) % A.nty;
                    hz   = it / (A.ntx * A.nty);
                } else {
                    const long off = (long)(it - A.ntiles) * CAP;
                    if (off >= tail_count) {
                        kind = CH_STOP;
                        break;
                    }
                    rem  = (int)min((long)CAP, (long)tail_count - off);
                    pbeg = (long)tail_start + off;
                    kind = CH_TAIL;
                }
            }
            if (kind == CH_STOP) {
                if (lane == 0) {
                    s.desc[st].kind = CH_STOP;
                    mbar_arrive(&s.full[st]);
                }
                break;
            }
            const int cnt = min(rem, CAP);
            if (lane == 0) {
                ChunkDesc d;
                d.kind = kind; d.cnt = cnt; d.hx = hx; d.hy = hy; d.hz = hz;
                s.desc[st] = d;
                const uint32_t bytes = (uint32_t)(((cnt + 1) & ~1) * 8);
                mbar_expect_tx(&s.full[st], 6 * bytes);
#pragma unroll
                for (int a = 0; a < 6; ++a) bulk_g2s(&s.st[st].dat[a][0], A.in[a] + pbeg, bytes, &s.full[st]);
            }
            if (kind == CH_TILE) {
                // E window of the tile as x-pairs: ep[c][kz][jy][ix] = (E_c(node ix), E_c(node ix+1)); node (0,0,0)
                // is the lower node of the tile's first cell, ghosted index = 4*h + nghost - 1
                const int gx0 = 4 * hx + A.m.nghost - 1, gy0 = 4 * hy + A.m.nghost - 1,
                          gz0 = 4 * hz + A.m.nghost - 1;
                for (int e = lane; e < EP_N; e += 32) {
                    const int ix = e & 3, jy = (e >> 2) % 5, kz = ((e >> 2) / 5) % 5, c = (e >> 2) / 25;
                    const int gx = gx0 + ix, gy = gy0 + jy, gz = gz0 + kz;
                    double2 v = make_double2(0.0, 0.0);
                    if (gy < A.m.ey && gz < A.m.ez) {
                        const long base = ((long)gx + (long)A.m.ex * (gy + (long)A.m.ey * gz)) * 3 + c;
                        if (gx < A.m.ex) v.x = __ldg(&A.ef[base]);
                        if (gx + 1 < A.m.ex) v.y = __ldg(&A.ef[base + 3]);
                    }
                    s.st[st].ep[e] = v;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.full[st]);
            rem -= cnt;
            pbeg += cnt;
        }
        return;
    }

    // ======================================= consumer warps =================================================
    for (unsigned seq = 0;; ++seq) {
        const int st = seq & 1;
        mbar_wait(&s.full[st], (seq >> 1) & 1);
        const ChunkDesc D = s.desc[st];
        if (D.kind == CH_STOP) break;
        typename S::Stage& G = s.st[st];
        const int cnt        = D.cnt;

        if (D.kind == CH_TAIL) {
            // unsorted particles (overflow of the previous step, migration arrivals): global gather, direct placement
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                const int slot = k * NT + t;
                if (slot < cnt) {
                    double r[3] = {G.dat[0][slot], G.dat[1][slot], G.dat[2][slot]};
                    double p[3] = {G.dat[3][slot], G.dat[4][slot], G.dat[5][slot]};
                    Cic c;
                    cic_setup(A.m, r[0], r[1], r[2], c);
                    double E[3];
                    gather_point<3>(A.m, c, A.ef, E);
                    push_particle(A.P, r, p, E);
                    Cic cn;
                    cic_setup(A.m, r[0], r[1], r[2], cn);
                    const int cc[3] = {cn.a[0] - A.m.nghost, cn.a[1] - A.m.nghost, cn.a[2] - A.m.nghost};
                    if (owned_by_me(A, r, cc)) place_direct(A, r, p, cc, cn.whi);
                    else place_exit(A, r, p);
                }
            }
            fence_proxy_async();
            consumer_sync<NT>();
            if (t == 0) mbar_arrive(&s.empty[st]);
            continue;
        }

        const int wox = D.hx * TILE - WH, woy = D.hy * TILE - WH, woz = D.hz * TILE - WH;
        // ---- P1: gather, push, bin by new cell ------------------------------------------------------------
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const int slot = k * NT + t;
            unsigned short loc = NOSLOT, rk = 0;
            if (slot < cnt) {
                double r[3] = {G.dat[0][slot], G.dat[1][slot], G.dat[2][slot]};
                double p[3] = {G.dat[3][slot], G.dat[4][slot], G.dat[5][slot]};
                Cic c;
                cic_setup(A.m, r[0], r[1], r[2], c);
                double E[3];
                gather_pairs(G.ep, c.a[0] - A.m.nghost - D.hx * TILE, c.a[1] - A.m.nghost - D.hy * TILE,
                             c.a[2] - A.m.nghost - D.hz * TILE, c.whi, E);
                push_particle(A.P, r, p, E);
                Cic cn;
                cic_setup(A.m, r[0], r[1], r[2], cn);
                const int cc[3] = {cn.a[0] - A.m.nghost, cn.a[1] - A.m.nghost, cn.a[2] - A.m.nghost};
                if (!owned_by_me(A, r, cc)) {
                    place_exit(A, r, p);
                } else {
                    const int wx = cc[0] - wox, wy = cc[1] - woy, wz = cc[2] - woz;
                    if ((unsigned)wx < (unsigned)WIN && (unsigned)wy < (unsigned)WIN && (unsigned)wz < (unsigned)WIN) {
                        int sx, lx, dx, sy, ly, dy, sz, lz, dz;
                        win_seg(wx, sx, lx, dx);
                        win_seg(wy, sy, ly, dy);
                        win_seg(wz, sz, lz, dz);
                        const int id = s.tsbase[sx + 3 * (sy + 3 * sz)] + (lz * dy + ly) * dx + lx;
                        loc          = (unsigned short)id;
                        rk           = (unsigned short)atomicAdd(&s.hist[id], 1);
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            G.dat[d][slot]     = r[d];
                            G.dat[3 + d][slot] = p[d];
                        }
                    } else {
                        place_direct(A, r, p, cc, cn.whi);
                    }
                }
            }
            s.local[slot] = loc;
            s.rank[slot]  = rk;
        }
        consumer_sync<NT>();
        // ---- P2: exclusive scan of the window histogram (tile-major ids) -----------------------------------
        {
            constexpr int IPT = (WIN_CELLS + NT - 1) / NT;
            int v[IPT], sum = 0;
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const int c = t * IPT + j;
                v[j]        = c < WIN_CELLS ? s.hist[c] : 0;
                sum += v[j];
            }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            if (lane == 31) s.warp_sums[warp] = inc;
            consumer_sync<NT>();
            if (warp == 0) {
                int w  = lane < NT / 32 ? s.warp_sums[lane] : 0;
                int wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, wi, o);
                    if (lane >= o) wi += y;
                }
                s.warp_sums[lane] = wi - w;  // exclusive
                if (lane == 31) s.total = wi;
            }
            consumer_sync<NT>();
            int run = s.warp_sums[warp] + inc - sum;
#pragma unroll
            for (int j = 0; j < IPT; ++j) {
                const int c = t * IPT + j;
                if (c < WIN_CELLS) {
                    s.prefix[c] = run;
                    if (v[j] > 0) {
                        s.list[atomicAdd(&s.nne, 1)] = (unsigned short)c;
                        s.hist[c]                    = 0;  // ready for the next chunk
                    }
                    run += v[j];
                }
            }
            if (t == 0) s.prefix[WIN_CELLS] = s.total;
        }
        consumer_sync<NT>();
        // ---- reserve one block per destination tile; P3: sorted position -> arrival slot ----------------
        if (t < NSLOT) {
            const int b0 = s.prefix[s.tsbase[t]], b1 = s.prefix[s.tsbase[t + 1]];
            const int n  = b1 - b0;
            int adj = 0, lim = 0, tadj = 0;
            if (n > 0) {
                const int tx = D.hx + (t % 3) - 1, ty = D.hy + ((t / 3) % 3) - 1, tz = D.hz + (t / 9) - 1;
                if (tx < 0 || tx >= A.ntx || ty < 0 || ty >= A.nty || tz < 0 || tz >= A.ntz) {
                    atomicOr(&A.misc[BM_FLAGS], IPPLB_FLAG_INTERNAL);
                    lim  = 0;  // everything of this block is dropped
                    tadj = INT_MIN;
                } else {
                    const int tile = tx + A.ntx * (ty + A.nty * tz);
                    const int base = atomicAdd(&A.cursor_out[tile], n);
                    const int cap  = A.cap_out[tile];
                    const int abs0 = A.start_out[tile] + base;
                    adj            = abs0 - b0;
                    lim            = A.start_out[tile] + cap;
                    const int g0   = max(lim, abs0);  // first absolute slot that does not fit
                    const int over = abs0 + n - g0;
                    if (over > 0) {
                        const int tb = A.state_out[BS_TAIL_START] + atomicAdd(&A.state_out[BS_TAIL_COUNT], over);
                        tadj         = tb - g0;
                        if ((long)tb + over > A.capacity) {
                            atomicOr(&A.misc[BM_FLAGS], IPPLB_FLAG_CAPACITY);
                            tadj = INT_MIN;
                        }
                    }
                }
            }
            s.adj[t]  = adj;
            s.lim[t]  = lim;
            s.tadj[t] = tadj;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int slot           = k * NT + t;
            const unsigned short loc = s.local[slot];
            if (loc != NOSLOT) s.perm[s.prefix[loc] + s.rank[slot]] = (unsigned short)slot;
        }
        consumer_sync<NT>();
        // ---- P4: coalesced write-out in sorted order; weights of the new position ------------------------
        const int ntot = s.total;
        int gdst[K], src[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int p = k * NT + t;
            gdst[k]     = -1;
            src[k]      = 0;
            if (p < ntot) {
                const int slot = s.perm[p];
                const int ts   = s.tsof[s.local[slot]];
                int g          = s.adj[ts] + p;
                if (g >= s.lim[ts]) {
                    const int ta = s.tadj[ts];
                    g            = ta == INT_MIN ? -1 : g + ta;
                }
                src[k]  = slot;
                gdst[k] = g;  // negative: dropped (flag already raised)
                if (g >= 0) {
                    A.out[3][g] = G.dat[3][slot];
                    A.out[4][g] = G.dat[4][slot];
                    A.out[5][g] = G.dat[5][slot];
                }
            }
        }
        consumer_sync<NT>();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int p = k * NT + t;
            if (p < ntot) {
                const int slot = src[k];
                const double r0 = G.dat[0][slot], r1 = G.dat[1][slot], r2 = G.dat[2][slot];
                const int g = gdst[k];
                if (g >= 0) {
                    A.out[0][g] = r0;
                    A.out[1][g] = r1;
                    A.out[2][g] = r2;
                }
                int idx;
                double w0, w1, w2;
                cic_axis(r0, A.m.origin[0], A.m.invdx[0], idx, w0);
                cic_axis(r1, A.m.origin[1], A.m.invdx[1], idx, w1);
                cic_axis(r2, A.m.origin[2], A.m.invdx[2], idx, w2);
                G.dat[3][p] = w0;
                G.dat[4][p] = w1;
                G.dat[5][p] = w2;
            }
        }
        consumer_sync<NT>();
        // ---- P5: deposit, two lanes per non-empty cell -----------------------------------------------------
        {
            const int nitems = s.nne * 2;
            const int nround = (nitems + 31) & ~31;
            for (int j = t; j < nround; j += NT) {
                const bool valid = j < nitems;
                const int id     = valid ? s.list[j >> 1] : 0;
                const int g      = j & 1;
                const int b = valid ? s.prefix[id] + g : 0, e = valid ? s.prefix[id + 1] : 0;
                double acc[8];
#pragma unroll
                for (int n = 0; n < 8; ++n) acc[n] = 0.0;
                for (int p = b; p < e; p += 2) {
                    const double w0 = G.dat[3][p], w1 = G.dat[4][p], w2 = G.dat[5][p];
                    const double u0 = 1.0 - w0, u1 = 1.0 - w1, u2 = 1.0 - w2;
                    const double t00 = w1 * w2, t10 = u1 * w2, t01 = w1 * u2, t11 = u1 * u2;
                    acc[0] += w0 * t00;
                    acc[1] += u0 * t00;
                    acc[2] += w0 * t10;
                    acc[3] += u0 * t10;
                    acc[4] += w0 * t01;
                    acc[5] += u0 * t01;
                    acc[6] += w0 * t11;
                    acc[7] += u0 * t11;
                }
                // pair exchange: lane g = 0 finishes nodes 0..3, lane g = 1 nodes 4..7
#pragma unroll
                for (int n = 0; n < 4; ++n) {
                    const double give = g ? acc[n] : acc[4 + n];
                    const double got  = __shfl_xor_sync(0xffffffffu, give, 1);
                    const double mine = (g ? acc[4 + n] : acc[n]) + got;
                    if (g) acc[4 + n] = mine; else acc[n] = mine;
                }
                if (valid) {
                    const unsigned xyz = s.cellxyz[id];
                    const int a[3]     = {(int)(xyz & 15) + wox + A.m.nghost, (int)((xyz >> 4) & 15) + woy + A.m.nghost,
                                          (int)(xyz >> 8) + woz + A.m.nghost};
#pragma unroll
                    for (int n = 0; n < 4; ++n) {
                        const int node = 4 * g + n;
                        atomicAdd(&A.rho[cic_node(A.m, a, node)], A.q * (g ? acc[4 + n] : acc[n]));
                    }
                }
            }
        }
        fence_proxy_async();
        consumer_sync<NT>();
        if (t == 0) {
            s.nne = 0;
            mbar_arrive(&s.empty[st]);
        }
    }
}

constexpr int F_NT = 384, F_K = 2, F_MINB = 2;

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_bins_step(ipplb_ctx* ctx, ipplb_bins* b, const ipplb_push* push, const ipplb_particles* cur,
                    ipplb_particles* nxt, const double* efield, double* rho, double* exit_buf,
                    int exit_cap, const double region_min[3], const double region_max[3]) {
    IPPLB_REQUIRE(ctx && b && push && cur && nxt && efield && rho, "bins_step: bad arguments");
    IPPLB_REQUIRE(cur->q == nullptr, "bins_step: per-particle charge arrays take the unfused path");
    IPPLB_REQUIRE(cur->capacity >= b->capacity && nxt->capacity >= b->capacity,
                  "bins_step: particle bundles smaller than the bins capacity");
    IPPLB_REQUIRE(b->built, "bins_step: call ipplb_bins_build first");
    const int i = b->cur, o = 1 - b->cur;
    StepArgs A;
    A.m = make_mesh_dev(&b->mesh);
    A.P = make_push_dev(&b->mesh, push);
    const double* in[6] = {cur->x, cur->y, cur->z, cur->px, cur->py, cur->pz};
    double* out[6]      = {nxt->x, nxt->y, nxt->z, nxt->px, nxt->py, nxt->pz};
    for (int a = 0; a < 6; ++a) {
        IPPLB_REQUIRE(in[a] && out[a] && in[a] != out[a], "bins_step: null or aliased particle arrays");
        IPPLB_REQUIRE(((uintptr_t)in[a] & 15) == 0, "bins_step: particle arrays must be 16-byte aligned");
        A.in[a]  = in[a];
        A.out[a] = out[a];
    }
    A.start_in   = b->start(i);
    A.count_in   = b->count(i);
    A.state_in   = b->state(i);
    A.start_out  = b->start(o);
    A.cap_out    = b->cap(o);
    A.cursor_out = b->count(o);
    A.state_out  = b->state(o);
    A.misc       = b->misc();
    A.ef         = efield;
    A.rho        = rho;
    A.q          = cur->q_scalar;
    A.exit_buf   = exit_buf;
    A.exit_cap   = exit_buf ? exit_cap : 0;
    A.capacity   = (int)b->capacity;
    A.ntx = b->ntx; A.nty = b->nty; A.ntz = b->ntz; A.ntiles = b->ntiles;
    A.check_owner = (region_min && region_max) ? 1 : 0;
    for (int d = 0; d < 3; ++d) {
        A.rmin[d] = region_min ? region_min[d] : 0.0;
        A.rmax[d] = region_max ? region_max[d] : 0.0;
    }
    IPPLB_CUDA(cudaMemsetAsync(b->misc(), 0, sizeof(int) * 4, ctx->stream));
    using S   = StepSmem<F_NT, F_K>;
    auto kern = fused_step_kernel<F_NT, F_K, F_MINB>;
    static bool attr_set = false;
    if (!attr_set) {
        IPPLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S)));
        attr_set = true;
    }
    kern<<<ctx->num_sms * F_MINB, F_NT + 32, sizeof(S), ctx->stream>>>(A);
    IPPLB_CHECK_LAUNCH(ctx);
    int rc = bins_plan(ctx, b, o);
    if (rc) return rc;
    b->cur        = o;
    nxt->q        = nullptr;
    nxt->q_scalar = cur->q_scalar;
    return IPPLB_OK;
}

}  // extern "C"
