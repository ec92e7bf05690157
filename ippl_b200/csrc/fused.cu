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
#include <cstdlib>

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
    int epi;         // which E window (tile chunks)
    int pad[2];      // pad[0]: last chunk of its tile
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
    double* exit_buf;  // [nranks][seg_cap][6]: leavers grouped by destination rank, one (x y z px py pz) record each
    int* exit_cnt;     // [nranks]
    int seg_cap;
    const double* regions;  // [nranks][6] (device) or nullptr on a single rank
    int nranks, me;
    int capacity;
    int ntx, nty, ntz, ntiles;
    int check_owner;
    int balance_chunks;  // equal chunks per tile instead of full chunks + a remainder (IPPLB_FUSED_CFG = 2xx: off)
    double rmin[3], rmax[3];
};

template <int NT, int K>
struct StepSmem {
    static constexpr int CAP = NT * K;
    struct Stage {
        double dat[6][CAP];
    } st[2];
    double2 ep[2][EP_N];  // E window of the current / next tile (x-pairs), indexed by ChunkDesc::epi
    ChunkDesc desc[2];
    unsigned long long full[2], empty[2];
    int hist[WIN_CELLS];
    int prefix[WIN_CELLS + 1];
    unsigned short local[CAP], rank[CAP], perm[CAP];
    unsigned short list[WIN_CELLS];
    unsigned short cellxyz[WIN_CELLS];  // window id -> packed window coords
    unsigned short winid[WIN_CELLS];    // window coords (wz*64 + wy*8 + wx) -> tile-major window id
    unsigned char tsof[WIN_CELLS];      // window id -> destination tile slot
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
// A particle that left the rank's region: destination rank by the reference's search (every other rank's strict
// region, then the inclusive fallback, else stay -- ParticleSpatialLayout.hpp:372-395; the own strict test has
// already failed), appended to that rank's segment of the exit buffer.  Lanes of a warp that leave for the same
// rank share one atomic.
__device__ __forceinline__ void place_exit(const StepArgs& A, const double r[3], const double p[3]) {
    int d = 0;
    if (A.regions) {
        d = -1;
        for (int k = 0; d < 0 && k < A.nranks; ++k) {
            const double* R = A.regions + 6 * k;
            if (r[0] > R[0] && r[1] > R[1] && r[2] > R[2] && r[0] <= R[3] && r[1] <= R[4] && r[2] <= R[5]) d = k;
        }
        for (int k = 0; d < 0 && k < A.nranks; ++k) {
            const double* R = A.regions + 6 * k;
            if (r[0] >= R[0] && r[1] >= R[1] && r[2] >= R[2] && r[0] <= R[3] && r[1] <= R[4] && r[2] <= R[5]) d = k;
        }
        if (d < 0) d = A.me;
    }
    const unsigned peers = __match_any_sync(__activemask(), d);
    const int leader     = __ffs(peers) - 1;
    const int lane       = threadIdx.x & 31;
    int base             = 0;
    if (lane == leader) base = atomicAdd(&A.exit_cnt[d], __popc(peers));
    base        = __shfl_sync(peers, base, leader);
    const int e = base + __popc(peers & ((1u << lane) - 1u));
    if (e < A.seg_cap) {
        // one 48-byte record per leaver: a destination's segment is one contiguous message
        double2* rec = reinterpret_cast<double2*>(A.exit_buf + ((size_t)d * A.seg_cap + e) * 6);
        rec[0]       = make_double2(r[0], r[1]);
        rec[1]       = make_double2(r[2], p[0]);
        rec[2]       = make_double2(p[1], p[2]);
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

// ---- producer warp: shared by both kernel generations ---------------------------------------------------
template <typename S>
__device__ __forceinline__ void producer_loop(const StepArgs& A, S& s, const int lane) {
    constexpr int CAP = S::CAP;
    const int tail_start = A.state_in[BS_TAIL_START];
    const int tail_count = A.state_in[BS_TAIL_COUNT];
    int rem = 0, kind = CH_STOP, hx = 0, hy = 0, hz = 0, epi = 1, chunk = CAP;
    bool fresh = false;  // first chunk of a tile: its E window has to be staged
    long pbeg = 0;
    // work items are fetched two deep so that neither the scheduler atomic nor the table loads sit on the
    // critical path: C = atomic issued (result pending in lane 0), B = item known, count/start loads in flight
    int c_it = 0, b_it = 0, b_rem = 0, b_start = 0;
    auto issue_c = [&]() {
        if (lane == 0) c_it = atomicAdd(&A.misc[BM_WORK], 1);
    };
    auto load_b = [&]() {
        b_it = __shfl_sync(0xffffffffu, c_it, 0);
        if (b_it < A.ntiles) {
            b_rem   = A.count_in[b_it];
            b_start = A.start_in[b_it];
        }
    };
    issue_c();
    load_b();
    issue_c();
    for (unsigned seq = 0;; ++seq) {
        const int st = seq & 1;
        mbar_wait(&s.empty[st], ((seq >> 1) & 1) ^ 1);
        while (rem == 0) {
            const int it = b_it;
            if (it < A.ntiles) {
                rem   = b_rem;
                pbeg  = b_start;
                kind  = CH_TILE;
                chunk = CAP;
                if (A.balance_chunks && rem > CAP) {  // same number of chunks, equal (even) sizes
                    const int nch = (rem + CAP - 1) / CAP;
                    chunk         = min(CAP, (((rem + nch - 1) / nch) + 1) & ~1);
                }
                fresh = rem > 0;
                hx   = it % A.ntx;
                hy   = (it / A.ntx) % A.nty;
                hz   = it / (A.ntx * A.nty);
            } else {
                const long off = (long)(it - A.ntiles) * CAP;
                if (off >= tail_count) {
                    kind = CH_STOP;
                    break;
                }
                rem   = (int)min((long)CAP, (long)tail_count - off);
                pbeg  = (long)tail_start + off;
                kind  = CH_TAIL;
                chunk = CAP;
            }
            load_b();
            issue_c();
        }
        if (kind == CH_STOP) {
            if (lane == 0) {
                s.desc[st].kind = CH_STOP;
                mbar_arrive(&s.full[st]);
            }
            break;
        }
        const int cnt = min(rem, chunk);
        if (lane == 0) {
            ChunkDesc d;
            d.kind = kind; d.cnt = cnt; d.hx = hx; d.hy = hy; d.hz = hz;
            d.epi  = epi ^ (fresh ? 1 : 0);
            d.pad[0] = (kind == CH_TILE && rem == cnt) ? 1 : 0;
            d.pad[1] = 0;
            s.desc[st] = d;
            const uint32_t bytes = (uint32_t)(((cnt + 1) & ~1) * 8);
            mbar_expect_tx(&s.full[st], 6 * bytes);
#pragma unroll
            for (int a = 0; a < 6; ++a) bulk_g2s(&s.st[st].dat[a][0], A.in[a] + pbeg, bytes, &s.full[st]);
        }
        if (kind == CH_TILE && fresh) {
            // (safe to overwrite: the window of two tiles ago is dead once the stage of chunk seq - 2 was released)
            epi ^= 1;
            fresh = false;
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
                s.ep[epi][e] = v;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.full[st]);
        rem -= cnt;
        pbeg += cnt;
    }
}

// unsorted particles (overflow of the previous step, migration arrivals): global gather, direct placement
template <int NT, int K>
__device__ __forceinline__ void tail_chunk(const StepArgs& A, const double (*dat)[NT * K], const int cnt, const int t) {
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const int slot = k * NT + t;
        if (slot < cnt) {
            double r[3] = {dat[0][slot], dat[1][slot], dat[2][slot]};
            double p[3] = {dat[3][slot], dat[4][slot], dat[5][slot]};
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
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], NT / 32);
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
        s.winid[c]    = (unsigned short)id;
        s.tsof[id]    = (unsigned char)ts;
        s.hist[c]     = 0;
    }
    __syncthreads();

    if (warp == NT / 32) {
        producer_loop<S>(A, s, lane);
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
            tail_chunk<NT, K>(A, G.dat, cnt, t);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.empty[st]);
            continue;
        }

        const int wox = D.hx * TILE - WH, woy = D.hy * TILE - WH, woz = D.hz * TILE - WH;
        // ---- P1: gather, push, bin by new cell ------------------------------------------------------------
        {
            const double2* ep = s.ep[D.epi];
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
                    gather_pairs(ep, c.a[0] - A.m.nghost - D.hx * TILE, c.a[1] - A.m.nghost - D.hy * TILE,
                                 c.a[2] - A.m.nghost - D.hz * TILE, c.whi, E);
                    push_particle(A.P, r, p, E);
                    Cic cn;
                    cic_setup(A.m, r[0], r[1], r[2], cn);
                    const int cc[3] = {cn.a[0] - A.m.nghost, cn.a[1] - A.m.nghost, cn.a[2] - A.m.nghost};
                    if (!owned_by_me(A, r, cc)) {
                        place_exit(A, r, p);
                    } else {
                        const int wx = cc[0] - wox, wy = cc[1] - woy, wz = cc[2] - woz;
                        if ((unsigned)wx < (unsigned)WIN && (unsigned)wy < (unsigned)WIN &&
                            (unsigned)wz < (unsigned)WIN) {
                            const int id = s.winid[(wz * WIN + wy) * WIN + wx];
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
        }
        consumer_sync<NT>();
        // ---- P2: exclusive scan of the window histogram (tile-major ids), 2 cells per thread ---------------
        static_assert(NT >= WIN_CELLS / 2, "the scan uses WIN_CELLS / 2 threads");
        // value = count | (count > 0) << 16: one scan yields the particle prefix and the ordered list of non-empty cells
        int v0 = 0, v1 = 0, inc = 0;
        if (t < WIN_CELLS / 2) {
            v0 = s.hist[2 * t];
            v1 = s.hist[2 * t + 1];
            v0 |= (v0 > 0) << 16;
            v1 |= (v1 > 0) << 16;
            inc = v0 + v1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            if (lane == 31) s.warp_sums[warp] = inc;
        }
        consumer_sync<NT>();
        if (t < WIN_CELLS / 2) {
            int run = inc - v0 - v1;
#pragma unroll
            for (int w = 0; w < WIN_CELLS / 64 - 1; ++w)
                if (w < warp) run += s.warp_sums[w];
            s.prefix[2 * t]     = run & 0xFFFF;
            s.prefix[2 * t + 1] = (run + v0) & 0xFFFF;
            if (v0 >> 16) {
                s.list[run >> 16] = (unsigned short)(2 * t);
                s.hist[2 * t]     = 0;  // ready for the next chunk
            }
            if (v1 >> 16) {
                s.list[(run + v0) >> 16] = (unsigned short)(2 * t + 1);
                s.hist[2 * t + 1]        = 0;
            }
            if (t == WIN_CELLS / 2 - 1) {
                s.prefix[WIN_CELLS] = (run + v0 + v1) & 0xFFFF;
                s.nne               = (run + v0 + v1) >> 16;
            }
        }
        consumer_sync<NT>();
        // ---- reserve one block per destination tile (global atomics, consumed only after the next barrier so
        //      their latency hides behind P3 and the P4 loads); P3: sorted position -> arrival slot -----------
        int rs_b0 = 0, rs_n = 0, rs_base = 0, rs_cap = 0, rs_start = 0;
        bool rs_bad = false;
        if (t < NSLOT) {
            rs_b0 = s.prefix[s.tsbase[t]];
            rs_n  = s.prefix[s.tsbase[t + 1]] - rs_b0;
            if (rs_n > 0) {
                const int tx = D.hx + (t % 3) - 1, ty = D.hy + ((t / 3) % 3) - 1, tz = D.hz + (t / 9) - 1;
                rs_bad = tx < 0 || tx >= A.ntx || ty < 0 || ty >= A.nty || tz < 0 || tz >= A.ntz;
                if (!rs_bad) {
                    const int tile = tx + A.ntx * (ty + A.nty * tz);
                    rs_base        = atomicAdd(&A.cursor_out[tile], rs_n);
                    rs_cap         = A.cap_out[tile];
                    rs_start       = A.start_out[tile];
                }
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int slot           = k * NT + t;
            const unsigned short loc = s.local[slot];
            if (loc != NOSLOT) s.perm[s.prefix[loc] + s.rank[slot]] = (unsigned short)slot;
        }
        consumer_sync<NT>();
        // ---- P4: sorted order -> registers (random shared-memory reads) ----------------------------------------
        const int ntot = s.prefix[WIN_CELLS];
        double pr[K][6];
        int pts[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int p = k * NT + t;
            pts[k]      = -1;
            if (p < ntot) {
                const int slot = s.perm[p];
                pts[k]         = s.tsof[s.local[slot]];
#pragma unroll
                for (int a = 0; a < 6; ++a) pr[k][a] = G.dat[a][slot];
            }
        }
        if (t < NSLOT) {
            int adj = 0, lim = 0, tadj = 0;
            if (rs_n > 0) {
                if (rs_bad) {
                    atomicOr(&A.misc[BM_FLAGS], IPPLB_FLAG_INTERNAL);
                    tadj = INT_MIN;  // lim = 0: everything of this block is dropped
                } else {
                    const int abs0 = rs_start + rs_base;
                    adj            = abs0 - rs_b0;
                    lim            = rs_start + rs_cap;
                    const int g0   = max(lim, abs0);  // first absolute slot that does not fit
                    const int over = abs0 + rs_n - g0;
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
        consumer_sync<NT>();
        // ---- coalesced store into next step's buckets; CIC weights of the new positions, sorted order --------
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int p = k * NT + t;
            if (pts[k] >= 0) {
                const int ts = pts[k];
                int g        = s.adj[ts] + p;
                if (g >= s.lim[ts]) {
                    const int ta = s.tadj[ts];
                    g            = ta == INT_MIN ? -1 : g + ta;
                }
                if (g >= 0) {  // negative: dropped (flag already raised)
#pragma unroll
                    for (int a = 0; a < 6; ++a) A.out[a][g] = pr[k][a];
                }
                int idx;
                double w0, w1, w2;
                cic_axis(pr[k][0], A.m.origin[0], A.m.invdx[0], idx, w0);
                cic_axis(pr[k][1], A.m.origin[1], A.m.invdx[1], idx, w1);
                cic_axis(pr[k][2], A.m.origin[2], A.m.invdx[2], idx, w2);
                G.dat[3][p] = w0;
                G.dat[4][p] = w1;
                G.dat[5][p] = w2;
            }
        }
        consumer_sync<NT>();
        // ---- P5: deposit, one lane per non-empty cell (list is in tile-major id order: the 64 home cells, which hold
        //      most particles, are neighbours in the list, so the lanes of a warp run similar trip counts).  The eight
        //      node sums follow from eight moments of the weights: 4 multiplies + 7 adds per particle.
        {
            const int nne = s.nne;
            for (int j = t; j < nne; j += NT) {
                const int id = s.list[j];
                const int b = s.prefix[id], e = s.prefix[id + 1];
                double s1 = 0.0, s2 = 0.0, s3 = 0.0, s12 = 0.0, s13 = 0.0, s23 = 0.0, s123 = 0.0;
                for (int p = b; p < e; ++p) {
                    const double w0 = G.dat[3][p], w1 = G.dat[4][p], w2 = G.dat[5][p];
                    const double w01 = w0 * w1;
                    s1 += w0;
                    s2 += w1;
                    s3 += w2;
                    s12 += w01;
                    s13 += w0 * w2;
                    s23 += w1 * w2;
                    s123 += w01 * w2;
                }
                const double cnt = (double)(e - b);
                double nd[8];  // node n: bit d set -> lower node along d (weight 1 - w_d)
                nd[0] = s123;
                nd[1] = s23 - s123;
                nd[2] = s13 - s123;
                nd[3] = (s3 - s13) - (s23 - s123);
                nd[4] = s12 - s123;
                nd[5] = (s2 - s12) - (s23 - s123);
                nd[6] = (s1 - s12) - (s13 - s123);
                nd[7] = ((cnt - s1) - (s2 - s12)) - ((s3 - s13) - (s23 - s123));
                const unsigned xyz = s.cellxyz[id];
                const int a[3]     = {(int)(xyz & 15) + wox + A.m.nghost, (int)((xyz >> 4) & 15) + woy + A.m.nghost,
                                      (int)(xyz >> 8) + woz + A.m.nghost};
#pragma unroll
                for (int n = 0; n < 8; ++n) atomicAdd(&A.rho[cic_node(A.m, a, n)], A.q * nd[n]);
            }
        }
        // every consumer releases the stage on its own (no CTA barrier between chunks): the next chunk's P1 only
        // touches the other stage, hist (already reset) and local / rank (dead since P4)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[st]);
    }
}


// ---- generation 3: sorted records + per-cell moment accumulators in tensor memory ----------------------------
// Same producer, same P1 arithmetic.  What changes:
//   * the pushed particle stays in registers across the scan and is written ONCE, as a 48-byte record
//     (x y | z px | py pz) at its sorted position of the (dead) stage: no write-back, no permutation arrays;
//   * the scan needs two barriers instead of three: each warp forms the 8 segment offsets with shuffles;
//   * the deposit no longer issues reductions per (chunk, cell).  Every window cell has a fixed owner thread; the
//     owner adds the chunk's 8 weight moments of its cell to accumulators that live in TENSOR MEMORY
//     (tcgen05.ld / tcgen05.st, 32x32b.x16: 16 columns = 8 doubles per lane, private to the warp, no shared-memory
//     or LSU traffic).  After the tile's last chunk every owner turns its cell's moments into the 8 node sums and
//     adds them to rho with one RED.F64 per node: ~0.5 reductions per particle instead of 1.6, and the end of a
//     tile needs no barrier and no staging (an earlier variant summed the cells into a 9x9x9 node lattice in shared
//     memory first -- 0.2 reductions per particle, but 9 CTA barriers per tile and 5.8 KB of shared memory, which is
//     what now pays for the 15th consumer warp).

template <int NT, int K>
struct Step3Smem {
    static constexpr int CAP = NT * K;
    struct Stage {
        double dat[6][CAP];  // as loaded: SoA planes; after the scan: CAP 48-byte records in sorted order
    } st[2];
    double2 ep[2][EP_N];
    ChunkDesc desc[2];
    unsigned long long full[2], empty[2];
    int hist[WIN_CELLS];
    int lpre[WIN_CELLS + 1];    // exclusive prefix inside the 64-cell scan segment
    int prefix[WIN_CELLS + 1];  // global particle prefix per window id (valid after barrier D)
    unsigned short cellxyz[WIN_CELLS];
    unsigned short winid[WIN_CELLS];
    unsigned char tsof[WIN_CELLS];
    unsigned char tsp[CAP];  // sorted position -> destination tile slot
    int tsbase[NSLOT + 1];
    int adj[NSLOT], lim[NSLOT], tadj[NSLOT];
    int seg_sums[WIN_CELLS / 64];
    int total;
    uint32_t tmem_base;
};

// tensor memory as lane-private scratch: 8 doubles per (warp, slot)
__device__ __forceinline__ void tmem_ld8(uint32_t addr, double v[8]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
        "[%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const double v[8]) {
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r[2 * i]     = (uint32_t)__double2loint(v[i]);
        r[2 * i + 1] = (uint32_t)__double2hiint(v[i]);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16};\n"
        "tcgen05.wait::st.sync.aligned;\n" ::"r"(addr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// a pushed particle (6 doubles = 12 columns) parked in the thread's own tensor-memory lane across the scan
__device__ __forceinline__ void tmem_st6(uint32_t addr, const double v[6]) {
    uint32_t r[12];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        r[2 * i]     = (uint32_t)__double2loint(v[i]);
        r[2 * i + 1] = (uint32_t)__double2hiint(v[i]);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
        "tcgen05.st.sync.aligned.32x32b.x4.b32 [%9], {%10, %11, %12, %13};\n" ::"r"(addr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(addr + 8), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11])
        : "memory");
}
__device__ __forceinline__ void tmem_ld6(uint32_t addr, double v[6]) {
    uint32_t r[12];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%12];\n"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [%13];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11])
        : "r"(addr), "r"(addr + 8)
        : "memory");
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
constexpr int tmem_cols_pow2(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int NT, int K, int MINB>
__global__ void __launch_bounds__(NT + 32, MINB) fused_step3_kernel(const StepArgs A) {
    using S            = Step3Smem<NT, K>;
    constexpr int CAP  = S::CAP;
    constexpr int NW   = NT / 32;
    constexpr int NSEG = WIN_CELLS / 64;
    constexpr int NSL  = (WIN_CELLS + NT - 1) / NT;  // window cells owned per thread
    constexpr int WCOLS = NSL * 16 + K * 16;  // tensor-memory columns per warp: accumulators + parked particles
    constexpr int TCOLS = tmem_cols_pow2(((NW + 3) / 4) * WCOLS);
    static_assert(NT >= WIN_CELLS / 2, "the scan uses WIN_CELLS / 2 threads");
    static_assert(NSEG <= 16, "segment offsets live in one warp");
    static_assert(TCOLS * MINB <= 512, "tensor memory columns of the co-resident CTAs");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    S& s           = *reinterpret_cast<S*>(smem_raw);
    const int t    = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;

    if (t == 0) {
        int run = 0;
        for (int ts = 0; ts < NSLOT; ++ts) {
            s.tsbase[ts] = run;
            const int wx = (ts % 3 == 1) ? 4 : 2, wy = ((ts / 3) % 3 == 1) ? 4 : 2, wz = (ts / 9 == 1) ? 4 : 2;
            run += wx * wy * wz;
        }
        s.tsbase[NSLOT]   = run;
        s.lpre[WIN_CELLS] = 0;
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s.full[i], 1);
            mbar_init(&s.empty[i], NW);
        }
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s.tmem_base)),
                     "n"(TCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    for (int c = t; c < WIN_CELLS; c += NT + 32) {
        const int wx = c & 7, wy = (c >> 3) & 7, wz = c >> 6;
        int sx, lx, dx, sy, ly, dy, sz, lz, dz;
        win_seg(wx, sx, lx, dx);
        win_seg(wy, sy, ly, dy);
        win_seg(wz, sz, lz, dz);
        const int ts = sx + 3 * (sy + 3 * sz);
        const int id = s.tsbase[ts] + (lz * dy + ly) * dx + lx;
        s.cellxyz[id] = (unsigned short)(wx | (wy << 4) | (wz << 8));
        s.winid[c]    = (unsigned short)id;
        s.tsof[id]    = (unsigned char)ts;
        s.hist[c]     = 0;
    }
    __syncthreads();

    if (warp == NW) {
        producer_loop<S>(A, s, lane);
        return;
    }
    // this warp's accumulator columns: lanes 32 * (warp % 4) of tensor memory, 16 columns per owned cell slot
    const uint32_t tm0 = s.tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * WCOLS);
    const uint32_t tmp = tm0 + NSL * 16;  // parked particles
    {
        const double z[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) tmem_st8(tm0 + sl * 16, z);
    }

    for (unsigned seq = 0;; ++seq) {
        const int st = seq & 1;
        mbar_wait(&s.full[st], (seq >> 1) & 1);
        const ChunkDesc D = s.desc[st];
        if (D.kind == CH_STOP) break;
        typename S::Stage& G = s.st[st];
        const int cnt        = D.cnt;

        if (D.kind == CH_TAIL) {
            tail_chunk<NT, K>(A, G.dat, cnt, t);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.empty[st]);
            continue;
        }

        const int wox = D.hx * TILE - WH, woy = D.hy * TILE - WH, woz = D.hz * TILE - WH;
        // ---- P1: gather, push, bin by new cell; the pushed particle stays in registers ----------------------
        int lr[K];  // window id | rank inside the cell << 16, or -1
        {
            const double2* ep = s.ep[D.epi];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int slot = k * NT + t;
                lr[k]          = -1;
                double pk[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                if (slot < cnt) {
                    double r[3] = {G.dat[0][slot], G.dat[1][slot], G.dat[2][slot]};
                    double p[3] = {G.dat[3][slot], G.dat[4][slot], G.dat[5][slot]};
                    Cic c;
                    cic_setup(A.m, r[0], r[1], r[2], c);
                    double E[3];
                    gather_pairs(ep, c.a[0] - A.m.nghost - D.hx * TILE, c.a[1] - A.m.nghost - D.hy * TILE,
                                 c.a[2] - A.m.nghost - D.hz * TILE, c.whi, E);
                    push_particle(A.P, r, p, E);
                    Cic cn;
                    cic_setup(A.m, r[0], r[1], r[2], cn);
                    const int cc[3] = {cn.a[0] - A.m.nghost, cn.a[1] - A.m.nghost, cn.a[2] - A.m.nghost};
                    if (!owned_by_me(A, r, cc)) {
                        place_exit(A, r, p);
                    } else {
                        const int wx = cc[0] - wox, wy = cc[1] - woy, wz = cc[2] - woz;
                        if ((unsigned)wx < (unsigned)WIN && (unsigned)wy < (unsigned)WIN &&
                            (unsigned)wz < (unsigned)WIN) {
                            const int id = s.winid[(wz * WIN + wy) * WIN + wx];
                            lr[k]        = id | (atomicAdd(&s.hist[id], 1) << 16);
                        } else {
                            place_direct(A, r, p, cc, cn.whi);
                        }
                    }
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        pk[d]     = r[d];
                        pk[3 + d] = p[d];
                    }
                }
                tmem_st6(tmp + k * 16, pk);  // parked in tensor memory until the scan is done
            }
            tmem_wait_st();
        }
        consumer_sync<NT>();  // ---- A: every input slot has been read, the histogram is complete
        // scan part 1: two cells per thread, exclusive inside the warp's 64-cell segment
        int v0 = 0, v1 = 0;
        if (t < WIN_CELLS / 2) {
            v0      = s.hist[2 * t];
            v1      = s.hist[2 * t + 1];
            int inc = v0 + v1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            s.lpre[2 * t]     = inc - v0 - v1;
            s.lpre[2 * t + 1] = inc - v1;
            if (lane == 31) s.seg_sums[warp] = inc;
            s.hist[2 * t]     = 0;  // ready for the next chunk (its P1 starts after barrier D)
            s.hist[2 * t + 1] = 0;
        }
        consumer_sync<NT>();  // ---- B
        // segment offsets: lane l < NSEG holds the exclusive offset of segment l, lane NSEG the grand total
        int segoff;
        {
            const int v = lane < NSEG ? s.seg_sums[lane] : 0;
            int inc     = v;
#pragma unroll
            for (int o = 1; o < 2 * NSEG; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            segoff = inc - v;  // lanes >= NSEG: v = 0, inc = total
        }
        // global prefix of window id c (c <= WIN_CELLS); every lane of the warp has to call it
        auto gpre = [&](int c) { return s.lpre[c] + __shfl_sync(0xffffffffu, segoff, c >> 6); };
        if (t < WIN_CELLS / 2) {  // published for P4 / P5 (after D)
            const int run       = gpre(2 * t);
            s.prefix[2 * t]     = run;
            s.prefix[2 * t + 1] = run + v0;
            if (t == WIN_CELLS / 2 - 1) {
                s.prefix[WIN_CELLS] = run + v0 + v1;
                s.total             = run + v0 + v1;
            }
        }
        // one reservation per destination tile (27 lanes of warp 0): the global atomics are in flight while the warp
        // writes its records
        int rs_b0 = 0, rs_n = 0, rs_base = 0, rs_cap = 0, rs_start = 0;
        bool rs_bad = false;
        if (warp == 0) {
            const int c0 = s.tsbase[min(lane, NSLOT)], c1 = s.tsbase[min(lane + 1, NSLOT)];
            rs_b0 = gpre(c0);
            rs_n  = gpre(c1) - rs_b0;
            if (lane < NSLOT && rs_n > 0) {
                const int tx = D.hx + (lane % 3) - 1, ty = D.hy + ((lane / 3) % 3) - 1, tz = D.hz + (lane / 9) - 1;
                rs_bad = tx < 0 || tx >= A.ntx || ty < 0 || ty >= A.nty || tz < 0 || tz >= A.ntz;
                if (!rs_bad) {
                    const int tile = tx + A.ntx * (ty + A.nty * tz);
                    rs_base        = atomicAdd(&A.cursor_out[tile], rs_n);
                    rs_cap         = A.cap_out[tile];
                    rs_start       = A.start_out[tile];
                }
            }
        }
        // records at their sorted position
        double2* rec = reinterpret_cast<double2*>(&G.dat[0][0]);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int id  = lr[k] < 0 ? 0 : (lr[k] & 0xFFFF);
            const int pre = gpre(id);
            double pk[6];
            tmem_ld6(tmp + k * 16, pk);
            if (lr[k] >= 0) {
                const int pos    = pre + (lr[k] >> 16);
                rec[3 * pos]     = make_double2(pk[0], pk[1]);
                rec[3 * pos + 1] = make_double2(pk[2], pk[3]);
                rec[3 * pos + 2] = make_double2(pk[4], pk[5]);
                s.tsp[pos]       = s.tsof[id];
            }
        }
        if (warp == 0 && lane < NSLOT) {
            int adj = 0, lim = 0, tadj = 0;
            if (rs_n > 0) {
                if (rs_bad) {
                    atomicOr(&A.misc[BM_FLAGS], IPPLB_FLAG_INTERNAL);
                    tadj = INT_MIN;  // lim = 0: everything of this block is dropped
                } else {
                    const int abs0 = rs_start + rs_base;
                    adj            = abs0 - rs_b0;
                    lim            = rs_start + rs_cap;
                    const int g0   = max(lim, abs0);  // first absolute slot that does not fit
                    const int over = abs0 + rs_n - g0;
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
            s.adj[lane]  = adj;
            s.lim[lane]  = lim;
            s.tadj[lane] = tadj;
        }
        consumer_sync<NT>();  // ---- D: records, prefix and the block placement are published
        // ---- P4: coalesced store of the sorted records into next step's buckets; the CIC weights of the new
        //      position replace the head of the (now dead) record: (w0 w1 | w2 .)
        {
            const int ntot = s.total;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int p = k * NT + t;
                if (p < ntot) {
                    const double2 a = rec[3 * p], b = rec[3 * p + 1], c = rec[3 * p + 2];
                    const int ts    = s.tsp[p];
                    int g           = s.adj[ts] + p;
                    if (g >= s.lim[ts]) {
                        const int ta = s.tadj[ts];
                        g            = ta == INT_MIN ? -1 : g + ta;
                    }
                    if (g >= 0) {  // negative: dropped (flag already raised)
                        A.out[0][g] = a.x;
                        A.out[1][g] = a.y;
                        A.out[2][g] = b.x;
                        A.out[3][g] = b.y;
                        A.out[4][g] = c.x;
                        A.out[5][g] = c.y;
                    }
                    int idx;
                    double w0, w1, w2;
                    cic_axis(a.x, A.m.origin[0], A.m.invdx[0], idx, w0);
                    cic_axis(a.y, A.m.origin[1], A.m.invdx[1], idx, w1);
                    cic_axis(b.x, A.m.origin[2], A.m.invdx[2], idx, w2);
                    rec[3 * p]       = make_double2(w0, w1);
                    rec[3 * p + 1].x = w2;
                }
            }
        }
        consumer_sync<NT>();  // ---- E: weights are in place
        // ---- P5: every window cell has a fixed owner (window id = slot * NT + thread); the owner adds the chunk's
        //      moments of its cell to the cell's accumulators in tensor memory
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
            if (sl * NT + warp * 32 < WIN_CELLS) {  // warp-uniform
                const int id = sl * NT + t;
                const int b = s.prefix[id], e = s.prefix[id + 1];
                if (__any_sync(0xffffffffu, e > b)) {
                    double m[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // 1 w0 w1 w2 w0w1 w0w2 w1w2 w0w1w2
                    for (int p = b; p < e; ++p) {
                        const double2 w01 = rec[3 * p];
                        const double w2   = rec[3 * p + 1].x;
                        const double p01  = w01.x * w01.y;
                        m[1] += w01.x;
                        m[2] += w01.y;
                        m[3] += w2;
                        m[4] += p01;
                        m[5] += w01.x * w2;
                        m[6] += w01.y * w2;
                        m[7] += p01 * w2;
                    }
                    m[0] = (double)(e - b);
                    double acc[8];
                    tmem_ld8(tm0 + sl * 16, acc);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i] += m[i];
                    tmem_st8(tm0 + sl * 16, acc);
                }
            }
        }
        // every consumer warp releases the stage on its own: the next chunk's P1 touches the other stage and hist
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.empty[st]);

        if (D.pad[0]) {
            // ---- last chunk of the tile: every owner turns its cell's moments into the 8 node sums and adds them to
            //      rho with one RED.F64 per node -- no barrier, no staging: the warps run on into the next tile
            const double z8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int sl = 0; sl < NSL; ++sl) {
                if (sl * NT + warp * 32 < WIN_CELLS) {
                    double a[8];
                    tmem_ld8(tm0 + sl * 16, a);
                    tmem_st8(tm0 + sl * 16, z8);
                    if (a[0] != 0.0) {
                        const double s1 = a[1], s2 = a[2], s3 = a[3], s12 = a[4], s13 = a[5], s23 = a[6], s123 = a[7];
                        double nd[8];  // node n: bit d set -> lower node along d (weight 1 - w_d)
                        nd[0] = s123;
                        nd[1] = s23 - s123;
                        nd[2] = s13 - s123;
                        nd[3] = (s3 - s13) - (s23 - s123);
                        nd[4] = s12 - s123;
                        nd[5] = (s2 - s12) - (s23 - s123);
                        nd[6] = (s1 - s12) - (s13 - s123);
                        nd[7] = ((a[0] - s1) - (s2 - s12)) - ((s3 - s13) - (s23 - s123));
                        const unsigned xyz = s.cellxyz[sl * NT + t];
                        // ghosted index of the cell's upper node = window coordinate + window origin + nghost
                        const long gx = (long)(xyz & 15) + wox + A.m.nghost, gy = (long)((xyz >> 4) & 15) + woy + A.m.nghost,
                                   gz = (long)(xyz >> 8) + woz + A.m.nghost;
#pragma unroll
                        for (int n = 0; n < 8; ++n)
                            atomicAdd(&A.rho[(gx - (n & 1)) + (long)A.m.ex * ((gy - ((n >> 1) & 1)) + (long)A.m.ey * (gz - ((n >> 2) & 1)))],
                                      A.q * nd[n]);
                    }
                }
            }
        }
    }
    consumer_sync<NT>();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(s.tmem_base), "n"(TCOLS) : "memory");
    }
}

template <int NT, int K, int MINB>
static int launch_fused(ipplb_ctx* ctx, const StepArgs& A) {
    using S   = StepSmem<NT, K>;
    auto kern = fused_step_kernel<NT, K, MINB>;
    static bool attr_set = false;
    if (!attr_set) {
        IPPLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S)));
        attr_set = true;
    }
    kern<<<ctx->num_sms * MINB, NT + 32, sizeof(S), ctx->stream>>>(A);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

template <int NT, int K, int MINB>
static int launch_fused3(ipplb_ctx* ctx, const StepArgs& A) {
    using S   = Step3Smem<NT, K>;
    auto kern = fused_step3_kernel<NT, K, MINB>;
    static bool attr_set = false;
    if (!attr_set) {
        IPPLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S)));
        attr_set = true;
    }
    kern<<<ctx->num_sms * MINB, NT + 32, sizeof(S), ctx->stream>>>(A);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

// tuning knob (experiments only): IPPLB_FUSED_CFG selects the CTA shape; the default is the measured best
static int fused_cfg() {
    static int cfg = -1;
    if (cfg < 0) {
        const char* e = getenv("IPPLB_FUSED_CFG");
        cfg           = e ? atoi(e) : 0;
    }
    return cfg;
}

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
        IPPLB_REQUIRE(((uintptr_t)exit_buf & 15) == 0, "bins_step: the exit buffer must be 16-byte aligned");
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
    const int nrk = ctx->d_regions ? ctx->nranks : 1;
    IPPLB_REQUIRE(nrk <= MAX_RANKS, "bins_step: too many ranks for the exit buffer segmentation");
    A.exit_buf   = exit_buf;
    A.exit_cnt   = b->d_exit_cnt;
    A.seg_cap    = exit_buf ? exit_cap / nrk : 0;
    A.regions    = ctx->d_regions;
    A.nranks     = nrk;
    A.me         = ctx->rank;
    b->exit_ranks = nrk;
    A.capacity   = (int)b->capacity;
    A.ntx = b->ntx; A.nty = b->nty; A.ntz = b->ntz; A.ntiles = b->ntiles;
    A.check_owner = (region_min && region_max) ? 1 : 0;
    A.balance_chunks = (fused_cfg() / 100) == 2 ? 0 : 1;  // on by default (4.25 -> 4.20 ms at C2); IPPLB_FUSED_CFG=2xx turns it off
    for (int d = 0; d < 3; ++d) {
        A.rmin[d] = region_min ? region_min[d] : 0.0;
        A.rmax[d] = region_max ? region_max[d] : 0.0;
    }
    IPPLB_CUDA(cudaMemsetAsync(b->misc(), 0, sizeof(int) * 4, ctx->stream));
    IPPLB_CUDA(cudaMemsetAsync(b->d_exit_cnt, 0, sizeof(int) * nrk, ctx->stream));
    int rc;
    switch (fused_cfg() % 100) {
        case 1: rc = launch_fused<256, 2, 3>(ctx, A); break;
        case 2: rc = launch_fused<512, 2, 1>(ctx, A); break;
        case 3: rc = launch_fused<768, 2, 1>(ctx, A); break;
        case 4: rc = launch_fused<256, 2, 2>(ctx, A); break;
        case 10: rc = launch_fused3<384, 2, 2>(ctx, A); break;
        case 11: rc = launch_fused3<256, 2, 3>(ctx, A); break;
        case 12: rc = launch_fused3<256, 3, 2>(ctx, A); break;
        case 13: rc = launch_fused3<512, 2, 1>(ctx, A); break;
        case 14: rc = launch_fused3<352, 2, 2>(ctx, A); break;
        case 15: rc = launch_fused3<320, 2, 2>(ctx, A); break;
        case 17: rc = launch_fused3<416, 2, 2>(ctx, A); break;
        case 5: rc = launch_fused<384, 2, 2>(ctx, A); break;  // generation 2
        case 16: rc = launch_fused3<448, 2, 2>(ctx, A); break;
        default: rc = launch_fused3<480, 2, 2>(ctx, A); break;  // generation 3, 15 consumer warps (measured best)
    }
    if (rc) return rc;
    rc = bins_plan(ctx, b, o, A.seg_cap);

    if (rc) return rc;
    b->cur        = o;
    nxt->q        = nullptr;
    nxt->q_scalar = cur->q_scalar;
    return IPPLB_OK;
}

}  // extern "C"
