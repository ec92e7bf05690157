// fused.cu -- the B200-first PIC step: ONE pass over bucketed particles (see include/ippl_b200.h, ipplb_bins).
//
// Storage: per-tile buckets (tile = 4x4x4 key cells) with slack, double-buffered; device-resident tables
// start/cap/count per tile, re-planned on the device after every step (bins.cu).
//
// Kernel (persistent, 2 CTAs per SM, dynamic tile scheduler, 15 consumer warps + 1 producer warp):
//   producer warp : walks (tile, chunk) work items two deep, streams the chunk's six SoA runs into a 2-stage
//                   shared-memory ring with bulk async copies (cp.async.bulk + mbarrier complete_tx, TMA 1-D)
//                   and stages the tile's 5x5x5-node E window as x-pairs so the gather is 12 LDS.128.
//   consumer warps: P1 gather E from shared memory, kick/kick/drift/BC in registers, bin by NEW cell with
//                      integer shared-memory atomics on an 8x8x8-cell window (tile-major ids); the pushed
//                      particle is parked in TENSOR MEMORY (tcgen05.st) across the scan;
//                   P2 scan of the window histogram (8 warps) while a ninth warp sums the histogram per
//                      destination tile and issues the <= 27 bucket reservations (global atomics);
//                   P3 every particle is written once, as a 48-byte record, at its sorted position;
//                   P4 coalesced store of the records into next step's buckets (overflow -> tail); the CIC
//                      weights of the new position replace the head of the record;
//                   P5 deposit: every window cell has a fixed owner thread that adds the chunk's 8 weight
//                      moments of its cell to accumulators in tensor memory; node sums go to rho once per
//                      tile (one RED.F64 per node and cell).
// HBM traffic: 48 B read + 48 B written per particle.  E never round-trips through HBM, rho is produced
// without re-reading the particles, and there is no sort pass.  Positions and momenta are bit-identical to
// the unfused path (same device functions, same operation order).
#include <climits>
#include <cstdint>
#include <cstdlib>

#include "bins.h"
#include "push.cuh"

// A/B knob of the build (scripts/build_variants.sh): rare paths in line or out of line (out of line costs ~18 %: the
// call ABI pins registers and the 64-register kernel spills)
#ifndef IPPLB_SLOW_ATTR
#define IPPLB_SLOW_ATTR __forceinline__
#endif
#ifdef IPPLB_NO_WRAP   // A/B: periodic aliasing compiled out
#define IPPLB_WRAP(A) false
#define IPPLB_WRAP_CT(S) false
#else
#define IPPLB_WRAP(A) ((A).wrap != 0)
#ifdef IPPLB_WRAP_RUNTIME   // A/B: runtime flag only
#define IPPLB_WRAP_CT(S) false
#else
#define IPPLB_WRAP_CT(S) (S)
#endif
#endif

namespace ipplb {

constexpr int WH        = 2;              // window halo (cells) around the home tile
constexpr int WIN       = TILE + 2 * WH;  // 8
constexpr int WIN_CELLS = WIN * WIN * WIN;
constexpr int NSLOT     = 27;             // destination tiles touched by the window
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
    double* const* peer_seg;  // peer mode: [nranks] the destination ranks' inboxes ([nranks][seg_cap][6] each); this rank
                              // writes segment `me` of inbox d over peer memory (NVLink) -- or nullptr
    int* exit_cnt;     // [nranks]
    int seg_cap;
    const double* regions;  // [nranks][6] (device) or nullptr on a single rank
    int nranks, me;
    int capacity;
    int ntx, nty, ntz, ntiles;
    int check_owner;
    int balance_chunks;  // equal chunks per tile instead of full chunks + a remainder
    int wrap;  // the rank owns the whole periodic domain: E reads and rho adds alias ghost nodes to the opposite interior layer
               // (no fillHalo(E) needed before the step, no accumulateHalo(rho) after it; rho's ghost layers stay untouched)
    double rmin[3], rmax[3];
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
// the same wait with a suspend-time hint (ns): the warp sleeps in hardware until the phase completes instead of
// re-issuing the test every ~40 ns (a spinning producer warp otherwise takes issue slots from its scheduler's consumers)
__device__ __forceinline__ void mbar_wait_hint(unsigned long long* b, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(b)),
        "r"(parity), "r"(ns)
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
// The two functions below are rare paths: they are kept OUT OF LINE (the kernel's parameter block is __grid_constant__,
// so its address can be passed) to keep the hot loop short -- the whole kernel has to stay resident in the instruction
// cache while two CTAs per SM run different phases of it.
static __device__ IPPLB_SLOW_ATTR void place_exit(const StepArgs& A, double x, double y, double z, double px, double py, double pz) {
    int d = 0;
    if (A.regions) {
        d = -1;
        for (int k = 0; d < 0 && k < A.nranks; ++k) {
            const double* R = A.regions + 6 * k;
            if (x > R[0] && y > R[1] && z > R[2] && x <= R[3] && y <= R[4] && z <= R[5]) d = k;
        }
        for (int k = 0; d < 0 && k < A.nranks; ++k) {
            const double* R = A.regions + 6 * k;
            if (x >= R[0] && y >= R[1] && z >= R[2] && x <= R[3] && y <= R[4] && z <= R[5]) d = k;
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
        double* base = A.peer_seg ? A.peer_seg[d] + (size_t)A.me * A.seg_cap * 6 : A.exit_buf + (size_t)d * A.seg_cap * 6;
        double2* rec = reinterpret_cast<double2*>(base + (size_t)e * 6);
        rec[0]       = make_double2(x, y);
        rec[1]       = make_double2(z, px);
        rec[2]       = make_double2(py, pz);
    }
}

// a particle whose destination is outside the chunk's window (or that sits in the unsorted tail): claim one
// slot of the destination bucket, scattered write, 8 reductions
static __device__ IPPLB_SLOW_ATTR void place_direct(const StepArgs& A, const double r[3], const double p[3], const int c[3],
                                                    const double whi[3]) {
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
    for (int n = 0; n < 8; ++n)
        atomicAdd(&A.rho[IPPLB_WRAP(A) ? cic_node_wrapped(A.m, a, n) : cic_node(A.m, a, n)], dmul(A.q, cic_weight(whi, n)));
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
template <typename S, bool HINT, bool WRAP_CT>
__device__ __forceinline__ void producer_loop(const StepArgs& A, S& s, const int lane) {
    constexpr int CAP = S::CAP;
    const int tail_start = A.state_in[BS_TAIL_START];
    const int tail_count = A.state_in[BS_TAIL_COUNT];
    int rem = 0, kind = CH_STOP, hx = 0, hy = 0, hz = 0, chunk = CAP;
    bool fresh = false;  // first chunk of a tile
    long pbeg = 0;
    // E windows: tile number k (counting the non-empty tiles this CTA works on) uses buffer k & 1 and signals it through
    // eready[k & 1].  The window of the NEXT tile is staged one tile ahead, behind the second chunk of the current tile
    // (when the previous tile's window is dead), so that a tile never starts by waiting for ~20 dependent loads per lane.
    int tile_seq = 0, buf_cur = 0, cur_tile = -1, staged_for = -1;
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
    // E window of a tile as x-pairs: ep[c][kz][jy][ix] = (E_c(node ix), E_c(node ix+1)); node (0,0,0) is the lower node
    // of the tile's first cell, ghosted index = 4*h + nghost - 1
    auto stage_window = [&](int tile, int buf) {
        const int tx = tile % A.ntx, ty = (tile / A.ntx) % A.nty, tz = tile / (A.ntx * A.nty);
        const int gx0 = 4 * tx + A.m.nghost - 1, gy0 = 4 * ty + A.m.nghost - 1, gz0 = 4 * tz + A.m.nghost - 1;
        for (int e = lane; e < EP_N; e += 32) {
            const int ix = e & 3, jy = (e >> 2) % 5, kz = ((e >> 2) / 5) % 5, c = (e >> 2) / 25;
            int gx = gx0 + ix, gx1 = gx + 1, gy = gy0 + jy, gz = gz0 + kz;
            double2 v = make_double2(0.0, 0.0);
            if (gy < A.m.ey && gz < A.m.ez) {
                const bool okx = gx < A.m.ex, okx1 = gx1 < A.m.ex;
                if (IPPLB_WRAP_CT(WRAP_CT) || IPPLB_WRAP(A)) {
                    gx  = wrap_axis(gx, A.m.nl[0], A.m.nghost);
                    gx1 = wrap_axis(gx1, A.m.nl[0], A.m.nghost);
                    gy  = wrap_axis(gy, A.m.nl[1], A.m.nghost);
                    gz  = wrap_axis(gz, A.m.nl[2], A.m.nghost);
                }
                const long row = (long)A.m.ex * (gy + (long)A.m.ey * gz);
                if (okx) v.x = __ldg(&A.ef[(gx + row) * 3 + c]);
                if (okx1) v.y = __ldg(&A.ef[(gx1 + row) * 3 + c]);
            }
            s.ep[buf][e] = v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s.eready[buf]);
    };
    issue_c();
    load_b();
    issue_c();
    for (unsigned seq = 0;; ++seq) {
        const int st = seq & 1;
        if (HINT) mbar_wait_hint(&s.empty[st], ((seq >> 1) & 1) ^ 1, 20000);
        else mbar_wait(&s.empty[st], ((seq >> 1) & 1) ^ 1);
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
                fresh    = rem > 0;
                cur_tile = it;
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
        const int cnt    = min(rem, chunk);
        const bool first = kind == CH_TILE && fresh;
        int first_word = 0;  // first chunk of a tile: 1 | parity of the window barrier << 1 (the consumers keep no state for it)
        if (first) {
            buf_cur    = tile_seq & 1;
            first_word = 1 | (((tile_seq >> 1) & 1) << 1);
            ++tile_seq;
            fresh = false;
        }
        if (lane == 0) {
            ChunkDesc d;
            d.kind = kind; d.cnt = cnt; d.hx = hx; d.hy = hy; d.hz = hz;
            d.epi  = buf_cur;
            d.pad[0] = (kind == CH_TILE && rem == cnt) ? 1 : 0;  // last chunk of its tile
            d.pad[1] = first_word;                               // first chunk: the consumers wait for the tile's E window
            s.desc[st] = d;
            const uint32_t bytes = (uint32_t)(((cnt + 1) & ~1) * 8);
            mbar_expect_tx(&s.full[st], 6 * bytes);
#pragma unroll
            for (int a = 0; a < 6; ++a) bulk_g2s(&s.st[st].dat[a][0], A.in[a] + pbeg, bytes, &s.full[st]);
            mbar_arrive(&s.full[st]);
        }
        if (first) {
            // not staged ahead (first tile of the CTA, a one-chunk predecessor, an empty tile in between): stage it now.
            // Safe to overwrite: the window of two tiles ago is dead once the stage of chunk seq - 2 was released.
            if (staged_for != cur_tile) stage_window(cur_tile, buf_cur);
        } else if (kind == CH_TILE && b_it < A.ntiles && b_rem > 0 && staged_for != b_it) {
            // second or later chunk of a tile: the previous tile's window (buffer tile_seq & 1) is dead -- its last chunk
            // was released before this chunk's stage became free -- so the NEXT tile's window can be staged now
            stage_window(b_it, tile_seq & 1);
            staged_for = b_it;
        }
        rem -= cnt;
        pbeg += cnt;
    }
}

// unsorted particles (overflow of the previous step, migration arrivals): global gather, direct placement (out of line)
template <int NT, int K>
__device__ IPPLB_SLOW_ATTR void tail_chunk(const StepArgs& A, const double (*dat)[NT * K], const int cnt, const int t) {
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const int slot = k * NT + t;
        if (slot < cnt) {
            double r[3] = {dat[0][slot], dat[1][slot], dat[2][slot]};
            double p[3] = {dat[3][slot], dat[4][slot], dat[5][slot]};
            Cic c;
            cic_setup(A.m, r[0], r[1], r[2], c);
            double E[3];
            if (IPPLB_WRAP(A)) gather_point_at<3>(c.whi, [&](int n) { return cic_node_wrapped(A.m, c.a, n); }, A.ef, E);
            else gather_point<3>(A.m, c, A.ef, E);
            push_particle(A.P, r, p, E);
            Cic cn;
            cic_setup(A.m, r[0], r[1], r[2], cn);
            const int cc[3] = {cn.a[0] - A.m.nghost, cn.a[1] - A.m.nghost, cn.a[2] - A.m.nghost};
            if (owned_by_me(A, r, cc)) place_direct(A, r, p, cc, cn.whi);
            else place_exit(A, r[0], r[1], r[2], p[0], p[1], p[2]);
        }
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
    unsigned long long full[2], empty[2], eready[2];
    int hist[WIN_CELLS];
    int lpre[WIN_CELLS + 1];    // exclusive prefix inside the 64-cell scan segment
    int prefix[WIN_CELLS + 1];  // global particle prefix per window id (valid after barrier D)
    unsigned short cellxyz[WIN_CELLS];
    unsigned short winid[WIN_CELLS];
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

// Kernel variants (template parameter VAR, bit mask; the default is chosen in ipplb_bins_step):
//   V_OWNER_TAIL  the 64 cells of the home tile, which receive ~70 % of a chunk's particles and make their owners' P5
//                 the longest, are owned by the LAST two consumer warps -- the warps whose second P1 / P4 round is empty
//                 when the producer cuts a tile into equal chunks (5 x 820 of 960 slots at 64 particles per cell);
//   V_HINT        mbarrier waits carry a suspend-time hint: a waiting warp sleeps in hardware instead of re-issuing the
//                 test every ~40 ns (the producer's spin was 4 % of all issued instructions);
//   V_LEAPFROG    the steady-state leapfrog push with every sub-step on (kick, kick, drift, periodic BC): no flag tests;
//   V_PERIODIC1   one rank that owns the whole periodic domain: no ownership test on the fast path (a wrapped particle
//                 is always inside; the test is kept on the out-of-window path) and the periodic aliasing of ghost
//                 nodes (StepArgs::wrap) compiled in unconditionally;
enum { V_OWNER_TAIL = 1, V_HINT = 2, V_LEAPFROG = 4, V_P5U2 = 8, V_PERIODIC1 = 16 };
constexpr int HOME_ID0 = 224;  // first tile-major window id of the home tile (slot 13): sum of the sizes of slots 0..12

template <int NT, int K, int MINB, int VAR>
__global__ void __launch_bounds__(NT + 32, MINB) fused_step3_kernel(const StepArgs A) {
    using S            = Step3Smem<NT, K>;
    constexpr int NW   = NT / 32;
    constexpr int NSEG = WIN_CELLS / 64;
    constexpr int NSL  = (WIN_CELLS + NT - 1) / NT;  // window cells owned per thread
    constexpr int WCOLS = NSL * 16 + K * 16;  // tensor-memory columns per warp: accumulators + parked particles
    constexpr int TCOLS = tmem_cols_pow2(((NW + 3) / 4) * WCOLS);
    constexpr bool OWNER_TAIL = (VAR & V_OWNER_TAIL) != 0;
    constexpr bool HINT       = (VAR & V_HINT) != 0;
    constexpr bool P5U2       = (VAR & V_P5U2) != 0;
    constexpr bool LEAPFROG   = (VAR & V_LEAPFROG) != 0;
    constexpr bool PERIODIC1  = (VAR & V_PERIODIC1) != 0;
    static_assert(NT >= WIN_CELLS / 2, "the scan uses WIN_CELLS / 2 threads");
    static_assert(NSEG <= 16, "segment offsets live in one warp");
    static_assert(TCOLS * MINB <= 512, "tensor memory columns of the co-resident CTAs");
    static_assert(!OWNER_TAIL || (NSL == 2 && NT >= HOME_ID0 + 64 + 32 && 2 * NT - WIN_CELLS >= 96),
                  "owner remap: ids NT.. go to the third-last warp");
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
            mbar_init(&s.eready[i], 1);
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
        s.winid[c]    = (unsigned short)(id | (ts << 9));  // window id (9 bits) | destination tile slot
        s.hist[c]     = 0;
    }
    __syncthreads();

    if (warp == NW) {
        producer_loop<S, HINT, PERIODIC1>(A, s, lane);
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
    // window cell this thread owns in slot sl of P5 (< 0: none; recomputed where needed instead of held in registers).
    // Plain map: id = sl * NT + t.  V_OWNER_TAIL: the ids are rotated so that the home tile's 64 cells land on the last
    // two warps, and the WIN_CELLS - NT ids of the second slot (far corner / edge cells, almost always empty) on the
    // warp before them.
    auto own = [&](int sl) -> int {
        if (OWNER_TAIL) {
            if (sl == 0) {
                const int v = t + (64 + HOME_ID0);
                return v >= NT ? v - NT : v;
            }
            return (t >= NT - 96 && t < NT - 96 + (WIN_CELLS - NT)) ? NT + (t - (NT - 96)) : -1;
        }
        return sl * NT + t < WIN_CELLS ? sl * NT + t : -1;
    };

    for (unsigned seq = 0;; ++seq) {
        const int st = seq & 1;
        if (HINT) mbar_wait_hint(&s.full[st], (seq >> 1) & 1, 20000);
        else mbar_wait(&s.full[st], (seq >> 1) & 1);
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

        if (D.pad[1]) {  // first chunk of a tile: its E window (staged one tile ahead as a rule) has to be there
            mbar_wait(&s.eready[D.epi], (unsigned)D.pad[1] >> 1);
        }
        const int wox = D.hx * TILE - WH, woy = D.hy * TILE - WH, woz = D.hz * TILE - WH;
        // ---- P1: gather, push, bin by new cell; the pushed particle stays in registers ----------------------
        int lr[K];  // window id (9 bits) | destination tile slot << 9 | rank inside the cell << 16, or -1
        auto bin_particle = [&](int wx, int wy, int wz) {
            const int idts = s.winid[(wz * WIN + wy) * WIN + wx];
            return idts | (atomicAdd(&s.hist[idts & 0x1FF], 1) << 16);
        };
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
                    if (LEAPFROG) push_leapfrog_full(A.P, r, p, E);
                    else push_particle(A.P, r, p, E);
                    Cic cn;
                    cic_setup(A.m, r[0], r[1], r[2], cn);
                    const int cc[3] = {cn.a[0] - A.m.nghost, cn.a[1] - A.m.nghost, cn.a[2] - A.m.nghost};
                    const int wx = cc[0] - wox, wy = cc[1] - woy, wz = cc[2] - woz;
                    const bool inwin = (unsigned)wx < (unsigned)WIN && (unsigned)wy < (unsigned)WIN &&
                                       (unsigned)wz < (unsigned)WIN;
                    if (PERIODIC1) {
                        // every window cell a wrapped particle can reach lies inside the periodic domain
                        if (inwin) {
                            lr[k] = bin_particle(wx, wy, wz);
                        } else if (owned_by_me(A, r, cc)) {
                            place_direct(A, r, p, cc, cn.whi);
                        } else {
                            place_exit(A, r[0], r[1], r[2], p[0], p[1], p[2]);
                        }
                    } else if (!owned_by_me(A, r, cc)) {
                        place_exit(A, r[0], r[1], r[2], p[0], p[1], p[2]);
                    } else if (inwin) {
                        lr[k] = bin_particle(wx, wy, wz);
                    } else {
                        place_direct(A, r, p, cc, cn.whi);
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
            const int id  = lr[k] < 0 ? 0 : (lr[k] & 0x1FF);
            const int pre = gpre(id);
            double pk[6];
            tmem_ld6(tmp + k * 16, pk);
            if (lr[k] >= 0) {
                const int pos    = pre + (lr[k] >> 16);
                rec[3 * pos]     = make_double2(pk[0], pk[1]);
                rec[3 * pos + 1] = make_double2(pk[2], pk[3]);
                rec[3 * pos + 2] = make_double2(pk[4], pk[5]);
                s.tsp[pos]       = (unsigned char)((lr[k] >> 9) & 31);
            }
        }
        int t_now = t;  // opaque copy: keeps the compiler from hoisting this test out of the chunk loop into a spilled predicate
        asm volatile("" : "+r"(t_now));
        if (t_now < NSLOT) {
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
        // ---- P5: every window cell has a fixed owner thread; the owner adds the chunk's moments of its cell to the
        //      cell's accumulators in tensor memory
#pragma unroll
        for (int sl = 0; sl < NSL; ++sl) {
            const int id = own(sl);
            if (__any_sync(0xffffffffu, id >= 0)) {
                const int b = id >= 0 ? s.prefix[id] : 0, e = id >= 0 ? s.prefix[id + 1] : 0;
                if (__any_sync(0xffffffffu, e > b)) {
                    double m[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // 1 w0 w1 w2 w0w1 w0w2 w1w2 w0w1w2
                    auto add_weights = [&](int p) {
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
                    };
                    if (P5U2) {
#pragma unroll 2
                        for (int p = b; p < e; ++p) add_weights(p);
                    } else {
                        for (int p = b; p < e; ++p) add_weights(p);
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
                if (__any_sync(0xffffffffu, own(sl) >= 0)) {
                    double a[8];
                    tmem_ld8(tm0 + sl * 16, a);
                    tmem_st8(tm0 + sl * 16, z8);
                    if (own(sl) >= 0 && a[0] != 0.0) {
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
                        const unsigned xyz = s.cellxyz[own(sl)];
                        // ghosted index of the cell's upper node = window coordinate + window origin + nghost
                        const long gx = (long)(xyz & 15) + wox + A.m.nghost, gy = (long)((xyz >> 4) & 15) + woy + A.m.nghost,
                                   gz = (long)(xyz >> 8) + woz + A.m.nghost;
                        // upper / lower node per axis (aliased to the opposite interior layer on a whole periodic domain)
                        long xs[2] = {gx, gx - 1}, ys[2] = {gy, gy - 1}, zs[2] = {gz, gz - 1};
                        if (IPPLB_WRAP_CT(PERIODIC1) || IPPLB_WRAP(A)) {
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                xs[i] = wrap_axis((int)xs[i], A.m.nl[0], A.m.nghost);
                                ys[i] = wrap_axis((int)ys[i], A.m.nl[1], A.m.nghost);
                                zs[i] = wrap_axis((int)zs[i], A.m.nl[2], A.m.nghost);
                            }
                        }
#pragma unroll
                        for (int n = 0; n < 8; ++n)
                            atomicAdd(&A.rho[xs[n & 1] + (long)A.m.ex * (ys[(n >> 1) & 1] + (long)A.m.ey * zs[(n >> 2) & 1])],
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

template <int NT, int K, int MINB, int VAR>
static int launch_fused3(ipplb_ctx* ctx, const StepArgs& A) {
    using S   = Step3Smem<NT, K>;
    auto kern = fused_step3_kernel<NT, K, MINB, VAR>;
    // the attribute is per device: one flag per device ordinal (a process may hold contexts on several GPUs)
    static bool attr_set[64] = {};
    const int dev = ctx->device & 63;
    if (!attr_set[dev]) {
        IPPLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S)));
        attr_set[dev] = true;
    }
    kern<<<ctx->num_sms * MINB, NT + 32, sizeof(S), ctx->stream>>>(A);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

double* const* mig_peer_table(ipplb_ctx* ctx, long* seg_cap);  // comm.cu

// tuning knob (experiments only): IPPLB_FUSED_VAR selects the kernel variant mask (V_*); the default is the measured best
static int fused_var() {
    int var = -1;
    {
        const char* e = getenv("IPPLB_FUSED_VAR");  // read on every call so that one process can A/B the variants
        var           = e ? (atoi(e) & 31) : (V_OWNER_TAIL | V_HINT | V_P5U2 | V_LEAPFROG | V_PERIODIC1);
    }
    return var;
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
    A.peer_seg   = nullptr;
    b->exit_p2p  = 0;
    if (!exit_buf && nrk > 1 && ctx->mig) {
        // peer mode (ipplb_migrate_connect): leavers go straight into the destination ranks' inboxes
        long seg = 0;
        A.peer_seg = mig_peer_table(ctx, &seg);
        IPPLB_REQUIRE(seg < (1L << 31), "bins_step: inbox segment too large");
        A.seg_cap   = (int)seg;
        b->exit_p2p = 1;
    }
    A.regions    = ctx->d_regions;
    A.nranks     = nrk;
    A.me         = ctx->rank;
    b->exit_ranks = nrk;
    A.capacity   = (int)b->capacity;
    A.ntx = b->ntx; A.nty = b->nty; A.ntz = b->ntz; A.ntiles = b->ntiles;
    A.check_owner = (region_min && region_max) ? 1 : 0;
    A.balance_chunks = 1;  // equal chunks per tile (4.25 -> 4.20 ms at C2)
    for (int d = 0; d < 3; ++d) {
        A.rmin[d] = region_min ? region_min[d] : 0.0;
        A.rmax[d] = region_max ? region_max[d] : 0.0;
    }
    // the step's scratch words and exit counters were re-armed by the last bins_plan (build or step)
    // One rank that owns the whole periodic domain: ghost nodes are aliased to the opposite interior layer inside the
    // kernel (E window reads, rho adds) -- what HaloCells::applyPeriodicSerialDim does in two extra passes over the
    // fields (src/Field/HaloCells.hpp:297-336).  efield's ghost layers are then not read and rho's receive nothing.
    const ipplb_mesh& M = b->mesh;
    const bool whole = M.nl[0] == M.ng[0] && M.nl[1] == M.ng[1] && M.nl[2] == M.ng[2];
    A.wrap = (nrk == 1 && !A.check_owner && whole && push->do_bc) ? 1 : 0;
    // the variant follows the call: V_PERIODIC1 when the aliasing applies, V_LEAPFROG for the steady-state leapfrog step
    const bool lf = push->kind == IPPLB_PUSH_LEAPFROG && push->do_kick2 && push->do_kick1 && push->do_drift && push->do_bc;
    int var = fused_var();
    if (!lf) var &= ~V_LEAPFROG;
    if (!A.wrap) var &= ~V_PERIODIC1;
    int rc;
    // instantiated: the tuned set (V_OWNER_TAIL | V_HINT | V_P5U2) with every combination of the two call-dependent
    // bits, and the plain kernel (0) as the A/B baseline; any other request runs the plain kernel
    constexpr int T = V_OWNER_TAIL | V_HINT | V_P5U2;
    const bool timed = b->timing && b->ev && b->ev_n < ipplb_bins::NEV;
    if (timed) IPPLB_CUDA(cudaEventRecord(b->ev[2 * b->ev_n], ctx->stream));
    switch (var) {
#define IPPLB_VAR_CASE(v) case v: rc = launch_fused3<480, 2, 2, v>(ctx, A); break;
        IPPLB_VAR_CASE(T)
        IPPLB_VAR_CASE(T | V_LEAPFROG)
        IPPLB_VAR_CASE(T | V_PERIODIC1)
        IPPLB_VAR_CASE(T | V_LEAPFROG | V_PERIODIC1)
        IPPLB_VAR_CASE(V_LEAPFROG | V_PERIODIC1)
#undef IPPLB_VAR_CASE
        default: rc = launch_fused3<480, 2, 2, 0>(ctx, A); break;
    }
    if (timed) {
        IPPLB_CUDA(cudaEventRecord(b->ev[2 * b->ev_n + 1], ctx->stream));
        ++b->ev_n;
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
