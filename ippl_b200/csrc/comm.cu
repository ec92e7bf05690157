// comm.cu -- multi-rank part: NCCL bootstrap, batched halo exchange, particle migration.
//
// Every collective is written as LOCAL PHASES (kernels of one rank) around a TRANSPORT step.  Two transports:
//   * NCCL over NVLink / NVSwitch, one process per GPU (the product path);
//   * "loop": all ranks of a small job held by ONE process on ONE device (ipplb_loop_*), device-to-device copies as the
//     transport.  It runs exactly the same kernels, tables and host logic, which makes the multi-rank path testable on a
//     single GPU (ownership, counts, particles, halo strips against the oracle's all-ranks simulation).
// Migration of the bucketed store goes over PEER MEMORY: the fused step writes every leaver straight into its destination
// rank's inbox (cudaIpc-mapped over NVLink; the other contexts' buffers in a loop), one small all-gather of the counts
// doubles as the barrier, and one kernel drops the arrivals into their buckets.  No host synchronisation.
//
// One process per GPU.  Halo traffic: ALL neighbour components of a rank are packed by one kernel
// into one send buffer (grouped by peer), exchanged with ONE grouped ncclSend/ncclRecv per peer
// (NVLink/NVSwitch), and unpacked by one kernel (= for fill, atomic += for accumulate).  The
// reference does 26 pack kernels + 26 MPI messages + 26 unpack kernels with a fence after each
// (src/Field/HaloCells.hpp:109-242).  Migration: ownership by the reference's fp region test
// (bit-exact), counts by one all-gather, SoA segments per peer, arrivals dropped into the holes.
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "bins.h"
#include "cic.cuh"
#include "layout.h"
#include "transport.h"

namespace ipplb {

#define IPPLB_NCCL(call)                                                                  \
    do {                                                                                  \
        ncclResult_t r__ = (call);                                                        \
        if (r__ != ncclSuccess) {                                                         \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
            return IPPLB_ERR_NCCL;                                                        \
        }                                                                                 \
    } while (0)

struct Region {
    int lo[3], n[3];  // local ghosted start + extents
    long off;         // cell offset inside the packed buffer (multiply by ncomp)
};

struct PeerSeg {
    int peer;
    long send_off, send_cells;  // fill direction: what I send (my interior strips)
    long recv_off, recv_cells;  // fill direction: what I receive (my ghost strips)
};

struct CommPlan {
    Layout L;
    double origin[3], h[3];
    ipplb_mesh mesh;  // this rank's mesh
    // fill direction regions; accumulate swaps the roles (HaloCells.hpp:157-183)
    std::vector<Region> send_regions, recv_regions;  // ordered by peer, then message tag
    std::vector<PeerSeg> peers;
    long send_cells = 0, recv_cells = 0;
    Region *d_send = nullptr, *d_recv = nullptr;
    long *d_send_prefix = nullptr, *d_recv_prefix = nullptr;  // cells prefix per region (+1)
    int serial_mask = 0;  // dims whose local extent equals the global one
    double* d_regions = nullptr;  // [nranks][6] physical regions
    // migration scratch
    int* d_dest = nullptr; long dest_cap = 0;
    int* d_counts = nullptr;       // [nranks + UPD_EXTRA] row of the count exchange, then cursor [nranks], offsets, counters
    int* d_matrix = nullptr;       // [nranks * (nranks + UPD_EXTRA)] every rank's row
    int *h_counts = nullptr, *h_matrix = nullptr;  // pinned
    // ipplb_update_plan -> ipplb_update_commit
    bool planned = false; long plan_n = 0, plan_nh = 0, plan_na = 0;
    const int* loop_row = nullptr;  // in-process rank group: this rank's row of the pending count exchange
};

// words appended to a rank's row of send counts in the count exchange, so that every rank can evaluate every rank's
// fit / error condition and all of them abort TOGETHER (a rank-local abort between the count exchange and the grouped
// send / receive would leave its peers hanging in the receive)
enum { UPD_N = 0, UPD_CAP_LO = 1, UPD_CAP_HI = 2, UPD_FLAGS = 3, UPD_EXTRA = 4 };

// peer-memory migration of the bucketed store (ipplb_migrate_connect)
struct MigBox {
    long seg_cap = 0;                            // records per (source, destination) segment
    double* inbox[2] = {nullptr, nullptr};       // [nranks][seg_cap][6], two step parities
    double** d_peer[2] = {nullptr, nullptr};     // device tables [nranks]: rank r's inbox of that parity as mapped HERE
    std::vector<void*> opened;                   // cudaIpc mappings to close
    int parity = 0;
};

// ---- halo pack / unpack ---------------------------------------------------------------------------
__device__ __forceinline__ int find_region(const long* __restrict__ prefix, int nreg, long t) {
    int lo = 0, hi = nreg;  // prefix[lo] <= t < prefix[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (prefix[mid] <= t) lo = mid; else hi = mid;
    }
    return lo;
}

// mode 0: buf[...] = field[region]            (pack)
// mode 1: field[region] = buf[...]            (unpack, fill)
// mode 2: atomicAdd(field[region], buf[...])  (unpack, accumulate: strips of different components
//                                              overlap on edge/corner cells)
__global__ void __launch_bounds__(256)
halo_copy_kernel(const Region* __restrict__ regs, const long* __restrict__ prefix, int nreg,
                 long total_cells, int ncomp, int e0, int e1, double* __restrict__ field,
                 double* __restrict__ buf, int mode) {
    const long nt = total_cells * ncomp;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nt;
         t += (long)gridDim.x * blockDim.x) {
        const long cell = t / ncomp;
        const int c     = (int)(t - cell * ncomp);
        const int r     = find_region(prefix, nreg, cell);
        const Region R  = regs[r];
        long l          = cell - prefix[r];
        const int i = (int)(l % R.n[0]), j = (int)((l / R.n[0]) % R.n[1]),
                  k = (int)(l / ((long)R.n[0] * R.n[1]));
        const long f = ((long)(R.lo[0] + i) + (long)e0 * ((R.lo[1] + j) + (long)e1 * (R.lo[2] + k))) *
                           ncomp + c;
        if (mode == 0) buf[t] = field[f];
        else if (mode == 1) field[f] = buf[t];
        else atomicAdd(&field[f], buf[t]);
    }
}

// ---- ownership -------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_region_strict(const double* __restrict__ R, double x, double y,
                                                 double z) {
    return x > R[0] && y > R[1] && z > R[2] && x <= R[3] && y <= R[4] && z <= R[5];
}
__device__ __forceinline__ bool in_region_incl(const double* __restrict__ R, double x, double y,
                                               double z) {
    return x >= R[0] && y >= R[1] && z >= R[2] && x <= R[3] && y <= R[4] && z <= R[5];
}

// destRankOf of ParticleSpatialLayout.hpp:372-395: own region, then every rank ascending (strict
// regions are disjoint, so the neighbour-first search order of the reference finds the same rank),
// then the inclusive fallback ascending, else stay.
__global__ void __launch_bounds__(256)
locate_kernel(const double* __restrict__ regions, int nranks, int me, long n,
              const double* __restrict__ x, const double* __restrict__ y,
              const double* __restrict__ z, int* __restrict__ dest, int* __restrict__ counts,
              int count_self = 0) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long)gridDim.x * blockDim.x) {
        const double px = x[i], py = y[i], pz = z[i];
        int d = -1;
        if (in_region_strict(regions + 6 * me, px, py, pz)) d = me;
        for (int r = 0; d < 0 && r < nranks; ++r)
            if (in_region_strict(regions + 6 * r, px, py, pz)) d = r;
        for (int r = 0; d < 0 && r < nranks; ++r)
            if (in_region_incl(regions + 6 * r, px, py, pz)) d = r;
        if (d < 0) d = me;
        dest[i] = d;
        if (d != me || count_self) atomicAdd(&counts[d], 1);
    }
}

constexpr int NATTR = 7;  // x y z px py pz q (q segment present only when q != NULL)

struct AttrPtrs {
    double* a[NATTR];
    int n;
};

// leavers -> send buffer.  Peer block p: nattr arrays of cnt[p] doubles at offset off[p]*nattr.
// The claimed slot also records the hole index.
__global__ void __launch_bounds__(256)
pack_leavers_kernel(long n, int me, const int* __restrict__ dest, const int* __restrict__ send_off,
                    const int* __restrict__ send_cnt, int* __restrict__ cursor, AttrPtrs A,
                    double* __restrict__ sendbuf, int* __restrict__ holes, int pack_self = 0) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long)gridDim.x * blockDim.x) {
        const int d = dest[i];
        if (d == me && !pack_self) continue;
        const int k   = atomicAdd(&cursor[d], 1);
        const long b  = (long)send_off[d] * A.n;
        const int cnt = send_cnt[d];
        for (int a = 0; a < A.n; ++a) sendbuf[b + (long)a * cnt + k] = A.a[a][i];
        holes[send_off[d] + k] = (int)i;
    }
}

// arrival j (global index over all sources) -> holes[j] if j < nholes else n_old + (j - nholes)
__global__ void __launch_bounds__(256)
unpack_arrivals_kernel(int nsrc, const int* __restrict__ recv_off, const int* __restrict__ recv_cnt,
                       long total, const double* __restrict__ recvbuf, AttrPtrs A,
                       const int* __restrict__ holes, int nholes, long n_old) {
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < total;
         j += (long)gridDim.x * blockDim.x) {
        int s = 0;
        while (s + 1 < nsrc && recv_off[s + 1] <= j) ++s;
        const long k   = j - recv_off[s];
        const long b   = (long)recv_off[s] * A.n;
        const int cnt  = recv_cnt[s];
        const long pos = j < nholes ? (long)holes[j] : n_old + (j - nholes);
        for (int a = 0; a < A.n; ++a) A.a[a][pos] = recvbuf[b + (long)a * cnt + k];
    }
}

// Fewer arrivals than holes: holes[na..nh) remain.  New count n' = n_old - (nh - na).  Remaining holes
// below n' are filled with the survivors at or above n'.
__global__ void mark_tail_holes_kernel(const int* __restrict__ holes, int first, int nh, long nprime,
                                       int* __restrict__ tail_flag, int* __restrict__ low_holes,
                                       int* __restrict__ counters) {
    for (int j = first + blockIdx.x * blockDim.x + threadIdx.x; j < nh; j += gridDim.x * blockDim.x) {
        const int h = holes[j];
        if (h >= nprime) tail_flag[h - nprime] = 1;
        else low_holes[atomicAdd(&counters[0], 1)] = h;
    }
}
__global__ void collect_tail_survivors_kernel(long nprime, int tail, const int* __restrict__ tail_flag,
                                              int* __restrict__ movers, int* __restrict__ counters) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < tail; t += gridDim.x * blockDim.x)
        if (!tail_flag[t]) movers[atomicAdd(&counters[1], 1)] = (int)(nprime + t);
}
__global__ void fill_low_holes_kernel(const int* __restrict__ low_holes, const int* __restrict__ movers,
                                      const int* __restrict__ counters, AttrPtrs A) {
    const int n = counters[0];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        for (int a = 0; a < A.n; ++a) A.a[a][low_holes[t]] = A.a[a][movers[t]];
}

static int grid1d(long n) {
    long g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

static void free_mig(MigBox* M) {
    if (!M) return;
    for (void* p : M->opened) cudaIpcCloseMemHandle(p);
    for (int i = 0; i < 2; ++i) { cudaFree(M->inbox[i]); cudaFree(M->d_peer[i]); }
    delete M;
}

static void free_plan(CommPlan* P) {
    if (!P) return;
    cudaFree(P->d_send); cudaFree(P->d_recv); cudaFree(P->d_send_prefix); cudaFree(P->d_recv_prefix);
    cudaFree(P->d_regions); cudaFree(P->d_dest); cudaFree(P->d_counts); cudaFree(P->d_matrix);
    cudaFreeHost(P->h_counts); cudaFreeHost(P->h_matrix);
    delete P;
}

// Build the per-peer message lists.  Message identity = (sender's component id).  A receiver's entry
// with component c' expects the sender's component matching(c') (FieldLayout::getMatchingIndex).  Both
// sides order the messages of one peer pair by that tag, so the concatenated buffers line up.
static void build_regions(const std::vector<NeighborEntry>& nb, int nranks, CommPlan* P) {
    struct Msg { int peer, tag, seq; Region r; };
    std::vector<Msg> s, r;
    int seq = 0;
    for (auto& e : nb) {
        Msg ms{e.peer, e.comp, seq, {}}, mr{e.peer, matching_component(e.comp), seq, {}};
        for (int d = 0; d < 3; ++d) {
            ms.r.lo[d] = e.send_lo[d]; ms.r.n[d] = e.send_hi[d] - e.send_lo[d];
            mr.r.lo[d] = e.recv_lo[d]; mr.r.n[d] = e.recv_hi[d] - e.recv_lo[d];
        }
        s.push_back(ms); r.push_back(mr);
        ++seq;
    }
    auto order = [](const Msg& a, const Msg& b) {
        if (a.peer != b.peer) return a.peer < b.peer;
        if (a.tag != b.tag) return a.tag < b.tag;
        return a.seq < b.seq;
    };
    std::sort(s.begin(), s.end(), order);
    std::sort(r.begin(), r.end(), order);
    P->send_regions.clear(); P->recv_regions.clear(); P->peers.clear();
    long so = 0, ro = 0;
    for (auto& m : s) { m.r.off = so; so += (long)m.r.n[0] * m.r.n[1] * m.r.n[2]; P->send_regions.push_back(m.r); }
    for (auto& m : r) { m.r.off = ro; ro += (long)m.r.n[0] * m.r.n[1] * m.r.n[2]; P->recv_regions.push_back(m.r); }
    P->send_cells = so; P->recv_cells = ro;
    for (int p = 0; p < nranks; ++p) {
        PeerSeg seg{p, -1, 0, -1, 0};
        for (size_t i = 0; i < s.size(); ++i)
            if (s[i].peer == p) {
                if (seg.send_off < 0) seg.send_off = P->send_regions[i].off;
                seg.send_cells += (long)s[i].r.n[0] * s[i].r.n[1] * s[i].r.n[2];
            }
        for (size_t i = 0; i < r.size(); ++i)
            if (r[i].peer == p) {
                if (seg.recv_off < 0) seg.recv_off = P->recv_regions[i].off;
                seg.recv_cells += (long)r[i].r.n[0] * r[i].r.n[1] * r[i].r.n[2];
            }
        if (seg.send_cells || seg.recv_cells) P->peers.push_back(seg);
    }
}

}  // namespace ipplb



namespace ipplb {

// ---- transport ----------------------------------------------------------------------------------------------------
int nccl_exchange(ipplb_ctx* ctx, const std::vector<Xfer>& x) {
    IPPLB_NCCL(ncclGroupStart());
    for (auto& t : x) {
        if (t.sbytes) IPPLB_NCCL(ncclSend(t.sptr, t.sbytes, ncclChar, t.peer, (ncclComm_t)ctx->nccl, ctx->stream));
        if (t.rbytes) IPPLB_NCCL(ncclRecv(t.rptr, t.rbytes, ncclChar, t.peer, (ncclComm_t)ctx->nccl, ctx->stream));
    }
    IPPLB_NCCL(ncclGroupEnd());
    ctx->launches++;
    return IPPLB_OK;
}

// all ranks in one process: rank a's send to b is matched with b's receive from a
int loop_exchange(ipplb_loop* L, const std::vector<std::vector<Xfer>>& all) {
    const int nr = (int)L->ctx.size();
    for (int a = 0; a < nr; ++a) IPPLB_CUDA(cudaStreamSynchronize(L->ctx[a]->stream));
    for (int a = 0; a < nr; ++a)
        for (auto& s : all[a]) {
            if (!s.sbytes) continue;
            const Xfer* r = nullptr;
            for (auto& c : all[s.peer])
                if (c.peer == a) r = &c;
            if (!r || r->rbytes != s.sbytes) {
                set_error("loop transport: rank %d sends %zu bytes to %d, which expects %zu", a, s.sbytes, s.peer, r ? r->rbytes : 0);
                return IPPLB_ERR_ARG;
            }
            IPPLB_CUDA(cudaMemcpyAsync(r->rptr, s.sptr, s.sbytes, cudaMemcpyDeviceToDevice, L->ctx[s.peer]->stream));
        }
    for (int a = 0; a < nr; ++a) IPPLB_CUDA(cudaStreamSynchronize(L->ctx[a]->stream));
    return IPPLB_OK;
}

// every rank's `words` ints -> every rank's matrix [nranks][words]
static int gather_rows(ipplb_ctx* ctx, const int* row, int* matrix, int words) {
    if (ctx->loop) {
        ipplb_loop* L = ctx->loop;
        // called once per rank by the loop drivers AFTER every rank has produced its row (they synchronise in between)
        for (size_t r = 0; r < L->ctx.size(); ++r) {
            CommPlan* Q = (CommPlan*)L->ctx[r]->plan;
            const int* src = L->ctx[r] == ctx ? row : (const int*)Q->loop_row;
            IPPLB_CUDA(cudaMemcpyAsync(matrix + r * words, src, sizeof(int) * words, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        return IPPLB_OK;
    }
    IPPLB_NCCL(ncclAllGather(row, matrix, words, ncclInt, (ncclComm_t)ctx->nccl, ctx->stream));
    ctx->launches++;
    return IPPLB_OK;
}

// ---- halo exchange: local phases ---------------------------------------------------------------------------------------
static int halo_pack(ipplb_ctx* ctx, double* field, int ncomp, int mode) {
    CommPlan* P = (CommPlan*)ctx->plan;
    const int e0 = P->mesh.nl[0] + 2 * P->mesh.nghost, e1 = P->mesh.nl[1] + 2 * P->mesh.nghost;
    // fill: pack my `send` strips, receive into my `recv` strips.  accumulate: roles swap.
    const bool fill      = mode == 0;
    const long out_cells = fill ? P->send_cells : P->recv_cells;
    const long in_cells  = fill ? P->recv_cells : P->send_cells;
    int rc;
    if ((rc = ensure(ctx, ctx->send, sizeof(double) * (size_t)(out_cells * ncomp + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->recv, sizeof(double) * (size_t)(in_cells * ncomp + 1)))) return rc;
    if (out_cells) {
        halo_copy_kernel<<<grid1d(out_cells * ncomp), 256, 0, ctx->stream>>>(
            fill ? P->d_send : P->d_recv, fill ? P->d_send_prefix : P->d_recv_prefix,
            (int)(fill ? P->send_regions.size() : P->recv_regions.size()), out_cells, ncomp, e0, e1, field,
            (double*)ctx->send.ptr, 0);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    return IPPLB_OK;
}

static void halo_xfers(ipplb_ctx* ctx, int ncomp, int mode, std::vector<Xfer>& x) {
    CommPlan* P = (CommPlan*)ctx->plan;
    const bool fill = mode == 0;
    double* sb = (double*)ctx->send.ptr;
    double* rb = (double*)ctx->recv.ptr;
    x.clear();
    for (auto& seg : P->peers) {
        const long so = fill ? seg.send_off : seg.recv_off, sc = fill ? seg.send_cells : seg.recv_cells;
        const long ro = fill ? seg.recv_off : seg.send_off, rc = fill ? seg.recv_cells : seg.send_cells;
        x.push_back({seg.peer, sc ? sb + so * ncomp : nullptr, sizeof(double) * (size_t)(sc * ncomp), rc ? rb + ro * ncomp : nullptr,
                     sizeof(double) * (size_t)(rc * ncomp)});
    }
}

static int halo_unpack(ipplb_ctx* ctx, double* field, int ncomp, int mode, bool exchanged) {
    CommPlan* P = (CommPlan*)ctx->plan;
    const int e0 = P->mesh.nl[0] + 2 * P->mesh.nghost, e1 = P->mesh.nl[1] + 2 * P->mesh.nghost;
    const bool fill     = mode == 0;
    const long in_cells = fill ? P->recv_cells : P->send_cells;
    if (exchanged && in_cells) {
        halo_copy_kernel<<<grid1d(in_cells * ncomp), 256, 0, ctx->stream>>>(
            fill ? P->d_recv : P->d_send, fill ? P->d_recv_prefix : P->d_send_prefix,
            (int)(fill ? P->recv_regions.size() : P->send_regions.size()), in_cells, ncomp, e0, e1, field,
            (double*)ctx->recv.ptr, fill ? 1 : 2);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    if (P->serial_mask) {
        return mode == 0 ? ipplb_halo_fill_periodic(ctx, &P->mesh, field, ncomp, P->serial_mask)
                         : ipplb_halo_accumulate_periodic(ctx, &P->mesh, field, ncomp, P->serial_mask);
    }
    return IPPLB_OK;
}

// ---- ParticleSpatialLayout::update on contiguous arrays: local phases --------------------------------------------------
// phase 1: destination rank of every particle + this rank's row of the count exchange
static int update_locate(ipplb_ctx* ctx, ipplb_particles* p) {
    CommPlan* P  = (CommPlan*)ctx->plan;
    const int nr = ctx->nranks, me = ctx->rank;
    const long n = p->n;
    if (P->dest_cap < n + 1) {
        if (P->d_dest) { IPPLB_CUDA(cudaStreamSynchronize(ctx->stream)); IPPLB_CUDA(cudaFree(P->d_dest)); }
        P->dest_cap = n + n / 4 + 1024;
        IPPLB_CUDA(cudaMalloc(&P->d_dest, sizeof(int) * P->dest_cap));
    }
    int* cnt = P->d_counts;
    IPPLB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (5 * nr + UPD_EXTRA + 32), ctx->stream));
    if (n > 0) {
        locate_kernel<<<grid1d(n), 256, 0, ctx->stream>>>(P->d_regions, nr, me, n, p->x, p->y, p->z, P->d_dest, cnt);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    const int extra[UPD_EXTRA] = {(int)n, (int)(p->capacity & 0x7fffffffL), (int)(p->capacity >> 31), 0};
    std::memcpy(P->h_counts, extra, sizeof(extra));
    IPPLB_CUDA(cudaMemcpyAsync(cnt + nr, P->h_counts, sizeof(extra), cudaMemcpyHostToDevice, ctx->stream));
    P->loop_row = cnt;
    return IPPLB_OK;
}

// phase 2 (after the rows are in d_matrix): host-side sizes; every rank checks EVERY rank's fit, so all abort together
static int update_sizes(ipplb_ctx* ctx, ipplb_particles* p, long* sent_host, long* recv_host) {
    CommPlan* P  = (CommPlan*)ctx->plan;
    const int nr = ctx->nranks, me = ctx->rank, W = nr + UPD_EXTRA;
    IPPLB_CUDA(cudaMemcpyAsync(P->h_matrix, P->d_matrix, sizeof(int) * nr * W, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int* M = P->h_matrix;
    long nh = 0, na = 0;
    for (int r = 0; r < nr; ++r) {
        const int sc = M[me * W + r], rc = M[r * W + me];
        nh += sc; na += rc;
        if (sent_host) sent_host[r] = sc;
        if (recv_host) recv_host[r] = rc;
    }
    P->plan_nh = nh; P->plan_na = na; P->plan_n = p->n;
    for (int r = 0; r < nr; ++r) {
        long out = 0, in = 0;
        for (int t = 0; t < nr; ++t) { out += M[r * W + t]; in += M[t * W + r]; }
        const long n_r = M[r * W + nr + UPD_N], cap_r = (long)M[r * W + nr + UPD_CAP_LO] + ((long)M[r * W + nr + UPD_CAP_HI] << 31);
        if (n_r - out + in > cap_r) {
            set_error("update: rank %d would hold %ld particles after the migration, its arrays have room for %ld "
                      "(every rank returns this error; ipplb_update_plan reports the size to reserve)", r, n_r - out + in, cap_r);
            return IPPLB_ERR_CAPACITY;
        }
    }
    return IPPLB_OK;
}

struct UpdateBufs {
    AttrPtrs A;
    int *holes, *low_holes, *movers, *tail_flag, *d_roff, *d_rcnt;
    std::vector<int> h_soff, h_roff, h_rcnt, h_scnt;
};

// phase 3: pack the leavers per destination; fills the transfer list
static int update_pack(ipplb_ctx* ctx, ipplb_particles* p, UpdateBufs& B, std::vector<Xfer>& x) {
    CommPlan* P  = (CommPlan*)ctx->plan;
    const int nr = ctx->nranks, me = ctx->rank, W = nr + UPD_EXTRA;
    const long n = p->n;
    const int* M = P->h_matrix;
    B.h_soff.assign(nr + 1, 0); B.h_roff.assign(nr + 1, 0); B.h_rcnt.assign(nr, 0); B.h_scnt.assign(nr, 0);
    for (int r = 0; r < nr; ++r) {
        B.h_scnt[r] = M[me * W + r];
        B.h_rcnt[r] = M[r * W + me];
        B.h_soff[r + 1] = B.h_soff[r] + B.h_scnt[r];
        B.h_roff[r + 1] = B.h_roff[r] + B.h_rcnt[r];
    }
    const int nh = B.h_soff[nr], na = B.h_roff[nr];
    B.A.n = 0;
    double* attrs[NATTR] = {p->x, p->y, p->z, p->px, p->py, p->pz, p->q};
    for (int a = 0; a < NATTR; ++a) if (attrs[a]) B.A.a[B.A.n++] = attrs[a];
    int rc;
    if ((rc = ensure(ctx, ctx->send, sizeof(double) * ((size_t)nh * B.A.n + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->recv, sizeof(double) * ((size_t)na * B.A.n + 1)))) return rc;
    // misc: holes[nh], low_holes[nh], movers[nh], tail_flag[nh], recv offsets/counts
    if ((rc = ensure(ctx, ctx->misc, sizeof(int) * ((size_t)4 * nh + 4 * nr + 64)))) return rc;
    B.holes = (int*)ctx->misc.ptr;
    B.low_holes = B.holes + nh; B.movers = B.low_holes + nh; B.tail_flag = B.movers + nh;
    B.d_roff = B.tail_flag + nh; B.d_rcnt = B.d_roff + nr + 1;
    int* cnt = P->d_counts;
    int* cursor = cnt + nr + UPD_EXTRA;
    int* soff = cursor + nr;
    double* sb = (double*)ctx->send.ptr; double* rb = (double*)ctx->recv.ptr;
    IPPLB_CUDA(cudaMemcpyAsync(soff, B.h_soff.data(), sizeof(int) * (nr + 1), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_CUDA(cudaMemcpyAsync(B.d_roff, B.h_roff.data(), sizeof(int) * (nr + 1), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_CUDA(cudaMemcpyAsync(B.d_rcnt, B.h_rcnt.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, ctx->stream));
    if (nh > 0) {
        pack_leavers_kernel<<<grid1d(n), 256, 0, ctx->stream>>>(n, me, P->d_dest, soff, cnt, cursor, B.A, sb, B.holes);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    x.clear();
    for (int r = 0; r < nr; ++r) {
        if (r == me) continue;
        if (!B.h_scnt[r] && !B.h_rcnt[r]) continue;
        x.push_back({r, sb + (size_t)B.h_soff[r] * B.A.n, sizeof(double) * (size_t)B.h_scnt[r] * B.A.n,
                     rb + (size_t)B.h_roff[r] * B.A.n, sizeof(double) * (size_t)B.h_rcnt[r] * B.A.n});
    }
    return IPPLB_OK;
}

// phase 4: arrivals into the holes / behind the end, remaining holes filled from the end
static int update_unpack(ipplb_ctx* ctx, ipplb_particles* p, UpdateBufs& B) {
    CommPlan* P  = (CommPlan*)ctx->plan;
    const int nr = ctx->nranks;
    const long n = p->n;
    const int nh = B.h_soff[nr], na = B.h_roff[nr];
    const long n_new = n - nh + na;
    int* counters = P->d_counts + 3 * nr + UPD_EXTRA + 8;
    double* rb = (double*)ctx->recv.ptr;
    if (na > 0) {
        unpack_arrivals_kernel<<<(unsigned)((na + 255) / 256), 256, 0, ctx->stream>>>(nr, B.d_roff, B.d_rcnt, na, rb, B.A, B.holes, nh, n);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    if (nh > na) {
        const int tail = nh - na;
        IPPLB_CUDA(cudaMemsetAsync(B.tail_flag, 0, sizeof(int) * tail, ctx->stream));
        mark_tail_holes_kernel<<<grid1d(tail), 256, 0, ctx->stream>>>(B.holes, na, nh, n_new, B.tail_flag, B.low_holes, counters);
        IPPLB_CHECK_LAUNCH(ctx);
        collect_tail_survivors_kernel<<<grid1d(tail), 256, 0, ctx->stream>>>(n_new, tail, B.tail_flag, B.movers, counters);
        IPPLB_CHECK_LAUNCH(ctx);
        fill_low_holes_kernel<<<grid1d(tail), 256, 0, ctx->stream>>>(B.low_holes, B.movers, counters, B.A);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    p->n = n_new;
    return IPPLB_OK;
}

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_ctx_destroy(ipplb_ctx* ctx) {
    if (!ctx) return IPPLB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_mig((MigBox*)ctx->mig);
    if (ctx->nccl) ncclCommDestroy((ncclComm_t)ctx->nccl);
    free_plan((CommPlan*)ctx->plan);
    Scratch* all[] = {&ctx->keys, &ctx->counts, &ctx->cub_tmp, &ctx->reduce, &ctx->send, &ctx->recv, &ctx->misc};
    for (Scratch* s : all)
        if (s->ptr) cudaFree(s->ptr);
    if (ctx->reduce_host) cudaFreeHost(ctx->reduce_host);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_in) {
        cudaStreamDestroy(ctx->s_in);
        cudaStreamDestroy(ctx->s_out);
        for (int s = 0; s < 2; ++s) {
            cudaEventDestroy(ctx->ev_in[s]);
            cudaEventDestroy(ctx->ev_comp[s]);
            cudaEventDestroy(ctx->ev_out[s]);
        }
    }
    delete ctx;
    return IPPLB_OK;
}

int ipplb_nccl_unique_id(char id_out[IPPLB_NCCL_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) <= IPPLB_NCCL_ID_BYTES, "ncclUniqueId larger than expected");
    ncclUniqueId id;
    IPPLB_NCCL(ncclGetUniqueId(&id));
    std::memset(id_out, 0, IPPLB_NCCL_ID_BYTES);
    std::memcpy(id_out, &id, sizeof(id));
    return IPPLB_OK;
}

int ipplb_comm_init(ipplb_ctx* ctx, int rank, int nranks, const char id[IPPLB_NCCL_ID_BYTES]) {
    IPPLB_REQUIRE(ctx && nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad arguments");
    ctx->rank   = rank;
    ctx->nranks = nranks;
    if (nranks == 1) return IPPLB_OK;
    IPPLB_REQUIRE(id, "comm_init: id is NULL");
    IPPLB_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    IPPLB_NCCL(ncclCommInitRank(&comm, nranks, uid, rank));
    ctx->nccl = (ncclComm*)comm;
    return IPPLB_OK;
}

// ---- all ranks of a small job in one process (tests; see the head of this file) -----------------------------------------
int ipplb_loop_create(ipplb_loop** out, ipplb_ctx* const* ctxs, int nranks) {
    IPPLB_REQUIRE(out && ctxs && nranks >= 1 && nranks <= MAX_RANKS, "loop_create: bad arguments");
    ipplb_loop* L = new ipplb_loop();
    for (int r = 0; r < nranks; ++r) {
        IPPLB_REQUIRE(ctxs[r] && !ctxs[r]->nccl && !ctxs[r]->loop, "loop_create: context is NULL or already in a communicator");
        ctxs[r]->rank = r; ctxs[r]->nranks = nranks; ctxs[r]->loop = L;
        L->ctx.push_back(ctxs[r]);
    }
    *out = L;
    return IPPLB_OK;
}

int ipplb_loop_destroy(ipplb_loop* L) {
    if (!L) return IPPLB_OK;
    for (ipplb_ctx* c : L->ctx) { c->loop = nullptr; c->nranks = 1; c->rank = 0; }
    delete L;
    return IPPLB_OK;
}

int ipplb_ctx_set_layout(ipplb_ctx* ctx, const ipplb_layout* l, const double origin[3],
                         const double h[3]) {
    IPPLB_REQUIRE(ctx && l && origin && h, "set_layout: bad arguments");
    IPPLB_REQUIRE((int)l->L.boxes.size() == ctx->nranks, "set_layout: layout rank count != communicator size");
    free_plan((CommPlan*)ctx->plan);
    ctx->plan   = nullptr;
    CommPlan* P = new CommPlan();
    P->L        = l->L;
    for (int d = 0; d < 3; ++d) { P->origin[d] = origin[d]; P->h[d] = h[d]; }
    ipplb_layout_mesh(l, ctx->rank, origin, h, &P->mesh);
    const int nr = ctx->nranks;
    build_regions(P->L.neighbors(ctx->rank), nr, P);
    P->serial_mask = 0;
    if (P->L.periodic)
        for (int d = 0; d < 3; ++d)
            if (P->mesh.nl[d] == P->mesh.ng[d]) P->serial_mask |= 1 << d;
    auto upload = [&](const std::vector<Region>& v, Region** dptr, long** dprefix) -> int {
        std::vector<long> prefix(v.size() + 1, 0);
        for (size_t i = 0; i < v.size(); ++i) prefix[i + 1] = prefix[i] + (long)v[i].n[0] * v[i].n[1] * v[i].n[2];
        IPPLB_CUDA(cudaMalloc(dptr, sizeof(Region) * (v.size() + 1)));
        IPPLB_CUDA(cudaMalloc(dprefix, sizeof(long) * prefix.size()));
        if (!v.empty()) IPPLB_CUDA(cudaMemcpy(*dptr, v.data(), sizeof(Region) * v.size(), cudaMemcpyHostToDevice));
        IPPLB_CUDA(cudaMemcpy(*dprefix, prefix.data(), sizeof(long) * prefix.size(), cudaMemcpyHostToDevice));
        return IPPLB_OK;
    };
    int rc;
    if ((rc = upload(P->send_regions, &P->d_send, &P->d_send_prefix))) return rc;
    if ((rc = upload(P->recv_regions, &P->d_recv, &P->d_recv_prefix))) return rc;
    std::vector<double> regs(6 * nr);
    P->L.regions(origin, h, regs.data());
    IPPLB_CUDA(cudaMalloc(&P->d_regions, sizeof(double) * 6 * nr));
    IPPLB_CUDA(cudaMemcpy(P->d_regions, regs.data(), sizeof(double) * 6 * nr, cudaMemcpyHostToDevice));
    ctx->d_regions = P->d_regions;
    const int W = nr + UPD_EXTRA;
    IPPLB_CUDA(cudaMalloc(&P->d_counts, sizeof(int) * (5 * nr + UPD_EXTRA + 32)));
    IPPLB_CUDA(cudaMalloc(&P->d_matrix, sizeof(int) * nr * W));
    IPPLB_CUDA(cudaMallocHost(&P->h_counts, sizeof(int) * (5 * nr + UPD_EXTRA + 32)));
    IPPLB_CUDA(cudaMallocHost(&P->h_matrix, sizeof(int) * nr * W));
    ctx->plan = P;
    return IPPLB_OK;
}

int ipplb_halo_exchange(ipplb_ctx* ctx, double* field, int ncomp, int mode) {
    IPPLB_REQUIRE(ctx && field && ncomp >= 1 && (mode == 0 || mode == 1), "halo_exchange: bad arguments");
    CommPlan* P = (CommPlan*)ctx->plan;
    IPPLB_REQUIRE(P, "halo_exchange: no layout bound (ipplb_ctx_set_layout)");
    IPPLB_REQUIRE(!ctx->loop, "halo_exchange: this context belongs to an in-process rank group (use ipplb_loop_halo_exchange)");
    const bool exch = ctx->nranks > 1 && (P->send_cells || P->recv_cells);
    int rc;
    if (exch) {
        IPPLB_REQUIRE(ctx->nccl, "halo_exchange: communicator not initialised");
        if ((rc = halo_pack(ctx, field, ncomp, mode))) return rc;
        std::vector<Xfer> x;
        halo_xfers(ctx, ncomp, mode, x);
        if ((rc = nccl_exchange(ctx, x))) return rc;
    }
    return halo_unpack(ctx, field, ncomp, mode, exch);
}

int ipplb_loop_halo_exchange(ipplb_loop* L, double* const* fields, int ncomp, int mode) {
    IPPLB_REQUIRE(L && fields && ncomp >= 1 && (mode == 0 || mode == 1), "loop_halo_exchange: bad arguments");
    const int nr = (int)L->ctx.size();
    std::vector<std::vector<Xfer>> all(nr);
    int rc;
    for (int r = 0; r < nr; ++r) {
        IPPLB_REQUIRE(L->ctx[r]->plan && fields[r], "loop_halo_exchange: no layout bound / NULL field");
        if ((rc = halo_pack(L->ctx[r], fields[r], ncomp, mode))) return rc;
        halo_xfers(L->ctx[r], ncomp, mode, all[r]);
    }
    if (nr > 1 && (rc = loop_exchange(L, all))) return rc;
    for (int r = 0; r < nr; ++r)
        if ((rc = halo_unpack(L->ctx[r], fields[r], ncomp, mode, nr > 1))) return rc;
    return IPPLB_OK;
}

// ---- ParticleSpatialLayout::update on contiguous arrays --------------------------------------------------------------------
int ipplb_update_plan(ipplb_ctx* ctx, ipplb_particles* p, long* n_after_host, long* sent_host, long* recv_host) {
    IPPLB_REQUIRE(ctx && p, "update_plan: bad arguments");
    const int nr = ctx->nranks;
    if (sent_host) std::fill(sent_host, sent_host + nr, 0L);
    if (recv_host) std::fill(recv_host, recv_host + nr, 0L);
    if (n_after_host) *n_after_host = p->n;
    if (nr < 2) return IPPLB_OK;  // ParticleSpatialLayout.hpp:128
    CommPlan* P = (CommPlan*)ctx->plan;
    IPPLB_REQUIRE(P && ctx->nccl && !ctx->loop, "update_plan: no layout / NCCL communicator bound");
    int rc;
    if ((rc = update_locate(ctx, p))) return rc;
    if ((rc = gather_rows(ctx, P->d_counts, P->d_matrix, nr + UPD_EXTRA))) return rc;
    rc = update_sizes(ctx, p, sent_host, recv_host);
    if (n_after_host) *n_after_host = p->n - P->plan_nh + P->plan_na;
    P->planned = rc == IPPLB_OK || rc == IPPLB_ERR_CAPACITY;
    return rc;
}

int ipplb_update_commit(ipplb_ctx* ctx, ipplb_particles* p) {
    IPPLB_REQUIRE(ctx && p, "update_commit: bad arguments");
    const int nr = ctx->nranks;
    if (nr < 2) return IPPLB_OK;
    CommPlan* P = (CommPlan*)ctx->plan;
    IPPLB_REQUIRE(P && P->planned && P->plan_n == p->n, "update_commit: call ipplb_update_plan on the same particles first");
    P->planned = false;
    // the caller may have grown its arrays since the plan: agree on the outcome once more (one int all-reduce), so that a
    // rank that still does not fit takes every rank out with it instead of leaving them in the receive
    const long n_new = p->n - P->plan_nh + P->plan_na;
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * 2048))) return rc;
    int* d_ok = (int*)((double*)ctx->reduce.ptr + 1800);
    int* h_ok = (int*)(ctx->reduce_host + 32);
    *h_ok     = n_new > p->capacity ? 1 : 0;
    IPPLB_CUDA(cudaMemcpyAsync(d_ok, h_ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_NCCL(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    IPPLB_CUDA(cudaMemcpyAsync(h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*h_ok) {
        set_error("update_commit: %d rank(s) cannot hold their particles after the migration (this rank: %ld of %ld)", *h_ok, n_new, p->capacity);
        return IPPLB_ERR_CAPACITY;
    }
    if (P->plan_nh == 0 && P->plan_na == 0) {
        // nothing moves here; other ranks may still exchange among themselves
    }
    UpdateBufs B;
    std::vector<Xfer> x;
    if ((rc = update_pack(ctx, p, B, x))) return rc;
    if (!x.empty() && (rc = nccl_exchange(ctx, x))) return rc;
    return update_unpack(ctx, p, B);
}

int ipplb_update(ipplb_ctx* ctx, ipplb_particles* p, long* sent_host, long* recv_host) {
    IPPLB_REQUIRE(ctx && p, "update: bad arguments");
    int rc;
    if ((rc = ipplb_update_plan(ctx, p, nullptr, sent_host, recv_host))) return rc;
    if (ctx->nranks < 2) return IPPLB_OK;
    const int nr = ctx->nranks;
    if (sent_host) sent_host[ctx->rank] = 0;   // particles that stay are not "sent" (reference: the own rank is skipped)
    if (recv_host) recv_host[ctx->rank] = 0;
    (void)nr;
    return ipplb_update_commit(ctx, p);
}

int ipplb_loop_update(ipplb_loop* L, ipplb_particles* parts, long* sent_host, long* recv_host) {
    IPPLB_REQUIRE(L && parts, "loop_update: bad arguments");
    const int nr = (int)L->ctx.size();
    if (nr < 2) return IPPLB_OK;
    int rc;
    for (int r = 0; r < nr; ++r) {
        IPPLB_REQUIRE(L->ctx[r]->plan, "loop_update: no layout bound");
        if ((rc = update_locate(L->ctx[r], &parts[r]))) return rc;
    }
    for (int r = 0; r < nr; ++r) IPPLB_CUDA(cudaStreamSynchronize(L->ctx[r]->stream));
    for (int r = 0; r < nr; ++r) {
        CommPlan* P = (CommPlan*)L->ctx[r]->plan;
        if ((rc = gather_rows(L->ctx[r], P->d_counts, P->d_matrix, nr + UPD_EXTRA))) return rc;
    }
    for (int r = 0; r < nr; ++r) {
        long* s = sent_host ? sent_host + (size_t)r * nr : nullptr;
        long* v = recv_host ? recv_host + (size_t)r * nr : nullptr;
        if ((rc = update_sizes(L->ctx[r], &parts[r], s, v))) return rc;
        if (s) s[r] = 0;
        if (v) v[r] = 0;
    }
    std::vector<UpdateBufs> B(nr);
    std::vector<std::vector<Xfer>> all(nr);
    for (int r = 0; r < nr; ++r)
        if ((rc = update_pack(L->ctx[r], &parts[r], B[r], all[r]))) return rc;
    if ((rc = loop_exchange(L, all))) return rc;
    for (int r = 0; r < nr; ++r)
        if ((rc = update_unpack(L->ctx[r], &parts[r], B[r]))) return rc;
    return IPPLB_OK;
}

}  // extern "C"

// ---- migration of the bucketed store -----------------------------------------------------------------------------------
// The fused step has applied the BC and the ownership test, found every leaver's destination rank (reference search
// order) and written it as one 48-byte record into that rank's segment -- of the caller's exit buffer (ipplb_bins_migrate:
// NCCL messages, host-synchronous) or, after ipplb_migrate_connect, of the destination rank's own inbox over peer memory
// (ipplb_bins_migrate_async: no message, no host synchronisation).
namespace ipplb {

// a rank's row of the NCCL migration's count exchange: leavers per destination, error flags, free slots behind the tail
__global__ void migrate_row_kernel(const int* __restrict__ exit_cnt, int nranks, const int* __restrict__ misc, int capacity,
                                   int* __restrict__ row) {
    const int t = threadIdx.x;
    if (t < nranks) row[t] = exit_cnt[t];
    if (t == 0) {
        row[nranks]     = misc[BM_ST_FLAGS];
        row[nranks + 1] = capacity - misc[BM_ST_TAIL_START] - misc[BM_ST_TAIL];
    }
}

__global__ void __launch_bounds__(256)
arrivals_kernel(MeshDev m, const double* __restrict__ recs, int na, long tail_pos, double* __restrict__ x,
                double* __restrict__ y, double* __restrict__ z, double* __restrict__ px, double* __restrict__ py,
                double* __restrict__ pz, double q, double* __restrict__ rho, int* __restrict__ state,
                int* __restrict__ misc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        state[BS_TAIL_COUNT] += na;
        misc[BM_ST_TOTAL] += na;
        misc[BM_ST_TAIL] += na;
    }
    if (i >= na) return;
    const double2* rec = reinterpret_cast<const double2*>(recs + (size_t)i * 6);
    const double2 a = rec[0], b = rec[1], c = rec[2];
    const long g = tail_pos + i;
    x[g] = a.x; y[g] = a.y; z[g] = b.x;
    px[g] = b.y; py[g] = c.x; pz[g] = c.y;
    if (rho) {
        Cic cc;
        cic_setup(m, a.x, a.y, b.x, cc);
#pragma unroll
        for (int n = 0; n < 8; ++n) atomicAdd(&rho[cic_node(m, cc.a, n)], ipplb::dmul(q, cic_weight(cc.whi, n)));
    }
}

// [host-emulation begin: arrivals_p2p_kernel]  (tests/test_kernel_text_cpu.py compiles the marked text for the host and runs
// it under a lock-step block emulator, tests/emu/emu_arrivals.cpp)
struct ArriveArgs {
    MeshDev m;
    const int* matrix;      // [nranks][nranks] leavers per (source, destination)
    int nranks, me;
    const double* inbox;    // [nranks][seg_cap][6]
    long seg_cap;
    const int *start, *cap;
    int *count, *state, *misc;
    int capacity, ntx, nty;
    double* out[6];
    double q;
    double* rho;
};

// Arrivals straight into their buckets: arrival j of source s sits at inbox[s][j]; its tile's cursor hands out the slot
// (overflow -> tail), the charge is deposited, the status words follow.  Fixed grid, device-side counts.
__global__ void __launch_bounds__(256) arrivals_p2p_kernel(const ArriveArgs a) {
    __shared__ long off[MAX_RANKS + 1];
    if (threadIdx.x == 0) {
        long run = 0;
        for (int s = 0; s < a.nranks; ++s) {
            off[s] = run;
            const long c = a.matrix[s * a.nranks + a.me];
            run += c < a.seg_cap ? c : a.seg_cap;  // (a longer segment was truncated and flagged by its sender)
        }
        off[a.nranks] = run;
    }
    __syncthreads();
    const long total = off[a.nranks];
    int n_bucket = 0, n_tail = 0;
    for (long j = (long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long)gridDim.x * blockDim.x) {
        int s = 0;
        while (off[s + 1] <= j) ++s;
        const double2* rec = reinterpret_cast<const double2*>(a.inbox + ((size_t)s * a.seg_cap + (j - off[s])) * 6);
        const double2 r0 = rec[0], r1 = rec[1], r2 = rec[2];
        Cic c;
        cic_setup(a.m, r0.x, r0.y, r1.x, c);
        const int cx = c.a[0] - a.m.nghost, cy = c.a[1] - a.m.nghost, cz = c.a[2] - a.m.nghost;
        // (v - v == 0 fails for NaN and infinities; the conversion to a cell index turns NaN into cell 0, which is inside
        //  the box of a rank that starts there)
        const bool finite = (r0.x - r0.x == 0.0) && (r0.y - r0.y == 0.0) && (r1.x - r1.x == 0.0);
        if (!finite || cx < 0 || cx > a.m.nl[0] || cy < 0 || cy > a.m.nl[1] || cz < 0 || cz > a.m.nl[2]) {
            // not a position of this rank's box (a non-finite position ends up here through the "stay" fallback of
            // the destination search): no bucket can take it -- flag it instead of indexing the tables with it
            atomicOr(&a.misc[BM_ST_FLAGS], IPPLB_FLAG_INTERNAL);
            continue;
        }
        const int tile = (cx >> 2) + a.ntx * ((cy >> 2) + a.nty * (cz >> 2));
        long g;
        const int slot = atomicAdd(&a.count[tile], 1);
        if (slot < a.cap[tile]) {
            g = (long)a.start[tile] + slot;
            ++n_bucket;
        } else {
            atomicSub(&a.count[tile], 1);
            g = (long)a.state[BS_TAIL_START] + atomicAdd(&a.state[BS_TAIL_COUNT], 1);
            if (g >= a.capacity) {
                atomicSub(&a.state[BS_TAIL_COUNT], 1);
                atomicOr(&a.misc[BM_ST_FLAGS], IPPLB_FLAG_CAPACITY);
                g = -1;
            } else {
                ++n_tail;
            }
        }
        if (g >= 0) {
            a.out[0][g] = r0.x; a.out[1][g] = r0.y; a.out[2][g] = r1.x;
            a.out[3][g] = r1.y; a.out[4][g] = r2.x; a.out[5][g] = r2.y;
            if (a.rho) {
#pragma unroll
                for (int n = 0; n < 8; ++n) atomicAdd(&a.rho[cic_node(a.m, c.a, n)], ipplb::dmul(a.q, cic_weight(c.whi, n)));
            }
        }
    }
    if (n_bucket) { atomicAdd(&a.misc[BM_ST_BUCKETED], n_bucket); atomicAdd(&a.misc[BM_ST_TOTAL], n_bucket); }
    if (n_tail) { atomicAdd(&a.misc[BM_ST_TAIL], n_tail); atomicAdd(&a.misc[BM_ST_TOTAL], n_tail); }
}
// [host-emulation end: arrivals_p2p_kernel]

static int arrivals_p2p(ipplb_ctx* ctx, ipplb_bins* b, ipplb_particles* cur, double* rho) {
    CommPlan* P = (CommPlan*)ctx->plan;
    MigBox* M   = (MigBox*)ctx->mig;
    ArriveArgs a;
    a.m = make_mesh_dev(&b->mesh);
    a.matrix = P->d_matrix; a.nranks = ctx->nranks; a.me = ctx->rank;
    a.inbox = M->inbox[M->parity]; a.seg_cap = M->seg_cap;
    const int o = b->cur;
    a.start = b->start(o); a.cap = b->cap(o); a.count = b->count(o); a.state = b->state(o); a.misc = b->misc();
    a.capacity = (int)b->capacity; a.ntx = b->ntx; a.nty = b->nty;
    double* out[6] = {cur->x, cur->y, cur->z, cur->px, cur->py, cur->pz};
    for (int k = 0; k < 6; ++k) a.out[k] = out[k];
    a.q = cur->q_scalar; a.rho = rho;
    // latency bound (record load -> cursor atomic -> dependent stores): as many resident warps as the 40-register kernel allows
    arrivals_p2p_kernel<<<ctx->num_sms * 6, 256, 0, ctx->stream>>>(a);
    IPPLB_CHECK_LAUNCH(ctx);
    M->parity ^= 1;
    return IPPLB_OK;
}

double* const* mig_peer_table(ipplb_ctx* ctx, long* seg_cap) {
    MigBox* M = (MigBox*)ctx->mig;
    *seg_cap  = M->seg_cap;
    return M->d_peer[M->parity];
}

static int mig_alloc(ipplb_ctx* ctx, long seg_cap) {
    free_mig((MigBox*)ctx->mig);
    ctx->mig  = nullptr;
    MigBox* M = new MigBox();
    M->seg_cap = seg_cap;
    const size_t bytes = sizeof(double) * 6 * (size_t)seg_cap * ctx->nranks;
    for (int i = 0; i < 2; ++i) {
        IPPLB_CUDA(cudaMalloc(&M->inbox[i], bytes));
        IPPLB_CUDA(cudaMalloc(&M->d_peer[i], sizeof(double*) * ctx->nranks));
    }
    ctx->mig = M;
    return IPPLB_OK;
}

}  // namespace ipplb

extern "C" {

int ipplb_migrate_connect(ipplb_ctx* ctx, long seg_cap) {
    IPPLB_REQUIRE(ctx && seg_cap > 0, "migrate_connect: bad arguments");
    IPPLB_REQUIRE(ctx->nranks >= 2 && ctx->nccl && !ctx->loop, "migrate_connect: needs an NCCL communicator of at least 2 ranks");
    IPPLB_CUDA(cudaSetDevice(ctx->device));
    const int nr = ctx->nranks, me = ctx->rank;
    int rc;
    if (ctx->mig) {
        // reconnecting: every rank unmaps its peers' old inboxes before any of them is freed (one all-reduce as barrier)
        MigBox* O = (MigBox*)ctx->mig;
        for (void* p : O->opened) cudaIpcCloseMemHandle(p);
        O->opened.clear();
        if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * 2048))) return rc;
        int* d_bar = (int*)((double*)ctx->reduce.ptr + 1900);
        IPPLB_NCCL(ncclAllReduce(d_bar, d_bar, 1, ncclInt, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
        IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    // One segment size for the whole job: a sender addresses segment `me` of the destination's inbox with ITS seg_cap, the
    // destination reads segment s with its own -- the two must be the same number.  Callers size seg_cap from rank-local
    // particle counts (unequal after an ORB repartition), so the ranks agree on the largest request here.
    {
        if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * 2048))) return rc;
        long long* d_seg = (long long*)((double*)ctx->reduce.ptr + 1904);
        long long h_seg  = seg_cap;
        IPPLB_CUDA(cudaMemcpyAsync(d_seg, &h_seg, sizeof(h_seg), cudaMemcpyHostToDevice, ctx->stream));
        IPPLB_NCCL(ncclAllReduce(d_seg, d_seg, 1, ncclInt64, ncclMax, (ncclComm_t)ctx->nccl, ctx->stream));
        IPPLB_CUDA(cudaMemcpyAsync(&h_seg, d_seg, sizeof(h_seg), cudaMemcpyDeviceToHost, ctx->stream));
        IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
        seg_cap = (long)h_seg;
    }
    if ((rc = mig_alloc(ctx, seg_cap))) return rc;
    MigBox* M = (MigBox*)ctx->mig;
    // exchange the two inbox handles of every rank (one all-gather of 128 bytes per rank), then map the peers' inboxes
    struct Handles { cudaIpcMemHandle_t h[2]; };
    static_assert(sizeof(Handles) == 128, "two cudaIpc handles");
    Handles mine;
    for (int i = 0; i < 2; ++i) IPPLB_CUDA(cudaIpcGetMemHandle(&mine.h[i], M->inbox[i]));
    Handles* d_all = nullptr;
    IPPLB_CUDA(cudaMalloc(&d_all, sizeof(Handles) * nr));
    IPPLB_CUDA(cudaMemcpyAsync(d_all + me, &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_NCCL(ncclAllGather(d_all + me, d_all, sizeof(Handles), ncclChar, (ncclComm_t)ctx->nccl, ctx->stream));
    std::vector<Handles> all(nr);
    IPPLB_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * nr, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    IPPLB_CUDA(cudaFree(d_all));
    for (int i = 0; i < 2; ++i) {
        std::vector<double*> peer(nr, nullptr);
        for (int r = 0; r < nr; ++r) {
            if (r == me) { peer[r] = M->inbox[i]; continue; }
            void* p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, all[r].h[i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                set_error("migrate_connect: cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
                free_mig(M);
                ctx->mig = nullptr;
                return IPPLB_ERR_CUDA;
            }
            M->opened.push_back(p);
            peer[r] = (double*)p;
        }
        IPPLB_CUDA(cudaMemcpy(M->d_peer[i], peer.data(), sizeof(double*) * nr, cudaMemcpyHostToDevice));
    }
    return IPPLB_OK;
}

int ipplb_loop_migrate_connect(ipplb_loop* L, long seg_cap) {
    IPPLB_REQUIRE(L && seg_cap > 0, "loop_migrate_connect: bad arguments");
    const int nr = (int)L->ctx.size();
    int rc;
    for (int r = 0; r < nr; ++r)
        if ((rc = mig_alloc(L->ctx[r], seg_cap))) return rc;
    for (int r = 0; r < nr; ++r)
        for (int i = 0; i < 2; ++i) {
            std::vector<double*> peer(nr);
            for (int t = 0; t < nr; ++t) peer[t] = ((MigBox*)L->ctx[t]->mig)->inbox[i];
            IPPLB_CUDA(cudaMemcpy(((MigBox*)L->ctx[r]->mig)->d_peer[i], peer.data(), sizeof(double*) * nr, cudaMemcpyHostToDevice));
        }
    return IPPLB_OK;
}

int ipplb_bins_migrate_async(ipplb_ctx* ctx, ipplb_bins* b, ipplb_particles* cur, double* rho) {
    IPPLB_REQUIRE(ctx && b && cur, "bins_migrate_async: bad arguments");
    const int nr = ctx->nranks;
    if (nr < 2) return IPPLB_OK;
    CommPlan* P = (CommPlan*)ctx->plan;
    IPPLB_REQUIRE(P && ctx->mig && ctx->nccl && !ctx->loop, "bins_migrate_async: call ipplb_ctx_set_layout and ipplb_migrate_connect first");
    IPPLB_REQUIRE(b->exit_ranks == nr && b->exit_p2p, "bins_migrate_async: the last step did not write into the peers' inboxes");
    // the counts every rank wrote for every destination; on the stream this all-gather is also the barrier behind which
    // every peer's records have landed in this rank's inbox (their fused kernels have completed before they contribute)
    int rc;
    if ((rc = gather_rows(ctx, b->d_exit_cnt + MAX_RANKS, P->d_matrix, nr))) return rc;
    return arrivals_p2p(ctx, b, cur, rho);
}

int ipplb_loop_bins_migrate(ipplb_loop* L, ipplb_bins* const* bins, ipplb_particles* cur, double* const* rho) {
    IPPLB_REQUIRE(L && bins && cur, "loop_bins_migrate: bad arguments");
    const int nr = (int)L->ctx.size();
    if (nr < 2) return IPPLB_OK;
    int rc;
    for (int r = 0; r < nr; ++r) {
        ipplb_ctx* c = L->ctx[r];
        IPPLB_REQUIRE(c->plan && c->mig && bins[r] && bins[r]->exit_ranks == nr && bins[r]->exit_p2p,
                      "loop_bins_migrate: layout / inbox / last step not in peer mode");
        ((CommPlan*)c->plan)->loop_row = bins[r]->d_exit_cnt + MAX_RANKS;
        IPPLB_CUDA(cudaStreamSynchronize(c->stream));
    }
    for (int r = 0; r < nr; ++r) {
        ipplb_ctx* c = L->ctx[r];
        if ((rc = gather_rows(c, bins[r]->d_exit_cnt + MAX_RANKS, ((CommPlan*)c->plan)->d_matrix, nr))) return rc;
    }
    for (int r = 0; r < nr; ++r)
        if ((rc = arrivals_p2p(L->ctx[r], bins[r], &cur[r], rho ? rho[r] : nullptr))) return rc;
    return IPPLB_OK;
}

int ipplb_migrate_counts(ipplb_ctx* ctx, long* sent_host, long* recv_host) {
    IPPLB_REQUIRE(ctx && ctx->plan, "migrate_counts: bad arguments");
    CommPlan* P  = (CommPlan*)ctx->plan;
    const int nr = ctx->nranks, me = ctx->rank;
    IPPLB_CUDA(cudaMemcpyAsync(P->h_matrix, P->d_matrix, sizeof(int) * nr * nr, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < nr; ++r) {
        if (sent_host) sent_host[r] = r == me ? 0 : P->h_matrix[me * nr + r];
        if (recv_host) recv_host[r] = r == me ? 0 : P->h_matrix[r * nr + me];
    }
    return IPPLB_OK;
}

int ipplb_bins_migrate(ipplb_ctx* ctx, ipplb_bins* b, ipplb_particles* cur, const double* exit_buf,
                       int exit_cap, double* rho, long* sent_host, long* recv_host) {
    IPPLB_REQUIRE(ctx && b && cur, "bins_migrate: bad arguments");
    const int nr = ctx->nranks, me = ctx->rank;
    if (sent_host) std::fill(sent_host, sent_host + nr, 0L);
    if (recv_host) std::fill(recv_host, recv_host + nr, 0L);
    CommPlan* P = (CommPlan*)ctx->plan;
    int* h = b->h_status;
    const int W = nr + 2;  // a rank's row: leavers per destination, its error flags, its free slots behind the tail
    if (nr >= 2) {
        IPPLB_REQUIRE(P && ctx->nccl && exit_buf, "bins_migrate: no layout/communicator/exit buffer bound");
        IPPLB_REQUIRE(b->exit_ranks == nr && !b->exit_p2p, "bins_migrate: the last step was not run with this layout and an exit buffer");
        IPPLB_REQUIRE(W * nr <= nr * (nr + UPD_EXTRA), "bins_migrate: count matrix too small");
        migrate_row_kernel<<<1, 64, 0, ctx->stream>>>(b->d_exit_cnt + MAX_RANKS, nr, b->misc(), (int)b->capacity, P->d_counts);
        IPPLB_CHECK_LAUNCH(ctx);
        IPPLB_NCCL(ncclAllGather(P->d_counts, P->d_matrix, W, ncclInt, (ncclComm_t)ctx->nccl, ctx->stream));
        ctx->launches++;
        IPPLB_CUDA(cudaMemcpyAsync(P->h_matrix, P->d_matrix, sizeof(int) * nr * W, cudaMemcpyDeviceToHost, ctx->stream));
    }
    IPPLB_CUDA(cudaMemcpyAsync(h, b->misc(), sizeof(int) * BM_WORDS, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    cur->n = h[BM_ST_TOTAL];
    if (nr < 2) {
        const int flags = h[BM_ST_FLAGS];
        if (flags & (IPPLB_FLAG_EXIT_OVERFLOW | IPPLB_FLAG_CAPACITY | IPPLB_FLAG_INTERNAL)) {
            set_error("bins_migrate: the last fused step raised flags 0x%x", flags);
            return IPPLB_ERR_CAPACITY;
        }
        IPPLB_REQUIRE(h[BM_ST_EXIT] == 0, "bins_migrate: leavers on a single rank");
        return IPPLB_OK;
    }
    // every rank sees every rank's flags and room: the decision to abort is the same everywhere (no rank is left in a
    // receive whose sender has returned an error)
    const int* M = P->h_matrix;
    for (int r = 0; r < nr; ++r) {
        long in = 0;
        for (int t = 0; t < nr; ++t) in += M[t * W + r];
        const int flags = M[r * W + nr];
        if ((flags & (IPPLB_FLAG_EXIT_OVERFLOW | IPPLB_FLAG_CAPACITY | IPPLB_FLAG_INTERNAL)) || in > M[r * W + nr + 1]) {
            set_error("bins_migrate: rank %d: flags 0x%x, %ld arrivals for %d free slots behind its tail (every rank returns this error)",
                      r, flags, in, M[r * W + nr + 1]);
            return IPPLB_ERR_CAPACITY;
        }
    }
    const int seg = exit_cap / nr;
    long na = 0, nh = 0;
    std::vector<long> roff(nr + 1, 0);
    for (int r = 0; r < nr; ++r) {
        const int sc = M[me * W + r], rc_ = M[r * W + me];
        roff[r + 1] = roff[r] + rc_;
        nh += sc;
        if (sent_host) sent_host[r] = r == me ? 0 : sc;
        if (recv_host) recv_host[r] = r == me ? 0 : rc_;
    }
    na = roff[nr];
    if (na == 0 && nh == 0) return IPPLB_OK;
    const long tail_pos = (long)h[BM_ST_TAIL_START] + h[BM_ST_TAIL];
    int rc;
    if ((rc = ensure(ctx, ctx->recv, sizeof(double) * (size_t)(6 * na + 2)))) return rc;
    double* stage = (double*)ctx->recv.ptr;
    std::vector<Xfer> x;
    for (int r = 0; r < nr; ++r) {
        const int sc = M[me * W + r], rc_ = M[r * W + me];
        if (r == me || (!sc && !rc_)) continue;
        x.push_back({r, exit_buf + (size_t)r * 6 * seg, sizeof(double) * 6 * (size_t)sc, stage + 6 * roff[r], sizeof(double) * 6 * (size_t)rc_});
    }
    if (!x.empty() && (rc = nccl_exchange(ctx, x))) return rc;
    const int self = M[me * W + me];  // inclusive-fallback hits that stay here
    if (self)
        IPPLB_CUDA(cudaMemcpyAsync(stage + 6 * roff[me], exit_buf + (size_t)me * 6 * seg, sizeof(double) * 6 * (size_t)self,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    if (na > 0) {
        arrivals_kernel<<<(unsigned)((na + 255) / 256), 256, 0, ctx->stream>>>(make_mesh_dev(&b->mesh), stage, (int)na, tail_pos, cur->x, cur->y,
                                                             cur->z, cur->px, cur->py, cur->pz, cur->q_scalar, rho,
                                                             b->state(b->cur), b->misc());
        IPPLB_CHECK_LAUNCH(ctx);
    }
    cur->n += na;
    return IPPLB_OK;
}

int ipplb_allreduce_sum_f64(ipplb_ctx* ctx, double* value_host) {
    IPPLB_REQUIRE(ctx && value_host, "allreduce: bad arguments");
    if (ctx->nranks < 2) return IPPLB_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * 2048))) return rc;
    double* d = (double*)ctx->reduce.ptr + 1500;
    ctx->reduce_host[8] = *value_host;
    IPPLB_CUDA(cudaMemcpyAsync(d, ctx->reduce_host + 8, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_NCCL(ncclAllReduce(d, d, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    ctx->launches++;
    IPPLB_CUDA(cudaMemcpyAsync(ctx->reduce_host + 8, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    *value_host = ctx->reduce_host[8];
    return IPPLB_OK;
}

int ipplb_allreduce_max_f64(ipplb_ctx* ctx, double* value_host) {
    IPPLB_REQUIRE(ctx && value_host, "allreduce: bad arguments");
    if (ctx->nranks < 2) return IPPLB_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * 2048))) return rc;
    double* d = (double*)ctx->reduce.ptr + 1700;
    ctx->reduce_host[24] = *value_host;
    IPPLB_CUDA(cudaMemcpyAsync(d, ctx->reduce_host + 24, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_NCCL(ncclAllReduce(d, d, 1, ncclDouble, ncclMax, (ncclComm_t)ctx->nccl, ctx->stream));
    ctx->launches++;
    IPPLB_CUDA(cudaMemcpyAsync(ctx->reduce_host + 24, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    *value_host = ctx->reduce_host[24];
    return IPPLB_OK;
}

int ipplb_allreduce_sum_i64(ipplb_ctx* ctx, long* value_host) {
    IPPLB_REQUIRE(ctx && value_host, "allreduce: bad arguments");
    if (ctx->nranks < 2) return IPPLB_OK;
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * 2048))) return rc;
    long* d = (long*)((double*)ctx->reduce.ptr + 1600);
    long* h = (long*)(ctx->reduce_host + 16);
    *h      = *value_host;
    IPPLB_CUDA(cudaMemcpyAsync(d, h, sizeof(long), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_NCCL(ncclAllReduce(d, d, 1, ncclInt64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream));
    ctx->launches++;
    IPPLB_CUDA(cudaMemcpyAsync(h, d, sizeof(long), cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    *value_host = *h;
    return IPPLB_OK;
}

}  // extern "C"
