// comm.cu -- multi-GPU part: NCCL bootstrap, batched halo exchange, particle migration.
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
    int* d_counts = nullptr;       // [nranks] send counts, then cursor copy [nranks], then misc[8]
    int* d_matrix = nullptr;       // [nranks*nranks]
    int *h_counts = nullptr, *h_matrix = nullptr;  // pinned
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

using namespace ipplb;

extern "C" {

int ipplb_ctx_destroy(ipplb_ctx* ctx) {
    if (!ctx) return IPPLB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
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
    IPPLB_CUDA(cudaMalloc(&P->d_counts, sizeof(int) * (4 * nr + 16)));
    IPPLB_CUDA(cudaMalloc(&P->d_matrix, sizeof(int) * nr * nr));
    IPPLB_CUDA(cudaMallocHost(&P->h_counts, sizeof(int) * (4 * nr + 16)));
    IPPLB_CUDA(cudaMallocHost(&P->h_matrix, sizeof(int) * nr * nr));
    ctx->plan = P;
    return IPPLB_OK;
}

int ipplb_halo_exchange(ipplb_ctx* ctx, double* field, int ncomp, int mode) {
    IPPLB_REQUIRE(ctx && field && ncomp >= 1 && (mode == 0 || mode == 1), "halo_exchange: bad arguments");
    CommPlan* P = (CommPlan*)ctx->plan;
    IPPLB_REQUIRE(P, "halo_exchange: no layout bound (ipplb_ctx_set_layout)");
    const int e0 = P->mesh.nl[0] + 2 * P->mesh.nghost, e1 = P->mesh.nl[1] + 2 * P->mesh.nghost;
    if (ctx->nranks > 1 && (P->send_cells || P->recv_cells)) {
        IPPLB_REQUIRE(ctx->nccl, "halo_exchange: communicator not initialised");
        // fill: pack my `send` strips, receive into my `recv` strips.  accumulate: roles swap.
        const bool fill         = mode == 0;
        const long out_cells    = fill ? P->send_cells : P->recv_cells;
        const long in_cells     = fill ? P->recv_cells : P->send_cells;
        int rc;
        if ((rc = ensure(ctx, ctx->send, sizeof(double) * (size_t)(out_cells * ncomp + 1)))) return rc;
        if ((rc = ensure(ctx, ctx->recv, sizeof(double) * (size_t)(in_cells * ncomp + 1)))) return rc;
        double* sb = (double*)ctx->send.ptr;
        double* rb = (double*)ctx->recv.ptr;
        if (out_cells) {
            halo_copy_kernel<<<grid1d(out_cells * ncomp), 256, 0, ctx->stream>>>(
                fill ? P->d_send : P->d_recv, fill ? P->d_send_prefix : P->d_recv_prefix,
                (int)(fill ? P->send_regions.size() : P->recv_regions.size()), out_cells, ncomp, e0, e1,
                field, sb, 0);
            IPPLB_CHECK_LAUNCH(ctx);
        }
        IPPLB_NCCL(ncclGroupStart());
        for (auto& seg : P->peers) {
            const long so = fill ? seg.send_off : seg.recv_off, sc = fill ? seg.send_cells : seg.recv_cells;
            const long ro = fill ? seg.recv_off : seg.send_off, rc2 = fill ? seg.recv_cells : seg.send_cells;
            if (sc) IPPLB_NCCL(ncclSend(sb + so * ncomp, (size_t)(sc * ncomp), ncclDouble, seg.peer, (ncclComm_t)ctx->nccl, ctx->stream));
            if (rc2) IPPLB_NCCL(ncclRecv(rb + ro * ncomp, (size_t)(rc2 * ncomp), ncclDouble, seg.peer, (ncclComm_t)ctx->nccl, ctx->stream));
        }
        IPPLB_NCCL(ncclGroupEnd());
        ctx->launches++;
        if (in_cells) {
            halo_copy_kernel<<<grid1d(in_cells * ncomp), 256, 0, ctx->stream>>>(
                fill ? P->d_recv : P->d_send, fill ? P->d_recv_prefix : P->d_send_prefix,
                (int)(fill ? P->recv_regions.size() : P->send_regions.size()), in_cells, ncomp, e0, e1,
                field, rb, fill ? 1 : 2);
            IPPLB_CHECK_LAUNCH(ctx);
        }
    }
    if (P->serial_mask) {
        return mode == 0 ? ipplb_halo_fill_periodic(ctx, &P->mesh, field, ncomp, P->serial_mask)
                         : ipplb_halo_accumulate_periodic(ctx, &P->mesh, field, ncomp, P->serial_mask);
    }
    return IPPLB_OK;
}

int ipplb_update(ipplb_ctx* ctx, ipplb_particles* p, long* sent_host, long* recv_host) {
    IPPLB_REQUIRE(ctx && p, "update: bad arguments");
    const int nr = ctx->nranks, me = ctx->rank;
    if (sent_host) std::fill(sent_host, sent_host + nr, 0L);
    if (recv_host) std::fill(recv_host, recv_host + nr, 0L);
    if (nr < 2) return IPPLB_OK;  // ParticleSpatialLayout.hpp:128
    CommPlan* P = (CommPlan*)ctx->plan;
    IPPLB_REQUIRE(P && ctx->nccl, "update: no layout/communicator bound");
    const long n = p->n;
    if (P->dest_cap < n + 1) {
        if (P->d_dest) { IPPLB_CUDA(cudaStreamSynchronize(ctx->stream)); IPPLB_CUDA(cudaFree(P->d_dest)); }
        P->dest_cap = n + n / 4 + 1024;
        IPPLB_CUDA(cudaMalloc(&P->d_dest, sizeof(int) * P->dest_cap));
    }
    int* cnt = P->d_counts;          // [nr] send counts
    int* cursor = cnt + nr;          // [nr]
    int* soff = cnt + 2 * nr;        // [nr+1] send offsets
    int* counters = cnt + 3 * nr + 8;  // [2] small counters
    IPPLB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (4 * nr + 16), ctx->stream));
    if (n > 0) {
        locate_kernel<<<grid1d(n), 256, 0, ctx->stream>>>(P->d_regions, nr, me, n, p->x, p->y, p->z,
                                                          P->d_dest, cnt);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    // counts: every rank contributes its row; the matrix gives both what I send and what I receive
    IPPLB_NCCL(ncclAllGather(cnt, P->d_matrix, nr, ncclInt, (ncclComm_t)ctx->nccl, ctx->stream));
    ctx->launches++;
    IPPLB_CUDA(cudaMemcpyAsync(P->h_matrix, P->d_matrix, sizeof(int) * nr * nr, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int> h_soff(nr + 1, 0), h_roff(nr + 1, 0), h_rcnt(nr, 0), h_scnt(nr, 0);
    for (int r = 0; r < nr; ++r) {
        h_scnt[r] = P->h_matrix[me * nr + r];
        h_rcnt[r] = P->h_matrix[r * nr + me];
        h_soff[r + 1] = h_soff[r] + h_scnt[r];
        h_roff[r + 1] = h_roff[r] + h_rcnt[r];
        if (sent_host) sent_host[r] = h_scnt[r];
        if (recv_host) recv_host[r] = h_rcnt[r];
    }
    const int nh = h_soff[nr], na = h_roff[nr];
    if (nh == 0 && na == 0) return IPPLB_OK;
    const long n_new = n - nh + na;
    if (n_new > p->capacity) {
        set_error("update: %ld particles after migration exceed the capacity %ld", n_new, p->capacity);
        return IPPLB_ERR_CAPACITY;
    }
    AttrPtrs A;
    A.n = 0;
    double* attrs[NATTR] = {p->x, p->y, p->z, p->px, p->py, p->pz, p->q};
    for (int a = 0; a < NATTR; ++a) if (attrs[a]) A.a[A.n++] = attrs[a];
    int rc;
    if ((rc = ensure(ctx, ctx->send, sizeof(double) * ((size_t)nh * A.n + 1)))) return rc;
    if ((rc = ensure(ctx, ctx->recv, sizeof(double) * ((size_t)na * A.n + 1)))) return rc;
    // misc: holes[nh], low_holes[nh], movers[nh], tail_flag[nh], recv offsets/counts
    if ((rc = ensure(ctx, ctx->misc, sizeof(int) * ((size_t)4 * nh + 4 * nr + 64)))) return rc;
    int* holes = (int*)ctx->misc.ptr;
    int* low_holes = holes + nh; int* movers = low_holes + nh; int* tail_flag = movers + nh;
    int* d_roff = tail_flag + nh; int* d_rcnt = d_roff + nr + 1;
    double* sb = (double*)ctx->send.ptr; double* rb = (double*)ctx->recv.ptr;
    IPPLB_CUDA(cudaMemcpyAsync(soff, h_soff.data(), sizeof(int) * (nr + 1), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_CUDA(cudaMemcpyAsync(d_roff, h_roff.data(), sizeof(int) * (nr + 1), cudaMemcpyHostToDevice, ctx->stream));
    IPPLB_CUDA(cudaMemcpyAsync(d_rcnt, h_rcnt.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, ctx->stream));
    if (nh > 0) {
        pack_leavers_kernel<<<grid1d(n), 256, 0, ctx->stream>>>(n, me, P->d_dest, soff, cnt, cursor, A, sb, holes);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    IPPLB_NCCL(ncclGroupStart());
    for (int r = 0; r < nr; ++r) {
        if (r == me) continue;
        if (h_scnt[r]) IPPLB_NCCL(ncclSend(sb + (size_t)h_soff[r] * A.n, (size_t)h_scnt[r] * A.n, ncclDouble, r, (ncclComm_t)ctx->nccl, ctx->stream));
        if (h_rcnt[r]) IPPLB_NCCL(ncclRecv(rb + (size_t)h_roff[r] * A.n, (size_t)h_rcnt[r] * A.n, ncclDouble, r, (ncclComm_t)ctx->nccl, ctx->stream));
    }
    IPPLB_NCCL(ncclGroupEnd());
    ctx->launches++;
    if (na > 0) {
        unpack_arrivals_kernel<<<(unsigned)((na + 255) / 256), 256, 0, ctx->stream>>>(nr, d_roff, d_rcnt, na, rb, A, holes, nh, n);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    if (nh > na) {
        const int tail = nh - na;
        IPPLB_CUDA(cudaMemsetAsync(tail_flag, 0, sizeof(int) * tail, ctx->stream));
        mark_tail_holes_kernel<<<grid1d(tail), 256, 0, ctx->stream>>>(holes, na, nh, n_new, tail_flag, low_holes, counters);
        IPPLB_CHECK_LAUNCH(ctx);
        collect_tail_survivors_kernel<<<grid1d(tail), 256, 0, ctx->stream>>>(n_new, tail, tail_flag, movers, counters);
        IPPLB_CHECK_LAUNCH(ctx);
        fill_low_holes_kernel<<<grid1d(tail), 256, 0, ctx->stream>>>(low_holes, movers, counters, A);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    p->n = n_new;
    return IPPLB_OK;
}

// Migration for the bucketed store.  The fused step already applied the BC and the ownership test, found every
// leaver's destination rank (reference search order) and appended it as one 48-byte record to that rank's segment
// of exit_buf [nranks][exit_cap / nranks][6].  Here: one all-gather of the per-destination counts, ONE host sync
// (counts + tail position), ONE ncclSend + ncclRecv per peer (a segment is one contiguous message) into a staging
// buffer, and ONE kernel that turns the arrived records into SoA slots behind the tail of `cur`, commits the tail
// count and deposits the arrivals into rho (the reference scatters after update(), so arrivals belong to this
// step's rho: AlpineManager.h:157-175).
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

}  // extern "C"  (kernel above has C++ linkage)
extern "C" {

int ipplb_bins_migrate(ipplb_ctx* ctx, ipplb_bins* b, ipplb_particles* cur, const double* exit_buf,
                       int exit_cap, double* rho, long* sent_host, long* recv_host) {
    IPPLB_REQUIRE(ctx && b && cur, "bins_migrate: bad arguments");
    const int nr = ctx->nranks, me = ctx->rank;
    if (sent_host) std::fill(sent_host, sent_host + nr, 0L);
    if (recv_host) std::fill(recv_host, recv_host + nr, 0L);
    CommPlan* P = (CommPlan*)ctx->plan;
    int* h = b->h_status;
    if (nr >= 2) {
        IPPLB_REQUIRE(P && ctx->nccl && exit_buf, "bins_migrate: no layout/communicator/exit buffer bound");
        IPPLB_REQUIRE(b->exit_ranks == nr, "bins_migrate: the last step was not run with this layout");
        IPPLB_NCCL(ncclAllGather(b->d_exit_cnt + MAX_RANKS, P->d_matrix, nr, ncclInt, (ncclComm_t)ctx->nccl, ctx->stream));
        ctx->launches++;
        IPPLB_CUDA(cudaMemcpyAsync(P->h_matrix, P->d_matrix, sizeof(int) * nr * nr, cudaMemcpyDeviceToHost, ctx->stream));
    }
    IPPLB_CUDA(cudaMemcpyAsync(h, b->misc(), sizeof(int) * BM_WORDS, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    const int flags = h[BM_ST_FLAGS];
    if (flags & (IPPLB_FLAG_EXIT_OVERFLOW | IPPLB_FLAG_CAPACITY | IPPLB_FLAG_INTERNAL)) {
        set_error("bins_migrate: the last fused step raised flags 0x%x (exit buffer %d, capacity %ld)", flags,
                  exit_cap, b->capacity);
        return IPPLB_ERR_CAPACITY;
    }
    cur->n = h[BM_ST_TOTAL];
    if (nr < 2) {
        IPPLB_REQUIRE(h[BM_ST_EXIT] == 0, "bins_migrate: leavers on a single rank");
        return IPPLB_OK;
    }
    const int seg = exit_cap / nr;
    long na = 0, nh = 0;
    std::vector<long> roff(nr + 1, 0);
    for (int r = 0; r < nr; ++r) {
        const int sc = P->h_matrix[me * nr + r], rc_ = P->h_matrix[r * nr + me];
        roff[r + 1] = roff[r] + rc_;
        nh += sc;
        if (sent_host) sent_host[r] = r == me ? 0 : sc;
        if (recv_host) recv_host[r] = r == me ? 0 : rc_;
    }
    na = roff[nr];
    if (na == 0 && nh == 0) return IPPLB_OK;
    const long tail_pos = (long)h[BM_ST_TAIL_START] + h[BM_ST_TAIL];
    if (tail_pos + na > b->capacity) {
        set_error("bins_migrate: %ld arrivals do not fit behind the tail (%ld of %ld used)", na, tail_pos, b->capacity);
        return IPPLB_ERR_CAPACITY;
    }
    int rc;
    if ((rc = ensure(ctx, ctx->recv, sizeof(double) * (size_t)(6 * na + 2)))) return rc;
    double* stage = (double*)ctx->recv.ptr;
    IPPLB_NCCL(ncclGroupStart());
    for (int r = 0; r < nr; ++r) {
        const int sc = P->h_matrix[me * nr + r], rc_ = P->h_matrix[r * nr + me];
        if (r == me) continue;
        if (sc) IPPLB_NCCL(ncclSend(exit_buf + (size_t)r * 6 * seg, (size_t)6 * sc, ncclDouble, r, (ncclComm_t)ctx->nccl, ctx->stream));
        if (rc_) IPPLB_NCCL(ncclRecv(stage + 6 * roff[r], (size_t)6 * rc_, ncclDouble, r, (ncclComm_t)ctx->nccl, ctx->stream));
    }
    IPPLB_NCCL(ncclGroupEnd());
    ctx->launches++;
    const int self = P->h_matrix[me * nr + me];  // inclusive-fallback hits that stay here
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
