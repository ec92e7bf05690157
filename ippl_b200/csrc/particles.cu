// particles.cu -- scatter / gather / push / periodic BC / cell sort kernels for sm_100a.
//
// Data layout: SoA fp64 particle arrays (coalesced 64/128-bit accesses), ghosted x-fastest fields.
// All kernels are HBM-bound streaming kernels; no tensor cores (nothing here is a contraction).
#include <cub/device/device_scan.cuh>

#include <cstdint>

#include "keys.cuh"
#include "push.cuh"

namespace ipplb {

// ------------------------------------------------------------------------------------------------
// Scatter, any particle order.  One thread per particle; lanes of a warp that sit in the same cell
// in a contiguous run (the common case for cell-sorted or nearly sorted input) are combined with a
// segmented shuffle reduction so only run heads issue the 8 fp64 reductions (RED.E.ADD.F64).
// Replaces the per-particle Kokkos::atomic_add of ParticleAttrib.hpp:169-184 / CIC.hpp:26-45.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
scatter_atomic_kernel(MeshDev m, long begin, long end, const double* __restrict__ x,
                      const double* __restrict__ y, const double* __restrict__ z,
                      const double* __restrict__ q, double q_scalar, const int* __restrict__ hash,
                      double* __restrict__ rho) {
    const unsigned lane = threadIdx.x & 31u;
    const long stride   = (long)gridDim.x * blockDim.x;
    // all lanes of a warp iterate together (shuffles need full participation)
    long base = begin + (long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
    for (; base < end; base += stride) {
        const long idx   = base + lane;
        const bool valid = idx < end;
        double w[8];
        int key = -1 - (int)lane;  // distinct negative keys for idle lanes
        Cic c;
        c.a[0] = c.a[1] = c.a[2] = 0;
        if (valid) {
            const long i = hash ? (long)hash[idx] : idx;
            cic_setup(m, x[i], y[i], z[i], c);
            const double val = q ? q[i] : q_scalar;
#pragma unroll
            for (int p = 0; p < 8; ++p) w[p] = dmul(val, cic_weight(c.whi, p));
            key = cell_key(m, c.a);
        } else {
#pragma unroll
            for (int p = 0; p < 8; ++p) w[p] = 0.0;
        }
        // run id = number of run heads at or before this lane
        const int prev      = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head     = (lane == 0) || (prev != key);
        const unsigned hm   = __ballot_sync(0xffffffffu, head);
        const int rid       = __popc(hm & (0xffffffffu >> (31 - lane)));
        if (hm != 0xffffffffu) {  // at least one run longer than 1: segmented reduction
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int orid = __shfl_down_sync(0xffffffffu, rid, off);
                const bool take = (lane + off < 32) && (orid == rid);
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const double o = __shfl_down_sync(0xffffffffu, w[p], off);
                    if (take) w[p] = dadd(w[p], o);
                }
            }
        }
        if (valid && head) {
#pragma unroll
            for (int p = 0; p < 8; ++p) atomicAdd(&rho[cic_node(m, c.a, p)], w[p]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Scatter for cell-sorted particles: one thread per (cell, stencil node).
// A CTA owns CPB consecutive cells of the sorted order; their particles form one contiguous range
// that is staged through shared memory in windows (coalesced loads, weights computed once per
// particle), then each (cell,node) thread walks its own cell's slice of the window accumulating in
// a register -- no atomics, no shuffles in the inner loop.  One RED.F64 per (cell,node) at the end:
// 8 per CELL instead of 8 per PARTICLE.
// ------------------------------------------------------------------------------------------------
constexpr int SC_CPB = 32;    // cells per CTA
constexpr int SC_WIN = 1024;  // particles per staging window

__global__ void __launch_bounds__(SC_CPB * 8)
scatter_sorted_kernel(MeshDev m, int ncells, const double* __restrict__ x,
                      const double* __restrict__ y, const double* __restrict__ z,
                      const double* __restrict__ q, double q_scalar,
                      const int* __restrict__ offsets, double* __restrict__ rho) {
    __shared__ double s_w0[SC_WIN], s_w1[SC_WIN], s_w2[SC_WIN], s_v[SC_WIN];
    const int tid  = threadIdx.x;
    const int node = tid & 7;
    for (int c0 = blockIdx.x * SC_CPB; c0 < ncells; c0 += gridDim.x * SC_CPB) {
        const int cend = min(c0 + SC_CPB, ncells);
        const int pbeg = offsets[c0], pend = offsets[cend];
        if (pbeg == pend) continue;  // uniform across the CTA
        const int cell = c0 + (tid >> 3);
        int mb = 0, me = 0;
        if (cell < cend) {
            mb = offsets[cell];
            me = offsets[cell + 1];
        }
        double acc = 0.0;
        for (int w = pbeg; w < pend; w += SC_WIN) {
            const int wn = min(SC_WIN, pend - w);
            for (int i = tid; i < wn; i += SC_CPB * 8) {
                const long g = (long)w + i;
                int idx;
                double a, b, c;
                cic_axis(x[g], m.origin[0], m.invdx[0], idx, a);
                cic_axis(y[g], m.origin[1], m.invdx[1], idx, b);
                cic_axis(z[g], m.origin[2], m.invdx[2], idx, c);
                s_w0[i] = a;
                s_w1[i] = b;
                s_w2[i] = c;
                s_v[i]  = q ? q[g] : q_scalar;
            }
            __syncthreads();
            const int lo = max(mb, w) - w, hi = min(me, w + wn) - w;
            for (int i = lo; i < hi; ++i) {
                double w0 = s_w0[i], w1 = s_w1[i], w2 = s_w2[i];
                if (node & 1) w0 = dsub(1.0, w0);
                if (node & 2) w1 = dsub(1.0, w1);
                if (node & 4) w2 = dsub(1.0, w2);
                acc = dadd(acc, dmul(s_v[i], dmul(w0, dmul(w1, w2))));
            }
            __syncthreads();
        }
        if (me > mb) {
            int a[3];
            key_to_args(m, cell, a);
            atomicAdd(&rho[cic_node(m, a, node)], acc);
        }
    }
}

template <int NCOMP>
__global__ void __launch_bounds__(256)
gather_kernel(MeshDev m, long n, const double* __restrict__ x, const double* __restrict__ y,
              const double* __restrict__ z, const double* __restrict__ f, double* __restrict__ o0,
              double* __restrict__ o1, double* __restrict__ o2, int add) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Cic c;
        cic_setup(m, x[i], y[i], z[i], c);
        double g[NCOMP];
        gather_point<NCOMP>(m, c, f, g);
        double* o[3] = {o0, o1, o2};
#pragma unroll
        for (int d = 0; d < NCOMP; ++d) o[d][i] = add ? dadd(o[d][i], g[d]) : g[d];
    }
}

// ------------------------------------------------------------------------------------------------
// Fused gather + push (+ periodic BC): reads R,P once, writes R,P once; E lives in registers.
//   leapfrog: P = P - c*E (kick2), P = P - c*E (kick1), R = R + dt*P, wrap     (c = 0.5*dt)
//   penning : Kick2, Kick1 of PenningTrapManager.h:313-333 / 256-272, drift, wrap
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_push_kernel(MeshDev m, PushDev P, long n, double* __restrict__ x, double* __restrict__ y,
                   double* __restrict__ z, double* __restrict__ px, double* __restrict__ py,
                   double* __restrict__ pz, const double* __restrict__ ef) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double r[3] = {x[i], y[i], z[i]};
        double p[3] = {px[i], py[i], pz[i]};
        Cic c;
        cic_setup(m, r[0], r[1], r[2], c);
        double E[3];
        gather_point<3>(m, c, ef, E);
        push_particle(P, r, p, E);
        x[i]  = r[0];
        y[i]  = r[1];
        z[i]  = r[2];
        px[i] = p[0];
        py[i] = p[1];
        pz[i] = p[2];
    }
}

// variant 2 of the two kernels above (ipplb_ctx_set_gather_variant): gather_point3_vec, otherwise the same
__global__ void __launch_bounds__(256)
gather3v_kernel(MeshDev m, long n, const double* __restrict__ x, const double* __restrict__ y,
                const double* __restrict__ z, const double* __restrict__ f, double* __restrict__ o0,
                double* __restrict__ o1, double* __restrict__ o2, int add) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Cic c;
        cic_setup(m, x[i], y[i], z[i], c);
        double g[3];
        gather_point3_vec(m, c, f, g);
        double* o[3] = {o0, o1, o2};
#pragma unroll
        for (int d = 0; d < 3; ++d) o[d][i] = add ? dadd(o[d][i], g[d]) : g[d];
    }
}

__global__ void __launch_bounds__(256)
gather_push_v_kernel(MeshDev m, PushDev P, long n, double* __restrict__ x, double* __restrict__ y,
                     double* __restrict__ z, double* __restrict__ px, double* __restrict__ py,
                     double* __restrict__ pz, const double* __restrict__ ef) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double r[3] = {x[i], y[i], z[i]};
        double p[3] = {px[i], py[i], pz[i]};
        Cic c;
        cic_setup(m, r[0], r[1], r[2], c);
        double E[3];
        gather_point3_vec(m, c, ef, E);
        push_particle(P, r, p, E);
        x[i]  = r[0];
        y[i]  = r[1];
        z[i]  = r[2];
        px[i] = p[0];
        py[i] = p[1];
        pz[i] = p[2];
    }
}

// unfused pieces --------------------------------------------------------------------------------
__global__ void axpy_kernel(long n, double a, const double* __restrict__ x, double* __restrict__ y) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = dadd(y[i], dmul(a, x[i]));
}

__global__ void periodic_bc_kernel(long n, double* __restrict__ x, double* __restrict__ y,
                                   double* __restrict__ z, PushDev P, int mask) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (mask & 1) x[i] = periodic_wrap(x[i], P.ext[0], P.mid[0]);
        if (mask & 2) y[i] = periodic_wrap(y[i], P.ext[1], P.mid[1]);
        if (mask & 4) z[i] = periodic_wrap(z[i], P.ext[2], P.mid[2]);
    }
}

__global__ void penning_kick_kernel(int which, PushDev P, long n, const double* __restrict__ x,
                                    const double* __restrict__ y, const double* __restrict__ z,
                                    double* __restrict__ px, double* __restrict__ py,
                                    double* __restrict__ pz, const double* __restrict__ ex,
                                    const double* __restrict__ ey, const double* __restrict__ ez) {
    const long stride = (long)gridDim.x * blockDim.x;
    P.kind     = IPPLB_PUSH_PENNING;
    P.do_kick1 = which == 1;
    P.do_kick2 = which == 2;
    P.do_drift = P.do_bc = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double r[3] = {x[i], y[i], z[i]};
        double p[3] = {px[i], py[i], pz[i]};
        double E[3] = {ex[i], ey[i], ez[i]};
        push_particle(P, r, p, E);
        px[i] = p[0];
        py[i] = p[1];
        pz[i] = p[2];
    }
}

// Counting sort by cell (integer keys): sort_keys_kernel lives in keys.cuh (shared with bins.cu).
__global__ void __launch_bounds__(256)
sort_move_kernel(long n, const int* __restrict__ keys, int* __restrict__ cursor,
                 const double* __restrict__ x, const double* __restrict__ y,
                 const double* __restrict__ z, const double* __restrict__ px,
                 const double* __restrict__ py, const double* __restrict__ pz,
                 const double* __restrict__ q, double* __restrict__ ox, double* __restrict__ oy,
                 double* __restrict__ oz, double* __restrict__ opx, double* __restrict__ opy,
                 double* __restrict__ opz, double* __restrict__ oq) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int slot = atomicAdd(&cursor[keys[i]], 1);
        ox[slot]  = x[i];
        oy[slot]  = y[i];
        oz[slot]  = z[i];
        if (px) {
            opx[slot] = px[i];
            opy[slot] = py[i];
            opz[slot] = pz[i];
        }
        if (q) oq[slot] = q[i];
    }
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

int ipplb_scatter_cic(ipplb_ctx* ctx, const ipplb_mesh* mesh, long begin, long end, const double* x,
                      const double* y, const double* z, const double* q, double q_scalar,
                      const int* hash, double* rho) {
    IPPLB_REQUIRE(ctx && mesh && rho && begin <= end, "scatter: bad arguments");
    if (end == begin) return IPPLB_OK;
    MeshDev m = make_mesh_dev(mesh);
    scatter_atomic_kernel<<<grid_for(ctx, end - begin, 256, 16), 256, 0, ctx->stream>>>(
        m, begin, end, x, y, z, q, q_scalar, hash, rho);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

long ipplb_sort_ncells(const ipplb_mesh* mesh) {
    return (long)tiles_along(mesh->nl[0]) * tiles_along(mesh->nl[1]) * tiles_along(mesh->nl[2]) * TILE_CELLS;
}

int ipplb_scatter_cic_sorted(ipplb_ctx* ctx, const ipplb_mesh* mesh, long n, const double* x,
                             const double* y, const double* z, const double* q, double q_scalar,
                             const int* cell_offsets, double* rho) {
    IPPLB_REQUIRE(ctx && mesh && rho && cell_offsets, "scatter_sorted: bad arguments");
    if (n == 0) return IPPLB_OK;
    MeshDev m        = make_mesh_dev(mesh);
    const int ncells = (int)ipplb_sort_ncells(mesh);
    int grid         = (ncells + SC_CPB - 1) / SC_CPB;
    scatter_sorted_kernel<<<grid, SC_CPB * 8, 0, ctx->stream>>>(m, ncells, x, y, z, q, q_scalar,
                                                                cell_offsets, rho);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_gather_cic(ipplb_ctx* ctx, const ipplb_mesh* mesh, long n, const double* x, const double* y,
                     const double* z, const double* field, int ncomp, double* const* out,
                     int add_to_attribute) {
    IPPLB_REQUIRE(ctx && mesh && field && out && (ncomp == 1 || ncomp == 3), "gather: bad arguments");
    if (n == 0) return IPPLB_OK;
    MeshDev m = make_mesh_dev(mesh);
    int grid  = grid_for(ctx, n, 256, 16);
    if (ncomp == 3 && ctx->gather_variant == 2 && ((uintptr_t)field & 15) == 0)
        gather3v_kernel<<<grid, 256, 0, ctx->stream>>>(m, n, x, y, z, field, out[0], out[1], out[2], add_to_attribute);
    else if (ncomp == 3)
        gather_kernel<3><<<grid, 256, 0, ctx->stream>>>(m, n, x, y, z, field, out[0], out[1], out[2],
                                                        add_to_attribute);
    else
        gather_kernel<1><<<grid, 256, 0, ctx->stream>>>(m, n, x, y, z, field, out[0], nullptr, nullptr,
                                                        add_to_attribute);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_gather_push(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push,
                      ipplb_particles* p, const double* efield) {
    IPPLB_REQUIRE(ctx && mesh && push && p && efield, "gather_push: bad arguments");
    if (p->n == 0) return IPPLB_OK;
    MeshDev m = make_mesh_dev(mesh);
    PushDev P = make_push_dev(mesh, push);
    if (ctx->gather_variant == 2 && ((uintptr_t)efield & 15) == 0)
        gather_push_v_kernel<<<grid_for(ctx, p->n, 256, 16), 256, 0, ctx->stream>>>(
            m, P, p->n, p->x, p->y, p->z, p->px, p->py, p->pz, efield);
    else
        gather_push_kernel<<<grid_for(ctx, p->n, 256, 16), 256, 0, ctx->stream>>>(
            m, P, p->n, p->x, p->y, p->z, p->px, p->py, p->pz, efield);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_ctx_set_gather_variant(ipplb_ctx* ctx, int variant) {
    IPPLB_REQUIRE(ctx && (variant == 1 || variant == 2), "ctx_set_gather_variant: variant is 1 or 2");
    ctx->gather_variant = variant;
    return IPPLB_OK;
}

int ipplb_axpy(ipplb_ctx* ctx, long n, double a, const double* x, double* y) {
    IPPLB_REQUIRE(ctx && x && y, "axpy: bad arguments");
    if (n == 0) return IPPLB_OK;
    axpy_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(n, a, x, y);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_apply_periodic_bc(ipplb_ctx* ctx, long n, double* x, double* y, double* z,
                            const double lo[3], const double hi[3], int mask) {
    IPPLB_REQUIRE(ctx && x && y && z, "apply_periodic_bc: bad arguments");
    if (n == 0) return IPPLB_OK;
    PushDev P;
    std::memset(&P, 0, sizeof(P));
    for (int d = 0; d < 3; ++d) {
        P.ext[d] = hi[d] - lo[d];
        P.mid[d] = (lo[d] + hi[d]) / 2;
    }
    periodic_bc_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(n, x, y, z, P, mask);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_penning_kick(ipplb_ctx* ctx, int which, const ipplb_push* push, long n, const double* x,
                       const double* y, const double* z, double* px, double* py, double* pz,
                       const double* ex, const double* ey, const double* ez) {
    IPPLB_REQUIRE(ctx && push && (which == 1 || which == 2), "penning_kick: bad arguments");
    if (n == 0) return IPPLB_OK;
    ipplb_mesh dummy;
    std::memset(&dummy, 0, sizeof(dummy));
    ipplb_push pp = *push;
    pp.kind       = IPPLB_PUSH_PENNING;
    PushDev P     = make_push_dev(&dummy, &pp);
    penning_kick_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(which, P, n, x, y, z, px,
                                                                            py, pz, ex, ey, ez);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_sort_by_cell(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_particles* in,
                       ipplb_particles* out, int* cell_offsets) {
    IPPLB_REQUIRE(ctx && mesh && in && out && cell_offsets, "sort: bad arguments");
    IPPLB_REQUIRE(out->capacity >= in->n, "sort: output capacity too small");
    const long n      = in->n;
    const long ncells = ipplb_sort_ncells(mesh);
    MeshDev m         = make_mesh_dev(mesh);
    int rc;
    if ((rc = ensure(ctx, ctx->keys, sizeof(int) * (size_t)(n > 0 ? n : 1)))) return rc;
    if ((rc = ensure(ctx, ctx->counts, sizeof(int) * (size_t)(ncells + 1)))) return rc;
    int* keys   = (int*)ctx->keys.ptr;
    int* counts = (int*)ctx->counts.ptr;
    IPPLB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * (ncells + 1), ctx->stream));
    if (n > 0) {
        sort_keys_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(m, n, in->x, in->y, in->z,
                                                                             keys, counts);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts, cell_offsets, (int)(ncells + 1),
                                  ctx->stream);
    if ((rc = ensure(ctx, ctx->cub_tmp, tmp_bytes))) return rc;
    IPPLB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.ptr, tmp_bytes, counts, cell_offsets,
                                             (int)(ncells + 1), ctx->stream));
    ctx->launches += 2;
    // cursor = copy of the offsets
    IPPLB_CUDA(cudaMemcpyAsync(counts, cell_offsets, sizeof(int) * (ncells + 1),
                               cudaMemcpyDeviceToDevice, ctx->stream));
    if (n > 0) {
        sort_move_kernel<<<grid_for(ctx, n, 256, 16), 256, 0, ctx->stream>>>(
            n, keys, counts, in->x, in->y, in->z, in->px, in->py, in->pz, in->q, out->x, out->y,
            out->z, out->px, out->py, out->pz, out->q);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    out->n        = n;
    out->q_scalar = in->q_scalar;
    return IPPLB_OK;
}

}  // extern "C"
