// diag.cu -- diagnostics of the alpine dumps (field energies, norms, kinetic energy) and the
// device part of orthogonal recursive bisection (plane sums) with its host state machine.
// All reductions are two-stage (block partials, then one block) so the result is deterministic for a given grid.
#include <algorithm>
#include <numeric>
#include <vector>

#include <nccl.h>

#include "bins.h"

namespace ipplb {

constexpr int RB = 256;  // reduction block

// block-wide sum / max of NV values per thread; thread 0 returns the result in v[]
template <int NV, int NMAX>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* sh) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        sh[threadIdx.x] = v[k];
        __syncthreads();
        for (int o = RB / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const double a = sh[threadIdx.x], b = sh[threadIdx.x + o];
                sh[threadIdx.x] = (k >= NV - NMAX) ? fmax(a, b) : a + b;
            }
            __syncthreads();
        }
        v[k] = sh[0];
        __syncthreads();
    }
}

// partial[block][NV] -> out[NV]; the last NMAX values are maxima
template <int NV, int NMAX>
__global__ void __launch_bounds__(RB) final_reduce_kernel(const double* __restrict__ partial, int nblocks,
                                                          double* __restrict__ out) {
    __shared__ double sh[RB];
    double v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += RB) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const double x = partial[(size_t)b * NV + k];
            v[k]           = (k >= NV - NMAX) ? fmax(v[k], x) : v[k] + x;
        }
    }
    block_reduce<NV, NMAX>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) out[k] = v[k];
    }
}

__device__ __forceinline__ long interior_index(const MeshDev& m, long t) {
    const int i = (int)(t % m.nl[0]) + m.nghost;
    const int j = (int)((t / m.nl[0]) % m.nl[1]) + m.nghost;
    const int k = (int)(t / ((long)m.nl[0] * m.nl[1])) + m.nghost;
    return i + (long)m.ex * (j + (long)m.ey * k);
}

// [sum Ex^2, sum Ey^2, sum Ez^2, sum dot(E,E), max|Ex|, max|Ey|, max|Ez|] per block
__global__ void __launch_bounds__(RB)
energy_stats_kernel(MeshDev m, const double* __restrict__ ef, double* __restrict__ partial) {
    __shared__ double sh[RB];
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    double v[7]   = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni; t += (long)gridDim.x * blockDim.x) {
        const long c = interior_index(m, t) * 3;
        const double ex = ef[c], ey = ef[c + 1], ez = ef[c + 2];
        const double xx = __dmul_rn(ex, ex), yy = __dmul_rn(ey, ey), zz = __dmul_rn(ez, ez);
        v[0] += xx;
        v[1] += yy;
        v[2] += zz;
        v[3] += __dadd_rn(__dadd_rn(xx, yy), zz);  // dot(E, E): left fold over the components
        v[4] = fmax(v[4], fabs(ex));
        v[5] = fmax(v[5], fabs(ey));
        v[6] = fmax(v[6], fabs(ez));
    }
    block_reduce<7, 3>(v, sh);
    if (threadIdx.x == 0)
        for (int k = 0; k < 7; ++k) partial[(size_t)blockIdx.x * 7 + k] = v[k];
}

__global__ void __launch_bounds__(RB)
norm_stats_kernel(MeshDev m, const double* __restrict__ f, double* __restrict__ partial) {
    __shared__ double sh[RB];
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    double v[2]   = {0.0, 0.0};
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni; t += (long)gridDim.x * blockDim.x) {
        const double x = f[interior_index(m, t)];
        v[0] += __dmul_rn(x, x);
        v[1] = fmax(v[1], fabs(x));
    }
    block_reduce<2, 1>(v, sh);
    if (threadIdx.x == 0) {
        partial[(size_t)blockIdx.x * 2]     = v[0];
        partial[(size_t)blockIdx.x * 2 + 1] = v[1];
    }
}

__global__ void __launch_bounds__(RB)
kinetic_kernel(long n, const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
               double* __restrict__ partial) {
    __shared__ double sh[RB];
    double v[1] = {0.0};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const double a = px[i], b = py[i], c = pz[i];
        v[0] += __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c));
    }
    block_reduce<1, 0>(v, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = v[0];
}

// the same over buckets (one block walks tiles blockIdx, blockIdx + grid, ...) and the tail (item ntiles)
__global__ void __launch_bounds__(RB)
kinetic_bins_kernel(int ntiles, const int* __restrict__ start, const int* __restrict__ count,
                    const int* __restrict__ state, const double* __restrict__ px, const double* __restrict__ py,
                    const double* __restrict__ pz, double* __restrict__ partial) {
    __shared__ double sh[RB];
    double v[1] = {0.0};
    for (int t = blockIdx.x; t <= ntiles; t += gridDim.x) {
        const long b = t < ntiles ? start[t] : state[BS_TAIL_START];
        const long e = b + (t < ntiles ? count[t] : state[BS_TAIL_COUNT]);
        for (long i = b + threadIdx.x; i < e; i += RB) {
            const double a = px[i], bb = py[i], c = pz[i];
            v[0] += __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(bb, bb)), __dmul_rn(c, c));
        }
    }
    block_reduce<1, 0>(v, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = v[0];
}

// one block per plane of the (clipped) domain along `axis`: sum of the plane's cells, block-reduced
__global__ void __launch_bounds__(RB)
plane_sum_kernel(MeshDev m, const double* __restrict__ f, int axis, int a0, int lo_u, int n_u, int lo_v, int n_v,
                 double* __restrict__ out) {
    __shared__ double sh[RB];
    // plane blockIdx.x: local ghosted coordinate a0 + blockIdx.x along axis; (u, v) = the other two axes in order
    const int u_ax = axis == 0 ? 1 : 0, v_ax = axis == 2 ? 1 : 2;
    const long cnt = (long)n_u * n_v;
    double v[1]    = {0.0};
    for (long t = threadIdx.x; t < cnt; t += RB) {
        int c[3];
        c[axis] = a0 + blockIdx.x;
        c[u_ax] = lo_u + (int)(t % n_u);
        c[v_ax] = lo_v + (int)(t / n_u);
        v[0] += f[c[0] + (long)m.ex * (c[1] + (long)m.ey * c[2])];
    }
    block_reduce<1, 0>(v, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = v[0];
}

static int reduce_grid(ipplb_ctx* ctx, long items) {
    const long want = (items + RB - 1) / RB;
    const long cap  = (long)ctx->num_sms * 8;
    return (int)std::max<long>(1, std::min(want, cap));
}

// runs `partial` (grid blocks x NV) -> final -> host
template <int NV, int NMAX>
static int finish_reduce(ipplb_ctx* ctx, double* partial, int grid, double* out_host) {
    double* d_out = partial + (size_t)grid * NV;
    final_reduce_kernel<NV, NMAX><<<1, RB, 0, ctx->stream>>>(partial, grid, d_out);
    IPPLB_CHECK_LAUNCH(ctx);
    IPPLB_CUDA(cudaMemcpyAsync(ctx->reduce_host + 32, d_out, sizeof(double) * NV, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < NV; ++k) out_host[k] = ctx->reduce_host[32 + k];
    return IPPLB_OK;
}

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_field_energy_stats(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* efield, double out_host[7]) {
    IPPLB_REQUIRE(ctx && mesh && efield && out_host, "field_energy_stats: bad arguments");
    const MeshDev m = make_mesh_dev(mesh);
    const long ni   = (long)m.nl[0] * m.nl[1] * m.nl[2];
    const int grid  = reduce_grid(ctx, ni);
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * ((size_t)grid * 7 + 16)))) return rc;
    double* partial = (double*)ctx->reduce.ptr;
    energy_stats_kernel<<<grid, RB, 0, ctx->stream>>>(m, efield, partial);
    IPPLB_CHECK_LAUNCH(ctx);
    double r[7];
    if ((rc = finish_reduce<7, 3>(ctx, partial, grid, r))) return rc;
    out_host[0] = r[0]; out_host[1] = r[1]; out_host[2] = r[2];
    out_host[3] = r[4]; out_host[4] = r[5]; out_host[5] = r[6];
    out_host[6] = r[3];
    return IPPLB_OK;
}

int ipplb_field_norm_stats(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* field, double out_host[2]) {
    IPPLB_REQUIRE(ctx && mesh && field && out_host, "field_norm_stats: bad arguments");
    const MeshDev m = make_mesh_dev(mesh);
    const long ni   = (long)m.nl[0] * m.nl[1] * m.nl[2];
    const int grid  = reduce_grid(ctx, ni);
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * ((size_t)grid * 2 + 16)))) return rc;
    double* partial = (double*)ctx->reduce.ptr;
    norm_stats_kernel<<<grid, RB, 0, ctx->stream>>>(m, field, partial);
    IPPLB_CHECK_LAUNCH(ctx);
    return finish_reduce<2, 1>(ctx, partial, grid, out_host);
}

int ipplb_particles_kinetic(ipplb_ctx* ctx, long n, const double* px, const double* py, const double* pz,
                            double* out_host) {
    IPPLB_REQUIRE(ctx && out_host && n >= 0 && (n == 0 || (px && py && pz)), "particles_kinetic: bad arguments");
    const int grid = reduce_grid(ctx, std::max<long>(n, 1));
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * ((size_t)grid + 16)))) return rc;
    double* partial = (double*)ctx->reduce.ptr;
    kinetic_kernel<<<grid, RB, 0, ctx->stream>>>(n, px, py, pz, partial);
    IPPLB_CHECK_LAUNCH(ctx);
    return finish_reduce<1, 0>(ctx, partial, grid, out_host);
}

int ipplb_bins_kinetic(ipplb_ctx* ctx, ipplb_bins* b, const ipplb_particles* cur, double* out_host) {
    IPPLB_REQUIRE(ctx && b && cur && out_host && b->built, "bins_kinetic: bad arguments");
    const int grid = std::min(b->ntiles + 1, ctx->num_sms * 8);
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * ((size_t)grid + 16)))) return rc;
    double* partial = (double*)ctx->reduce.ptr;
    kinetic_bins_kernel<<<grid, RB, 0, ctx->stream>>>(b->ntiles, b->start(b->cur), b->count(b->cur), b->state(b->cur),
                                                      cur->px, cur->py, cur->pz, partial);
    IPPLB_CHECK_LAUNCH(ctx);
    return finish_reduce<1, 0>(ctx, partial, grid, out_host);
}


int ipplb_orb_plane_sums(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* field, int axis, const int dom_lo[3],
                         const int dom_hi[3], double* out_host) {
    IPPLB_REQUIRE(ctx && mesh && field && dom_lo && dom_hi && out_host && axis >= 0 && axis < 3,
                  "orb_plane_sums: bad arguments");
    const MeshDev m = make_mesh_dev(mesh);
    const int len   = dom_hi[axis] - dom_lo[axis] + 1;
    IPPLB_REQUIRE(len >= 1 && len <= 1 << 20, "orb_plane_sums: bad domain");
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * ((size_t)len + 16)))) return rc;
    double* d_out = (double*)ctx->reduce.ptr;
    IPPLB_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * len, ctx->stream));
    // clip the domain to the local box (perpendicularReduction :121-160); empty intersection -> all zeros
    int lo[3], n[3];
    bool empty = false;
    for (int d = 0; d < 3; ++d) {
        const int first = m.first[d], last = m.first[d] + m.nl[d] - 1;
        const int inf = std::max(first, dom_lo[d]), sup = std::min(last, dom_hi[d]);
        if (sup < inf) empty = true;
        lo[d] = inf - first + m.nghost;
        n[d]  = sup - inf + 1;
    }
    if (!empty) {
        const int array_start = std::max(0, m.first[axis] - dom_lo[axis]);
        const int u_ax = axis == 0 ? 1 : 0, v_ax = axis == 2 ? 1 : 2;
        plane_sum_kernel<<<n[axis], RB, 0, ctx->stream>>>(m, field, axis, lo[axis], lo[u_ax], n[u_ax], lo[v_ax], n[v_ax],
                                                          d_out + array_start);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    if (ctx->nranks > 1) {
        if (ncclAllReduce(d_out, d_out, len, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream) != ncclSuccess) {
            set_error("orb_plane_sums: ncclAllReduce failed");
            return IPPLB_ERR_NCCL;
        }
        ctx->launches++;
    }
    IPPLB_CUDA(cudaMemcpyAsync(out_host, d_out, sizeof(double) * len, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return IPPLB_OK;
}

int ipplb_orb_repartition(ipplb_ctx* ctx, const ipplb_mesh* mesh, int nranks, const double* weight_field,
                          int* boxes_out, int* ok) {
    IPPLB_REQUIRE(ctx && mesh && weight_field && boxes_out && ok, "orb_repartition: bad arguments");
    IPPLB_REQUIRE(nranks == ctx->nranks || ctx->nranks == 1, "orb_repartition: nranks differs from the communicator");
    ipplb_orb* o = nullptr;
    int rc;
    if ((rc = ipplb_orb_begin(&o, mesh->ng, nranks))) return rc;
    std::vector<double> w;
    for (;;) {
        int lo[3], hi[3], axis, pending;
        if ((rc = ipplb_orb_next(o, lo, hi, &axis, &pending))) break;
        if (!pending) break;
        w.assign((size_t)(hi[axis] - lo[axis] + 1), 0.0);
        if ((rc = ipplb_orb_plane_sums(ctx, mesh, weight_field, axis, lo, hi, w.data()))) break;
        if ((rc = ipplb_orb_cut(o, w.data(), (int)w.size()))) break;
    }
    if (rc) {
        ipplb_orb_destroy(o);
        return rc;
    }
    return ipplb_orb_finish(o, boxes_out, ok);
}

}  // extern "C"
