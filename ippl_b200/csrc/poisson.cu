// poisson.cu -- periodic FFT Poisson solve on cuFFT (NON-OWNED stage, timed separately).
// Restates FFTPeriodicPoissonSolver::solve, GRAD output
// (src/PoissonSolvers/FFTPeriodicPoissonSolver.hpp:53-169): E = irfft( rho_hat/N * -(i k_gd / |k|^2) ).
#include <cufft.h>
#include <nccl.h>

#include <cmath>
#include <vector>

#include "common.cuh"
#include "layout.h"
#include "poisson.h"

namespace ipplb {

__global__ void pack_interior_kernel(MeshDev m, const double* __restrict__ f, double* __restrict__ o) {
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni;
         t += (long)gridDim.x * blockDim.x) {
        int i = (int)(t % m.nl[0]) + m.nghost;
        int j = (int)((t / m.nl[0]) % m.nl[1]) + m.nghost;
        int k = (int)(t / ((long)m.nl[0] * m.nl[1])) + m.nghost;
        o[t]  = f[i + (long)m.ex * (j + (long)m.ey * k)];
    }
}

// one pass produces the three gradient spectra: spec_g = (rho_hat / N) * -(i * k_g * factor)
__global__ void kspace_kernel(int nxh, int ny, int nz, double inv_n, const double* __restrict__ kx,
                              const double* __restrict__ ky, const double* __restrict__ kz,
                              const cufftDoubleComplex* __restrict__ rh,
                              cufftDoubleComplex* __restrict__ g0, cufftDoubleComplex* __restrict__ g1,
                              cufftDoubleComplex* __restrict__ g2) {
    const long nh = (long)nxh * ny * nz;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nh;
         t += (long)gridDim.x * blockDim.x) {
        int i = (int)(t % nxh), j = (int)((t / nxh) % ny), k = (int)(t / ((long)nxh * ny));
        const double k0 = kx[i], k1 = ky[j], k2 = kz[k];
        double Dr = 0;
        Dr += k0 * k0;
        Dr += k1 * k1;
        Dr += k2 * k2;
        const bool nzr      = (Dr != 0.0);
        const double factor = nzr ? (1.0 / Dr) : 0.0;
        const double a = rh[t].x * inv_n, b = rh[t].y * inv_n;
        // (a + b i) * (0 - c i) = b c - a c i
        double c = k0 * factor;
        g0[t]    = make_cuDoubleComplex(b * c, -(a * c));
        c        = k1 * factor;
        g1[t]    = make_cuDoubleComplex(b * c, -(a * c));
        c        = k2 * factor;
        g2[t]    = make_cuDoubleComplex(b * c, -(a * c));
    }
}

__global__ void unpack_e_kernel(MeshDev m, const double* __restrict__ r, double* __restrict__ ef) {
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni;
         t += (long)gridDim.x * blockDim.x) {
        int i = (int)(t % m.nl[0]) + m.nghost;
        int j = (int)((t / m.nl[0]) % m.nl[1]) + m.nghost;
        int k = (int)(t / ((long)m.nl[0] * m.nl[1])) + m.nghost;
        long l = (i + (long)m.ex * (j + (long)m.ey * k)) * 3;
        ef[l]     = r[t];
        ef[l + 1] = r[ni + t];
        ef[l + 2] = r[2 * ni + t];
    }
}

// like the reference, the inverse transform lands in rho's storage (:153); keep that side effect:
// rho interior <- last gradient component
__global__ void clobber_rho_kernel(MeshDev m, const double* __restrict__ r, double* __restrict__ f) {
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni;
         t += (long)gridDim.x * blockDim.x) {
        int i = (int)(t % m.nl[0]) + m.nghost;
        int j = (int)((t / m.nl[0]) % m.nl[1]) + m.nghost;
        int k = (int)(t / ((long)m.nl[0] * m.nl[1])) + m.nghost;
        f[i + (long)m.ex * (j + (long)m.ey * k)] = r[2 * ni + t];
    }
}

// stage (boxes back to back, x fastest inside a box) -> the global real array
__global__ void assemble_boxes_kernel(int nranks, const int* __restrict__ boxes, const long* __restrict__ off, int gx,
                                      int gy, const double* __restrict__ stage, double* __restrict__ real) {
    const long n = off[nranks];
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        int r = 0;
        while (t >= off[r + 1]) ++r;
        const int* b  = boxes + 6 * r;
        const int bx = b[3] - b[0] + 1, by = b[4] - b[1] + 1;
        const long l = t - off[r];
        const int i = (int)(l % bx) + b[0], j = (int)((l / bx) % by) + b[1], k = (int)(l / ((long)bx * by)) + b[2];
        real[i + (long)gx * (j + (long)gy * k)] = stage[t];
    }
}

// this rank's box of the three global gradient planes -> ghosted AoS-3 field interior (and rho <- last plane)
__global__ void unpack_box_kernel(MeshDev m, int gx, int gy, long N, const double* __restrict__ r, double* __restrict__ ef,
                                  double* __restrict__ rho) {
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni; t += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(t % m.nl[0]), cj = (int)((t / m.nl[0]) % m.nl[1]), ck = (int)(t / ((long)m.nl[0] * m.nl[1]));
        const long gidx = (ci + m.first[0]) + (long)gx * ((cj + m.first[1]) + (long)gy * (ck + m.first[2]));
        const long c    = (ci + m.nghost) + (long)m.ex * ((cj + m.nghost) + (long)m.ey * (ck + m.nghost));
        ef[3 * c]     = r[gidx];
        ef[3 * c + 1] = r[N + gidx];
        ef[3 * c + 2] = r[2 * N + gidx];
        rho[c]        = r[2 * N + gidx];
    }
}

// kVec[d] = notMid * 2 * pi / Len * (iVec[d] - shift * N[d]), Len = rmax - origin,
// rmax = origin + N*h  (FFTPeriodicPoissonSolver.hpp:66-70, 127-137)
std::vector<double> poisson_k_tables(const int ng[3], const double origin[3], const double h[3]) {
    const int cnt[3] = {ng[0] / 2 + 1, ng[1], ng[2]};
    std::vector<double> kh((size_t)cnt[0] + cnt[1] + cnt[2]);
    const double pi = M_PI;
    int off = 0;
    for (int d = 0; d < 3; ++d) {
        const int Nd      = ng[d];
        const double rmax = origin[d] + (Nd * h[d]);
        const double Len  = rmax - origin[d];
        for (int i = 0; i < cnt[d]; ++i) {
            bool shift  = (i > (Nd / 2));
            bool notMid = (i != (Nd / 2));
            kh[off + i] = notMid * 2 * pi / Len * (i - shift * Nd);
        }
        off += cnt[d];
    }
    return kh;
}

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_poisson_create(ipplb_ctx* ctx, const ipplb_mesh* mesh, ipplb_poisson** out) {
    IPPLB_REQUIRE(ctx && mesh && out, "poisson_create: bad arguments");
    for (int d = 0; d < 3; ++d)
        IPPLB_REQUIRE(mesh->nl[d] == mesh->ng[d] && mesh->first[d] == 0,
                      "poisson_create: single-GPU solver needs the whole domain on this rank (use ipplb_poisson_create_dist)");
    ipplb_poisson* s = new ipplb_poisson();
    s->ctx = ctx;
    s->m   = make_mesh_dev(mesh);
    s->nx = mesh->ng[0]; s->ny = mesh->ng[1]; s->nz = mesh->ng[2];
    s->nxh = s->nx / 2 + 1;
    const long N = (long)s->nx * s->ny * s->nz, Nh = (long)s->nxh * s->ny * s->nz;
    IPPLB_CUDA(cudaMalloc(&s->real, sizeof(double) * 3 * N));
    IPPLB_CUDA(cudaMalloc(&s->spec, sizeof(cufftDoubleComplex) * 4 * Nh));
    IPPLB_CUDA(cudaMalloc(&s->kx, sizeof(double) * (s->nxh + s->ny + s->nz)));
    s->ky = s->kx + s->nxh;
    s->kz = s->ky + s->ny;
    const std::vector<double> kh = poisson_k_tables(mesh->ng, mesh->origin, mesh->h);
    IPPLB_CUDA(cudaMemcpy(s->kx, kh.data(), sizeof(double) * kh.size(), cudaMemcpyHostToDevice));
    int dims[3] = {s->nz, s->ny, s->nx};
    IPPLB_CUFFT(cufftPlan3d(&s->fwd, s->nz, s->ny, s->nx, CUFFT_D2Z));
    IPPLB_CUFFT(cufftSetStream(s->fwd, ctx->stream));
    IPPLB_CUFFT(cufftPlanMany(&s->inv, 3, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 3));
    IPPLB_CUFFT(cufftSetStream(s->inv, ctx->stream));
    s->plans = true;
    *out = s;
    return IPPLB_OK;
}

int ipplb_poisson_create_dist(ipplb_ctx* ctx, const ipplb_layout* layout, const double origin[3], const double h[3],
                              ipplb_poisson** out) {
    IPPLB_REQUIRE(ctx && layout && origin && h && out, "poisson_create_dist: bad arguments");
    const int nr = (int)layout->L.boxes.size();
    IPPLB_REQUIRE(nr == ctx->nranks, "poisson_create_dist: layout rank count != communicator size");
    ipplb_mesh whole{};
    for (int d = 0; d < 3; ++d) {
        whole.ng[d] = whole.nl[d] = layout->L.ng[d];
        whole.first[d]            = 0;
        whole.origin[d]           = origin[d];
        whole.h[d]                = h[d];
    }
    whole.nghost = layout->L.nghost;
    int rc;
    if ((rc = ipplb_poisson_create(ctx, &whole, out))) return rc;
    if (nr == 1) return IPPLB_OK;
    IPPLB_REQUIRE(ctx->nccl, "poisson_create_dist: communicator not initialised (ipplb_comm_init)");
    ipplb_poisson* s = *out;
    s->dist   = true;
    s->nranks = nr;
    s->g      = s->m;
    ipplb_mesh mine;
    if ((rc = ipplb_layout_mesh(layout, ctx->rank, origin, h, &mine))) return rc;
    s->m = make_mesh_dev(&mine);
    std::vector<int> boxes(6 * (size_t)nr);
    s->off.assign(nr + 1, 0);
    for (int r = 0; r < nr; ++r) {
        const IBox& b = layout->L.boxes[r];
        long n = 1;
        for (int d = 0; d < 3; ++d) {
            boxes[6 * r + d]     = b.lo[d];
            boxes[6 * r + 3 + d] = b.hi[d];
            n *= b.hi[d] - b.lo[d] + 1;
        }
        s->off[r + 1] = s->off[r] + n;
    }
    const long N = (long)s->nx * s->ny * s->nz;
    IPPLB_REQUIRE(s->off[nr] == N, "poisson_create_dist: the rank boxes do not tile the domain");
    IPPLB_CUDA(cudaMalloc(&s->stage, sizeof(double) * N));
    IPPLB_CUDA(cudaMalloc(&s->d_boxes, sizeof(int) * boxes.size()));
    IPPLB_CUDA(cudaMalloc(&s->d_off, sizeof(long) * s->off.size()));
    IPPLB_CUDA(cudaMemcpy(s->d_boxes, boxes.data(), sizeof(int) * boxes.size(), cudaMemcpyHostToDevice));
    IPPLB_CUDA(cudaMemcpy(s->d_off, s->off.data(), sizeof(long) * s->off.size(), cudaMemcpyHostToDevice));
    return IPPLB_OK;
}

int ipplb_poisson_solve(ipplb_poisson* s, double* rho, double* efield) {
    IPPLB_REQUIRE(s && rho && efield, "poisson_solve: bad arguments");
    if (s->slab) return slab_solve(s, rho, efield);
    ipplb_ctx* ctx = s->ctx;
    const long N = (long)s->nx * s->ny * s->nz, Nh = (long)s->nxh * s->ny * s->nz;
    const int g  = (int)((N + 255) / 256 < 148 * 16 ? (N + 255) / 256 : 148 * 16);
    if (s->dist) {
        // every rank's rho interior to every rank: one grouped broadcast per box (boxes may differ in size: ORB)
        const int me = ctx->rank;
        pack_interior_kernel<<<g, 256, 0, ctx->stream>>>(s->m, rho, s->stage + s->off[me]);
        IPPLB_CHECK_LAUNCH(ctx);
        bool ok = ncclGroupStart() == ncclSuccess;
        for (int r = 0; ok && r < s->nranks; ++r)
            ok = ncclBroadcast(s->stage + s->off[r], s->stage + s->off[r], (size_t)(s->off[r + 1] - s->off[r]), ncclDouble, r,
                               (ncclComm_t)ctx->nccl, ctx->stream) == ncclSuccess;
        ok = (ncclGroupEnd() == ncclSuccess) && ok;
        if (!ok) {
            set_error("poisson_solve: NCCL broadcast of the rho boxes failed");
            return IPPLB_ERR_NCCL;
        }
        ctx->launches++;
        assemble_boxes_kernel<<<g, 256, 0, ctx->stream>>>(s->nranks, s->d_boxes, s->d_off, s->nx, s->ny, s->stage, s->real);
        IPPLB_CHECK_LAUNCH(ctx);
    } else {
        pack_interior_kernel<<<g, 256, 0, ctx->stream>>>(s->m, rho, s->real);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    IPPLB_CUFFT(cufftExecD2Z(s->fwd, s->real, s->spec));
    kspace_kernel<<<g, 256, 0, ctx->stream>>>(s->nxh, s->ny, s->nz, 1.0 / (double)N, s->kx, s->ky,
                                              s->kz, s->spec, s->spec + Nh, s->spec + 2 * Nh,
                                              s->spec + 3 * Nh);
    IPPLB_CHECK_LAUNCH(ctx);
    IPPLB_CUFFT(cufftExecZ2D(s->inv, s->spec + Nh, s->real));
    ctx->launches += 2;
    if (s->dist) {
        unpack_box_kernel<<<g, 256, 0, ctx->stream>>>(s->m, s->nx, s->ny, N, s->real, efield, rho);
        IPPLB_CHECK_LAUNCH(ctx);
        return IPPLB_OK;
    }
    unpack_e_kernel<<<g, 256, 0, ctx->stream>>>(s->m, s->real, efield);
    IPPLB_CHECK_LAUNCH(ctx);
    clobber_rho_kernel<<<g, 256, 0, ctx->stream>>>(s->m, s->real, rho);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_poisson_destroy(ipplb_poisson* s) {
    if (!s) return IPPLB_OK;
    if (s->slab) slab_free(s->slab);
    if (s->plans) {
        cufftDestroy(s->fwd);
        cufftDestroy(s->inv);
    }
    cudaFree(s->real);
    cudaFree(s->spec);
    cudaFree(s->kx);
    cudaFree(s->stage);
    cudaFree(s->d_boxes);
    cudaFree(s->d_off);
    delete s;
    return IPPLB_OK;
}

}  // extern "C"
