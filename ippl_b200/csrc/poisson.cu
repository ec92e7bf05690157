// poisson.cu -- periodic FFT Poisson solve on cuFFT (NON-OWNED stage, timed separately).
// Restates FFTPeriodicPoissonSolver::solve, GRAD output
// (src/PoissonSolvers/FFTPeriodicPoissonSolver.hpp:53-169): E = irfft( rho_hat/N * -(i k_gd / |k|^2) ).
#include <cufft.h>

#include <cmath>
#include <vector>

#include "common.cuh"

struct ipplb_poisson {
    ipplb_ctx* ctx = nullptr;
    ipplb::MeshDev m;
    int nx = 0, ny = 0, nz = 0, nxh = 0;
    cufftHandle fwd = 0, inv = 0;
    double* real     = nullptr;          // 3 * N (component planes after the inverse)
    cufftDoubleComplex* spec = nullptr;  // 4 * Nh: [0] rho_hat, [1..3] gradient spectra
    double* kx = nullptr;                // kx[nxh] ky[ny] kz[nz]
    double *ky = nullptr, *kz = nullptr;
};

namespace ipplb {

#define IPPLB_CUFFT(call)                                                          \
    do {                                                                           \
        cufftResult r__ = (call);                                                  \
        if (r__ != CUFFT_SUCCESS) {                                                \
            set_error("%s:%d: %s -> cufft error %d", __FILE__, __LINE__, #call, (int)r__); \
            return IPPLB_ERR_CUFFT;                                                \
        }                                                                          \
    } while (0)

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

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_poisson_create(ipplb_ctx* ctx, const ipplb_mesh* mesh, ipplb_poisson** out) {
    IPPLB_REQUIRE(ctx && mesh && out, "poisson_create: bad arguments");
    for (int d = 0; d < 3; ++d)
        IPPLB_REQUIRE(mesh->nl[d] == mesh->ng[d] && mesh->first[d] == 0,
                      "poisson_create: single-GPU solver needs the whole domain on this rank");
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
    // kVec[d] = notMid * 2 * pi / Len * (iVec[d] - shift * N[d]), Len = rmax - origin,
    // rmax = origin + N*h  (FFTPeriodicPoissonSolver.hpp:66-70, 127-137)
    std::vector<double> kh(s->nxh + s->ny + s->nz);
    const double pi = M_PI;
    int off = 0;
    const int cnt[3] = {s->nxh, s->ny, s->nz};
    for (int d = 0; d < 3; ++d) {
        const int Nd      = mesh->ng[d];
        const double rmax = mesh->origin[d] + (Nd * mesh->h[d]);
        const double Len  = rmax - mesh->origin[d];
        for (int i = 0; i < cnt[d]; ++i) {
            bool shift  = (i > (Nd / 2));
            bool notMid = (i != (Nd / 2));
            kh[off + i] = notMid * 2 * pi / Len * (i - shift * Nd);
        }
        off += cnt[d];
    }
    IPPLB_CUDA(cudaMemcpy(s->kx, kh.data(), sizeof(double) * kh.size(), cudaMemcpyHostToDevice));
    int dims[3] = {s->nz, s->ny, s->nx};
    IPPLB_CUFFT(cufftPlan3d(&s->fwd, s->nz, s->ny, s->nx, CUFFT_D2Z));
    IPPLB_CUFFT(cufftSetStream(s->fwd, ctx->stream));
    IPPLB_CUFFT(cufftPlanMany(&s->inv, 3, dims, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 3));
    IPPLB_CUFFT(cufftSetStream(s->inv, ctx->stream));
    *out = s;
    return IPPLB_OK;
}

int ipplb_poisson_solve(ipplb_poisson* s, double* rho, double* efield) {
    IPPLB_REQUIRE(s && rho && efield, "poisson_solve: bad arguments");
    ipplb_ctx* ctx = s->ctx;
    const long N = (long)s->nx * s->ny * s->nz, Nh = (long)s->nxh * s->ny * s->nz;
    const int g  = (int)((N + 255) / 256 < 148 * 16 ? (N + 255) / 256 : 148 * 16);
    pack_interior_kernel<<<g, 256, 0, ctx->stream>>>(s->m, rho, s->real);
    IPPLB_CHECK_LAUNCH(ctx);
    IPPLB_CUFFT(cufftExecD2Z(s->fwd, s->real, s->spec));
    kspace_kernel<<<g, 256, 0, ctx->stream>>>(s->nxh, s->ny, s->nz, 1.0 / (double)N, s->kx, s->ky,
                                              s->kz, s->spec, s->spec + Nh, s->spec + 2 * Nh,
                                              s->spec + 3 * Nh);
    IPPLB_CHECK_LAUNCH(ctx);
    IPPLB_CUFFT(cufftExecZ2D(s->inv, s->spec + Nh, s->real));
    ctx->launches += 2;
    unpack_e_kernel<<<g, 256, 0, ctx->stream>>>(s->m, s->real, efield);
    IPPLB_CHECK_LAUNCH(ctx);
    clobber_rho_kernel<<<g, 256, 0, ctx->stream>>>(s->m, s->real, rho);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_poisson_destroy(ipplb_poisson* s) {
    if (!s) return IPPLB_OK;
    cufftDestroy(s->fwd);
    cufftDestroy(s->inv);
    cudaFree(s->real);
    cudaFree(s->spec);
    cudaFree(s->kx);
    delete s;
    return IPPLB_OK;
}

}  // extern "C"
