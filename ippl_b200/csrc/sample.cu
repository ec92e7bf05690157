// sample.cu -- particle initialisation on the device: inverse-transform sampling of the alpine
// position distributions (src/Random/InverseTransformSampling.h, Distribution.h, Utility.h NewtonRaphson) and the
// Gaussian velocity sampler (src/Random/Randn.h), on a counter-based uniform stream (Philox4x32-10) so that every
// particle's numbers depend only on (seed, particle id, dimension): reproducible on any decomposition and on the
// host (the oracle restates the same generator), which Kokkos::Random_XorShift64_Pool is not.
#include <cmath>

#include "common.cuh"
#include "philox.h"

namespace ipplb {

struct DistDev {
    int kind[3];
    double par[6];
};

__host__ __device__ inline double dist_cdf(const DistDev& D, int d, double x) {
    switch (D.kind[d]) {
        case IPPLB_DIST_COSINE: return x + (D.par[2 * d] / D.par[2 * d + 1]) * sin(D.par[2 * d + 1] * x);
        case IPPLB_DIST_NORMAL: return 0.5 * (1 + erf((x - D.par[2 * d]) / (D.par[2 * d + 1] * sqrt(2.0))));
        default: return x;
    }
}
__host__ __device__ inline double dist_pdf(const DistDev& D, int d, double x) {
    switch (D.kind[d]) {
        case IPPLB_DIST_COSINE: return 1.0 + D.par[2 * d] * cos(D.par[2 * d + 1] * x);
        case IPPLB_DIST_NORMAL: {
            const double pi = 3.14159265358979323846, mean = D.par[2 * d], sd = D.par[2 * d + 1];
            // src/Random/NormalDistribution.h:31-36
            return (1.0 / (sd * sqrt(2 * pi))) * exp(-(x - mean) * (x - mean) / (2 * sd * sd));
        }
        default: return 1.0;
    }
}
__host__ __device__ inline double dist_estimate(const DistDev& D, int d, double u) {
    return D.kind[d] == IPPLB_DIST_NORMAL ? D.par[2 * d] + 0. * u * D.par[2 * d + 1] : u + D.par[d] * 0.;
}

static DistDev make_dist(const ipplb_dist* d) {
    DistDev D;
    for (int k = 0; k < 3; ++k) D.kind[k] = d->kind[k];
    for (int k = 0; k < 6; ++k) D.par[k] = d->par[k];
    return D;
}

struct Bounds3 {
    double lo[3], hi[3];
};

__global__ void __launch_bounds__(256)
sample_positions_kernel(DistDev D, Bounds3 U, unsigned long long seed, long first_id, long n, double* __restrict__ x,
                        double* __restrict__ y, double* __restrict__ z) {
    double* out[3] = {x, y, z};
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            // InverseTransformSampling::fill_random::operator(), InverseTransformSampling.h:199-215
            const double u01 = philox_uniform(seed, (unsigned long long)(first_id + i), (unsigned)d, 0);
            double u         = U.lo[d] + (U.hi[d] - U.lo[d]) * u01;  // rand_gen.drand(umin, umax)
            double s         = dist_estimate(D, d, u);
            // NewtonRaphson::solve, src/Random/Utility.h:52-59 (atol 1e-12, max_iter 20)
            unsigned iter = 0;
            while (iter < 20u && fabs(dist_cdf(D, d, s) - u) > 1e-12) {
                s = s - ((dist_cdf(D, d, s) - u) / dist_pdf(D, d, s));
                iter += 1;
            }
            out[d][i] = s;
        }
    }
}

__global__ void __launch_bounds__(256)
sample_normal_kernel(Bounds3 MS, unsigned long long seed, long first_id, long n, double* __restrict__ px,
                     double* __restrict__ py, double* __restrict__ pz) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        double g[3];
        philox_normal3(seed, (unsigned long long)(first_id + i), g);
        // randn::operator(), src/Random/Randn.h:82-94: v(i)[d] = mu[d] + sd[d] * normal(0, 1)
        px[i] = MS.lo[0] + MS.hi[0] * g[0];
        py[i] = MS.lo[1] + MS.hi[1] * g[1];
        pz[i] = MS.lo[2] + MS.hi[2] * g[2];
    }
}

__global__ void __launch_bounds__(256) fill_pdf_kernel(MeshDev m, DistDev D, Bounds3 H, double* __restrict__ f) {
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni; t += (long)gridDim.x * blockDim.x) {
        const int c[3] = {(int)(t % m.nl[0]), (int)((t / m.nl[0]) % m.nl[1]), (int)(t / ((long)m.nl[0] * m.nl[1]))};
        // xvec = (args + lDom.first() - nghost + 0.5) * hr + origin, LandauDampingManager.h:193-194;
        // getFullPdf = product over the dimensions starting from 1.0, src/Random/Distribution.h:104-112
        double total = 1.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double xv = ((double)(c[d] + m.first[d]) + 0.5) * H.hi[d] + H.lo[d];
            total *= dist_pdf(D, d, xv);
        }
        f[(c[0] + m.nghost) + (long)m.ex * ((c[1] + m.nghost) + (long)m.ey * (c[2] + m.nghost))] = total;
    }
}

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_sample_counts(const ipplb_dist* dist, const double rmin[3], const double rmax[3], const double* regions,
                        int nranks, long ntotal, long* nlocal_out, double* ubounds_out) {
    IPPLB_REQUIRE(dist && rmin && rmax && regions && nranks >= 1 && ntotal >= 0 && nlocal_out,
                  "sample_counts: bad arguments");
    const DistDev D = make_dist(dist);
    unsigned long nglobal = 0;
    for (int r = 0; r < nranks; ++r) {
        // updateBounds(rmax, rmin, locrmax, locrmin), InverseTransformSampling.h:106-131
        double pnr = 1.0, pdr = 1.0;
        for (int d = 0; d < 3; ++d) {
            const double lmin = regions[r * 6 + d], lmax = regions[r * 6 + 3 + d];
            const double nr = dist_cdf(D, d, lmax) - dist_cdf(D, d, lmin);
            const double dr = dist_cdf(D, d, rmax[d]) - dist_cdf(D, d, rmin[d]);
            pnr *= nr;  // std::accumulate(..., 1.0, multiplies): ((1 * a) * b) * c
            pdr *= dr;
            if (ubounds_out) {
                ubounds_out[r * 6 + d]     = dist_cdf(D, d, lmin);
                ubounds_out[r * 6 + 3 + d] = dist_cdf(D, d, lmax);
            }
        }
        const double factor = pnr / pdr;
        nlocal_out[r]       = (long)(unsigned long)(factor * ntotal);
        nglobal += (unsigned long)nlocal_out[r];
    }
    const int rest = (int)((unsigned long)ntotal - nglobal);
    for (int r = 0; r < nranks; ++r)
        if (r < rest) ++nlocal_out[r];
    return IPPLB_OK;
}

int ipplb_sample_positions(ipplb_ctx* ctx, const ipplb_dist* dist, const double umin[3], const double umax[3],
                           uint64_t seed, long first_id, long n, double* x, double* y, double* z) {
    IPPLB_REQUIRE(ctx && dist && umin && umax && n >= 0 && (n == 0 || (x && y && z)), "sample_positions: bad arguments");
    if (n == 0) return IPPLB_OK;
    Bounds3 U;
    for (int d = 0; d < 3; ++d) {
        U.lo[d] = umin[d];
        U.hi[d] = umax[d];
    }
    const int grid = (int)std::min<long>((n + 255) / 256, (long)ctx->num_sms * 16);
    sample_positions_kernel<<<grid, 256, 0, ctx->stream>>>(make_dist(dist), U, seed, first_id, n, x, y, z);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_sample_normal(ipplb_ctx* ctx, const double mu[3], const double sd[3], uint64_t seed, long first_id,
                        long n, double* px, double* py, double* pz) {
    IPPLB_REQUIRE(ctx && mu && sd && n >= 0 && (n == 0 || (px && py && pz)), "sample_normal: bad arguments");
    if (n == 0) return IPPLB_OK;
    Bounds3 MS;
    for (int d = 0; d < 3; ++d) {
        MS.lo[d] = mu[d];
        MS.hi[d] = sd[d];
    }
    const int grid = (int)std::min<long>((n + 255) / 256, (long)ctx->num_sms * 16);
    sample_normal_kernel<<<grid, 256, 0, ctx->stream>>>(MS, seed, first_id, n, px, py, pz);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_field_fill_pdf(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_dist* dist, double* field) {
    IPPLB_REQUIRE(ctx && mesh && dist && field, "field_fill_pdf: bad arguments");
    const MeshDev m = make_mesh_dev(mesh);
    Bounds3 H;
    for (int d = 0; d < 3; ++d) {
        H.lo[d] = mesh->origin[d];
        H.hi[d] = mesh->h[d];
    }
    const long ni  = (long)m.nl[0] * m.nl[1] * m.nl[2];
    const int grid = (int)std::min<long>((ni + 255) / 256, (long)ctx->num_sms * 16);
    fill_pdf_kernel<<<grid, 256, 0, ctx->stream>>>(m, make_dist(dist), H, field);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

}  // extern "C"
