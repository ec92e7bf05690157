// fftdist.cu -- the slab-decomposed periodic FFT Poisson solve (NON-OWNED stage): executes the host plan of slabplan.cpp.
// Replaces heFFTe's distributed r2c transform between the FieldLayout boxes (src/FFT/FFT.hpp:118-193,
// src/PoissonSolvers/FFTPeriodicPoissonSolver.hpp:53-169) with batched cuFFT transforms of whole planes / lines and four
// message exchanges; the k-space step is the arithmetic of poisson.cu's kspace_kernel on the y-slab layout.
#include <cufft.h>
#include <nccl.h>

#include <vector>

#include "common.cuh"
#include "layout.h"
#include "poisson.h"
#include "slabplan.h"
#include "transport.h"

namespace ipplb {

// [host-emulation begin: slab structs]  (tests/test_kernel_text_cpu.py compiles the marked text for the host, tests/emu/emu_slab.cpp)
struct CopyDev {
    long src_off, dst_off;
    long ss[3], ds[3];
    int n[3];
    int src_buf, dst_buf, elem;
};

struct BufTable {
    double* p[SB_COUNT];
};
// [host-emulation end: slab structs]

struct SlabState {
    SlabPlan plan;
    double* buf[SB_COUNT] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // REAL .. RECV owned
    CopyDev* d_copies = nullptr;      // all lists back to back
    int first[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}}, count[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};  // [phase][pre / post]
    long biggest[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}};                                                 // elements of the largest copy
    cufftHandle fwd2d = 0, inv2d = 0, z1d = 0;
    bool have2d = false, have1d = false;
};

// [host-emulation begin: slab kernels]
// a list of strided 3-D sub-box copies: blockIdx.y picks the copy, the x dimension strides over its elements
__global__ void __launch_bounds__(256) slab_copy_kernel(const CopyDev* __restrict__ list, const BufTable B) {
    const CopyDev c = list[blockIdx.y];
    const long n0 = c.n[0], n01 = (long)c.n[0] * c.n[1], total = n01 * c.n[2];
    if (c.elem == 2) {
        const double2* src = reinterpret_cast<const double2*>(B.p[c.src_buf]) + c.src_off;
        double2* dst       = reinterpret_cast<double2*>(B.p[c.dst_buf]) + c.dst_off;
        for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
            const long i = t % n0, j = (t / n0) % c.n[1], k = t / n01;
            dst[i * c.ds[0] + j * c.ds[1] + k * c.ds[2]] = src[i * c.ss[0] + j * c.ss[1] + k * c.ss[2]];
        }
    } else {
        const double* src = B.p[c.src_buf] + c.src_off;
        double* dst       = B.p[c.dst_buf] + c.dst_off;
        for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
            const long i = t % n0, j = (t / n0) % c.n[1], k = t / n01;
            dst[i * c.ds[0] + j * c.ds[1] + k * c.ds[2]] = src[i * c.ss[0] + j * c.ss[1] + k * c.ss[2]];
        }
    }
}

// poisson.cu's k-space step on the y-slab layout [nz][nyl][nxh]: spec_g = (rho_hat / N) * -(i * k_g * factor)
__global__ void kspace_slab_kernel(int nxh, int nyl, int nz, int ys, double inv_n, const double* __restrict__ kx,
                                   const double* __restrict__ ky, const double* __restrict__ kz,
                                   const cufftDoubleComplex* __restrict__ rh, cufftDoubleComplex* __restrict__ g0,
                                   cufftDoubleComplex* __restrict__ g1, cufftDoubleComplex* __restrict__ g2) {
    const long nh = (long)nxh * nyl * nz;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nh; t += (long)gridDim.x * blockDim.x) {
        const int i = (int)(t % nxh), j = (int)((t / nxh) % nyl), k = (int)(t / ((long)nxh * nyl));
        const double k0 = kx[i], k1 = ky[ys + j], k2 = kz[k];
        double Dr = 0;
        Dr += k0 * k0;
        Dr += k1 * k1;
        Dr += k2 * k2;
        const bool nzr      = (Dr != 0.0);
        const double factor = nzr ? (1.0 / Dr) : 0.0;
        const double a = rh[t].x * inv_n, b = rh[t].y * inv_n;
        double c = k0 * factor;
        g0[t]    = make_cuDoubleComplex(b * c, -(a * c));
        c        = k1 * factor;
        g1[t]    = make_cuDoubleComplex(b * c, -(a * c));
        c        = k2 * factor;
        g2[t]    = make_cuDoubleComplex(b * c, -(a * c));
    }
}
// [host-emulation end: slab kernels]

void slab_free(SlabState* st) {
    if (!st) return;
    for (int b = SB_REAL; b < SB_COUNT; ++b)
        if (st->buf[b]) cudaFree(st->buf[b]);
    if (st->d_copies) cudaFree(st->d_copies);
    if (st->have2d) {
        cufftDestroy(st->fwd2d);
        cufftDestroy(st->inv2d);
    }
    if (st->have1d) cufftDestroy(st->z1d);
    delete st;
}

// ---- the local steps of one rank ------------------------------------------------------------------------------------
static int run_copies(ipplb_poisson* s, int phase, int which, double* rho, double* ef) {
    SlabState* st = s->slab;
    const int n   = st->count[phase][which];
    if (n == 0) return IPPLB_OK;
    BufTable B;
    for (int b = 0; b < SB_COUNT; ++b) B.p[b] = st->buf[b];
    B.p[SB_RHO] = rho;
    B.p[SB_EF]  = ef;
    long gx = (st->biggest[phase][which] + 255) / 256;
    const long cap = (long)s->ctx->num_sms * 8;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    slab_copy_kernel<<<dim3((unsigned)gx, (unsigned)n), 256, 0, s->ctx->stream>>>(st->d_copies + st->first[phase][which], B);
    IPPLB_CHECK_LAUNCH(s->ctx);
    return IPPLB_OK;
}

// the messages of a phase; a message to myself is a device copy on my stream
static int phase_xfers(ipplb_poisson* s, int phase, std::vector<Xfer>& x) {
    SlabState* st = s->slab;
    x.clear();
    for (const SlabMsg& m : st->plan.phase[phase].msgs) {
        if (m.peer == st->plan.me) {
            if (m.scount)
                IPPLB_CUDA(cudaMemcpyAsync(st->buf[SB_RECV] + m.roff, st->buf[SB_SEND] + m.soff, sizeof(double) * m.scount,
                                           cudaMemcpyDeviceToDevice, s->ctx->stream));
            continue;
        }
        x.push_back(Xfer{m.peer, st->buf[SB_SEND] + m.soff, sizeof(double) * (size_t)m.scount, st->buf[SB_RECV] + m.roff,
                         sizeof(double) * (size_t)m.rcount});
    }
    return IPPLB_OK;
}

// the transforms between the exchanges: step 0 after P0, 1 after P1, 2 after P2
static int transforms(ipplb_poisson* s, int step) {
    SlabState* st     = s->slab;
    const SlabPlan& P = st->plan;
    ipplb_ctx* ctx    = s->ctx;
    const int nzl = P.ze - P.zs, nyl = P.ye - P.ys;
    const long S2 = (long)nzl * P.ng[1] * P.nxh, SZ = (long)P.ng[2] * nyl * P.nxh;
    cufftDoubleComplex* spec2d = reinterpret_cast<cufftDoubleComplex*>(st->buf[SB_SPEC2D]);
    cufftDoubleComplex* specz  = reinterpret_cast<cufftDoubleComplex*>(st->buf[SB_SPECZ]);
    (void)S2;
    if (step == 0) {
        if (nzl > 0) {
            IPPLB_CUFFT(cufftExecD2Z(st->fwd2d, st->buf[SB_REAL], spec2d));
            ctx->launches++;
        }
    } else if (step == 1) {
        if (nyl > 0) {
            IPPLB_CUFFT(cufftExecZ2Z(st->z1d, specz, specz, CUFFT_FORWARD));
            const long N  = (long)P.ng[0] * P.ng[1] * P.ng[2];
            const long g0 = (SZ + 255) / 256;
            const int g   = (int)(g0 < (long)ctx->num_sms * 16 ? g0 : (long)ctx->num_sms * 16);
            kspace_slab_kernel<<<g, 256, 0, ctx->stream>>>(P.nxh, nyl, P.ng[2], P.ys, 1.0 / (double)N, s->kx, s->ky, s->kz, specz,
                                                           specz + SZ, specz + 2 * SZ, specz + 3 * SZ);
            IPPLB_CHECK_LAUNCH(ctx);
            for (int c = 0; c < 3; ++c) IPPLB_CUFFT(cufftExecZ2Z(st->z1d, specz + (1 + c) * SZ, specz + (1 + c) * SZ, CUFFT_INVERSE));
            ctx->launches += 4;
        }
    } else {
        if (nzl > 0) {
            IPPLB_CUFFT(cufftExecZ2D(st->inv2d, spec2d, st->buf[SB_REAL]));
            ctx->launches++;
        }
    }
    return IPPLB_OK;
}

int slab_solve(ipplb_poisson* s, double* rho, double* efield) {
    ipplb_ctx* ctx = s->ctx;
    IPPLB_REQUIRE(ctx->nccl && !ctx->loop, "poisson_solve (slab): needs an NCCL communicator (in-process rank groups: ipplb_loop_poisson_solve)");
    int rc;
    std::vector<Xfer> x;
    for (int ph = 0; ph < 4; ++ph) {
        if ((rc = run_copies(s, ph, 0, rho, efield))) return rc;
        if ((rc = phase_xfers(s, ph, x))) return rc;
        if (!x.empty() && (rc = nccl_exchange(ctx, x))) return rc;
        if ((rc = run_copies(s, ph, 1, rho, efield))) return rc;
        if (ph < 3 && (rc = transforms(s, ph))) return rc;
    }
    return IPPLB_OK;
}

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_poisson_create_slab(ipplb_ctx* ctx, const ipplb_layout* layout, const double origin[3], const double h[3],
                              ipplb_poisson** out) {
    IPPLB_REQUIRE(ctx && layout && origin && h && out, "poisson_create_slab: bad arguments");
    const int nr = (int)layout->L.boxes.size();
    IPPLB_REQUIRE(nr == ctx->nranks, "poisson_create_slab: layout rank count != communicator size");
    if (nr == 1) return ipplb_poisson_create_dist(ctx, layout, origin, h, out);
    IPPLB_REQUIRE(ctx->nccl || ctx->loop, "poisson_create_slab: no communicator (ipplb_comm_init / ipplb_loop_create)");
    IPPLB_CUDA(cudaSetDevice(ctx->device));
    ipplb_poisson* s = new ipplb_poisson();
    s->ctx    = ctx;
    s->nranks = nr;
    s->nx = layout->L.ng[0]; s->ny = layout->L.ng[1]; s->nz = layout->L.ng[2];
    s->nxh = s->nx / 2 + 1;
    SlabState* st = new SlabState();
    s->slab       = st;
    auto fail = [&](int rc) {
        ipplb_poisson_destroy(s);
        return rc;
    };
    int rc;
    if ((rc = st->plan.build(layout->L, ctx->rank))) return fail(rc);
    const SlabPlan& P = st->plan;
    for (int b = SB_REAL; b < SB_COUNT; ++b) {
        const size_t bytes = sizeof(double) * (size_t)(P.size[b] > 0 ? P.size[b] : 2);
        cudaError_t e      = cudaMalloc(&st->buf[b], bytes);
        if (e != cudaSuccess) {
            set_error("poisson_create_slab: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
            return fail(IPPLB_ERR_CUDA);
        }
    }
    // k tables (same host arithmetic as the single-GPU solver)
    {
        const std::vector<double> kh = poisson_k_tables(layout->L.ng, origin, h);
        cudaError_t e = cudaMalloc(&s->kx, sizeof(double) * kh.size());
        if (e == cudaSuccess) e = cudaMemcpy(s->kx, kh.data(), sizeof(double) * kh.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("poisson_create_slab: k tables -> %s", cudaGetErrorString(e));
            return fail(IPPLB_ERR_CUDA);
        }
        s->ky = s->kx + s->nxh;
        s->kz = s->ky + s->ny;
    }
    // copy lists -> one device table
    {
        std::vector<CopyDev> all;
        for (int ph = 0; ph < 4; ++ph)
            for (int w = 0; w < 2; ++w) {
                const std::vector<SlabCopy>& src = w == 0 ? P.phase[ph].pre : P.phase[ph].post;
                st->first[ph][w] = (int)all.size();
                st->count[ph][w] = (int)src.size();
                for (const SlabCopy& c : src) {
                    CopyDev d;
                    d.src_off = c.src_off; d.dst_off = c.dst_off;
                    for (int a = 0; a < 3; ++a) { d.ss[a] = c.ss[a]; d.ds[a] = c.ds[a]; d.n[a] = c.n[a]; }
                    d.src_buf = c.src_buf; d.dst_buf = c.dst_buf; d.elem = c.elem;
                    all.push_back(d);
                    const long tot = (long)c.n[0] * c.n[1] * c.n[2];
                    if (tot > st->biggest[ph][w]) st->biggest[ph][w] = tot;
                }
            }
        const size_t bytes = sizeof(CopyDev) * (all.empty() ? 1 : all.size());
        cudaError_t e      = cudaMalloc(&st->d_copies, bytes);
        if (e == cudaSuccess && !all.empty()) e = cudaMemcpy(st->d_copies, all.data(), sizeof(CopyDev) * all.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("poisson_create_slab: copy tables -> %s", cudaGetErrorString(e));
            return fail(IPPLB_ERR_CUDA);
        }
    }
    // transforms: whole (y, x) planes of my z-slab; lines along z of my y-slab (stride = one plane of the slab)
    const int nzl = P.ze - P.zs, nyl = P.ye - P.ys;
    auto cufft_fail = [&](cufftResult r, const char* what) {
        set_error("poisson_create_slab: %s -> cufft error %d", what, (int)r);
        return fail(IPPLB_ERR_CUFFT);
    };
    if (nzl > 0) {
        int n2[2] = {s->ny, s->nx};
        cufftResult r = cufftPlanMany(&st->fwd2d, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_D2Z, nzl);
        if (r != CUFFT_SUCCESS) return cufft_fail(r, "2-D forward plan");
        r = cufftPlanMany(&st->inv2d, 2, n2, nullptr, 1, 0, nullptr, 1, 0, CUFFT_Z2D, 3 * nzl);
        if (r != CUFFT_SUCCESS) {
            cufftDestroy(st->fwd2d);
            return cufft_fail(r, "2-D inverse plan");
        }
        st->have2d = true;
        if ((r = cufftSetStream(st->fwd2d, ctx->stream)) != CUFFT_SUCCESS) return cufft_fail(r, "set stream");
        if ((r = cufftSetStream(st->inv2d, ctx->stream)) != CUFFT_SUCCESS) return cufft_fail(r, "set stream");
    }
    if (nyl > 0) {
        int n1[1]    = {s->nz};
        int embed[1] = {s->nz};
        const int stride = nyl * s->nxh;
        cufftResult r = cufftPlanMany(&st->z1d, 1, n1, embed, stride, 1, embed, stride, 1, CUFFT_Z2Z, stride);
        if (r != CUFFT_SUCCESS) return cufft_fail(r, "1-D plan along z");
        st->have1d = true;
        if ((r = cufftSetStream(st->z1d, ctx->stream)) != CUFFT_SUCCESS) return cufft_fail(r, "set stream");
    }
    *out = s;
    return IPPLB_OK;
}

int ipplb_loop_poisson_solve(ipplb_loop* L, ipplb_poisson* const* solvers, double* const* rho, double* const* efield) {
    IPPLB_REQUIRE(L && solvers && rho && efield, "loop_poisson_solve: bad arguments");
    const int nr = (int)L->ctx.size();
    for (int r = 0; r < nr; ++r)
        IPPLB_REQUIRE(solvers[r] && solvers[r]->slab && solvers[r]->ctx == L->ctx[r] && rho[r] && efield[r],
                      "loop_poisson_solve: solver r must be the slab solver of rank r's context");
    int rc;
    std::vector<std::vector<Xfer>> all(nr);
    for (int ph = 0; ph < 4; ++ph) {
        for (int r = 0; r < nr; ++r) {
            IPPLB_CUDA(cudaSetDevice(L->ctx[r]->device));
            if ((rc = run_copies(solvers[r], ph, 0, rho[r], efield[r]))) return rc;
            if ((rc = phase_xfers(solvers[r], ph, all[r]))) return rc;
        }
        if ((rc = loop_exchange(L, all))) return rc;
        for (int r = 0; r < nr; ++r) {
            if ((rc = run_copies(solvers[r], ph, 1, rho[r], efield[r]))) return rc;
            if (ph < 3 && (rc = transforms(solvers[r], ph))) return rc;
        }
    }
    return IPPLB_OK;
}

}  // extern "C"
