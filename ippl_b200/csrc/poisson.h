// poisson.h -- the solver object shared by poisson.cu (single-GPU and replicated solve) and fftdist.cu (slab-decomposed solve)
#pragma once
#include <cufft.h>

#include <vector>

#include "common.cuh"
#include "slabplan.h"

namespace ipplb {
struct SlabState;   // fftdist.cu
}

struct ipplb_poisson {
    ipplb_ctx* ctx = nullptr;
    ipplb::MeshDev m;
    int nx = 0, ny = 0, nz = 0, nxh = 0;
    cufftHandle fwd = 0, inv = 0;
    bool plans = false;              // fwd / inv exist (single-GPU and replicated solve)
    double* real     = nullptr;          // 3 * N (component planes after the inverse)
    cufftDoubleComplex* spec = nullptr;  // 4 * Nh: [0] rho_hat, [1..3] gradient spectra
    double* kx = nullptr;                // kx[nxh] ky[ny] kz[nz]
    double *ky = nullptr, *kz = nullptr;
    // multi-rank (replicated solve): every rank gathers all rho boxes, solves the whole domain, keeps its box of E
    bool dist       = false;
    ipplb::MeshDev g;               // the whole domain as one box (m is this rank's box)
    int nranks      = 1;
    double* stage   = nullptr;      // all boxes back to back, box r at off[r]
    int* d_boxes    = nullptr;      // [nranks][6] lo, hi (inclusive)
    long* d_off     = nullptr;      // [nranks + 1]
    std::vector<long> off;
    ipplb::SlabState* slab = nullptr;   // set by ipplb_poisson_create_slab: ipplb_poisson_solve / _destroy dispatch on it
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


// kx[nx / 2 + 1] ky[ny] kz[nz] of FFTPeriodicPoissonSolver.hpp:66-70, 127-137 (host)
std::vector<double> poisson_k_tables(const int ng[3], const double origin[3], const double h[3]);
// fftdist.cu
int slab_solve(ipplb_poisson* s, double* rho, double* efield);
void slab_free(SlabState* st);

}  // namespace ipplb
