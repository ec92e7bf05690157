// common.cuh -- shared declarations for the B200 (sm_100a) particle-mesh kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/ippl_b200.h"

namespace ipplb {

void set_error(const char* fmt, ...);

#define IPPLB_CUDA(call)                                                                     \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            ipplb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                   \
                             cudaGetErrorString(e__));                                       \
            return (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver)          \
                       ? IPPLB_ERR_NO_DEVICE                                                 \
                       : IPPLB_ERR_CUDA;                                                     \
        }                                                                                    \
    } while (0)

#define IPPLB_CHECK_LAUNCH(ctx)                                                              \
    do {                                                                                     \
        (ctx)->launches++;                                                                   \
        IPPLB_CUDA(cudaGetLastError());                                                      \
    } while (0)

#define IPPLB_REQUIRE(cond, msg)                                                             \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ipplb::set_error("%s:%d: %s", __FILE__, __LINE__, msg);                          \
            return IPPLB_ERR_ARG;                                                            \
        }                                                                                    \
    } while (0)

// Device-side mesh constants, passed by value to kernels.
struct MeshDev {
    double origin[3];
    double invdx[3];  // 1.0 / h, formed once on the host like ParticleAttrib.hpp:153
    int first[3];     // lDom.first()
    int nl[3];
    int ng[3];
    int nghost;
    int ex, ey, ez;  // ghosted extents
};

inline MeshDev make_mesh_dev(const ipplb_mesh* m) {
    MeshDev d;
    for (int k = 0; k < 3; ++k) {
        d.origin[k] = m->origin[k];
        d.invdx[k]  = 1.0 / m->h[k];
        d.first[k]  = m->first[k];
        d.nl[k]     = m->nl[k];
        d.ng[k]     = m->ng[k];
    }
    d.nghost = m->nghost;
    d.ex     = m->nl[0] + 2 * m->nghost;
    d.ey     = m->nl[1] + 2 * m->nghost;
    d.ez     = m->nl[2] + 2 * m->nghost;
    return d;
}

struct Scratch {
    void* ptr    = nullptr;
    size_t bytes = 0;
};

}  // namespace ipplb

struct ncclComm;
struct ipplb_loop;

struct ipplb_ctx {
    int device          = 0;
    cudaStream_t stream = nullptr;
    bool own_stream     = false;
    long launches       = 0;
    int num_sms         = 148;
    int gather_variant  = 1;  // ipplb_ctx_set_gather_variant: 1 eight-byte loads per node component, 2 sixteen-byte loads per x-pair
    // scratch pools (grown on demand, never shrunk -- same policy as the reference's
    // BufferHandler, src/Communicate/BufferHandler.hpp:14-34)
    ipplb::Scratch keys, counts, cub_tmp, reduce, send, recv, misc;
    double* reduce_host = nullptr;  // pinned
    // multi-GPU
    ncclComm* nccl = nullptr;
    int rank = 0, nranks = 1;
    void* plan = nullptr;  // ipplb::CommPlan*
    void* mig  = nullptr;  // ipplb::MigBox*: peer-memory inboxes of the bucketed store's migration (ipplb_migrate_connect)
    ipplb_loop* loop = nullptr;  // set when this context is one rank of an in-process rank group (ipplb_loop_create)
    const double* d_regions = nullptr;  // [nranks][6] physical regions (device), set by ipplb_ctx_set_layout
    // copy streams + events of the host-buffer pipeline (ipplb_pic_step_host_batches), created on first use
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
};

namespace ipplb {
int ensure(ipplb_ctx* ctx, Scratch& s, size_t bytes);
}
