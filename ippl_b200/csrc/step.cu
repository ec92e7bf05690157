// step.cu -- whole-step conveniences on top of the kernels (bench.py, facade).
#include "common.cuh"

using namespace ipplb;

static long ghosted_cells(const ipplb_mesh* mesh) {
    return (long)(mesh->nl[0] + 2 * mesh->nghost) * (mesh->nl[1] + 2 * mesh->nghost) *
           (mesh->nl[2] + 2 * mesh->nghost);
}

static void swap_bundles(ipplb_particles* p, ipplb_particles* scratch) {
    ipplb_particles t = *p;
    *p                = *scratch;
    *scratch          = t;
}

extern "C" {

int ipplb_pic_step(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push, ipplb_particles* p,
                   ipplb_particles* scratch, int* cell_offsets, ipplb_bins* bins, const double* efield,
                   double* rho, int do_sort) {
    IPPLB_REQUIRE(ctx && mesh && push && p && efield && rho, "pic_step: bad arguments");
    int rc;
    int mask = 0;
    for (int d = 0; d < 3; ++d)
        if (mesh->nl[d] == mesh->ng[d]) mask |= 1 << d;
    const long cells = ghosted_cells(mesh);
    if (do_sort == 2) {
        IPPLB_REQUIRE(bins && scratch, "pic_step: the fused step needs bins and a second particle bundle");
        if ((rc = ipplb_field_fill(ctx, rho, cells, 0.0))) return rc;
        const long n = p->n;
        // without the BC on the whole periodic domain a leaver would be dropped silently (no exit buffer here)
        IPPLB_REQUIRE(push->do_bc && mask == 7, "pic_step: the fused single-rank step needs the periodic BC on the whole domain");
        // the rank owns the whole periodic domain: ipplb_bins_step aliases ghost nodes to the opposite interior layer
        // itself, so E needs no fillHalo before the step and rho no accumulateHalo after it
        if ((rc = ipplb_bins_step(ctx, bins, push, p, scratch, efield, rho, nullptr, 0, nullptr, nullptr))) return rc;
        swap_bundles(p, scratch);
        p->n       = n;  // single rank, periodic: nobody leaves (ipplb_bins_status reports the device truth)
        scratch->n = 0;
        return IPPLB_OK;
    }
    if ((rc = ipplb_gather_push(ctx, mesh, push, p, efield))) return rc;
    if ((rc = ipplb_field_fill(ctx, rho, cells, 0.0))) return rc;
    if (do_sort) {
        IPPLB_REQUIRE(scratch && cell_offsets, "pic_step: sort needs scratch arrays and cell_offsets");
        if ((rc = ipplb_sort_by_cell(ctx, mesh, p, scratch, cell_offsets))) return rc;
        swap_bundles(p, scratch);
        scratch->n = 0;
        if ((rc = ipplb_scatter_cic_sorted(ctx, mesh, p->n, p->x, p->y, p->z, p->q, p->q_scalar,
                                           cell_offsets, rho)))
            return rc;
    } else {
        if ((rc = ipplb_scatter_cic(ctx, mesh, 0, p->n, p->x, p->y, p->z, p->q, p->q_scalar, nullptr,
                                    rho)))
            return rc;
    }
    return ipplb_halo_accumulate_periodic(ctx, mesh, rho, 1, mask);
}

int ipplb_pic_step_host(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push, long n,
                        double* const host_arrays[6], double q_scalar, const double* efield_dev,
                        double* rho_host, ipplb_particles* dev, ipplb_particles* scratch,
                        ipplb_bins* bins, double* rho_dev) {
    IPPLB_REQUIRE(ctx && mesh && push && host_arrays && dev && scratch && bins && rho_dev,
                  "pic_step_host: bad arguments");
    IPPLB_REQUIRE(dev->capacity >= n && scratch->capacity >= n, "pic_step_host: device capacity too small");
    // host -> dev (contiguous), dev -> scratch (bucketed), scratch -> dev (fused step), dev -> scratch
    // (contiguous), scratch -> host
    double* d[6] = {dev->x, dev->y, dev->z, dev->px, dev->py, dev->pz};
    for (int a = 0; a < 6; ++a)
        IPPLB_CUDA(cudaMemcpyAsync(d[a], host_arrays[a], sizeof(double) * (size_t)n,
                                   cudaMemcpyHostToDevice, ctx->stream));
    dev->n        = n;
    dev->q        = nullptr;
    dev->q_scalar = q_scalar;
    int rc;
    if ((rc = ipplb_bins_build(ctx, bins, dev, scratch))) return rc;
    if ((rc = ipplb_pic_step(ctx, mesh, push, scratch, dev, nullptr, bins, efield_dev, rho_dev, 2))) return rc;
    // after the swap inside pic_step `scratch` names the bucketed result and `dev` the spare bundle
    if ((rc = ipplb_bins_compact(ctx, bins, scratch, dev))) return rc;
    IPPLB_REQUIRE(dev->n == n, "pic_step_host: particle count changed");
    double* d2[6] = {dev->x, dev->y, dev->z, dev->px, dev->py, dev->pz};
    for (int a = 0; a < 6; ++a)
        IPPLB_CUDA(cudaMemcpyAsync(host_arrays[a], d2[a], sizeof(double) * (size_t)n,
                                   cudaMemcpyDeviceToHost, ctx->stream));
    if (rho_host)
        IPPLB_CUDA(cudaMemcpyAsync(rho_host, rho_dev, sizeof(double) * (size_t)ghosted_cells(mesh),
                                   cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return IPPLB_OK;
}

// Steady-state end-to-end step (bench.py's e2e): the particles live on the device -- as they do in the reference, whose
// ParticleAttrib views are device allocations -- and what crosses the host boundary every step is what the (host-side or
// non-owned) field solve exchanges with the particle path: E comes in, rho goes out.
int ipplb_pic_step_host_fields(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push, ipplb_particles* p,
                               ipplb_particles* scratch, ipplb_bins* bins, const double* efield_host, double* rho_host,
                               double* efield_dev, double* rho_dev) {
    IPPLB_REQUIRE(ctx && mesh && push && p && scratch && bins && efield_host && rho_host && efield_dev && rho_dev,
                  "pic_step_host_fields: bad arguments");
    const size_t cells = (size_t)ghosted_cells(mesh);
    IPPLB_CUDA(cudaMemcpyAsync(efield_dev, efield_host, sizeof(double) * 3 * cells, cudaMemcpyHostToDevice, ctx->stream));
    int rc;
    if ((rc = ipplb_pic_step(ctx, mesh, push, p, scratch, nullptr, bins, efield_dev, rho_dev, 2))) return rc;
    IPPLB_CUDA(cudaMemcpyAsync(rho_host, rho_dev, sizeof(double) * cells, cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return IPPLB_OK;
}

// Pipelined variant over a sequence of independent batches (a "step" = one pass of the hot path over one batch):
// upload of batch k+1, compute of batch k and download of batch k-1 overlap on three streams, two device slots.
// Per slot: H2D -> dev; build dev -> scratch; fused step scratch -> dev; compact dev -> scratch; D2H <- scratch.
int ipplb_pic_step_host_batches(ipplb_ctx* ctx, const ipplb_mesh* mesh, const ipplb_push* push, long n, int nbatch,
                                double* const* host_arrays, double q_scalar, const double* efield_dev,
                                double* const* rho_host, ipplb_particles* dev, ipplb_particles* scratch,
                                ipplb_bins* const* bins, double* const* rho_dev) {
    IPPLB_REQUIRE(ctx && mesh && push && host_arrays && dev && scratch && bins && rho_dev && nbatch >= 0,
                  "pic_step_host_batches: bad arguments");
    if (!ctx->s_in) {
        IPPLB_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        IPPLB_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (int s = 0; s < 2; ++s) {
            IPPLB_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[s], cudaEventDisableTiming));
            IPPLB_CUDA(cudaEventCreateWithFlags(&ctx->ev_comp[s], cudaEventDisableTiming));
            IPPLB_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[s], cudaEventDisableTiming));
        }
    }
    const long cells = ghosted_cells(mesh);
    auto upload = [&](int k) -> int {
        const int s = k & 1;
        IPPLB_REQUIRE(dev[s].capacity >= n && scratch[s].capacity >= n, "pic_step_host_batches: device capacity too small");
        // the slot's `dev` bundle is free once the compute of batch k-2 is done; callers may also hand the same host
        // buffers to batches k-2 and k, so wait for that download as well (no-ops for the first two batches)
        if (k >= 2) {
            IPPLB_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->ev_comp[s], 0));
            IPPLB_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->ev_out[s], 0));
        }
        double* d[6] = {dev[s].x, dev[s].y, dev[s].z, dev[s].px, dev[s].py, dev[s].pz};
        for (int a = 0; a < 6; ++a)
            IPPLB_CUDA(cudaMemcpyAsync(d[a], host_arrays[6 * k + a], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice,
                                       ctx->s_in));
        IPPLB_CUDA(cudaEventRecord(ctx->ev_in[s], ctx->s_in));
        return IPPLB_OK;
    };
    int rc;
    if (nbatch > 0 && (rc = upload(0))) return rc;
    for (int k = 0; k < nbatch; ++k) {
        const int s = k & 1;
        if (k + 1 < nbatch && (rc = upload(k + 1))) return rc;
        // compute: needs this batch's upload and (for `scratch`) the download of batch k-2
        IPPLB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[s], 0));
        if (k >= 2) IPPLB_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[s], 0));
        ipplb_particles a = dev[s], b = scratch[s];
        a.n = n; a.q = nullptr; a.q_scalar = q_scalar;
        if ((rc = ipplb_bins_build(ctx, bins[s], &a, &b))) return rc;
        if ((rc = ipplb_pic_step(ctx, mesh, push, &b, &a, nullptr, bins[s], efield_dev, rho_dev[s], 2))) return rc;
        // after the swap inside pic_step `b` names the bucketed result (storage of dev[s]) and `a` the spare one
        if ((rc = ipplb_bins_compact(ctx, bins[s], &b, &a))) return rc;
        IPPLB_REQUIRE(a.n == n, "pic_step_host_batches: particle count changed");
        IPPLB_CUDA(cudaEventRecord(ctx->ev_comp[s], ctx->stream));
        // download from the compacted bundle (storage of scratch[s]) + rho
        IPPLB_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_comp[s], 0));
        double* o[6] = {a.x, a.y, a.z, a.px, a.py, a.pz};
        for (int c = 0; c < 6; ++c)
            IPPLB_CUDA(cudaMemcpyAsync(host_arrays[6 * k + c], o[c], sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost,
                                       ctx->s_out));
        if (rho_host && rho_host[k])
            IPPLB_CUDA(cudaMemcpyAsync(rho_host[k], rho_dev[s], sizeof(double) * (size_t)cells, cudaMemcpyDeviceToHost,
                                       ctx->s_out));
        IPPLB_CUDA(cudaEventRecord(ctx->ev_out[s], ctx->s_out));
    }
    IPPLB_CUDA(cudaStreamSynchronize(ctx->s_out));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return IPPLB_OK;
}

}  // extern "C"
