// field.cu -- context management + ghosted-field kernels (fill, interior sum, density, periodic halo).
#include "common.cuh"

#include <cmath>

namespace ipplb {

static thread_local std::string g_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
}

int ensure(ipplb_ctx* ctx, Scratch& s, size_t bytes) {
    if (s.bytes >= bytes && s.ptr) return IPPLB_OK;
    if (s.ptr) {
        IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
        IPPLB_CUDA(cudaFree(s.ptr));
        s.ptr   = nullptr;
        s.bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;  // over-allocate like the reference's buffer pool
    want        = (want + 4095) & ~(size_t)4095;
    IPPLB_CUDA(cudaMalloc(&s.ptr, want));
    s.bytes = want;
    return IPPLB_OK;
}

__global__ void fill_kernel(double* __restrict__ f, long n, double v) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f[i] = v;
}

// interior reduction: block partial sums, final pass on one block (deterministic order)
__global__ void __launch_bounds__(256)
interior_sum_kernel(MeshDev m, const double* __restrict__ f, double* __restrict__ partial) {
    __shared__ double s[256];
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    double acc    = 0.0;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni;
         t += (long)gridDim.x * blockDim.x) {
        int i = (int)(t % m.nl[0]) + m.nghost;
        int j = (int)((t / m.nl[0]) % m.nl[1]) + m.nghost;
        int k = (int)(t / ((long)m.nl[0] * m.nl[1])) + m.nghost;
        acc += f[i + (long)m.ex * (j + (long)m.ey * k)];
    }
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

__global__ void __launch_bounds__(256) final_sum_kernel(const double* __restrict__ partial, int n,
                                                        double* __restrict__ out) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0];
}

__global__ void __launch_bounds__(256)
ex_stats_kernel(MeshDev m, const double* __restrict__ ef, double* __restrict__ partial) {
    __shared__ double s2[256], sm[256];
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    double a2 = 0.0, amax = 0.0;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni;
         t += (long)gridDim.x * blockDim.x) {
        int i = (int)(t % m.nl[0]) + m.nghost;
        int j = (int)((t / m.nl[0]) % m.nl[1]) + m.nghost;
        int k = (int)(t / ((long)m.nl[0] * m.nl[1])) + m.nghost;
        double v = ef[(i + (long)m.ex * (j + (long)m.ey * k)) * 3];
        a2 += v * v;
        amax = fmax(amax, fabs(v));
    }
    s2[threadIdx.x] = a2;
    sm[threadIdx.x] = amax;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s2[threadIdx.x] += s2[threadIdx.x + o];
            sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x]     = s2[0];
        partial[2 * blockIdx.x + 1] = sm[0];
    }
}

__global__ void __launch_bounds__(256)
ex_stats_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double s2[256], sm[256];
    double a2 = 0.0, amax = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a2 += partial[2 * i];
        amax = fmax(amax, partial[2 * i + 1]);
    }
    s2[threadIdx.x] = a2;
    sm[threadIdx.x] = amax;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s2[threadIdx.x] += s2[threadIdx.x + o];
            sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = s2[0];
        out[1] = sm[0];
    }
}

// rho = rho / cellVolume; rho = rho - shift   (two roundings, AlpineManager.h:225-245)
__global__ void density_kernel(MeshDev m, double* __restrict__ f, double cell_volume, double shift) {
    const long ni = (long)m.nl[0] * m.nl[1] * m.nl[2];
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < ni;
         t += (long)gridDim.x * blockDim.x) {
        int i = (int)(t % m.nl[0]) + m.nghost;
        int j = (int)((t / m.nl[0]) % m.nl[1]) + m.nghost;
        int k = (int)(t / ((long)m.nl[0] * m.nl[1])) + m.nghost;
        long l = i + (long)m.ex * (j + (long)m.ey * k);
        f[l]   = __dsub_rn(__ddiv_rn(f[l], cell_volume), shift);
    }
}

// HaloPeriodicFunctor (HaloCells.hpp:59-87) for one dimension d: one thread per (ghost layer i,
// other-dims coordinate incl. ghosts, component).  accumulate: right += glow; left += gup.
// fill: glow = right; gup = left.   Launched for d = 0,1,2 in order so edges/corners cascade exactly
// like the reference's three sequential kernels.
__global__ void halo_periodic_kernel(double* __restrict__ v, int e0, int e1, int e2, int ncomp,
                                     int nghost, int d, int accumulate) {
    int ext[3]    = {e0, e1, e2};
    const int N   = ext[d] - 1;
    int end[3]    = {e0, e1, e2};
    end[d]        = nghost;
    const long nt = (long)end[0] * end[1] * end[2] * ncomp;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nt;
         t += (long)gridDim.x * blockDim.x) {
        int c    = (int)(t % ncomp);
        long r   = t / ncomp;
        int co[3];
        co[0] = (int)(r % end[0]);
        co[1] = (int)((r / end[0]) % end[1]);
        co[2] = (int)(r / ((long)end[0] * end[1]));
        const int i = co[d];
        int t3[3]   = {co[0], co[1], co[2]};
        auto at     = [&](int cd) {
            t3[d] = cd;
            return ((long)t3[0] + (long)e0 * (t3[1] + (long)e1 * t3[2])) * ncomp + c;
        };
        const long left = at(nghost + i), right = at(N - nghost - i);
        const long glow = at(nghost - 1 - i), gup = at(N - nghost + 1 + i);
        if (accumulate) {
            v[right] = __dadd_rn(v[right], v[glow]);
            v[left]  = __dadd_rn(v[left], v[gup]);
        } else {
            v[glow] = v[right];
            v[gup]  = v[left];
        }
    }
}

static int small_grid(long n) {
    long g = (n + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

static int halo_periodic(ipplb_ctx* ctx, const ipplb_mesh* mesh, double* field, int ncomp,
                         int serial_mask, int accumulate) {
    IPPLB_REQUIRE(ctx && mesh && field && ncomp >= 1, "halo_periodic: bad arguments");
    MeshDev m = make_mesh_dev(mesh);
    for (int d = 0; d < 3; ++d) {
        if (!(serial_mask & (1 << d))) continue;
        int ext[3] = {m.ex, m.ey, m.ez};
        // with nl[d] == 1 and nghost == 1 left == right: the two updates would race
        IPPLB_REQUIRE(ext[d] - 2 * m.nghost >= 2 * m.nghost || !accumulate,
                      "halo_periodic: local extent smaller than 2*nghost");
        long nt = (long)ncomp * m.nghost * ext[(d + 1) % 3] * ext[(d + 2) % 3];
        halo_periodic_kernel<<<small_grid(nt), 256, 0, ctx->stream>>>(field, m.ex, m.ey, m.ez, ncomp,
                                                                      m.nghost, d, accumulate);
        IPPLB_CHECK_LAUNCH(ctx);
    }
    return IPPLB_OK;
}

}  // namespace ipplb

using namespace ipplb;

extern "C" {

const char* ipplb_last_error(void) { return g_error.c_str(); }
const char* ipplb_version(void) { return "ippl_b200 0.1 (sm_100a)"; }

int ipplb_ctx_create(ipplb_ctx** out, int device, void* stream, int create_stream) {
    IPPLB_REQUIRE(out, "ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); ippl_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return IPPLB_ERR_NO_DEVICE;
    }
    IPPLB_REQUIRE(device >= 0 && device < count, "ctx_create: bad device index");
    IPPLB_CUDA(cudaSetDevice(device));
    ipplb_ctx* ctx = new ipplb_ctx();
    ctx->device    = device;
    if (create_stream) {
        IPPLB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    } else {
        ctx->stream = (cudaStream_t)stream;  // NULL = default stream
    }
    cudaDeviceProp prop;
    IPPLB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    IPPLB_CUDA(cudaMallocHost(&ctx->reduce_host, 64 * sizeof(double)));
    *out = ctx;
    return IPPLB_OK;
}

int ipplb_ctx_destroy(ipplb_ctx* ctx);  // defined in comm.cu (needs NCCL teardown)

int ipplb_sync(ipplb_ctx* ctx) {
    IPPLB_REQUIRE(ctx, "sync: ctx is NULL");
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    return IPPLB_OK;
}

void* ipplb_ctx_stream(ipplb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
long ipplb_launch_count(ipplb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ipplb_field_fill(ipplb_ctx* ctx, double* field, long count, double value) {
    IPPLB_REQUIRE(ctx && field && count >= 0, "field_fill: bad arguments");
    if (count == 0) return IPPLB_OK;
    if (value == 0.0) {
        IPPLB_CUDA(cudaMemsetAsync(field, 0, sizeof(double) * (size_t)count, ctx->stream));
        ctx->launches++;
        return IPPLB_OK;
    }
    fill_kernel<<<small_grid(count), 256, 0, ctx->stream>>>(field, count, value);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_field_sum(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* field, double* out_host) {
    IPPLB_REQUIRE(ctx && mesh && field && out_host, "field_sum: bad arguments");
    MeshDev m    = make_mesh_dev(mesh);
    const int nb = 592;
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * (2 * nb + 8)))) return rc;
    double* partial = (double*)ctx->reduce.ptr;
    interior_sum_kernel<<<nb, 256, 0, ctx->stream>>>(m, field, partial);
    IPPLB_CHECK_LAUNCH(ctx);
    final_sum_kernel<<<1, 256, 0, ctx->stream>>>(partial, nb, partial + 2 * nb);
    IPPLB_CHECK_LAUNCH(ctx);
    IPPLB_CUDA(cudaMemcpyAsync(ctx->reduce_host, partial + 2 * nb, sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_host = ctx->reduce_host[0];
    return IPPLB_OK;
}

int ipplb_field_ex_stats(ipplb_ctx* ctx, const ipplb_mesh* mesh, const double* efield,
                         double* out_host) {
    IPPLB_REQUIRE(ctx && mesh && efield && out_host, "field_ex_stats: bad arguments");
    MeshDev m    = make_mesh_dev(mesh);
    const int nb = 592;
    int rc;
    if ((rc = ensure(ctx, ctx->reduce, sizeof(double) * (2 * nb + 8)))) return rc;
    double* partial = (double*)ctx->reduce.ptr;
    ex_stats_kernel<<<nb, 256, 0, ctx->stream>>>(m, efield, partial);
    IPPLB_CHECK_LAUNCH(ctx);
    ex_stats_final_kernel<<<1, 256, 0, ctx->stream>>>(partial, nb, partial + 2 * nb);
    IPPLB_CHECK_LAUNCH(ctx);
    IPPLB_CUDA(cudaMemcpyAsync(ctx->reduce_host, partial + 2 * nb, 2 * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->stream));
    IPPLB_CUDA(cudaStreamSynchronize(ctx->stream));
    out_host[0] = ctx->reduce_host[0];
    out_host[1] = ctx->reduce_host[1];
    return IPPLB_OK;
}

int ipplb_field_density(ipplb_ctx* ctx, const ipplb_mesh* mesh, double* field, double cell_volume,
                        double shift) {
    IPPLB_REQUIRE(ctx && mesh && field, "field_density: bad arguments");
    MeshDev m = make_mesh_dev(mesh);
    long ni   = (long)m.nl[0] * m.nl[1] * m.nl[2];
    density_kernel<<<small_grid(ni), 256, 0, ctx->stream>>>(m, field, cell_volume, shift);
    IPPLB_CHECK_LAUNCH(ctx);
    return IPPLB_OK;
}

int ipplb_halo_accumulate_periodic(ipplb_ctx* ctx, const ipplb_mesh* mesh, double* field, int ncomp,
                                   int serial_mask) {
    return halo_periodic(ctx, mesh, field, ncomp, serial_mask, 1);
}
int ipplb_halo_fill_periodic(ipplb_ctx* ctx, const ipplb_mesh* mesh, double* field, int ncomp,
                             int serial_mask) {
    return halo_periodic(ctx, mesh, field, ncomp, serial_mask, 0);
}

}  // extern "C"
