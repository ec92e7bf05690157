// ipplb_mock.cpp -- TEST INFRASTRUCTURE, NOT THE PRODUCT.  A CPU stand-in for the part of the C-ABI (include/ippl_b200.h)
// that the C++ facade calls on ONE rank, implemented on the oracle (oracle/ippl_oracle.cpp) plus a small DFT, together
// with stand-ins for the handful of CUDA runtime calls the facade makes ("device" memory = host memory).
//
// Purpose: run the reference's UNCHANGED alpine drivers through include/ippl/compat + include/ippl/KokkosShim.cuh
// (host-emulation mode) on a machine without a GPU, so that the host logic of that layer -- managers' call sequences,
// samplers, reductions, CSV dumps -- is checked against the reference's known-answer file before it ever meets a GPU.
// Nothing under ippl_b200/, demos/*.cpp or bench.py links this; the product library has no CPU path.
// Built by oracle/Makefile (target `mock`) into oracle/_build/mock/; used only by tests/test_ref_drivers_host_cpu.py.
#include <cmath>
#include <complex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>
#include <unistd.h>

#include <chrono>
#include <fstream>
#include <thread>

#include "../../include/ippl_b200.h"
#include "../../include/ippl/philox.h"
#include "../../ippl_b200/csrc/layout.h"

// ---- the oracle's C functions (oracle/ippl_oracle.cpp, compiled into the same library) -------------------------------
extern "C" {
struct orc_mesh {
    int ng[3];
    int first[3];
    int nl[3];
    int nghost;
    double origin[3];
    double h[3];
};
void orc_scatter_cic(const orc_mesh* m, long begin, long end, const double* x, const double* y, const double* z, const double* q,
                     double q_scalar, const int* hash, double* rho, int parallel);
void orc_gather_cic(const orc_mesh* m, long n, const double* x, const double* y, const double* z, const double* efield, int ncomp,
                    double** out, int add_to_attribute, int parallel);
void orc_periodic_bc(long n, double* x, double lo, double hi, int parallel);
void orc_halo_periodic(double* v, const int ext[3], int ncomp, int nghost, const int serial[3], int mode);
double orc_field_sum(const double* v, const int ext[3], int nghost);
void orc_density(double* v, const int ext[3], int nghost, double cell_volume, double q_over_size);
void orc_halo_exchange(const int ng[3], int nranks, const int* boxes, int nghost, int periodic, double** fields, int ncomp, int mode);
void orc_regions(const int ng[3], int nranks, const int* boxes, const double origin[3], const double h[3], double* regions);
void orc_locate(int nranks, const double* regions, int my, long n, const double* x, const double* y, const double* z, int* dest);
}

struct ipplb_ctx {
    int rank = 0, nranks = 1;
    // several ranks = several processes; the "communicator" is a directory (IPPLB_MOCK_DIR) through which every collective
    // is an all-gather of byte strings: small test sizes only
    std::string dir;
    long seq = 0;
    ipplb::Layout L;
    bool have_layout = false;
    double origin[3] = {0, 0, 0}, h[3] = {1, 1, 1};
    std::vector<int> boxes;        // [nranks][6]
    std::vector<double> regions;   // [nranks][6]
    // ipplb_update_plan -> ipplb_update_commit
    std::vector<int> dest;
    std::vector<double> arrivals;  // records of 7 doubles
    long plan_n = -1;
};
struct ipplb_poisson {
    ipplb_mesh m;        // the whole domain
    std::vector<double> k[3];
    ipplb_ctx* ctx = nullptr;
    bool dist = false;
    ipplb_mesh mine;     // this rank's box (dist)
};
struct ipplb_bins {
    ipplb_mesh m;
    long capacity = 0;
    long nexit    = 0;   // leavers of the last step, as 6-double records at the start of exit_buf
};

namespace {
std::string g_err;
}
namespace ipplb {   // what layout.cpp reports errors through (defined in field.cu in the product)
void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
}  // namespace ipplb
namespace {
int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
orc_mesh to_orc(const ipplb_mesh* m) {
    orc_mesh o;
    for (int d = 0; d < 3; ++d) {
        o.ng[d] = m->ng[d]; o.first[d] = m->first[d]; o.nl[d] = m->nl[d]; o.origin[d] = m->origin[d]; o.h[d] = m->h[d];
    }
    o.nghost = m->nghost;
    return o;
}
void ext_of(const ipplb_mesh* m, int e[3]) {
    for (int d = 0; d < 3; ++d) e[d] = m->nl[d] + 2 * m->nghost;
}
using cplx = std::complex<double>;
// plain O(n^2) DFT along one axis of a [n2][n1][n0] array (small grids only)
void dft_axis(std::vector<cplx>& a, const int n[3], int axis, int sign) {
    const long s[3] = {1, n[0], (long)n[0] * n[1]};
    const int len = n[axis];
    std::vector<cplx> w(len), line(len), out(len);
    for (int t = 0; t < len; ++t) w[t] = std::polar(1.0, sign * 2.0 * M_PI * t / len);
    const int o1 = (axis + 1) % 3, o2 = (axis + 2) % 3;
    for (int p = 0; p < n[o2]; ++p)
        for (int q = 0; q < n[o1]; ++q) {
            const long base = p * s[o2] + q * s[o1];
            for (int t = 0; t < len; ++t) line[t] = a[base + t * s[axis]];
            for (int f = 0; f < len; ++f) {
                cplx acc = 0;
                for (int t = 0; t < len; ++t) acc += line[t] * w[(int)(((long)f * t) % len)];
                out[f] = acc;
            }
            for (int t = 0; t < len; ++t) a[base + t * s[axis]] = out[t];
        }
}
}  // namespace

// ---- several ranks: every collective is an all-gather of byte strings through a directory ------------------------------------
namespace {
using Bytes = std::vector<char>;
std::vector<Bytes> allgather(ipplb_ctx* c, const void* mine, size_t bytes) {
    std::vector<Bytes> all(c->nranks);
    if (c->nranks == 1) {
        all[0].assign((const char*)mine, (const char*)mine + bytes);
        return all;
    }
    const long q = c->seq++;
    auto name = [&](long seq, int r) { return c->dir + "/m" + std::to_string(seq) + "_" + std::to_string(r); };
    {
        const std::string tmp = name(q, c->rank) + ".tmp";
        std::ofstream f(tmp, std::ios::binary);
        f.write((const char*)mine, (std::streamsize)bytes);
        f.close();
        std::rename(tmp.c_str(), name(q, c->rank).c_str());
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < c->nranks; ++r) {
        for (;;) {
            std::ifstream f(name(q, r), std::ios::binary | std::ios::ate);
            if (f) {
                const std::streamsize n = f.tellg();
                all[r].resize((size_t)n);
                f.seekg(0);
                f.read(all[r].data(), n);
                break;
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) {
                std::fprintf(stderr, "ipplb mock: rank %d waited 120 s for rank %d in collective %ld\n", c->rank, r, q);
                std::abort();
            }
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
    }
    // everybody has read collective q - 2 before anybody could finish q - 1: my file of q - 2 can go
    if (q >= 2) std::remove(name(q - 2, c->rank).c_str());
    return all;
}
template <class T, class Op>
void allreduce(ipplb_ctx* c, T* v, Op op) {
    const auto all = allgather(c, v, sizeof(T));
    T acc;
    std::memcpy(&acc, all[0].data(), sizeof(T));
    for (int r = 1; r < c->nranks; ++r) {
        T x;
        std::memcpy(&x, all[r].data(), sizeof(T));
        acc = op(acc, x);
    }
    *v = acc;
}
int box_len(const ipplb_ctx* c, int r, int d) { return c->boxes[r * 6 + 3 + d] - c->boxes[r * 6 + d] + 1; }

// FFTPeriodicPoissonSolver::solve, GRAD output, on the whole domain: rho_g [nz][ny][nx] -> E_g[c] [nz][ny][nx]
void solve_global(const ipplb_poisson* s, const std::vector<double>& rho_g, std::vector<double> E_g[3]) {
    const ipplb_mesh& m = s->m;
    const int n[3] = {m.ng[0], m.ng[1], m.ng[2]};
    const long N = (long)n[0] * n[1] * n[2];
    std::vector<cplx> rh(N);
    for (long l = 0; l < N; ++l) rh[l] = rho_g[l];
    for (int a = 0; a < 3; ++a) dft_axis(rh, n, a, -1);
    for (int c = 0; c < 3; ++c) {
        std::vector<cplx> t(N);
        for (int k = 0; k < n[2]; ++k)
            for (int j = 0; j < n[1]; ++j)
                for (int i = 0; i < n[0]; ++i) {
                    const long l        = i + (long)n[0] * (j + (long)n[1] * k);
                    const double kk[3]  = {s->k[0][i], s->k[1][j], s->k[2][k]};
                    const double Dr     = kk[0] * kk[0] + kk[1] * kk[1] + kk[2] * kk[2];
                    const double factor = Dr != 0.0 ? 1.0 / Dr : 0.0;
                    t[l] = (rh[l] / (double)N) * cplx(0.0, -(kk[c] * factor));
                }
        for (int a = 0; a < 3; ++a) dft_axis(t, n, a, +1);
        E_g[c].resize(N);
        for (long l = 0; l < N; ++l) E_g[c][l] = t[l].real();
    }
}
}  // namespace

extern "C" {

// ---- CUDA runtime stand-ins: device memory is host memory --------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t bytes) {
    *p = std::calloc(bytes ? bytes : 1, 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void* p) {
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    std::memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) {
    std::memset(d, v, n);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "mock CUDA runtime (oracle/mock)"; }

// ---- context -----------------------------------------------------------------------------------------------------------
const char* ipplb_last_error(void) { return g_err.c_str(); }
int ipplb_ctx_create(ipplb_ctx** out, int, void*, int) {
    *out = new ipplb_ctx();
    return IPPLB_OK;
}
int ipplb_ctx_destroy(ipplb_ctx* c) {
    delete c;
    return IPPLB_OK;
}
int ipplb_sync(ipplb_ctx*) { return IPPLB_OK; }
void* ipplb_ctx_stream(ipplb_ctx*) { return nullptr; }

// ---- particle <-> mesh ----------------------------------------------------------------------------------------------------
int ipplb_scatter_cic(ipplb_ctx*, const ipplb_mesh* mesh, long begin, long end, const double* x, const double* y, const double* z,
                      const double* q, double q_scalar, const int* hash, double* rho) {
    const orc_mesh m = to_orc(mesh);
    orc_scatter_cic(&m, begin, end, x, y, z, q, q_scalar, hash, rho, 1);
    return IPPLB_OK;
}
int ipplb_gather_cic(ipplb_ctx*, const ipplb_mesh* mesh, long n, const double* x, const double* y, const double* z,
                     const double* field, int ncomp, double* const* out, int add) {
    const orc_mesh m = to_orc(mesh);
    double* o[3] = {out[0], ncomp > 1 ? out[1] : nullptr, ncomp > 2 ? out[2] : nullptr};
    orc_gather_cic(&m, n, x, y, z, field, ncomp, o, add, 1);
    return IPPLB_OK;
}
// y = y + a * x  (ParticleAttrib::operator=(Expression) for the alpine kick / drift)
int ipplb_axpy(ipplb_ctx*, long n, double a, const double* x, double* y) {
    for (long i = 0; i < n; ++i) y[i] = y[i] + a * x[i];
    return IPPLB_OK;
}
int ipplb_apply_periodic_bc(ipplb_ctx*, long n, double* x, double* y, double* z, const double lo[3], const double hi[3], int mask) {
    double* c[3] = {x, y, z};
    for (int d = 0; d < 3; ++d)
        if (mask & (1 << d)) orc_periodic_bc(n, c[d], lo[d], hi[d], 1);
    return IPPLB_OK;
}

// ---- fields ----------------------------------------------------------------------------------------------------------------
int ipplb_field_fill(ipplb_ctx*, double* f, long count, double v) {
    for (long i = 0; i < count; ++i) f[i] = v;
    return IPPLB_OK;
}
int ipplb_field_sum(ipplb_ctx*, const ipplb_mesh* mesh, const double* f, double* out) {
    int e[3];
    ext_of(mesh, e);
    *out = orc_field_sum(f, e, mesh->nghost);
    return IPPLB_OK;
}
int ipplb_field_density(ipplb_ctx*, const ipplb_mesh* mesh, double* f, double cell_volume, double shift) {
    int e[3];
    ext_of(mesh, e);
    orc_density(f, e, mesh->nghost, cell_volume, shift);
    return IPPLB_OK;
}
static int halo(const ipplb_mesh* mesh, double* f, int ncomp, int mask, int mode) {
    int e[3];
    ext_of(mesh, e);
    const int serial[3] = {mask & 1, (mask >> 1) & 1, (mask >> 2) & 1};
    orc_halo_periodic(f, e, ncomp, mesh->nghost, serial, mode);
    return IPPLB_OK;
}
int ipplb_halo_accumulate_periodic(ipplb_ctx*, const ipplb_mesh* mesh, double* f, int ncomp, int mask) { return halo(mesh, f, ncomp, mask, 1); }
int ipplb_halo_fill_periodic(ipplb_ctx*, const ipplb_mesh* mesh, double* f, int ncomp, int mask) { return halo(mesh, f, ncomp, mask, 0); }

// ---- periodic Poisson solve, GRAD output (FFTPeriodicPoissonSolver.hpp:53-169), full complex DFTs ------------------------------
int ipplb_poisson_create(ipplb_ctx*, const ipplb_mesh* mesh, ipplb_poisson** out) {
    for (int d = 0; d < 3; ++d)
        if (mesh->nl[d] != mesh->ng[d]) return fail(IPPLB_ERR_ARG, "mock poisson: single rank only");
    ipplb_poisson* s = new ipplb_poisson();
    s->m = *mesh;
    for (int d = 0; d < 3; ++d) {
        const int N      = mesh->ng[d];
        const double Len = (mesh->origin[d] + N * mesh->h[d]) - mesh->origin[d];
        s->k[d].resize(N);
        for (int i = 0; i < N; ++i) {
            const bool shift = i > N / 2, notMid = i != N / 2;
            s->k[d][i] = notMid * 2 * M_PI / Len * (i - shift * N);
        }
    }
    *out = s;
    return IPPLB_OK;
}
int ipplb_poisson_create_dist(ipplb_ctx* ctx, const ipplb_layout* layout, const double origin[3], const double h[3], ipplb_poisson** out) {
    const int nr = ipplb_layout_nranks(layout);
    ipplb_mesh whole{};
    for (int d = 0; d < 3; ++d) {
        whole.ng[d] = whole.nl[d] = layout->L.ng[d];
        whole.origin[d] = origin[d];
        whole.h[d]      = h[d];
    }
    whole.nghost = layout->L.nghost;
    const int rc = ipplb_poisson_create(ctx, &whole, out);
    if (rc || nr == 1) return rc;
    if (nr != ctx->nranks) return fail(IPPLB_ERR_ARG, "mock poisson: layout rank count != communicator size");
    (*out)->ctx  = ctx;
    (*out)->dist = true;
    ipplb_layout_mesh(layout, ctx->rank, origin, h, &(*out)->mine);
    return IPPLB_OK;
}
int ipplb_poisson_solve(ipplb_poisson* s, double* rho, double* ef) {
    const ipplb_mesh& w = s->m;
    const ipplb_mesh& m = s->dist ? s->mine : s->m;
    const int g = m.nghost;
    const long ex = m.nl[0] + 2 * g, ey = m.nl[1] + 2 * g, N = (long)w.ng[0] * w.ng[1] * w.ng[2];
    // my interior, x fastest
    std::vector<double> mine((size_t)m.nl[0] * m.nl[1] * m.nl[2]);
    for (int k = 0; k < m.nl[2]; ++k)
        for (int j = 0; j < m.nl[1]; ++j)
            for (int i = 0; i < m.nl[0]; ++i) mine[i + (long)m.nl[0] * (j + (long)m.nl[1] * k)] = rho[(i + g) + ex * ((j + g) + ey * (k + g))];
    std::vector<double> rho_g(N);
    if (s->dist) {   // replicated solve: every rank assembles the whole rho
        ipplb_ctx* c = s->ctx;
        const auto all = allgather(c, mine.data(), sizeof(double) * mine.size());
        for (int r = 0; r < c->nranks; ++r) {
            const int* b = &c->boxes[r * 6];
            const int bx = box_len(c, r, 0), by = box_len(c, r, 1), bz = box_len(c, r, 2);
            const double* src = (const double*)all[r].data();
            for (int k = 0; k < bz; ++k)
                for (int j = 0; j < by; ++j)
                    for (int i = 0; i < bx; ++i)
                        rho_g[(i + b[0]) + (long)w.ng[0] * ((j + b[1]) + (long)w.ng[1] * (k + b[2]))] = src[i + (long)bx * (j + (long)by * k)];
        }
    } else {
        rho_g = mine;
    }
    std::vector<double> E_g[3];
    solve_global(s, rho_g, E_g);
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < m.nl[2]; ++k)
            for (int j = 0; j < m.nl[1]; ++j)
                for (int i = 0; i < m.nl[0]; ++i) {
                    const long cell = (i + g) + ex * ((j + g) + ey * (k + g));
                    const long gl   = (i + m.first[0]) + (long)w.ng[0] * ((j + m.first[1]) + (long)w.ng[1] * (k + m.first[2]));
                    ef[3 * cell + c] = E_g[c][gl];
                    if (c == 2) rho[cell] = E_g[c][gl];   // the inverse lands in rho's storage (:153)
                }
    return IPPLB_OK;
}
int ipplb_poisson_destroy(ipplb_poisson* s) {
    delete s;
    return IPPLB_OK;
}

// ---- the drivers' dump reductions and the PenningTrap kicks (used by demos/ref_lambdas.cu as the other side of its checks) ----
struct orc_penning {
    double origin[3], length[3], V0, alpha, Bext, DrInv;
};
void orc_penning_kick1(const orc_penning* pp, long n, const double* x, const double* y, const double* z, double* px, double* py,
                       double* pz, const double* ex, const double* ey, const double* ez);
void orc_penning_kick2(const orc_penning* pp, long n, const double* x, const double* y, const double* z, double* px, double* py,
                       double* pz, const double* ex, const double* ey, const double* ez);
int ipplb_penning_kick(ipplb_ctx*, int which, const ipplb_push* push, long n, const double* x, const double* y, const double* z,
                       double* px, double* py, double* pz, const double* ex, const double* ey, const double* ez) {
    orc_penning pp;
    for (int d = 0; d < 3; ++d) { pp.origin[d] = push->origin[d]; pp.length[d] = push->length[d]; }
    pp.V0 = push->V0; pp.alpha = push->alpha; pp.Bext = push->Bext; pp.DrInv = push->DrInv;
    (which == 1 ? orc_penning_kick1 : orc_penning_kick2)(&pp, n, x, y, z, px, py, pz, ex, ey, ez);
    return IPPLB_OK;
}
int ipplb_field_energy_stats(ipplb_ctx*, const ipplb_mesh* mesh, const double* ef, double out[7]) {
    int e[3];
    ext_of(mesh, e);
    const int g = mesh->nghost;
    for (int i = 0; i < 7; ++i) out[i] = 0.0;
    for (int k = g; k < e[2] - g; ++k)
        for (int j = g; j < e[1] - g; ++j)
            for (int i = g; i < e[0] - g; ++i) {
                const double* v = ef + 3 * (i + (long)e[0] * (j + (long)e[1] * k));
                double dot = 0.0;
                for (int d = 0; d < 3; ++d) {
                    out[d] += v[d] * v[d];
                    out[3 + d] = std::fabs(v[d]) > out[3 + d] ? std::fabs(v[d]) : out[3 + d];
                    dot += v[d] * v[d];
                }
                out[6] += dot;
            }
    return IPPLB_OK;
}
int ipplb_field_ex_stats(ipplb_ctx* c, const ipplb_mesh* mesh, const double* ef, double* out) {
    double all[7];
    ipplb_field_energy_stats(c, mesh, ef, all);
    out[0] = all[0];
    out[1] = all[3];
    return IPPLB_OK;
}
int ipplb_field_norm_stats(ipplb_ctx*, const ipplb_mesh* mesh, const double* f, double out[2]) {
    int e[3];
    ext_of(mesh, e);
    const int g = mesh->nghost;
    out[0] = out[1] = 0.0;
    for (int k = g; k < e[2] - g; ++k)
        for (int j = g; j < e[1] - g; ++j)
            for (int i = g; i < e[0] - g; ++i) {
                const double v = f[i + (long)e[0] * (j + (long)e[1] * k)];
                out[0] += v * v;
                out[1] = std::fabs(v) > out[1] ? std::fabs(v) : out[1];
            }
    return IPPLB_OK;
}
int ipplb_particles_kinetic(ipplb_ctx*, long n, const double* px, const double* py, const double* pz, double* out) {
    double s = 0.0;
    for (long i = 0; i < n; ++i) s += px[i] * px[i] + py[i] * py[i] + pz[i] * pz[i];
    *out = s;
    return IPPLB_OK;
}

// ---- particle initialisation (host restatement of ippl_b200/csrc/sample.cu: same formulas, same Philox stream) ----------------
static double m_cdf(const ipplb_dist* D, int d, double x) {
    switch (D->kind[d]) {
        case IPPLB_DIST_COSINE: return x + (D->par[2 * d] / D->par[2 * d + 1]) * std::sin(D->par[2 * d + 1] * x);
        case IPPLB_DIST_NORMAL: return 0.5 * (1 + std::erf((x - D->par[2 * d]) / (D->par[2 * d + 1] * std::sqrt(2.0))));
        default: return x;
    }
}
static double m_pdf(const ipplb_dist* D, int d, double x) {
    switch (D->kind[d]) {
        case IPPLB_DIST_COSINE: return 1.0 + D->par[2 * d] * std::cos(D->par[2 * d + 1] * x);
        case IPPLB_DIST_NORMAL: {
            const double pi = 3.14159265358979323846, mean = D->par[2 * d], sd = D->par[2 * d + 1];
            return (1.0 / (sd * std::sqrt(2 * pi))) * std::exp(-(x - mean) * (x - mean) / (2 * sd * sd));
        }
        default: return 1.0;
    }
}
int ipplb_sample_counts(const ipplb_dist* D, const double rmin[3], const double rmax[3], const double* regions, int nranks,
                        long ntotal, long* nlocal, double* ub) {
    unsigned long nglobal = 0;
    for (int r = 0; r < nranks; ++r) {
        double pnr = 1.0, pdr = 1.0;
        for (int d = 0; d < 3; ++d) {
            const double lmin = regions[r * 6 + d], lmax = regions[r * 6 + 3 + d];
            pnr *= m_cdf(D, d, lmax) - m_cdf(D, d, lmin);
            pdr *= m_cdf(D, d, rmax[d]) - m_cdf(D, d, rmin[d]);
            if (ub) {
                ub[r * 6 + d]     = m_cdf(D, d, lmin);
                ub[r * 6 + 3 + d] = m_cdf(D, d, lmax);
            }
        }
        nlocal[r] = (long)(unsigned long)((pnr / pdr) * ntotal);
        nglobal += (unsigned long)nlocal[r];
    }
    const int rest = (int)((unsigned long)ntotal - nglobal);
    for (int r = 0; r < nranks; ++r)
        if (r < rest) ++nlocal[r];
    return IPPLB_OK;
}
int ipplb_sample_positions(ipplb_ctx*, const ipplb_dist* D, const double umin[3], const double umax[3], uint64_t seed, long first_id,
                           long n, double* x, double* y, double* z) {
    double* out[3] = {x, y, z};
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            const double u01 = philox_uniform(seed, (unsigned long long)(first_id + i), (unsigned)d, 0);
            const double u   = umin[d] + (umax[d] - umin[d]) * u01;
            double s = D->kind[d] == IPPLB_DIST_NORMAL ? D->par[2 * d] + 0. * u * D->par[2 * d + 1] : u + D->par[d] * 0.;
            unsigned iter = 0;
            while (iter < 20u && std::fabs(m_cdf(D, d, s) - u) > 1e-12) {
                s = s - ((m_cdf(D, d, s) - u) / m_pdf(D, d, s));
                iter += 1;
            }
            out[d][i] = s;
        }
    return IPPLB_OK;
}
int ipplb_sample_normal(ipplb_ctx*, const double mu[3], const double sd[3], uint64_t seed, long first_id, long n, double* px,
                        double* py, double* pz) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        double g[3];
        philox_normal3(seed, (unsigned long long)(first_id + i), g);
        px[i] = mu[0] + sd[0] * g[0];
        py[i] = mu[1] + sd[1] * g[1];
        pz[i] = mu[2] + sd[2] * g[2];
    }
    return IPPLB_OK;
}
int ipplb_field_fill_pdf(ipplb_ctx*, const ipplb_mesh* m, const ipplb_dist* D, double* f) {
    const int g = m->nghost;
    const long ex = m->nl[0] + 2 * g, ey = m->nl[1] + 2 * g;
    for (int k = 0; k < m->nl[2]; ++k)
        for (int j = 0; j < m->nl[1]; ++j)
            for (int i = 0; i < m->nl[0]; ++i) {
                const int c[3] = {i, j, k};
                double total = 1.0;
                for (int d = 0; d < 3; ++d) total *= m_pdf(D, d, ((double)(c[d] + m->first[d]) + 0.5) * m->h[d] + m->origin[d]);
                f[(i + g) + ex * ((j + g) + ey * (k + g))] = total;
            }
    return IPPLB_OK;
}
// the bucketed store behind the fused step, emulated: "buckets" are plain contiguous arrays, one step = the reference-order
// sequence gather, kick(s), drift, periodic BC, scatter on the oracle (leapfrog only) -- enough to exercise the HOST logic that
// drives ipplb_bins_* (the facade's lazy-fusion engine); the real store and kernel are CUDA only.
int ipplb_bins_create(ipplb_ctx*, const ipplb_mesh* mesh, long capacity, ipplb_bins** out) {
    ipplb_bins* b = new ipplb_bins();
    b->m          = *mesh;
    b->capacity   = capacity;
    *out          = b;
    return IPPLB_OK;
}
int ipplb_bins_destroy(ipplb_bins* b) {
    delete b;
    return IPPLB_OK;
}
static void copy6(const ipplb_particles* in, ipplb_particles* out, long n) {
    const double* s[6] = {in->x, in->y, in->z, in->px, in->py, in->pz};
    double* d[6]       = {out->x, out->y, out->z, out->px, out->py, out->pz};
    for (int a = 0; a < 6; ++a) std::memcpy(d[a], s[a], sizeof(double) * n);
}
int ipplb_bins_build(ipplb_ctx*, ipplb_bins* b, const ipplb_particles* in, ipplb_particles* out) {
    if (in->n > b->capacity || out->capacity < in->n) return fail(IPPLB_ERR_CAPACITY, "mock bins_build: capacity");
    copy6(in, out, in->n);
    out->n        = in->n;
    out->q_scalar = in->q_scalar;
    return IPPLB_OK;
}
int ipplb_bins_step(ipplb_ctx*, ipplb_bins* b, const ipplb_push* push, const ipplb_particles* cur, ipplb_particles* nxt, const double* efield,
                    double* rho, double* exit_buf, int exit_cap, const double* rmin, const double* rmax) {
    if (push->kind != IPPLB_PUSH_LEAPFROG) return fail(IPPLB_ERR_ARG, "mock bins_step: leapfrog only");
    const long n = cur->n;
    const orc_mesh m = to_orc(&b->m);
    std::vector<double> w[6], e0(n), e1(n), e2(n);
    const double* src[6] = {cur->x, cur->y, cur->z, cur->px, cur->py, cur->pz};
    for (int a = 0; a < 6; ++a) w[a].assign(src[a], src[a] + n);
    double* E[3] = {e0.data(), e1.data(), e2.data()};
    orc_gather_cic(&m, n, w[0].data(), w[1].data(), w[2].data(), efield, 3, E, 0, 1);
    const double c = 0.5 * push->dt;
    for (int k = 0; k < (push->do_kick2 ? 1 : 0) + (push->do_kick1 ? 1 : 0); ++k)
        for (int d = 0; d < 3; ++d)
            for (long i = 0; i < n; ++i) w[3 + d][i] = w[3 + d][i] - c * E[d][i];
    for (int d = 0; d < 3; ++d) {
        if (push->do_drift)
            for (long i = 0; i < n; ++i) w[d][i] = w[d][i] + push->dt * w[3 + d][i];
        if (push->do_bc) orc_periodic_bc(n, w[d].data(), 0 * b->m.h[d] + b->m.origin[d], b->m.ng[d] * b->m.h[d] + b->m.origin[d], 1);
    }
    // ownership (positionInRegion: min < x <= max): stayers go to nxt and are deposited, leavers become records in exit_buf
    double* dst[6] = {nxt->x, nxt->y, nxt->z, nxt->px, nxt->py, nxt->pz};
    long ns = 0;
    b->nexit = 0;
    for (long i = 0; i < n; ++i) {
        bool mine = true;
        if (rmin && rmax)
            for (int d = 0; d < 3; ++d) mine = mine && w[d][i] > rmin[d] && w[d][i] <= rmax[d];
        if (mine) {
            for (int a = 0; a < 6; ++a) dst[a][ns] = w[a][i];
            ++ns;
        } else {
            if (!exit_buf || b->nexit >= exit_cap) return fail(IPPLB_ERR_CAPACITY, "mock bins_step: exit buffer too small");
            for (int a = 0; a < 6; ++a) exit_buf[6 * b->nexit + a] = w[a][i];
            ++b->nexit;
        }
    }
    orc_scatter_cic(&m, 0, ns, nxt->x, nxt->y, nxt->z, nullptr, cur->q_scalar, nullptr, rho, 1);
    nxt->n = ns;
    return IPPLB_OK;
}
int ipplb_bins_compact(ipplb_ctx*, ipplb_bins*, const ipplb_particles* cur, ipplb_particles* out) {
    if (out->capacity < cur->n) return fail(IPPLB_ERR_CAPACITY, "mock bins_compact: capacity");
    copy6(cur, out, cur->n);
    out->n = cur->n;
    return IPPLB_OK;
}
int ipplb_bins_migrate(ipplb_ctx* c, ipplb_bins* b, ipplb_particles* cur, const double* exit_buf, int, double* rho, long*, long*);   // below (needs the transport)
int ipplb_bins_kinetic(ipplb_ctx*, ipplb_bins*, const ipplb_particles* cur, double* out) {
    return ipplb_particles_kinetic(nullptr, cur->n, cur->px, cur->py, cur->pz, out);
}

// ---- the communicator entry points (one rank: trivial; several ranks: processes exchanging through IPPLB_MOCK_DIR) ----------------
int ipplb_nccl_unique_id(char* id) {
    std::memset(id, 0, IPPLB_NCCL_ID_BYTES);
    return IPPLB_OK;
}
int ipplb_comm_init(ipplb_ctx* c, int rank, int nranks, const char*) {
    c->rank   = rank;
    c->nranks = nranks;
    if (nranks > 1) {
        const char* d = std::getenv("IPPLB_MOCK_DIR");
        if (!d) return fail(IPPLB_ERR_ARG, "mock: several ranks need IPPLB_MOCK_DIR (a directory shared by the rank processes)");
        c->dir = d;
    }
    return IPPLB_OK;
}
int ipplb_allreduce_sum_f64(ipplb_ctx* c, double* v) {
    allreduce(c, v, [](double a, double b) { return a + b; });
    return IPPLB_OK;
}
int ipplb_allreduce_sum_i64(ipplb_ctx* c, long* v) {
    allreduce(c, v, [](long a, long b) { return a + b; });
    return IPPLB_OK;
}
int ipplb_allreduce_max_f64(ipplb_ctx* c, double* v) {
    allreduce(c, v, [](double a, double b) { return a > b ? a : b; });
    return IPPLB_OK;
}
int ipplb_ctx_set_layout(ipplb_ctx* c, const ipplb_layout* l, const double origin[3], const double h[3]) {
    if (ipplb_layout_nranks(l) != c->nranks) return fail(IPPLB_ERR_ARG, "mock set_layout: layout rank count != communicator size");
    c->L           = l->L;
    c->have_layout = true;
    c->boxes.resize(6 * (size_t)c->nranks);
    ipplb_layout_boxes(l, c->boxes.data());
    c->regions.resize(6 * (size_t)c->nranks);
    for (int d = 0; d < 3; ++d) { c->origin[d] = origin[d]; c->h[d] = h[d]; }
    orc_regions(c->L.ng, c->nranks, c->boxes.data(), origin, h, c->regions.data());
    return IPPLB_OK;
}
// BareField::accumulateHalo / fillHalo over ranks: every rank sees every rank's ghosted field, runs the oracle's all-ranks
// exchange and keeps its own; then the in-rank periodic wrap of the dimensions it owns entirely (oracle.halo_full)
int ipplb_halo_exchange(ipplb_ctx* c, double* field, int ncomp, int mode) {
    if (!c->have_layout) return fail(IPPLB_ERR_ARG, "mock halo_exchange: no layout bound");
    const int g = c->L.nghost, me = c->rank;
    size_t cells = 1;
    for (int d = 0; d < 3; ++d) cells *= (size_t)(box_len(c, me, d) + 2 * g);
    auto all = allgather(c, field, sizeof(double) * cells * ncomp);
    std::vector<double*> f(c->nranks);
    for (int r = 0; r < c->nranks; ++r) f[r] = (double*)all[r].data();
    if (c->nranks > 1) orc_halo_exchange(c->L.ng, c->nranks, c->boxes.data(), g, c->L.periodic, f.data(), ncomp, mode);
    if (c->L.periodic) {
        int ext[3], serial[3];
        for (int d = 0; d < 3; ++d) {
            ext[d]    = box_len(c, me, d) + 2 * g;
            serial[d] = box_len(c, me, d) == c->L.ng[d];
        }
        orc_halo_periodic(f[me], ext, ncomp, g, serial, mode);
    }
    std::memcpy(field, f[me], sizeof(double) * cells * ncomp);
    return IPPLB_OK;
}
// ParticleSpatialLayout::update behind the facade's two-phase migrate (the BC has been applied by the caller): the plan locates
// and exchanges the leavers' records, the commit removes the leavers with the reference's hole filling and appends the arrivals
// in ascending source rank (oracle.update)
int ipplb_update_plan(ipplb_ctx* c, ipplb_particles* p, long* n_after, long* sent, long* recv) {
    if (n_after) *n_after = p->n;
    if (c->nranks < 2) return IPPLB_OK;
    if (!c->have_layout) return fail(IPPLB_ERR_ARG, "mock update_plan: no layout bound");
    const long n = p->n;
    c->dest.assign((size_t)n, c->rank);
    orc_locate(c->nranks, c->regions.data(), c->rank, n, p->x, p->y, p->z, c->dest.data());
    std::vector<double> out;   // records: dest, x, y, z, px, py, pz, q
    for (long i = 0; i < n; ++i)
        if (c->dest[i] != c->rank) {
            const double rec[8] = {(double)c->dest[i], p->x[i], p->y[i], p->z[i], p->px ? p->px[i] : 0.0, p->py ? p->py[i] : 0.0,
                                   p->pz ? p->pz[i] : 0.0, p->q ? p->q[i] : p->q_scalar};
            out.insert(out.end(), rec, rec + 8);
        }
    const auto all = allgather(c, out.data(), sizeof(double) * out.size());
    c->arrivals.clear();
    for (int r = 0; r < c->nranks; ++r) {
        const double* rec = (const double*)all[r].data();
        const size_t cnt  = all[r].size() / (8 * sizeof(double));
        long from_r = 0;
        for (size_t k = 0; k < cnt; ++k)
            if ((int)rec[8 * k] == c->rank && r != c->rank) {
                c->arrivals.insert(c->arrivals.end(), rec + 8 * k + 1, rec + 8 * k + 8);
                ++from_r;
            }
        if (recv) recv[r] = from_r;
    }
    if (sent)
        for (int r = 0; r < c->nranks; ++r) {
            sent[r] = 0;
            for (long i = 0; i < n; ++i) sent[r] += c->dest[i] == r && r != c->rank;
        }
    const long nleave = (long)(out.size() / 8), narrive = (long)(c->arrivals.size() / 7);
    c->plan_n = n;
    if (n_after) *n_after = n - nleave + narrive;
    long too_small = (n - nleave + narrive) > p->capacity ? 1 : 0;   // the outcome is collective, like the product's
    allreduce(c, &too_small, [](long a, long b) { return a + b; });
    return too_small ? fail(IPPLB_ERR_CAPACITY, "mock update_plan: %ld rank(s) cannot hold their particles after the migration", too_small) : IPPLB_OK;
}
int ipplb_update_commit(ipplb_ctx* c, ipplb_particles* p) {
    if (c->nranks < 2) return IPPLB_OK;
    if (c->plan_n != p->n) return fail(IPPLB_ERR_ARG, "mock update_commit: call ipplb_update_plan on the same particles first");
    const long n = p->n;
    long nd = 0;
    for (long i = 0; i < n; ++i) nd += c->dest[i] != c->rank;
    const long keep = n - nd, narrive = (long)(c->arrivals.size() / 7);
    long too_small = keep + narrive > p->capacity ? 1 : 0;
    allreduce(c, &too_small, [](long a, long b) { return a + b; });
    if (too_small) return fail(IPPLB_ERR_CAPACITY, "mock update_commit: capacity");
    double* arr[7] = {p->x, p->y, p->z, p->px, p->py, p->pz, p->q};
    // ParticleBase::internalDestroy: the i-th hole among the first n - nd slots takes the i-th survivor of the tail
    std::vector<long> holes, fill;
    for (long i = 0; i < keep; ++i)
        if (c->dest[i] != c->rank) holes.push_back(i);
    for (long i = keep; i < n; ++i)
        if (c->dest[i] == c->rank) fill.push_back(i);
    for (size_t k = 0; k < holes.size(); ++k)
        for (double* a : arr)
            if (a) a[holes[k]] = a[fill[k]];
    for (long k = 0; k < narrive; ++k)
        for (int a = 0; a < 7; ++a)
            if (arr[a]) arr[a][keep + k] = c->arrivals[7 * k + a];
    p->n      = keep + narrive;
    c->plan_n = -1;
    return IPPLB_OK;
}
int ipplb_update(ipplb_ctx* c, ipplb_particles* p, long* sent, long* recv) {
    const int rc = ipplb_update_plan(c, p, nullptr, sent, recv);
    return rc ? rc : ipplb_update_commit(c, p);
}
// ParticleSpatialLayout::update for the emulated bucketed store: destination of every leaver, exchange, append, deposit
int ipplb_bins_migrate(ipplb_ctx* c, ipplb_bins* b, ipplb_particles* cur, const double* exit_buf, int, double* rho, long*, long*) {
    if (c->nranks < 2) return IPPLB_OK;
    const long ne = b->nexit;
    std::vector<double> x(ne), y(ne), z(ne);
    for (long i = 0; i < ne; ++i) { x[i] = exit_buf[6 * i]; y[i] = exit_buf[6 * i + 1]; z[i] = exit_buf[6 * i + 2]; }
    std::vector<int> dest((size_t)ne, c->rank);
    orc_locate(c->nranks, c->regions.data(), c->rank, ne, x.data(), y.data(), z.data(), dest.data());
    std::vector<double> out;   // dest + 6 doubles
    for (long i = 0; i < ne; ++i) {
        out.push_back((double)dest[i]);
        out.insert(out.end(), exit_buf + 6 * i, exit_buf + 6 * i + 6);
    }
    const auto all = allgather(c, out.data(), sizeof(double) * out.size());
    double* dst[6] = {cur->x, cur->y, cur->z, cur->px, cur->py, cur->pz};
    const long n0 = cur->n;
    for (int r = 0; r < c->nranks; ++r) {
        const double* rec = (const double*)all[r].data();
        const size_t cnt  = all[r].size() / (7 * sizeof(double));
        for (size_t k = 0; k < cnt; ++k)
            if ((int)rec[7 * k] == c->rank) {
                if (cur->n >= cur->capacity) return fail(IPPLB_ERR_CAPACITY, "mock bins_migrate: capacity");
                for (int a = 0; a < 6; ++a) dst[a][cur->n] = rec[7 * k + 1 + a];
                ++cur->n;
            }
    }
    const orc_mesh m = to_orc(&b->m);
    if (rho && cur->n > n0) orc_scatter_cic(&m, n0, cur->n, cur->x, cur->y, cur->z, nullptr, cur->q_scalar, nullptr, rho, 1);
    b->nexit = 0;
    return IPPLB_OK;
}
// OrthogonalRecursiveBisection::binaryRepartition: the product's host state machine (ippl_b200/csrc/orb.cpp) fed with plane sums
// of every rank's interior, summed over the ranks
int ipplb_orb_repartition(ipplb_ctx* c, const ipplb_mesh* mesh, int nranks, const double* w, int* boxes_out, int* ok) {
    ipplb_orb* o = nullptr;
    int rc = ipplb_orb_begin(&o, mesh->ng, nranks);
    if (rc) return rc;
    const int g = mesh->nghost;
    const long ex = mesh->nl[0] + 2 * g, ey = mesh->nl[1] + 2 * g;
    for (;;) {
        int lo[3], hi[3], axis, pending;
        if ((rc = ipplb_orb_next(o, lo, hi, &axis, &pending)) || !pending) break;
        std::vector<double> red((size_t)(hi[axis] - lo[axis] + 1), 0.0);
        for (int k = 0; k < mesh->nl[2]; ++k)
            for (int j = 0; j < mesh->nl[1]; ++j)
                for (int i = 0; i < mesh->nl[0]; ++i) {
                    const int gi[3] = {i + mesh->first[0], j + mesh->first[1], k + mesh->first[2]};
                    bool in = true;
                    for (int d = 0; d < 3; ++d) in = in && gi[d] >= lo[d] && gi[d] <= hi[d];
                    if (in) red[gi[axis] - lo[axis]] += w[(i + g) + ex * ((j + g) + ey * (k + g))];
                }
        const auto all = allgather(c, red.data(), sizeof(double) * red.size());
        std::vector<double> tot(red.size(), 0.0);
        for (int r = 0; r < c->nranks; ++r)
            for (size_t t = 0; t < red.size(); ++t) tot[t] += ((const double*)all[r].data())[t];
        if ((rc = ipplb_orb_cut(o, tot.data(), (int)tot.size()))) break;
    }
    if (rc) {
        ipplb_orb_destroy(o);
        return rc;
    }
    return ipplb_orb_finish(o, boxes_out, ok);
}

}  // extern "C"
