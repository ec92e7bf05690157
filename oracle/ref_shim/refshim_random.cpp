// refshim_random.cpp -- TEST INFRASTRUCTURE.  extern "C" entry points around the REAL reference headers of the sampling
// path, compiled in place from /root/reference/src (never copied):
//   Random/NormalDistribution.h   normal_cdf_func / normal_pdf_func / normal_estimate_func, NormalDistribution
//   Random/Distribution.h         Distribution<T, Dim, DimP, Functions>: getCdf / getPdf / getEstimate / getObjFunc /
//                                 getDerObjFunc / getFullPdf
//   Random/Utility.h              detail::NewtonRaphson::solve
//   Random/InverseTransformSampling.h   the constructor (updateBounds: rank counts + CDF bounds), generate() and
//                                 fill_random::operator() (u = drand(umin, umax); estimate; Newton)
//   Random/Randn.h                randn::operator()
// The uniform / normal draws are REPLAYED from caller-supplied arrays (Kokkos_Random.hpp stand-in): what is compared
// with the restatement is the reference's arithmetic on identical random numbers.  The alpine managers' custom
// distribution functors live in demos/alpine/*Manager.h, which cannot be included (they pull in the whole framework):
// struct CustomDistributionFunctions of LandauDampingManager.h (:21-44) and of BumponTailInstabilityManager.h (:23-52)
// are cut out of those files at build time and compiled here unchanged.
#include <Kokkos_Core.hpp>
#include <Kokkos_Random.hpp>

#include <cstddef>
#include <functional>
#include <numeric>

#include "Types/Vector.h"
#include "Types/ViewTypes.h"

namespace Kokkos {   // serial stand-ins for the two launch forms generate() uses
    template <class... P> struct RangePolicy {
        std::size_t b, e;
        RangePolicy(std::size_t b_, std::size_t e_) : b(b_), e(e_) {}
    };
    template <class F> void parallel_for(std::size_t n, const F& f) { for (std::size_t i = 0; i < n; ++i) f(i); }
    template <class... P, class F> void parallel_for(RangePolicy<P...> r, const F& f) { for (std::size_t i = r.b; i < r.e; ++i) f(i); }
}  // namespace Kokkos

namespace refshim { inline int r_rank = 0; inline std::size_t r_nglobal = 0; }
namespace ippl {
    namespace detail { using size_type = std::size_t; }
    namespace mpi {
        // one process plays every rank in turn: allreduce of the local sample count returns the preset global sum
        struct RandomComm {
            int rank() const { return refshim::r_rank; }
            template <typename T, class Op> void allreduce(T*, T* out, int, Op) { *out = (T)refshim::r_nglobal; }
        };
    }
    inline mpi::RandomComm* Comm = new mpi::RandomComm();
}  // namespace ippl

#include "Random/Distribution.h"
#include "Random/InverseTransformSampling.h"
#include "Random/NormalDistribution.h"
#include "Random/Randn.h"
#include "Random/UniformDistribution.h"

// the managers' own functor structs, cut out of the reference files at build time (gen_snippets.py -> a temporary include directory)
namespace ref_landau {
#include "landau_dist.inc"
}
namespace ref_bumpontail {
constexpr unsigned Dim = 3;   // the driver-level constant the struct refers to (demos/alpine/BumponTailInstability.cpp)
#include "bumpontail_dist.inc"
}

namespace {
    using CosineFunctions = ref_landau::CustomDistributionFunctions;
    using BumpDist        = ippl::random::Distribution<double, 3, 6, ref_bumpontail::CustomDistributionFunctions>;
    using CosDist  = ippl::random::Distribution<double, 3, 6, CosineFunctions>;
    using NormDist = ippl::random::NormalDistribution<double, 3>;

    // the shape InverseTransformSampling's constructor asks of a RegionLayout: regions(rank)[d].min() / .max()
    struct FakeRegion {
        double lo, hi;
        double min() const { return lo; }
        double max() const { return hi; }
    };
    struct FakeRegions {
        const double* r;  // [nranks][6] = min[3], max[3]
        struct Row {
            const double* p;
            FakeRegion operator[](unsigned d) const { return FakeRegion{p[d], p[3 + d]}; }
        };
        Row operator()(int rank) const { return Row{r + 6 * rank}; }
    };
    struct FakeRegionLayout {
        using host_mirror_type = FakeRegions;
        FakeRegions regs;
        host_mirror_type gethLocalRegions() const { return regs; }
    };

    // the real constructor (updateBounds) for one rank, then the real generate() with replayed uniforms
    template <class Dist>
    void run_sampling(Dist& dist, const double* rmin, const double* rmax, const double* regions, int nranks, long ntotal,
                      long* nlocal_out, double* ubounds_out, int gen_rank, const double* u01, double* out) {
        using view_type = Kokkos::View<ippl::Vector<double, 3>*>;
        using ITS       = ippl::random::InverseTransformSampling<double, 3, Kokkos::Serial, Dist>;
        ippl::Vector<double, 3> lo, hi;
        for (int d = 0; d < 3; ++d) { lo[d] = rmin[d]; hi[d] = rmax[d]; }
        FakeRegionLayout rl{FakeRegions{regions}};
        std::size_t nt = (std::size_t)ntotal;
        // pass 1: raw counts (global sum preset to ntotal -> no remainder is handed out); pass 2: with the true sum
        std::size_t sum = 0;
        refshim::r_nglobal = nt;
        for (int r = 0; r < nranks; ++r) {
            refshim::r_rank = r;
            ITS s(dist, hi, lo, rl, nt);
            sum += s.getLocalSamplesNum();
        }
        refshim::r_nglobal = sum;
        for (int r = 0; r < nranks; ++r) {
            refshim::r_rank = r;
            ITS s(dist, hi, lo, rl, nt);
            nlocal_out[r] = (long)s.getLocalSamplesNum();
            for (int d = 0; d < 3; ++d) {
                ubounds_out[6 * r + d]     = s.umin_m[d];
                ubounds_out[6 * r + 3 + d] = s.umax_m[d];
            }
            if (r == gen_rank && out) {
                const std::size_t n = s.getLocalSamplesNum();
                view_type x("x", n);
                std::size_t cursor = 0;
                Kokkos::Random_XorShift64_Pool<> pool{u01, &cursor};   // u01[d][n]: generate() makes one pass per dimension
                s.generate(x, pool);
                for (std::size_t i = 0; i < n; ++i)
                    for (int d = 0; d < 3; ++d) out[(std::size_t)d * n + i] = x(i)[d];
            }
        }
        refshim::r_rank = 0;
    }
}  // namespace

extern "C" {

// kind 1: LandauDamping's functors, 2: NormalDistribution (PenningTrap), 3: BumponTail's functors.  which: 0 cdf, 1 pdf, 2 estimate, 3 objective(x, u = aux), 4 d objective
double refrand_eval(int kind, const double* par, int which, int d, double x, double aux) {
    if (kind == 3) {
        BumpDist D(par);
        switch (which) { case 0: return D.getCdf(x, d); case 1: return D.getPdf(x, d); case 2: return D.getEstimate(x, d);
                         case 3: return D.getObjFunc(x, d, aux); default: return D.getDerObjFunc(x, d); }
    }
    if (kind == 2) {
        NormDist D(par);
        switch (which) { case 0: return D.getCdf(x, d); case 1: return D.getPdf(x, d); case 2: return D.getEstimate(x, d);
                         case 3: return D.getObjFunc(x, d, aux); default: return D.getDerObjFunc(x, d); }
    }
    CosDist D(par);
    switch (which) { case 0: return D.getCdf(x, d); case 1: return D.getPdf(x, d); case 2: return D.getEstimate(x, d);
                     case 3: return D.getObjFunc(x, d, aux); default: return D.getDerObjFunc(x, d); }
}

double refrand_full_pdf(int kind, const double* par, const double x[3]) {
    ippl::Vector<double, 3> v;
    for (int d = 0; d < 3; ++d) v[d] = x[d];
    if (kind == 2) { NormDist D(par); return D.getFullPdf(v); }
    if (kind == 3) { BumpDist D(par); return D.getFullPdf(v); }
    CosDist D(par);
    return D.getFullPdf(v);
}

// NewtonRaphson::solve on one value; returns the solution
double refrand_newton(int kind, const double* par, int d, double x0, double u) {
    double x = x0;
    if (kind == 2) { NormDist D(par); ippl::random::detail::NewtonRaphson<double, NormDist> s(D); s.solve(d, x, u); return x; }
    if (kind == 3) { BumpDist D(par); ippl::random::detail::NewtonRaphson<double, BumpDist> s(D); s.solve(d, x, u); return x; }
    CosDist D(par);
    ippl::random::detail::NewtonRaphson<double, CosDist> s(D);
    s.solve(d, x, u);
    return x;
}

// InverseTransformSampling(dist, rmax, rmin, rlayout, ntotal) for every rank (counts + CDF bounds, :48-61, 106-131) and,
// for rank gen_rank, generate() (:235-244) with the uniforms u01[d][nlocal] replayed; out[d][nlocal] (may be NULL)
void refrand_sampling(int kind, const double* par, const double* rmin, const double* rmax, const double* regions, int nranks,
                      long ntotal, long* nlocal_out, double* ubounds_out, int gen_rank, const double* u01, double* out) {
    if (kind == 2) { NormDist D(par); run_sampling(D, rmin, rmax, regions, nranks, ntotal, nlocal_out, ubounds_out, gen_rank, u01, out); return; }
    if (kind == 3) { BumpDist D(par); run_sampling(D, rmin, rmax, regions, nranks, ntotal, nlocal_out, ubounds_out, gen_rank, u01, out); return; }
    CosDist D(par);
    run_sampling(D, rmin, rmax, regions, nranks, ntotal, nlocal_out, ubounds_out, gen_rank, u01, out);
}

// randn::operator(): v(i)[d] = mu[d] + sd[d] * normal(0, 1) with the standard normals replayed from g[n][3]
void refrand_randn(const double* mu, const double* sd, long n, const double* g, double* out) {
    using view_type = Kokkos::View<ippl::Vector<double, 3>*>;
    using pool_type = Kokkos::Random_XorShift64_Pool<>;
    view_type v("v", (std::size_t)n);
    std::size_t cursor = 0;
    pool_type pool{g, &cursor};
    double m[3] = {mu[0], mu[1], mu[2]}, s[3] = {sd[0], sd[1], sd[2]};
    ippl::random::randn<double, 3> f(v, pool, m, s);
    for (long i = 0; i < n; ++i) f((std::size_t)i);
    for (long i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) out[(std::size_t)i * 3 + d] = v(i)[d];
}

}  // extern "C"
