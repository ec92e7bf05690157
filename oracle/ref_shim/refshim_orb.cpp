// refshim_orb.cpp -- TEST INFRASTRUCTURE.  Runs the REAL OrthogonalRecursiveBisection of the reference
// (Decomposition/OrthogonalRecursiveBisection.h + .hpp, compiled in place from /root/reference/src, never copied):
// findCutAxis, findMedian, cutDomain, perpendicularReduction, binaryRepartition and FieldLayout::updateLayout execute
// the reference's own code; what is replaced is what they stand on -- Kokkos (serial stand-ins below and in this
// directory), the communicator (one process plays the whole communicator: the field it hands to ORB covers the global
// domain, so the all-reduce of the per-rank plane sums is the identity) and the Field class (a plain ghosted array with
// the handful of methods ORB calls).  Particle/ParticleAttrib.h and Particle/ParticleSpatialLayout.h, which the ORB
// header includes but does not use by name, are skipped through their include guards.
#include <Kokkos_Core.hpp>

#include <array>
#include <cstddef>
#include <functional>
#include <numeric>
#include <vector>

#define IPPL_PARTICLE_ATTRIB_H
#define IPPL_PARTICLE_SPATIAL_LAYOUT_H

#include "Utility/IpplException.h"
#include "Types/Vector.h"

#include "Index/NDIndex.h"
#include "Interpolation/CIC.h"
#include "Region/NDRegion.h"

#include "FieldLayout/FieldLayout.h"

namespace Kokkos {
    template <typename T, std::size_t N> struct Array {
        T v[N];
        T& operator[](std::size_t i) { return v[i]; }
        const T& operator[](std::size_t i) const { return v[i]; }
    };
    template <typename T> struct Sum {
        T& ref;
        explicit Sum(T& r) : ref(r) {}
    };
    template <class A, class B> struct SpaceAccessibility { static constexpr bool accessible = true; };
    template <class... P> struct RangePolicy {
        std::size_t b, e;
        RangePolicy(std::size_t b_, std::size_t e_) : b(b_), e(e_) {}
    };
    template <class... P, class F> void parallel_for(const char*, RangePolicy<P...> r, const F& f) {
        for (std::size_t i = r.b; i < r.e; ++i) f(i);
    }
}  // namespace Kokkos

namespace ippl {
    // what OrthogonalRecursiveBisection.hpp uses of Utility/ParallelDispatch.h (ippl::apply is the reference's own,
    // Expression/IpplOperations.h)
    template <unsigned Dim, class... P> struct RangePolicy {
        using index_type       = long;
        using index_array_type = ippl::Vector<long, Dim>;
        Kokkos::Array<long, Dim> lo, hi;
    };
    template <unsigned Dim, class... P>
    RangePolicy<Dim, P...> createRangePolicy(const Kokkos::Array<long, Dim>& b, const Kokkos::Array<long, Dim>& e) {
        return RangePolicy<Dim, P...>{b, e};
    }
    template <class Policy, class F, class T> void parallel_reduce(const char*, const Policy& p, const F& f, Kokkos::Sum<T> red) {
        T acc = T(0);
        typename Policy::index_array_type a;
        for (a[2] = p.lo[2]; a[2] < p.hi[2]; ++a[2])
            for (a[1] = p.lo[1]; a[1] < p.hi[1]; ++a[1])
                for (a[0] = p.lo[0]; a[0] < p.hi[0]; ++a[0]) f(a, acc);
        red.ref = acc;
    }
}  // namespace ippl

namespace {
    struct View3 {
        static constexpr unsigned rank = 3;
        using value_type               = double;
        double* p;
        long e0, e1;
        double& operator()(std::size_t i, std::size_t j, std::size_t k) const { return p[i + e0 * (j + e1 * k)]; }
    };
    struct FakeMesh {
        using vector_type = ippl::Vector<double, 3>;
        vector_type hx, origin;
        const vector_type& getMeshSpacing() const { return hx; }
        const vector_type& getOrigin() const { return origin; }
    };
    // the Field ORB is instantiated with: a ghosted array over the WHOLE domain (this process is every rank at once)
    struct FakeField {
        static constexpr unsigned dim = 3;
        using Mesh_t          = FakeMesh;
        using value_type      = double;
        using execution_space = Kokkos::Serial;
        using memory_space    = Kokkos::HostSpace;
        std::vector<double> data;
        ippl::NDIndex<3> owned;
        FakeMesh* mesh           = nullptr;
        ippl::FieldLayout<3>* fl = nullptr;
        int nghost               = 1;
        void initialize(FakeMesh& m, ippl::FieldLayout<3>& l) {
            mesh  = &m;
            fl    = &l;
            owned = l.getDomain();
            data.assign((std::size_t)(owned[0].length() + 2) * (owned[1].length() + 2) * (owned[2].length() + 2), 0.0);
        }
        FakeField& operator=(double v) {
            std::fill(data.begin(), data.end(), v);
            return *this;
        }
        ippl::NDIndex<3> getOwned() const { return owned; }
        int getNghost() const { return nghost; }
        View3 getView() const { return View3{const_cast<double*>(data.data()), (long)owned[0].length() + 2, (long)owned[1].length() + 2}; }
        const FakeMesh& get_mesh() const { return *mesh; }
        const ippl::FieldLayout<3>& getLayout() const { return *fl; }
        void updateLayout(ippl::FieldLayout<3>&) {}  // the weights are consumed before the layout changes
        void accumulateHalo() {}
    };
    struct Positions {  // what scatterR asks of a particle attribute
        using memory_space = Kokkos::HostSpace;
        struct V {
            ippl::Vector<double, 3>* p;
            ippl::Vector<double, 3>& operator()(std::size_t i) const { return p[i]; }
        };
        std::vector<ippl::Vector<double, 3>> r;
        V getView() const { return V{const_cast<ippl::Vector<double, 3>*>(r.data())}; }
        std::size_t getParticleCount() const { return r.size(); }
    };
    using NoParticles = Positions;  // binaryRepartition below runs with isFirstRepartition = true: scatterR is not executed
}  // namespace

#include "Decomposition/OrthogonalRecursiveBisection.h"

extern "C" {

// findMedian(w), OrthogonalRecursiveBisection.hpp:185-216
int reforb_find_median(const double* w, int n) {
    ippl::OrthogonalRecursiveBisection<FakeField, double> orb;
    std::vector<double> v(w, w + n);
    return orb.findMedian(v);
}

// binaryRepartition on the global interior weights w[z][y][x] for `nranks` ranks: boxes_out[nranks][6] = lo[3], hi[3];
// returns 1 when the repartition was accepted (no box with an axis of length 1), 0 otherwise
int reforb_repartition(const int ng[3], int nranks, const double* w, int* boxes_out) {
    refshim::g_rank = 0;
    refshim::g_size = nranks;
    ippl::Index ix(ng[0]), iy(ng[1]), iz(ng[2]);
    ippl::NDIndex<3> domain(ix, iy, iz);
    std::array<bool, 3> par = {true, true, true};
    ippl::FieldLayout<3> fl(ippl::mpi::Communicator(), domain, par, true, 1);
    FakeMesh mesh;
    ippl::OrthogonalRecursiveBisection<FakeField, double> orb;
    orb.bf_m.initialize(mesh, fl);
    View3 v = orb.bf_m.getView();
    for (int k = 0; k < ng[2]; ++k)
        for (int j = 0; j < ng[1]; ++j)
            for (int i = 0; i < ng[0]; ++i) v(i + 1, j + 1, k + 1) = w[i + (std::size_t)ng[0] * (j + (std::size_t)ng[1] * k)];
    NoParticles none;
    const bool first = true;
    const bool ok    = orb.binaryRepartition(none, fl, first);
    const auto& doms = fl.getHostLocalDomains();
    for (int r = 0; r < nranks; ++r)
        for (int d = 0; d < 3; ++d) {
            boxes_out[6 * r + d]     = doms(r)[d].first();
            boxes_out[6 * r + 3 + d] = doms(r)[d].last();
        }
    refshim::g_size = 1;
    return ok ? 1 : 0;
}

// scatterR(R), OrthogonalRecursiveBisection.hpp:234-300: the reference's own particle loop -- l = (R - origin) * invdx +
// 0.5, index = (int) l, whi = l - index, wlo = 1 - whi, args = index - lDom.first() + nghost, scatterToField with weight
// 1 -- on a single-rank layout of ng cells; field_out is the ghosted (ng + 2)^3 array, x fastest, BEFORE any halo
// accumulation.  The same loop body as ParticleAttrib::scatter (ParticleAttrib.hpp:167-184).
void reforb_scatterR(const int ng[3], const double origin[3], const double h[3], long n, const double* x, const double* y,
                     const double* z, double* field_out) {
    refshim::g_rank = 0;
    refshim::g_size = 1;
    ippl::Index ix(ng[0]), iy(ng[1]), iz(ng[2]);
    ippl::NDIndex<3> domain(ix, iy, iz);
    std::array<bool, 3> par = {true, true, true};
    ippl::FieldLayout<3> fl(ippl::mpi::Communicator(), domain, par, true, 1);
    FakeMesh mesh;
    for (int d = 0; d < 3; ++d) { mesh.hx[d] = h[d]; mesh.origin[d] = origin[d]; }
    ippl::OrthogonalRecursiveBisection<FakeField, double> orb;
    orb.bf_m.initialize(mesh, fl);
    Positions R;
    R.r.resize((std::size_t)n);
    for (long i = 0; i < n; ++i) { R.r[i][0] = x[i]; R.r[i][1] = y[i]; R.r[i][2] = z[i]; }
    orb.scatterR(R);
    for (std::size_t i = 0; i < orb.bf_m.data.size(); ++i) field_out[i] = orb.bf_m.data[i];
}

}  // extern "C"
