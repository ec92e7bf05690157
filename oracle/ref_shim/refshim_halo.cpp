// refshim_halo.cpp -- TEST INFRASTRUCTURE.  Runs the REAL in-rank periodic halo code of the reference
// (Field/HaloCells.h + .hpp compiled in place from /root/reference/src, never copied): detail::HaloCells<double, 3>::
// applyPeriodicSerialDim<Op> with its HaloPeriodicFunctor and the reference's own assign / rhs_plus_assign operators --
// what BareField::fillHalo / accumulateHalo run for the un-split dimensions (BareField.hpp:152-172).  The inter-rank
// exchange code of the same header is parsed but never executed here (no MPI); Kokkos is replaced by serial stand-ins,
// Communicate/Archive.h and Utility/ParallelDispatch.h are skipped through their include guards.
#include <Kokkos_Core.hpp>

#include <array>
#include <cstddef>
#include <memory>
#include <utility>
#include <vector>

#define IPPL_ARCHIVE_H
#define IPPL_PARALLEL_DISPATCH_H   // Utility/ParallelDispatch.h pulls in the whole framework: its three names are below
#include "Types/IpplTypes.h"
namespace ippl { namespace detail {
    template <class MemorySpace> struct Archive {
        template <class V> void serialize(V&, size_type) {}
        template <class V> void deserialize(V&, size_type) {}
    };
}}  // namespace ippl::detail

#include "Utility/IpplException.h"
#include "Types/Vector.h"
#include "Types/ViewTypes.h"

#include "Index/NDIndex.h"
#include "FieldLayout/FieldLayout.h"

namespace refshim {
    // ghosted rank-3 array, x fastest (the layout BareField's Kokkos::View<T***, LayoutLeft> has)
    template <typename T> struct View3D {
        using value_type      = T;
        using execution_space = Kokkos::Serial;
        using memory_space    = Kokkos::HostSpace;
        static constexpr unsigned rank = 3;
        T* p = nullptr;
        long e[3] = {0, 0, 0};
        long extent(std::size_t d) const { return e[d]; }
        T& operator()(std::size_t i, std::size_t j, std::size_t k) const { return p[i + e[0] * (j + e[1] * k)]; }
    };
}  // namespace refshim
namespace ippl { namespace detail {
    template <typename T, class... P> struct ViewType<T, 3, P...> { using view_type = refshim::View3D<T>; };
}}  // namespace ippl::detail

namespace Kokkos {
    template <typename T, std::size_t N> struct Array {
        T v[N];
        T& operator[](std::size_t i) { return v[i]; }
        const T& operator[](std::size_t i) const { return v[i]; }
    };
    template <class A, class B> std::pair<A, B> make_pair(A a, B b) { return {a, b}; }
    template <class V, class... R> V subview(const V& v, R...) { return v; }   // parsed only
}  // namespace Kokkos

namespace ippl {
    inline mpi::Communicator* Comm = new mpi::Communicator();
    template <unsigned Dim, class... P> struct RangePolicy {
        using index_type       = long;
        using index_array_type = ippl::Vector<long, Dim>;
        Kokkos::Array<long, Dim> lo, hi;
    };
    template <unsigned Dim, class... P>
    RangePolicy<Dim, P...> createRangePolicy(const Kokkos::Array<long, Dim>& b, const Kokkos::Array<long, Dim>& e) {
        return RangePolicy<Dim, P...>{b, e};
    }
    template <class View> RangePolicy<3> getRangePolicy(const View& v, int shift = 0) {
        RangePolicy<3> p;
        for (int d = 0; d < 3; ++d) { p.lo[d] = shift; p.hi[d] = v.extent(d) - shift; }
        return p;
    }
    // ippl::parallel_for over a rank-3 box: the functor takes the index array by value (HaloPeriodicFunctor modifies it)
    template <class Policy, class F> void parallel_for(const char*, const Policy& p, const F& f) {
        typename Policy::index_array_type a;
        for (long k = p.lo[2]; k < p.hi[2]; ++k)
            for (long j = p.lo[1]; j < p.hi[1]; ++j)
                for (long i = p.lo[0]; i < p.hi[0]; ++i) {
                    a[0] = i; a[1] = j; a[2] = k;
                    f(a);
                }
    }
}  // namespace ippl

#include "Field/HaloCells.h"

extern "C" {

// applyPeriodicSerialDim on a ghosted scalar field (ext = local extents + 2 * nghost, x fastest) of a single-rank,
// all-periodic layout of ng cells: mode 0 = fillHalo's operator (assign), 1 = accumulateHalo's (rhs_plus_assign)
void refhalo_periodic(const int ng[3], int nghost, int mode, double* field) {
    refshim::g_rank = 0;
    refshim::g_size = 1;
    ippl::Index ix(ng[0]), iy(ng[1]), iz(ng[2]);
    ippl::NDIndex<3> domain(ix, iy, iz);
    std::array<bool, 3> par = {true, true, true};
    ippl::FieldLayout<3> fl(ippl::mpi::Communicator(), domain, par, true, nghost);
    using Halo = ippl::detail::HaloCells<double, 3>;
    Halo::view_type v;
    v.p = field;
    for (int d = 0; d < 3; ++d) v.e[d] = ng[d] + 2 * nghost;
    Halo h;
    if (mode == 0) h.applyPeriodicSerialDim<Halo::assign>(v, &fl, nghost);
    else h.applyPeriodicSerialDim<Halo::rhs_plus_assign>(v, &fl, nghost);
}

}  // extern "C"
