// refshim_halo.cpp -- TEST INFRASTRUCTURE.  Runs the REAL in-rank periodic halo code of the reference
// (Field/HaloCells.h + .hpp compiled in place from /root/reference/src, never copied): detail::HaloCells<double, 3>::
// applyPeriodicSerialDim<Op> with its HaloPeriodicFunctor and the reference's own assign / rhs_plus_assign operators --
// what BareField::fillHalo / accumulateHalo run for the un-split dimensions (BareField.hpp:152-172) -- and the inter-rank
// exchange of the same header, HaloCells::exchangeBoundaries with pack / unpack (HaloCells.hpp:109-285): every rank runs
// as a thread of this process over an in-process mailbox (Communicate/Communicator.h stand-in), so the reference's own
// code decides which strip goes to whom, with which tag, and how it lands (= or +=).  Kokkos is replaced by serial stand-ins,
// Communicate/Archive.h and Utility/ParallelDispatch.h are skipped through their include guards.
#include <Kokkos_Core.hpp>

#include <array>
#include <cstddef>
#include <memory>
#include <thread>
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
    // ghosted rank-3 array, x fastest (the layout BareField's Kokkos::View<T***, LayoutLeft> has); sub-views share storage
    template <typename T> struct View3D {
        using value_type      = T;
        using execution_space = Kokkos::Serial;
        using memory_space    = Kokkos::HostSpace;
        static constexpr unsigned rank = 3;
        T* p = nullptr;
        long e[3] = {0, 0, 0};
        long s[3] = {0, 0, 0};
        void set(T* data, long e0, long e1, long e2) {
            p = data;
            e[0] = e0; e[1] = e1; e[2] = e2;
            s[0] = 1; s[1] = e0; s[2] = e0 * e1;
        }
        long extent(std::size_t d) const { return e[d]; }
        std::size_t size() const { return (std::size_t)(e[0] * e[1] * e[2]); }
        T& operator()(std::size_t i, std::size_t j, std::size_t k) const { return p[i * s[0] + j * s[1] + k * s[2]]; }
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
    template <class T, class R0, class R1, class R2>
    refshim::View3D<T> subview(const refshim::View3D<T>& v, R0 r0, R1 r1, R2 r2) {
        refshim::View3D<T> o = v;
        o.p    = v.p + (long)r0.first * v.s[0] + (long)r1.first * v.s[1] + (long)r2.first * v.s[2];
        o.e[0] = (long)r0.second - (long)r0.first;
        o.e[1] = (long)r1.second - (long)r1.first;
        o.e[2] = (long)r2.second - (long)r2.first;
        return o;
    }
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
    v.set(field, ng[0] + 2 * nghost, ng[1] + 2 * nghost, ng[2] + 2 * nghost);
    Halo h;
    if (mode == 0) h.applyPeriodicSerialDim<Halo::assign>(v, &fl, nghost);
    else h.applyPeriodicSerialDim<Halo::rhs_plus_assign>(v, &fl, nghost);
}

}  // extern "C"

// BareField::fillHalo (mode 0) / accumulateHalo (mode 1) for EVERY rank of a FieldLayout of ng cells on nranks ranks
// (boxes != NULL: after FieldLayout::updateLayout(boxes), [nranks][6] lo, hi inclusive): fields[r] is rank r's ghosted
// array with ncomp (1 or 3) doubles per cell, x fastest.  One thread per rank runs the reference's exchangeBoundaries
// over the in-process mailbox, then applyPeriodicSerialDim, exactly the two calls of BareField.hpp:152-172.
template <typename T>
static void run_rank(const int* ng, int nranks, const int* boxes, int nghost, int mode, int r, double* field) {
    refshim::g_rank = r;   // thread-local
    ippl::Index ix(ng[0]), iy(ng[1]), iz(ng[2]);
    ippl::NDIndex<3> domain(ix, iy, iz);
    std::array<bool, 3> par = {true, true, true};
    ippl::FieldLayout<3> fl(ippl::mpi::Communicator(), domain, par, true, nghost);
    if (boxes) {
        std::vector<ippl::NDIndex<3>> doms(nranks);
        for (int q = 0; q < nranks; ++q)
            doms[q] = ippl::NDIndex<3>(ippl::Index(boxes[6 * q], boxes[6 * q + 3]), ippl::Index(boxes[6 * q + 1], boxes[6 * q + 4]),
                                       ippl::Index(boxes[6 * q + 2], boxes[6 * q + 5]));
        fl.updateLayout(doms);
    }
    using Halo = ippl::detail::HaloCells<T, 3>;
    const auto& ld = fl.getLocalNDIndex();
    typename Halo::view_type v;
    v.set(reinterpret_cast<T*>(field), ld[0].length() + 2 * nghost, ld[1].length() + 2 * nghost, ld[2].length() + 2 * nghost);
    Halo h;
    if (mode == 0) {
        if (fl.comm.size() > 1) h.fillHalo(v, &fl, nghost);
        h.template applyPeriodicSerialDim<typename Halo::assign>(v, &fl, nghost);
    } else {
        if (fl.comm.size() > 1) h.accumulateHalo(v, &fl, nghost);
        h.template applyPeriodicSerialDim<typename Halo::rhs_plus_assign>(v, &fl, nghost);
    }
}

extern "C" {

void refhalo_exchange(const int ng[3], int nranks, const int* boxes, int nghost, int ncomp, int mode, double* const* fields) {
    refshim::g_size = nranks;
    std::vector<std::thread> th;
    for (int r = 0; r < nranks; ++r)
        th.emplace_back([=] {
            if (ncomp == 3) run_rank<ippl::Vector<double, 3>>(ng, nranks, boxes, nghost, mode, r, fields[r]);
            else run_rank<double>(ng, nranks, boxes, nghost, mode, r, fields[r]);
        });
    for (auto& t : th) t.join();
    refshim::g_size = 1;
    refshim::g_rank = 0;
}

}  // extern "C"
