// refshim_region.cpp -- TEST INFRASTRUCTURE.  The physical region of every rank as the REAL reference computes it:
// detail::RegionLayout<double, 3, UniformCartesian<double, 3>>(FieldLayout, mesh) -> convertNDIndex ->
// UniformCartesian::getVertexPosition (Region/RegionLayout.h/.hpp, Meshes/UniformCartesian.h/.hpp, Meshes/Mesh.h/.hpp
// compiled in place from /root/reference/src, never copied).  These doubles decide particle ownership
// (ParticleSpatialLayout::positionInRegion), so the restatement and the product's ipplb_layout_regions have to
// reproduce them bit for bit.  Field/BareField.h, Field/Field.h and Utility/IpplInfo.h, which UniformCartesian.hpp
// includes for methods that are not used here, and Utility/TypeUtils.h are skipped through their include guards.
#include <Kokkos_Core.hpp>

#include <array>
#include <cstddef>
#include <iostream>
#include <vector>

#define IPPL_BARE_FIELD_H
#define IPPL_FIELD_H
#define IPPL_INFO_H
#define IPPL_TYPE_UTILS_H   // Utility/TypeUtils.h needs most of Kokkos; the one alias RegionLayout.h takes from it is below

#include "Utility/IpplException.h"
#include "Types/IpplTypes.h"
#include "Types/Vector.h"
#include "Types/ViewTypes.h"

#include "Index/NDIndex.h"
#include "Region/NDRegion.h"
#include "FieldLayout/FieldLayout.h"

namespace ippl {
    inline mpi::Communicator* Comm = new mpi::Communicator();
    namespace detail {
        template <template <typename...> class Type, typename View> struct CreateUniformType { using type = Type<>; };
    }
}  // namespace ippl

#include "Meshes/UniformCartesian.h"
#include "Region/RegionLayout.h"

extern "C" {

// regions_out[nranks][6] = min[3], max[3] of rank r's region for the reference's FieldLayout of ng cells on nranks ranks
// (or, when boxes != NULL, after FieldLayout::updateLayout with those boxes: [nranks][6] lo, hi inclusive)
void refregion_regions(const int ng[3], int nranks, const int* boxes, const double origin[3], const double h[3],
                       double* regions_out) {
    refshim::g_rank = 0;
    refshim::g_size = nranks;
    ippl::Index ix(ng[0]), iy(ng[1]), iz(ng[2]);
    ippl::NDIndex<3> domain(ix, iy, iz);
    std::array<bool, 3> par = {true, true, true};
    ippl::FieldLayout<3> fl(ippl::mpi::Communicator(), domain, par, true, 1);
    if (boxes) {
        std::vector<ippl::NDIndex<3>> doms(nranks);
        for (int r = 0; r < nranks; ++r)
            doms[r] = ippl::NDIndex<3>(ippl::Index(boxes[6 * r], boxes[6 * r + 3]), ippl::Index(boxes[6 * r + 1], boxes[6 * r + 4]),
                                       ippl::Index(boxes[6 * r + 2], boxes[6 * r + 5]));
        fl.updateLayout(doms);
    }
    using Mesh = ippl::UniformCartesian<double, 3>;
    ippl::Vector<double, 3> hx, org;
    for (int d = 0; d < 3; ++d) { hx[d] = h[d]; org[d] = origin[d]; }
    Mesh mesh(domain, hx, org);
    ippl::detail::RegionLayout<double, 3, Mesh> rl(fl, mesh);
    const auto regs = rl.gethLocalRegions();
    for (int r = 0; r < nranks; ++r)
        for (int d = 0; d < 3; ++d) {
            regions_out[6 * r + d]     = regs(r)[d].min();
            regions_out[6 * r + 3 + d] = regs(r)[d].max();
        }
    refshim::g_size = 1;
}

}  // extern "C"
