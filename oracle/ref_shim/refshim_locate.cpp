// refshim_locate.cpp -- TEST INFRASTRUCTURE.  Particle ownership as the reference decides it: the destRankOf lambda of
// ParticleSpatialLayout::locateParticlesPacked (src/Particle/ParticleSpatialLayout.hpp:372-395: own region, cached
// neighbours, every rank, then the inclusive fallback) and the return statements of positionInRegion /
// positionInRegionInclusive (:316-330) are cut out of the reference file at build time (gen_snippets.py -> psl_*.inc in a
// temporary include directory) and compiled here unchanged; regions are the reference's NDRegion / PRegion (Region/*.h, included
// in place).  The layout header itself cannot be included (it needs the whole particle framework).
#include <Kokkos_Core.hpp>

#include <cstddef>
#include <utility>
#include <vector>

#include "Utility/IpplException.h"
#include "Types/IpplTypes.h"
#include "Types/Vector.h"
#include "Region/NDRegion.h"

namespace {
    using vector_type = ippl::Vector<double, 3>;
    using region_type = ippl::NDRegion<double, 3>;
    using size_type   = ippl::detail::size_type;

    template <std::size_t... Idx>
    constexpr bool positionInRegion(const std::index_sequence<Idx...>&, const vector_type& pos, const region_type& region) {
#include "psl_in_region.inc"
    }
    template <std::size_t... Idx>
    constexpr bool positionInRegionInclusive(const std::index_sequence<Idx...>&, const vector_type& pos, const region_type& region) {
#include "psl_in_region_inclusive.inc"
    }
    struct RegionView {
        const region_type* p;
        std::size_t n;
        const region_type& operator()(std::size_t r) const { return p[r]; }
        std::size_t extent(int) const { return n; }
    };
    struct PosView {
        const vector_type* p;
        const vector_type& operator()(std::size_t i) const { return p[i]; }
    };
    struct IntView {
        const int* p;
        int operator()(std::size_t j) const { return p[j]; }
    };
}  // namespace

extern "C" {

// dest_out[i] = destination rank of particle i for rank `my`; regions[nranks][6] = min[3], max[3]; neighbours[nn] = the cached
// neighbour ranks searched before the global scan
void reflocate_dest_rank(int nranks, const double* regions, int my, const int* neighbours, int nn, long n, const double* R,
                         int* dest_out) {
    std::vector<region_type> regs((std::size_t)nranks);
    for (int r = 0; r < nranks; ++r)
        for (int d = 0; d < 3; ++d) regs[r][d] = ippl::PRegion<double>(regions[6 * r + d], regions[6 * r + 3 + d]);
    const RegionView Regions{regs.data(), (std::size_t)nranks};
    const PosView positions{reinterpret_cast<const vector_type*>(R)};
    const IntView neighbours_d{neighbours};
    const size_type neighbors_used = (size_type)nn;
    const size_type myRank         = (size_type)my;
    const auto is                  = std::make_index_sequence<3>{};
    auto destRankOf = [=](const std::size_t i) -> size_type {
#include "psl_dest_rank.inc"
    };
    for (long i = 0; i < n; ++i) dest_out[i] = (int)destRankOf((std::size_t)i);
}

}  // extern "C"
