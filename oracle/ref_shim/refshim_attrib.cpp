// refshim_attrib.cpp -- TEST INFRASTRUCTURE.  ParticleAttrib::scatter and ::gather as the reference writes them: the
// bodies of the two Kokkos lambdas (src/Particle/ParticleAttrib.hpp:167-184 and :229-244 -- optional hash remap, position
// -> l -> index truncation -> whi / wlo -> args, then detail::scatterToField / gatherFromField, replace or add) are cut
// out of the reference file at build time (gen_snippets.py -> attrib_{scatter,gather}.inc in a temporary include directory) and compiled here
// unchanged inside a plain loop; the interpolation itself is the reference's Interpolation/CIC.h, included in place.
// The ParticleAttrib header cannot be included (it needs the whole particle framework).
#include <Kokkos_Core.hpp>

#include <cstddef>
#include <utility>

#include "Utility/IpplException.h"
#include "Types/Vector.h"
#include "Index/NDIndex.h"
#include "Interpolation/CIC.h"

namespace {
    template <typename T> struct View3 {
        static constexpr unsigned rank = 3;
        using value_type               = T;
        T* p;
        long e0, e1;
        T& operator()(std::size_t i, std::size_t j, std::size_t k) const { return p[i + e0 * (j + e1 * k)]; }
    };
    template <typename T> struct View1 {
        T* p;
        std::size_t n;
        T& operator()(std::size_t i) const { return p[i]; }
        std::size_t extent(int) const { return n; }
    };
    struct FieldTag { static constexpr unsigned dim = 3; };
}  // namespace

extern "C" {

// scatter: particles [begin, end) (through hash[] when nhash > 0) of R[n][3] with values q[n] into the ghosted scalar field
// (x fastest, extents ext[3]); first = lDom.first()
void refattrib_scatter(long begin, long end, const double* R, const double* q, const int* hash, long nhash, const double origin_[3],
                       const double h[3], const int first[3], int nghost_, const int ext[3], double* field) {
    using namespace ippl;
    using Field        = FieldTag;
    using PositionType = double;
    using vector_type  = Vector<double, 3>;
    using value_type   = double;
    View3<double> view{field, ext[0], ext[1]};
    vector_type dx, origin;
    for (int d = 0; d < 3; ++d) { dx[d] = h[d]; origin[d] = origin_[d]; }
    const vector_type invdx = 1.0 / dx;                       // ParticleAttrib.hpp:153
    Index ix(first[0], first[0] + ext[0] - 2 * nghost_ - 1), iy(first[1], first[1] + ext[1] - 2 * nghost_ - 1),
        iz(first[2], first[2] + ext[2] - 2 * nghost_ - 1);
    const NDIndex<3> lDom(ix, iy, iz);
    const int nghost = nghost_;
    View1<const int> hash_array{hash, (std::size_t)nhash};
    const bool useHashView = hash_array.extent(0) > 0;       // :160
    View1<const double> dview{q, 0};
    View1<Vector<double, 3>> ppview{reinterpret_cast<Vector<double, 3>*>(const_cast<double*>(R)), 0};
    for (std::size_t idx = (std::size_t)begin; idx < (std::size_t)end; ++idx) {
#include "attrib_scatter.inc"
    }
}

// gather of a Vector<double,3> field (AoS-3 per cell) into E[n][3], replace or add
void refattrib_gather(long n, const double* R, const double origin_[3], const double h[3], const int first[3], int nghost_,
                      const int ext[3], const double* efield, int add, double* E) {
    using namespace ippl;
    using Field        = FieldTag;
    using PositionType = double;
    using vector_type  = Vector<double, 3>;
    using value_type   = Vector<double, 3>;
    const View3<Vector<double, 3>> view{reinterpret_cast<Vector<double, 3>*>(const_cast<double*>(efield)), ext[0], ext[1]};
    vector_type dx, origin;
    for (int d = 0; d < 3; ++d) { dx[d] = h[d]; origin[d] = origin_[d]; }
    const vector_type invdx = 1.0 / dx;                       // :217
    Index ix(first[0], first[0] + ext[0] - 2 * nghost_ - 1), iy(first[1], first[1] + ext[1] - 2 * nghost_ - 1),
        iz(first[2], first[2] + ext[2] - 2 * nghost_ - 1);
    const NDIndex<3> lDom(ix, iy, iz);
    const int nghost          = nghost_;
    const bool addToAttribute = add != 0;
    View1<Vector<double, 3>> dview{reinterpret_cast<Vector<double, 3>*>(E), 0};
    View1<Vector<double, 3>> ppview{reinterpret_cast<Vector<double, 3>*>(const_cast<double*>(R)), 0};
    for (std::size_t idx = 0; idx < (std::size_t)n; ++idx) {
#include "attrib_gather.inc"
    }
}

}  // extern "C"
