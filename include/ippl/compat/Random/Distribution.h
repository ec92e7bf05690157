// compat/Random/Distribution.h -- ippl::random::Distribution in the reference's shape (src/Random/Distribution.h:60-112): a
// parameter block plus the caller's CDF / PDF / Estimate functors, usable on the host and inside device lambdas
#ifndef IPPL_COMPAT_DISTRIBUTION_H
#define IPPL_COMPAT_DISTRIBUTION_H
#include "Ippl.h"
namespace ippl {
namespace random {
    template <typename T, unsigned Dim, unsigned DimP, typename Functions>
    class Distribution {
    public:
        T par_m[DimP];
        KOKKOS_INLINE_FUNCTION Distribution(const T* par_p) {
            for (unsigned i = 0; i < DimP; ++i) par_m[i] = par_p[i];
        }
        KOKKOS_INLINE_FUNCTION T getPdf(T x, unsigned d) const { return typename Functions::PDF()(x, d, par_m); }
        KOKKOS_INLINE_FUNCTION T getCdf(T x, unsigned d) const { return typename Functions::CDF()(x, d, par_m); }
        KOKKOS_INLINE_FUNCTION T getEstimate(T u, unsigned d) const { return typename Functions::Estimate()(u, d, par_m); }
        KOKKOS_INLINE_FUNCTION T getObjFunc(T x, unsigned d, T u) const { return getCdf(x, d) - u; }
        KOKKOS_INLINE_FUNCTION T getDerObjFunc(T x, unsigned d) const { return getPdf(x, d); }
        // product of the per-dimension densities
        KOKKOS_INLINE_FUNCTION T getFullPdf(const ippl::Vector<T, Dim>& x) const {
            T p = 1.0;
            for (unsigned d = 0; d < Dim; ++d) p *= getPdf(x[d], d);
            return p;
        }
    };
}  // namespace random
}  // namespace ippl
#endif
