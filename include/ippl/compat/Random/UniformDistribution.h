// compat/Random/UniformDistribution.h -- the uniform CDF / PDF / estimate helpers (src/Random/UniformDistribution.h:17-30)
#ifndef IPPL_COMPAT_UNIFORM_DISTRIBUTION_H
#define IPPL_COMPAT_UNIFORM_DISTRIBUTION_H
#include "Ippl.h"
namespace ippl {
namespace random {
    template <typename T>
    KOKKOS_FUNCTION T uniform_cdf_func(T x) { return x; }
    template <typename T>
    KOKKOS_FUNCTION T uniform_pdf_func() { return 1.; }
    template <typename T>
    KOKKOS_FUNCTION T uniform_estimate_func(T u) { return u; }
}  // namespace random
}  // namespace ippl
#endif
