// compat/Random/Utility.h -- the Newton iteration of the inverse-transform sampler (src/Random/Utility.h:27-60)
#ifndef IPPL_COMPAT_RANDOM_UTILITY_H
#define IPPL_COMPAT_RANDOM_UTILITY_H
#include "Ippl.h"
namespace ippl {
namespace random {
namespace detail {
    // while iter < max_iter && |cdf(x) - u| > atol: x -= (cdf(x) - u) / pdf(x)
    template <typename T, class Dist>
    KOKKOS_INLINE_FUNCTION void newton_raphson(const Dist& dist, unsigned d, T& x, T u, int max_iter = 20, T atol = 1e-12) {
        int it = 0;
        while (it < max_iter && Kokkos::fabs(dist.getObjFunc(x, d, u)) > atol) {
            x = x - dist.getObjFunc(x, d, u) / dist.getDerObjFunc(x, d);
            ++it;
        }
    }
}  // namespace detail
}  // namespace random
}  // namespace ippl
#endif
