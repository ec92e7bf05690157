// compat/Random/NormalDistribution.h -- ippl::random::NormalDistribution (src/Random/NormalDistribution.h:11-60): Gaussian
// per dimension, parameters (mean, standard deviation) per dimension
#ifndef IPPL_COMPAT_NORMAL_DISTRIBUTION_H
#define IPPL_COMPAT_NORMAL_DISTRIBUTION_H
#include "Random/Distribution.h"
namespace ippl {
namespace random {
    template <typename T>
    struct normal_functions {
        struct PDF {
            KOKKOS_INLINE_FUNCTION T operator()(T x, unsigned d, const T* p) const {
                const T mu = p[2 * d], sd = p[2 * d + 1], z = (x - mu) / sd;
                return (1.0 / (sd * 2.5066282746310002)) * Kokkos::exp(-0.5 * z * z);
            }
        };
        struct CDF {
            KOKKOS_INLINE_FUNCTION T operator()(T x, unsigned d, const T* p) const {
                const T mu = p[2 * d], sd = p[2 * d + 1];
                return 0.5 * (1.0 + Kokkos::erf((x - mu) / (sd * 1.4142135623730951)));
            }
        };
        struct Estimate {
            KOKKOS_INLINE_FUNCTION T operator()(T u, unsigned d, const T* p) const { return p[2 * d] + 0. * u; }
        };
    };
    template <typename T, unsigned Dim>
    class NormalDistribution : public Distribution<T, Dim, 2 * Dim, normal_functions<T>> {
    public:
        KOKKOS_INLINE_FUNCTION NormalDistribution(const T* par_p) : Distribution<T, Dim, 2 * Dim, normal_functions<T>>(par_p) {}
    };
}  // namespace random
}  // namespace ippl
#endif
