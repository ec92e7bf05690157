// compat/Random/Randn.h -- ippl::random::randn (src/Random/Randn.h:30-94): the functor the drivers hand to
// Kokkos::parallel_for to fill a velocity attribute with mu[d] + sd[d] * N(0, 1); normals by Box-Muller on the
// counter-based stream (include/ippl/philox.h)
#ifndef IPPL_COMPAT_RANDN_H
#define IPPL_COMPAT_RANDN_H
#include <Kokkos_Random.hpp>
#include "Random/Distribution.h"
#include "ippl/philox.h"
namespace ippl {
namespace random {
    template <typename T, unsigned Dim>
    struct randn {
        static_assert(Dim == 3, "the B200 path is three-dimensional");
        using view_type     = typename ippl::detail::ViewType<ippl::Vector<double, Dim>, 1>::view_type;
        using GeneratorPool = Kokkos::Random_XorShift64_Pool<>;
        view_type v;
        std::uint64_t seed;
        T mu[Dim], sd[Dim];
        randn(view_type v_, GeneratorPool pool, T* mu_p, T* sd_p) : v(v_), seed(pool.seed) {
            for (unsigned d = 0; d < Dim; ++d) {
                mu[d] = mu_p[d];
                sd[d] = sd_p[d];
            }
        }
        randn(view_type v_, GeneratorPool pool) : v(v_), seed(pool.seed) {
            for (unsigned d = 0; d < Dim; ++d) {
                mu[d] = 0.0;
                sd[d] = 1.0;
            }
        }
        KOKKOS_INLINE_FUNCTION const T& getMu(unsigned d) const { return mu[d]; }
        KOKKOS_INLINE_FUNCTION const T& getSd(unsigned d) const { return sd[d]; }
        KOKKOS_INLINE_FUNCTION void operator()(const size_t i) const {
            double g[3];
            philox_normal3(seed, (std::uint64_t)i, g);
            for (unsigned d = 0; d < Dim; ++d) v(i)[d] = mu[d] + sd[d] * g[d];
        }
    };
}  // namespace random
}  // namespace ippl
#endif
