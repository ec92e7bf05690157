// compat/Random/InverseTransformSampling.h -- ippl::random::InverseTransformSampling in the reference's shape
// (src/Random/InverseTransformSampling.h:30-256): how many samples each rank draws (the share of the CDF volume of its
// region, remainder to the first ranks) and the per-dimension inverse transform u -> x by Newton iteration, evaluated on
// the device with the CALLER'S distribution functors.  The uniform stream is the counter-based one of the C-ABI sampler
// (include/ippl/philox.h: key = seed of the "pool", counter = (sample index, dimension)).
#ifndef IPPL_COMPAT_INVERSE_TRANSFORM_SAMPLING_H
#define IPPL_COMPAT_INVERSE_TRANSFORM_SAMPLING_H
#include <Kokkos_Random.hpp>
#include <numeric>
#include "Random/Utility.h"
#include "ippl/philox.h"
namespace ippl {
namespace random {
    template <typename T, unsigned Dim, class DeviceType, class Distribution>
    class InverseTransformSampling {
    public:
        using view_type = typename ippl::detail::ViewType<Vector<T, Dim>, 1>::view_type;
        using size_type = ippl::detail::size_type;
        Distribution dist_m;
        size_type ntotal_m;
        Vector<T, Dim> umin_m, umax_m;

        // share of every rank from its region (RegionLayout), like updateBounds + the allreduce of the reference: the
        // regions of all ranks are known everywhere, so the remainder rule needs no communication
        template <class RegionLayout>
        InverseTransformSampling(Distribution& dist, Vector<T, Dim>& rmax, Vector<T, Dim>& rmin, RegionLayout& rlayout, size_type& ntotal)
            : dist_m(dist), ntotal_m(ntotal) {
            const int nr = ippl::Comm->size(), me = ippl::Comm->rank();
            const std::vector<double>& regs = rlayout.regions();   // [rank][min 3, max 3]
            T pdr = 1.0;
            for (unsigned d = 0; d < Dim; ++d) pdr *= dist_m.getCdf(rmax[d], d) - dist_m.getCdf(rmin[d], d);
            std::vector<size_type> n(nr);
            size_type sum = 0;
            for (int r = 0; r < nr; ++r) {
                T pnr = 1.0;
                for (unsigned d = 0; d < Dim; ++d)
                    pnr *= dist_m.getCdf(regs[6 * r + 3 + d], d) - dist_m.getCdf(regs[6 * r + d], d);
                n[r] = (size_type)(pnr / pdr * ntotal_m);
                sum += n[r];
            }
            const int rest = (int)(ntotal_m - sum);
            nlocal_m       = n[me] + (me < rest ? 1 : 0);
            for (unsigned d = 0; d < Dim; ++d) {
                umin_m[d] = dist_m.getCdf(regs[6 * me + d], d);
                umax_m[d] = dist_m.getCdf(regs[6 * me + 3 + d], d);
            }
        }
        // the whole domain on this rank
        InverseTransformSampling(Distribution& dist, Vector<T, Dim>& rmax, Vector<T, Dim>& rmin, size_type& ntotal)
            : dist_m(dist), ntotal_m(ntotal), nlocal_m(ntotal) {
            for (unsigned d = 0; d < Dim; ++d) {
                umin_m[d] = dist_m.getCdf(rmin[d], d);
                umax_m[d] = dist_m.getCdf(rmax[d], d);
            }
        }
        size_type getLocalSamplesNum() const { return nlocal_m; }
        void setLocalSamplesNum(size_type n) { nlocal_m = n; }

        void generate(view_type view, Kokkos::Random_XorShift64_Pool<> pool) { generate(view, 0, nlocal_m, pool); }
        void generate(view_type view, size_type startIndex, size_type endIndex, Kokkos::Random_XorShift64_Pool<> pool) {
            const Vector<T, Dim> lo = umin_m, hi = umax_m;
            const Distribution dist = dist_m;
            const std::uint64_t seed = pool.seed;
            Kokkos::parallel_for(
                "InverseTransformSampling::generate", Kokkos::RangePolicy<>((long)startIndex, (long)endIndex), KOKKOS_LAMBDA(const size_t i) {
                    for (unsigned d = 0; d < Dim; ++d) {
                        const T u = lo[d] + (hi[d] - lo[d]) * philox_uniform(seed, (std::uint64_t)i, d, 0u);
                        T x       = dist.getEstimate(u, d);
                        detail::newton_raphson<T, Distribution>(dist, d, x, u);
                        view(i)[d] = x;
                    }
                });
            Kokkos::fence();
        }

    private:
        size_type nlocal_m;
    };
}  // namespace random
}  // namespace ippl
#endif
