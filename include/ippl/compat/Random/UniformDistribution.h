// compat/Random/UniformDistribution.h -- B200 facade stand-in for the reference header of the same name
// (src/Random/UniformDistribution.h:17-30): on the unit interval the CDF is the identity, the density is one and the
// Newton start value is the uniform deviate itself.  BumponTail's CustomDistributionFunctions uses them for x and y.
#ifndef IPPL_COMPAT_UNIFORM_DISTRIBUTION_H
#define IPPL_COMPAT_UNIFORM_DISTRIBUTION_H
#include "Ippl.h"
namespace ippl::random {
#define IPPLC_UNIT_INTERVAL_FN(NAME, ARGS, VALUE) \
    template <class Real>                         \
    IPPL_HD inline Real NAME ARGS {               \
        return VALUE;                             \
    }
IPPLC_UNIT_INTERVAL_FN(uniform_cdf_func, (Real x), x)
IPPLC_UNIT_INTERVAL_FN(uniform_pdf_func, (), Real(1))
IPPLC_UNIT_INTERVAL_FN(uniform_estimate_func, (Real u), u)
#undef IPPLC_UNIT_INTERVAL_FN
}  // namespace ippl::random
#endif
