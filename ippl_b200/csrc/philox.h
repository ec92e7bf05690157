// philox.h -- the counter-based uniform / normal stream of the device sampler lives in include/ippl/philox.h (the
// driver-side sampling kernels of include/ippl/compat/Random use the same stream)
#pragma once
#include "../../include/ippl/philox.h"
