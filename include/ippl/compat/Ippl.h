// compat/Ippl.h -- "Ippl.h" as the reference's drivers include it: the B200 facade + the device-lambda layer + the few
// framework names the drivers touch beyond them (MPI stand-ins, field boundary conditions, Inform to a file, ...).
#ifndef IPPL_COMPAT_IPPL_H
#define IPPL_COMPAT_IPPL_H

#define IPPL_B200_REFERENCE_SHAPED_RANDOM 1   // compat/Random/*.h provide ippl::random in the reference's (functor) shape
#include "ippl/KokkosShim.cuh"

#include <cstring>
#include <filesystem>
#include <variant>

#include "compat_detail.h"

#endif
