// compat/Kokkos_MathematicalFunctions.hpp -- provided by include/ippl/KokkosShim.cuh
#pragma once
#include "ippl/KokkosShim.cuh"
