// compat/Kokkos_Random.hpp -- Kokkos::Random_XorShift64_Pool as the drivers use it: a seed holder.  The device sampler
// behind ippl::random (compat/Random/*.h) is counter based (Philox4x32-10, the stream of ipplb_sample_positions), so a
// "pool" is its seed.
#pragma once
#include <cstdint>
#include "ippl/KokkosShim.cuh"
namespace Kokkos {
template <class Device = DefaultExecutionSpace>
struct Random_XorShift64_Pool {
    std::uint64_t seed = 0;
    Random_XorShift64_Pool() = default;
    explicit Random_XorShift64_Pool(std::uint64_t s) : seed(s) {}
};
}  // namespace Kokkos
