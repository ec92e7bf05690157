// stand-in for <Kokkos_MathematicalConstants.hpp> (oracle/ref_shim): the constants the Random headers use
#pragma once
namespace Kokkos { namespace numbers {
    template <typename T> inline constexpr T pi_v = static_cast<T>(3.141592653589793238462643383279502884L);
}}  // namespace Kokkos::numbers
