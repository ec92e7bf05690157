#pragma once
#include <cmath>
namespace Kokkos {
    using std::abs; using std::fabs; using std::pow; using std::sqrt; using std::sin; using std::cos;
    using std::exp; using std::log; using std::floor; using std::ceil; using std::tan; using std::asin;
    using std::acos; using std::atan; using std::sinh; using std::cosh; using std::tanh; using std::erf;
    using std::fmod; using std::isnan; using std::isinf; using std::isfinite; using std::cbrt;
    using std::hypot; using std::atan2; using std::round; using std::trunc; using std::log10; using std::log2; using std::copysign;
}  // namespace Kokkos
