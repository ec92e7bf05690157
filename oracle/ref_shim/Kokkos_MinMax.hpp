#pragma once
#include <algorithm>
namespace Kokkos {
    using std::max;
    using std::min;
}  // namespace Kokkos
