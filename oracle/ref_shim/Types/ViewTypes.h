// stand-in for Types/ViewTypes.h (oracle/ref_shim): only the rank-1 view FieldLayout needs.
#pragma once
#include <Kokkos_Core.hpp>
namespace ippl { namespace detail {
    template <typename T, unsigned Dim, class... Properties> struct ViewType;
    template <typename T, class... Properties> struct ViewType<T, 1, Properties...> {
        using view_type = Kokkos::View<T*, Properties...>;
    };
}}  // namespace ippl::detail
