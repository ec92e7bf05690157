// Minimal stand-in for <Kokkos_Core.hpp> -- TEST INFRASTRUCTURE (oracle/ref_shim), host only.
#pragma once
#include <cstddef>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "Kokkos_Macros.hpp"
#include "Kokkos_MathematicalFunctions.hpp"
#include "Kokkos_MinMax.hpp"
namespace Kokkos {
    struct HostSpace {};
    struct Serial { using memory_space = HostSpace; };
    using DefaultExecutionSpace     = Serial;
    using DefaultHostExecutionSpace = Serial;
    template <typename T> inline void atomic_add(T* dst, const T& v) { *dst += v; }
    inline void fence() {}
    // rank-1 shared-storage view: just enough for FieldLayout's rank-box tables
    template <typename DataType, typename... Props> class View;
    template <typename T, typename... Props> class View<T*, Props...> {
    public:
        using host_mirror_type = View<T*, Props...>;
        using HostMirror       = host_mirror_type;
        using value_type       = T;
        using size_type        = std::size_t;
        using memory_space     = HostSpace;
        using execution_space  = Serial;
        View() : d_(std::make_shared<std::vector<T>>()) {}
        View(const std::string&, std::size_t n) : d_(std::make_shared<std::vector<T>>(n)) {}
        T& operator()(std::size_t i) const { return (*d_)[i]; }
        T& operator[](std::size_t i) const { return (*d_)[i]; }
        std::size_t size() const { return d_->size(); }
        std::size_t extent(int) const { return d_->size(); }
        T* data() const { return d_->data(); }
        std::shared_ptr<std::vector<T>> d_;
    };
    template <typename V> inline V create_mirror_view(const V& v) {
        V m; m.d_->resize(v.size()); return m;
    }
    template <typename V> inline void resize(V& v, std::size_t n) { v.d_->resize(n); }
    template <typename V> inline void realloc(V& v, std::size_t n) { v.d_->assign(n, typename V::value_type{}); }
    template <typename V> inline void deep_copy(const V& dst, const V& src) { *dst.d_ = *src.d_; }
}  // namespace Kokkos
