// KokkosShim.cuh -- the part of Kokkos the alpine DRIVERS use for kernels of their own, on top of the B200 facade.
//
// The reference's mini-apps contain device code that is not part of IPPL: KOKKOS_LAMBDA bodies handed to
// Kokkos::parallel_for / parallel_reduce over `attrib.getView()` (demos/alpine/PenningTrapManager.h:256-272, 313-333:
// the Kick1 / Kick2 of the Boris-type push; LandauDampingManager.h:401-412: energy from the particles) and to
// ippl::parallel_reduce over a field's range policy (LandauDampingManager.h:346-360: "Ex stats").  A host-only header
// cannot run those.  This header can, when the driver's translation unit is compiled by nvcc (`-x cu --extended-lambda`,
// sm_100a): it provides
//   KOKKOS_LAMBDA / KOKKOS_INLINE_FUNCTION / KOKKOS_FUNCTION,
//   Kokkos::{RangePolicy, parallel_for, parallel_reduce, Sum, Max, Min, fence, DefaultExecutionSpace, View (1-D, for
//            scratch the drivers allocate), pow, sin, cos, sqrt, exp, fabs, numbers::pi_v},
//   ippl::{RangePolicy<Dim>::index_array_type, getRangePolicy, parallel_for, parallel_reduce, apply}
// with launches on the facade context's stream; `getView()` of ParticleAttrib / Field (include/ippl/Ippl.h) hands out
// view(i)[d] / view(i, j, k) proxies over the SoA particle arrays and the ghosted field arrays.
// Not a Kokkos re-implementation: no layouts, no teams, no execution spaces other than the context's stream.
// demos/ref_lambdas.cu compiles the reference drivers' own lambda bodies -- cut out of the reference tree at build time
// -- on it, unchanged, and checks them against the C-ABI kernels.
#pragma once
// IPPL_SHIM_HOST_EMULATION (tests only, never the product): the driver is compiled by the host compiler and linked against
// the CPU mock of the C-ABI under oracle/mock, where "device" memory is host memory; kernels become host loops.  It exists
// to run the reference's unchanged drivers through this header and include/ippl/compat on a machine without a GPU.
#if !defined(__CUDACC__) && !defined(IPPL_SHIM_HOST_EMULATION)
#error "include/ippl/KokkosShim.cuh needs nvcc (device lambdas): compile the driver with -x cu --extended-lambda"
#endif
#if defined(IPPL_SHIM_HOST_EMULATION) && !defined(__CUDACC__)
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#include <cfloat>
#include <string>

#include "Ippl.h"

#define KOKKOS_LAMBDA [=] __host__ __device__
#define KOKKOS_CLASS_LAMBDA [ =, *this ] __host__ __device__
#define KOKKOS_INLINE_FUNCTION __host__ __device__ inline
#define KOKKOS_FUNCTION __host__ __device__
#define KOKKOS_FORCEINLINE_FUNCTION __host__ __device__ __forceinline__

namespace Kokkos {

struct DefaultExecutionSpace {};
struct DefaultHostExecutionSpace {};

template <class... Properties>
struct RangePolicy {
    long b = 0, e = 0;
    RangePolicy() = default;
    RangePolicy(long begin_, long end_) : b(begin_), e(end_) {}
    template <class Space>
    RangePolicy(const Space&, long begin_, long end_) : b(begin_), e(end_) {}
    long begin() const { return b; }
    long end() const { return e; }
};

namespace numbers {
    template <class T>
    inline constexpr T pi_v = static_cast<T>(3.141592653589793238462643383279502884L);
}

KOKKOS_INLINE_FUNCTION double pow(double a, double b) { return ::pow(a, b); }
// integer powers by repeated multiplication (what the drivers use pow for: squares) -- exact for b = 2
KOKKOS_INLINE_FUNCTION double pow(double a, int b) {
    if (b < 0 || b > 8) return ::pow(a, (double)b);
    double r = 1.0;
    for (int i = 0; i < b; ++i) r *= a;
    return r;
}
KOKKOS_INLINE_FUNCTION double sin(double a) { return ::sin(a); }
KOKKOS_INLINE_FUNCTION double cos(double a) { return ::cos(a); }
KOKKOS_INLINE_FUNCTION double sqrt(double a) { return ::sqrt(a); }
KOKKOS_INLINE_FUNCTION double exp(double a) { return ::exp(a); }
KOKKOS_INLINE_FUNCTION double erf(double a) { return ::erf(a); }
KOKKOS_INLINE_FUNCTION double fabs(double a) { return ::fabs(a); }
KOKKOS_INLINE_FUNCTION double abs(double a) { return ::fabs(a); }

inline void fence() { ippl::b200::check(ipplb_sync(ippl::b200::ctx()), "Kokkos::fence"); }
inline void fence(const std::string&) { fence(); }

// [host-emulation begin: shim reducers]  (tests/test_kernel_text_cpu.py compiles the marked text of this header for the
// host and runs the reduction kernels under a lock-step block emulator, tests/emu/emu_shim_reduce.cpp)
// reducers: the result lands in the referenced host scalar when parallel_reduce returns (blocking, like Kokkos with a
// scalar result)
template <class T>
struct Sum {
    using value_type = T;
    T& ref;
    explicit Sum(T& r) : ref(r) {}
    static __host__ __device__ T identity() { return T(0); }
    static __host__ __device__ void join(T& a, const T& b) { a += b; }
};
template <class T>
struct Max {
    using value_type = T;
    T& ref;
    explicit Max(T& r) : ref(r) {}
    static __host__ __device__ T identity() { return -DBL_MAX; }
    static __host__ __device__ void join(T& a, const T& b) { a = b > a ? b : a; }
};
template <class T>
struct Min {
    using value_type = T;
    T& ref;
    explicit Min(T& r) : ref(r) {}
    static __host__ __device__ T identity() { return DBL_MAX; }
    static __host__ __device__ void join(T& a, const T& b) { a = b < a ? b : a; }
};
// [host-emulation end: shim reducers]

namespace shim {
    inline cudaStream_t stream() { return (cudaStream_t)ipplb_ctx_stream(ippl::b200::ctx()); }
    inline int grid_for(long n, int block) {
        long g = (n + block - 1) / block;
        return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
    }

#ifndef IPPL_SHIM_HOST_EMULATION
    // [host-emulation begin: shim kernels]
    template <class F>
    __global__ void __launch_bounds__(256) for_kernel(long b, long e, F f) {
        for (long i = b + (long)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (long)gridDim.x * blockDim.x) f((std::size_t)i);
    }

    // fp64 atomic combine through compare-and-swap: exact for max / min, and used for the sum as well so that one code
    // path serves every reducer (a few hundred combines per launch: one per block)
    template <class R>
    __device__ void atomic_join(double* addr, double v) {
        unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
        unsigned long long old = *a, assumed;
        do {
            assumed  = old;
            double cur = __longlong_as_double((long long)assumed);
            R::join(cur, v);
            old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(cur));
        } while (assumed != old);
    }

    template <class R>
    __device__ void block_join(double v, double* out) {
        __shared__ double part[8];
        for (int o = 16; o > 0; o >>= 1) {
            double w = __shfl_xor_sync(0xffffffffu, v, o);
            R::join(v, w);
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __syncthreads();   // `part` may still be read by the previous reducer of this launch
        if (lane == 0) part[warp] = v;
        __syncthreads();
        if (warp == 0) {
            double t = lane < (int)(blockDim.x >> 5) ? part[lane] : R::identity();
            for (int o = 4; o > 0; o >>= 1) {
                double w = __shfl_xor_sync(0xffffffffu, t, o);
                R::join(t, w);
            }
            if (lane == 0) atomic_join<R>(out, t);
        }
    }

    template <class F, class R0>
    __global__ void __launch_bounds__(256) reduce1_kernel(long b, long e, F f, double* out) {
        double a0 = R0::identity();
        for (long i = b + (long)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (long)gridDim.x * blockDim.x) f((std::size_t)i, a0);
        block_join<R0>(a0, out);
    }
    template <class F, class R0, class R1>
    __global__ void __launch_bounds__(256) reduce2_kernel(long b, long e, F f, double* out) {
        double a0 = R0::identity(), a1 = R1::identity();
        for (long i = b + (long)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (long)gridDim.x * blockDim.x)
            f((std::size_t)i, a0, a1);
        block_join<R0>(a0, out);
        block_join<R1>(a1, out + 1);
    }
    // [host-emulation end: shim kernels]

    // result slots in device memory, initialised with the reducers' identities; read back when the launch is done
    struct Slots {
        double* d = nullptr;
        Slots(const double* init, int n) {
            ippl::b200::cuda_check(cudaMalloc(&d, sizeof(double) * n), "parallel_reduce");
            ippl::b200::cuda_check(cudaMemcpyAsync(d, init, sizeof(double) * n, cudaMemcpyHostToDevice, stream()), "parallel_reduce");
        }
        void fetch(double* out, int n) {
            ippl::b200::cuda_check(cudaMemcpyAsync(out, d, sizeof(double) * n, cudaMemcpyDeviceToHost, stream()), "parallel_reduce");
            ippl::b200::cuda_check(cudaStreamSynchronize(stream()), "parallel_reduce");
        }
        ~Slots() { cudaFree(d); }
    };
#endif  // !IPPL_SHIM_HOST_EMULATION
}  // namespace shim

// ---- parallel_for --------------------------------------------------------------------------------------------------------
template <class... P, class F>
void parallel_for(const std::string&, const RangePolicy<P...>& policy, const F& f) {
    const long n = policy.end() - policy.begin();
    if (n <= 0) return;
#ifdef IPPL_SHIM_HOST_EMULATION
    for (long i = policy.begin(); i < policy.end(); ++i) f((std::size_t)i);
#else
    shim::for_kernel<<<shim::grid_for(n, 256), 256, 0, shim::stream()>>>(policy.begin(), policy.end(), f);
    ippl::b200::cuda_check(cudaGetLastError(), "Kokkos::parallel_for");
#endif
}
template <class... P, class F>
void parallel_for(const RangePolicy<P...>& policy, const F& f) { parallel_for(std::string(), policy, f); }
template <class F>
void parallel_for(const std::string& name, std::size_t n, const F& f) { parallel_for(name, RangePolicy<>(0, (long)n), f); }
template <class F>
void parallel_for(std::size_t n, const F& f) { parallel_for(std::string(), RangePolicy<>(0, (long)n), f); }

// ---- parallel_reduce (one or two reducers over doubles) -----------------------------------------------------------------------
template <class... P, class F, class R0>
void parallel_reduce(const std::string&, const RangePolicy<P...>& policy, const F& f, const R0& r0) {
    double v[1] = {R0::identity()};
#ifdef IPPL_SHIM_HOST_EMULATION
    for (long i = policy.begin(); i < policy.end(); ++i) f((std::size_t)i, v[0]);
#else
    shim::Slots slots(v, 1);
    const long n = policy.end() - policy.begin();
    if (n > 0) {
        shim::reduce1_kernel<F, R0><<<shim::grid_for(n, 256), 256, 0, shim::stream()>>>(policy.begin(), policy.end(), f, slots.d);
        ippl::b200::cuda_check(cudaGetLastError(), "Kokkos::parallel_reduce");
    }
    slots.fetch(v, 1);
#endif
    r0.ref = v[0];
}
template <class... P, class F, class R0, class R1>
void parallel_reduce(const std::string&, const RangePolicy<P...>& policy, const F& f, const R0& r0, const R1& r1) {
    double v[2] = {R0::identity(), R1::identity()};
#ifdef IPPL_SHIM_HOST_EMULATION
    for (long i = policy.begin(); i < policy.end(); ++i) f((std::size_t)i, v[0], v[1]);
#else
    shim::Slots slots(v, 2);
    const long n = policy.end() - policy.begin();
    if (n > 0) {
        shim::reduce2_kernel<F, R0, R1><<<shim::grid_for(n, 256), 256, 0, shim::stream()>>>(policy.begin(), policy.end(), f, slots.d);
        ippl::b200::cuda_check(cudaGetLastError(), "Kokkos::parallel_reduce");
    }
    slots.fetch(v, 2);
#endif
    r0.ref = v[0];
    r1.ref = v[1];
}
// a particle count instead of a policy (PenningTrapManager.h:354-355)
template <class F, class R0>
void parallel_reduce(const std::string& name, std::size_t n, const F& f, const R0& r0) {
    parallel_reduce(name, RangePolicy<>(0, (long)n), f, r0);
}
// parallel_reduce(name, policy, f, double& result): a plain scalar means Sum
template <class... P, class F>
void parallel_reduce(const std::string& name, const RangePolicy<P...>& policy, const F& f, double& result) {
    parallel_reduce(name, policy, f, Sum<double>(result));
}

// ---- View: 1-D device scratch the drivers allocate themselves (e.g. Kokkos::View<double*> tmp("tmp", n)) -----------------------
template <class T>
struct View;
template <class T>
struct View<T*> {
    using execution_space = DefaultExecutionSpace;
    T* d = nullptr;
    std::size_t n = 0;
    std::shared_ptr<void> own;   // host-side ownership; device copies of the view carry only (d, n)
    View() = default;
    View(const std::string&, std::size_t count) : n(count) {
        ippl::b200::cuda_check(cudaMalloc(&d, sizeof(T) * (count ? count : 1)), "Kokkos::View");
        ippl::b200::cuda_check(cudaMemsetAsync(d, 0, sizeof(T) * (count ? count : 1), shim::stream()), "Kokkos::View");
        own = std::shared_ptr<void>(d, [](void* p) { cudaFree(p); });
    }
    __host__ __device__ T& operator()(std::size_t i) const { return d[i]; }
    __host__ __device__ std::size_t extent(int) const { return n; }
    __host__ __device__ std::size_t size() const { return n; }
    __host__ __device__ T* data() const { return d; }
};

}  // namespace Kokkos

// ---- dot(a, b).apply() on particle / field element references (src/Types/Vector.hpp expression form) ---------------------------
namespace ippl {
namespace detail {
    struct DotValue {
        double v;
        __host__ __device__ double apply() const { return v; }
        __host__ __device__ operator double() const { return v; }
    };
    template <class A, class B>
    __host__ __device__ DotValue dot3(const A& a, const B& b) {
        return DotValue{a[0] * b[0] + a[1] * b[1] + a[2] * b[2]};
    }
    __host__ __device__ inline DotValue dot(const SoARef3& a, const SoARef3& b) { return dot3(a, b); }
    __host__ __device__ inline DotValue dot(const AoSRef3& a, const AoSRef3& b) { return dot3(a, b); }
}  // namespace detail
__host__ __device__ inline detail::DotValue dot(const Vector<double, 3>& a, const Vector<double, 3>& b) { return detail::dot3(a, b); }
}  // namespace ippl

// ---- ippl::parallel_for / parallel_reduce over a field's index range (src/Utility/ParallelDispatch.h) --------------------------
namespace ippl {

// getRangePolicy(view, shift): the view's index range without `shift` ghost layers on every side
template <class View>
typename RangePolicy<3>::policy_type getRangePolicy(const View& v, int shift = 0) {
    typename RangePolicy<3>::policy_type p;
    for (int d = 0; d < 3; ++d) {
        p.lo[d] = shift;
        p.hi[d] = v.extent(d) - shift;
    }
    return p;
}

// ippl::apply(view, args): the view's element at an index array (src/Expression/IpplOperations.h)
template <class View>
__host__ __device__ auto apply(const View& v, const Vector<long, 3>& a) -> decltype(v(0L, 0L, 0L)) {
    return v(a[0], a[1], a[2]);
}

namespace detail {
    // flat index of the box -> index array, x fastest
    struct Unflatten {
        long lo[3], n0, n1;
        __host__ __device__ Vector<long, 3> operator()(std::size_t t) const {
            Vector<long, 3> a;
            a[0] = lo[0] + (long)(t % (std::size_t)n0);
            a[1] = lo[1] + (long)((t / (std::size_t)n0) % (std::size_t)n1);
            a[2] = lo[2] + (long)(t / ((std::size_t)n0 * (std::size_t)n1));
            return a;
        }
    };
    inline Unflatten unflatten(const RangePolicy<3>::policy_type& p) {
        Unflatten u;
        for (int d = 0; d < 3; ++d) u.lo[d] = p.lo[d];
        u.n0 = p.hi[0] - p.lo[0];
        u.n1 = p.hi[1] - p.lo[1];
        return u;
    }
}  // namespace detail

template <class F>
void parallel_for(const std::string& name, const RangePolicy<3>::policy_type& p, const F& f) {
    const detail::Unflatten u = detail::unflatten(p);
    Kokkos::parallel_for(name, Kokkos::RangePolicy<>(0, p.count()), [=] __host__ __device__(std::size_t t) { f(u(t)); });
}
template <class F, class R0>
void parallel_reduce(const std::string& name, const RangePolicy<3>::policy_type& p, const F& f, const R0& r0) {
    const detail::Unflatten u = detail::unflatten(p);
    Kokkos::parallel_reduce(name, Kokkos::RangePolicy<>(0, p.count()), [=] __host__ __device__(std::size_t t, double& a0) { f(u(t), a0); }, r0);
}
template <class F, class R0, class R1>
void parallel_reduce(const std::string& name, const RangePolicy<3>::policy_type& p, const F& f, const R0& r0, const R1& r1) {
    const detail::Unflatten u = detail::unflatten(p);
    Kokkos::parallel_reduce(
        name, Kokkos::RangePolicy<>(0, p.count()), [=] __host__ __device__(std::size_t t, double& a0, double& a1) { f(u(t), a0, a1); }, r0, r1);
}

}  // namespace ippl
