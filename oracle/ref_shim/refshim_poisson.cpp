// refshim_poisson.cpp -- TEST INFRASTRUCTURE.  The k-space step of the reference's periodic Poisson solve with gradient
// output: the body of the lambda "Gradient FFTPeriodicPoissonSolver" (src/PoissonSolvers/FFTPeriodicPoissonSolver.hpp:
// 115-150: wave numbers with the shift and the Nyquist ("notMid") rule, 1/|k|^2 with the k = 0 guard, multiplication by
// -(i k_gd factor)) is cut out of the reference file at build time (gen_snippets.py -> oracle/_ref/
// poisson_grad_lambda.inc) and compiled here unchanged, applied to every index of a complex array.  The solver header
// itself needs heFFTe and cannot be included; the transforms around this step are not the reference's here.
#include <Kokkos_Core.hpp>
#include <Kokkos_MathematicalConstants.hpp>

#include <cstddef>

#include "Types/Vector.h"
#include "Index/NDIndex.h"

namespace Kokkos {
    // Kokkos::complex arithmetic is the plain textbook formula (no NaN recovery): the operators the lambda uses
    template <typename T> struct complex {
        T re, im;
        complex() : re(0), im(0) {}
        complex(T r, T i) : re(r), im(i) {}
        complex& operator*=(const complex& o) {
            const T r = re * o.re - im * o.im, i = re * o.im + im * o.re;
            re = r; im = i;
            return *this;
        }
    };
    template <typename T> complex<T> operator*(const complex<T>& a, const T& s) { return complex<T>(a.re * s, a.im * s); }
    template <typename T> complex<T> operator-(const complex<T>& a) { return complex<T>(-a.re, -a.im); }
}  // namespace Kokkos

namespace {
    struct CView3 {
        static constexpr unsigned rank = 3;
        Kokkos::complex<double>* p;
        long e0, e1;
        Kokkos::complex<double>& operator()(std::size_t i, std::size_t j, std::size_t k) const { return p[i + e0 * (j + e1 * k)]; }
    };
}  // namespace

extern "C" {

// spec_out[c] = spec_in[c] * -(i k_gd(c) / |k(c)|^2) for every index c of an nx x ny x nz complex array (x fastest,
// interleaved re / im), wave numbers of the periodic box [origin, origin + N h)
void refpoisson_grad_kspace(const int ng[3], const double origin_[3], const double h[3], int gd_, const double* spec_in,
                            double* spec_out) {
    constexpr unsigned Dim = 3;
    using scalar_type      = double;
    using Vector_t         = ippl::Vector<double, Dim>;
    using Complex_t        = Kokkos::complex<double>;
    using index_array_type = ippl::Vector<long, Dim>;
    const std::size_t n    = (std::size_t)ng[0] * ng[1] * ng[2];
    CView3 view{reinterpret_cast<Complex_t*>(const_cast<double*>(spec_in)), ng[0], ng[1]};
    CView3 tempview{reinterpret_cast<Complex_t*>(spec_out), ng[0], ng[1]};
    for (std::size_t i = 0; i < 2 * n; ++i) spec_out[i] = 0.0;
    const int nghost = 0;
    scalar_type pi   = Kokkos::numbers::pi_v<scalar_type>;
    ippl::Index ix(ng[0]), iy(ng[1]), iz(ng[2]);
    ippl::NDIndex<Dim> lDomComplex(ix, iy, iz);
    Vector_t origin, hx, rmax;
    ippl::Vector<int, Dim> N;
    for (std::size_t d = 0; d < Dim; ++d) {   // FFTPeriodicPoissonSolver.hpp:66-70
        origin[d] = origin_[d];
        hx[d]     = h[d];
        N[d]      = ng[d];
        rmax[d]   = origin[d] + (N[d] * hx[d]);
    }
    Complex_t imag = {0.0, 1.0};
    const std::size_t gd = (std::size_t)gd_;
    index_array_type args;
    for (args[2] = 0; args[2] < ng[2]; ++args[2])
        for (args[1] = 0; args[1] < ng[1]; ++args[1])
            for (args[0] = 0; args[0] < ng[0]; ++args[0]) {
                using ippl::apply;
                using ippl::Vector;
#include "poisson_grad_lambda.inc"
            }
}

}  // extern "C"
