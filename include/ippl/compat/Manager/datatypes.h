// compat/Manager/datatypes.h -- the type aliases of src/Manager/datatypes.h on the B200 facade.  The periodic FFT solver is
// the facade's (cuFFT, non-owned stage); every other solver the drivers' FieldSolver can name (CG / PCG, truncated Green,
// open boundaries, the two FEM solvers, the null solver) is a stand-in that satisfies the types and throws when it is
// selected at run time: they are outside the hot path (out of scope).
#ifndef IPPL_COMPAT_DATATYPES_H
#define IPPL_COMPAT_DATATYPES_H
#include <variant>
#include "Ippl.h"

namespace ippl {
namespace detail {
    // the finite-element space handed to the FEM particle <-> mesh transfers
    struct NoSpace {
        template <class L>
        void updateLayout(const L&) {}
    };
    template <int Tag, class Lhs, class Rhs>
    class UnavailableSolver {
    public:
        enum OutputType { SOL = 1, GRAD = 2, SOL_AND_GRAD = 3 };
        enum Algorithm { HOCKNEY = 1, VICO = 2, BIHARMONIC = 3, DCT_VICO = 4 };
        void mergeParameters(const ParameterList&) {}
        template <class F> void setRhs(F&) {}
        template <class F> void setLhs(F&) {}
        template <class F> void setGradient(F&) {}
        [[noreturn]] void solve() { throw IpplException("solver", "only the periodic FFT solver (\"FFT\") is wired to the B200 path"); }
        int getIterationCount() const { return 0; }
        double getResidue() const { return 0.0; }
        NoSpace& getSpace() { return space_; }

    private:
        NoSpace space_;
    };
    template <bool B, class T>
    using ConditionalType = T;   // every alias below is used with Dim == 3
    template <class... T>
    using VariantFromConditionalTypes = std::variant<T...>;
}  // namespace detail
template <class Lhs, class Rhs> using PoissonCG = detail::UnavailableSolver<0, Lhs, Rhs>;
template <class Lhs, class Rhs> using NullSolver = detail::UnavailableSolver<1, Lhs, Rhs>;
template <class Lhs, class Rhs> using FFTTruncatedGreenPeriodicPoissonSolver = detail::UnavailableSolver<2, Lhs, Rhs>;
template <class Lhs, class Rhs> using FFTOpenPoissonSolver = detail::UnavailableSolver<3, Lhs, Rhs>;
template <class Lhs, class Rhs> using FEMPoissonSolver = detail::UnavailableSolver<4, Lhs, Rhs>;
template <class Lhs, class Rhs> using PreconditionedFEMPoissonSolver = detail::UnavailableSolver<5, Lhs, Rhs>;
}  // namespace ippl

template <unsigned Dim>
using Mesh_t = ippl::UniformCartesian<double, Dim>;
template <typename T, unsigned Dim>
using PLayout_t = typename ippl::ParticleSpatialLayout<T, Dim, Mesh_t<Dim>>;
template <unsigned Dim>
using Centering_t = typename Mesh_t<Dim>::DefaultCentering;
template <unsigned Dim>
using FieldLayout_t = ippl::FieldLayout<Dim>;
using size_type = ippl::detail::size_type;
template <typename T, unsigned Dim>
using Vector = ippl::Vector<T, Dim>;
template <typename T, unsigned Dim = 3, class... ViewArgs>
using Field = ippl::Field<T, Dim, Mesh_t<Dim>, Centering_t<Dim>>;
template <typename T = double, unsigned Dim = 3>
using ORB = ippl::OrthogonalRecursiveBisection<Field<double, Dim>, T>;
template <typename T>
using ParticleAttrib = ippl::ParticleAttrib<T>;
template <typename T, unsigned Dim>
using Vector_t = ippl::Vector<T, Dim>;
template <unsigned Dim, class... ViewArgs>
using Field_t = Field<double, Dim>;
template <typename T = double, unsigned Dim = 3, class... ViewArgs>
using VField_t = Field<Vector_t<T, Dim>, Dim>;
template <typename T = double, unsigned Dim = 3>
using CGSolver_t = ippl::PoissonCG<Field<T, Dim>, Field_t<Dim>>;
template <typename T = double, unsigned Dim = 3>
using NullSolver_t = ippl::NullSolver<VField_t<T, Dim>, Field_t<Dim>>;
using ippl::detail::ConditionalType, ippl::detail::VariantFromConditionalTypes;
template <typename T = double, unsigned Dim = 3>
using FFTSolver_t = ConditionalType<Dim == 2 || Dim == 3, ippl::FFTPeriodicPoissonSolver<VField_t<T, Dim>, Field_t<Dim>>>;
template <typename T = double, unsigned Dim = 3>
using FFTTruncatedGreenSolver_t = ConditionalType<Dim == 3, ippl::FFTTruncatedGreenPeriodicPoissonSolver<VField_t<T, Dim>, Field_t<Dim>>>;
template <typename T = double, unsigned Dim = 3>
using OpenSolver_t = ConditionalType<Dim == 3, ippl::FFTOpenPoissonSolver<VField_t<T, Dim>, Field_t<Dim>>>;
template <typename T = double, unsigned Dim = 3>
using FEMSolver_t = ippl::FEMPoissonSolver<Field<T, Dim>, Field<T, Dim>>;
template <typename T = double, unsigned Dim = 3>
using FEMPreconSolver_t = ippl::PreconditionedFEMPoissonSolver<Field<T, Dim>, Field<T, Dim>>;
template <typename T = double, unsigned Dim = 3>
using Solver_t = VariantFromConditionalTypes<CGSolver_t<T, Dim>, FFTSolver_t<T, Dim>, FFTTruncatedGreenSolver_t<T, Dim>,
                                             OpenSolver_t<T, Dim>, NullSolver_t<T, Dim>, FEMSolver_t<T, Dim>,
                                             FEMPreconSolver_t<T, Dim>>;
extern const char* TestName;
#endif
