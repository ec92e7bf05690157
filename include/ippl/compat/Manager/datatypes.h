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
}  // namespace detail
template <class Lhs, class Rhs> using PoissonCG = detail::UnavailableSolver<0, Lhs, Rhs>;
template <class Lhs, class Rhs> using NullSolver = detail::UnavailableSolver<1, Lhs, Rhs>;
template <class Lhs, class Rhs> using FFTTruncatedGreenPeriodicPoissonSolver = detail::UnavailableSolver<2, Lhs, Rhs>;
template <class Lhs, class Rhs> using FFTOpenPoissonSolver = detail::UnavailableSolver<3, Lhs, Rhs>;
template <class Lhs, class Rhs> using FEMPoissonSolver = detail::UnavailableSolver<4, Lhs, Rhs>;
template <class Lhs, class Rhs> using PreconditionedFEMPoissonSolver = detail::UnavailableSolver<5, Lhs, Rhs>;
}  // namespace ippl

// ---- the global aliases the drivers spell (names fixed by the reference's src/Manager/datatypes.h) -------------------
// geometry and layouts
template <unsigned D> using Mesh_t = ippl::UniformCartesian<double, D>;
template <unsigned D> using Centering_t = typename Mesh_t<D>::DefaultCentering;
template <unsigned D> using FieldLayout_t = ippl::FieldLayout<D>;
template <typename Real, unsigned D> using PLayout_t = ippl::ParticleSpatialLayout<Real, D, Mesh_t<D>>;
using size_type = ippl::detail::size_type;

// values, attributes, fields (trailing packs: the reference forwards Kokkos view arguments, the SoA facade has none)
template <typename V, unsigned D> using Vector = ippl::Vector<V, D>;
template <typename V, unsigned D> using Vector_t = ippl::Vector<V, D>;
template <typename V> using ParticleAttrib = ippl::ParticleAttrib<V>;
template <typename V, unsigned D = 3, class... Unused> using Field = ippl::Field<V, D, Mesh_t<D>, Centering_t<D>>;
template <unsigned D, class... Unused> using Field_t = Field<double, D>;
template <typename Real = double, unsigned D = 3, class... Unused> using VField_t = Field<Vector_t<Real, D>, D>;
template <typename Real = double, unsigned D = 3> using ORB = ippl::OrthogonalRecursiveBisection<Field<double, D>, Real>;

// solvers: IPPLC_SOLVER(alias, class, lhs, rhs) declares `alias<Real, D>`
#define IPPLC_SOLVER(ALIAS, CLASS, LHS, RHS) \
    template <typename Real = double, unsigned D = 3> using ALIAS = ippl::CLASS<LHS, RHS>
#define IPPLC_SCALAR Field<Real, D>
#define IPPLC_GRAD VField_t<Real, D>
IPPLC_SOLVER(FFTSolver_t, FFTPeriodicPoissonSolver, IPPLC_GRAD, Field_t<D>);   // the one wired to cuFFT
IPPLC_SOLVER(CGSolver_t, PoissonCG, IPPLC_SCALAR, Field_t<D>);
IPPLC_SOLVER(NullSolver_t, NullSolver, IPPLC_GRAD, Field_t<D>);
IPPLC_SOLVER(FFTTruncatedGreenSolver_t, FFTTruncatedGreenPeriodicPoissonSolver, IPPLC_GRAD, Field_t<D>);
IPPLC_SOLVER(OpenSolver_t, FFTOpenPoissonSolver, IPPLC_GRAD, Field_t<D>);
IPPLC_SOLVER(FEMSolver_t, FEMPoissonSolver, IPPLC_SCALAR, IPPLC_SCALAR);
IPPLC_SOLVER(FEMPreconSolver_t, PreconditionedFEMPoissonSolver, IPPLC_SCALAR, IPPLC_SCALAR);
#undef IPPLC_SOLVER
// alternative order = the order the drivers' std::get<> / holds_alternative<> calls were written against
template <typename Real = double, unsigned D = 3>
using Solver_t = std::variant<CGSolver_t<Real, D>, FFTSolver_t<Real, D>, FFTTruncatedGreenSolver_t<Real, D>, OpenSolver_t<Real, D>,
                              NullSolver_t<Real, D>, FEMSolver_t<Real, D>, FEMPreconSolver_t<Real, D>>;
#undef IPPLC_SCALAR
#undef IPPLC_GRAD

extern const char* TestName;   // every driver defines it
#endif
