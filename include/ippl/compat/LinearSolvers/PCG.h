// compat/LinearSolvers/PCG.h -- the preconditioner defaults the drivers' FieldSolver parses its command line against
// (values as in the reference's src/LinearSolvers; the CG solver itself is a throwing stand-in, see Manager/datatypes.h)
#pragma once
namespace ippl {
namespace pcg_preconditioner_defaults {
    inline constexpr int newton_level          = 5;
    inline constexpr int chebyshev_degree      = 31;
    inline constexpr int richardson_iterations = 4;
    inline constexpr int gauss_seidel_inner    = 2;
    inline constexpr int gauss_seidel_outer    = 2;
    inline constexpr int communication         = 0;
    inline constexpr double ssor_omega         = 1.57079632679;
    inline constexpr int mg_pre_smooth         = 2;
    inline constexpr int mg_post_smooth        = 2;
    inline constexpr double mg_omega           = 0.8;
    inline constexpr int mg_min_cells          = 4;
}  // namespace pcg_preconditioner_defaults
}  // namespace ippl
