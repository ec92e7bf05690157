// compat/FEM/FEMInterpolate.hpp -- the finite-element particle <-> mesh transfers are outside the B200 path (SURVEY 8:
// out of scope); the drivers only reach them with the FEM solvers, which throw when selected
#pragma once
#include "Ippl.h"
namespace ippl {
template <class... A>
void interpolate_grad_to_diracs(A&&...) { throw IpplException("interpolate_grad_to_diracs", "FEM solvers are not part of the B200 path"); }
template <class... A>
void assemble_rhs_from_particles(A&&...) { throw IpplException("assemble_rhs_from_particles", "FEM solvers are not part of the B200 path"); }
}  // namespace ippl
using ippl::assemble_rhs_from_particles;
using ippl::interpolate_grad_to_diracs;
