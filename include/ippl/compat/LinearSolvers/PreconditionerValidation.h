// compat/LinearSolvers/PreconditionerValidation.h -- stand-ins: preconditioned solvers are outside the B200 path
#pragma once
#include <string>
#include "Ippl.h"
namespace ippl {
namespace preconditioner_validation {
    inline void throwIfUnknownType(const std::string&, const char*) {}
    template <class... A>
    void sanitizeParams(A&&...) {}
}  // namespace preconditioner_validation
}  // namespace ippl
