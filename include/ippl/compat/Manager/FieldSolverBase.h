// compat/Manager/FieldSolverBase.h -- what the drivers' FieldSolver derives from (the reference keeps it in
// src/Manager/FieldSolverBase.h): it remembers the solver's name as given on the command line and owns the variant that
// holds whichever solver object initSolver() emplaces.  On the B200 path only "FFT" resolves to a working solver
// (Manager/datatypes.h).
#ifndef IPPL_COMPAT_FIELD_SOLVER_BASE_H
#define IPPL_COMPAT_FIELD_SOLVER_BASE_H
#include <string>
#include <utility>
#include "Manager/datatypes.h"
namespace ippl {
template <typename Real, unsigned D>
class FieldSolverBase {
    using variant_t = Solver_t<Real, D>;
    std::string name_;   // "FFT", "CG", ... as typed by the user
    variant_t active_;   // default-constructed alternative until initSolver() emplaces the chosen one

public:
    FieldSolverBase(std::string name) : name_(std::move(name)), active_() {}
    virtual ~FieldSolverBase() {}

    // driver hooks
    virtual void initSolver() = 0;
    virtual void runSolver() = 0;

    const std::string& getStype() const { return name_; }
    void setStype(const std::string& name) { name_ = name; }
    variant_t& getSolver() { return active_; }
};
}  // namespace ippl
#endif
