// compat/Manager/FieldSolverBase.h -- ippl::FieldSolverBase (src/Manager/FieldSolverBase.h): solver type string + the
// variant of solver objects the drivers' FieldSolver fills
#ifndef IPPL_COMPAT_FIELD_SOLVER_BASE_H
#define IPPL_COMPAT_FIELD_SOLVER_BASE_H
#include <memory>
#include <string>
#include "Manager/BaseManager.h"
#include "Manager/datatypes.h"
namespace ippl {
template <typename T, unsigned Dim>
class FieldSolverBase {
public:
    explicit FieldSolverBase(std::string solver) : stype_m(std::move(solver)) {}
    virtual ~FieldSolverBase() = default;
    virtual void initSolver() = 0;
    virtual void runSolver()  = 0;
    std::string getStype() const { return stype_m; }
    void setStype(const std::string solver) { stype_m = solver; }
    Solver_t<T, Dim>& getSolver() { return solver_m; }

private:
    std::string stype_m;
    Solver_t<T, Dim> solver_m;
};
}  // namespace ippl
#endif
