// compat/Manager/PicManager.h -- ippl::PicManager (src/Manager/PicManager.h:31-155): a BaseManager that owns the particle
// container(s), the field container, the field solver and the load balancer of a PIC mini-app.  The drivers' managers
// reach into the protected members by name (this->pcontainer_m, fcontainer_m, fsolver_m, loadbalancer_m), so those
// names are part of the interface; everything else is this repo's own arrangement.
#ifndef IPPL_COMPAT_PIC_MANAGER_H
#define IPPL_COMPAT_PIC_MANAGER_H
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>
#include "Decomposition/OrthogonalRecursiveBisection.h"
#include "Manager/BaseManager.h"
#include "Manager/FieldSolverBase.h"

// one owned component: get<Name>() / set<Name>(shared_ptr) over a protected member
#define IPPL_COMPAT_OWNED(Type, Name, member)                      \
    std::shared_ptr<Type> get##Name() { return this->member; }     \
    void set##Name(std::shared_ptr<Type> v_) { this->member = std::move(v_); }

namespace ippl {

template <typename T, unsigned Dim, class pc, class fc, class orb>
class PicManager : public BaseManager {
    using Solver_t = ippl::FieldSolverBase<T, Dim>;
    using Bunch_t  = std::shared_ptr<pc>;

protected:
    // (declared first: the accessors below refer to them)
    std::shared_ptr<fc> fcontainer_m;
    Bunch_t pcontainer_m;                    // the primary bunch == pcontainers_m[0] once one is set
    std::vector<Bunch_t> pcontainers_m;      // every bunch, in the order it was added
    std::shared_ptr<orb> loadbalancer_m;
    std::shared_ptr<Solver_t> fsolver_m;

public:
    virtual ~PicManager() = default;

    // the two halves of a PIC step every mini-app supplies
    virtual void par2grid() = 0;
    virtual void grid2par() = 0;

    IPPL_COMPAT_OWNED(fc, FieldContainer, fcontainer_m)
    IPPL_COMPAT_OWNED(Solver_t, FieldSolver, fsolver_m)
    IPPL_COMPAT_OWNED(orb, LoadBalancer, loadbalancer_m)

    // bunches: slot 0 is "the" particle container
    Bunch_t getParticleContainer() { return pcontainer_m; }
    Bunch_t getParticleContainer(size_t slot) { return pcontainers_m.at(slot); }
    const std::vector<Bunch_t>& getParticleContainers() const { return pcontainers_m; }
    size_t getNumParticleContainers() const { return pcontainers_m.size(); }
    void setParticleContainer(Bunch_t bunch) {
        if (pcontainers_m.empty()) pcontainers_m.resize(1);
        pcontainers_m.front() = bunch;
        pcontainer_m          = std::move(bunch);
    }
    size_t addParticleContainer(Bunch_t bunch) {
        const size_t slot = pcontainers_m.size();
        if (slot == 0) pcontainer_m = bunch;
        pcontainers_m.push_back(std::move(bunch));
        return slot;
    }
};

}  // namespace ippl
#undef IPPL_COMPAT_OWNED
#endif
