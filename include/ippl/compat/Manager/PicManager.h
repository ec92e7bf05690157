// compat/Manager/PicManager.h -- ippl::PicManager (src/Manager/PicManager.h:31-155): a BaseManager that owns the particle
// container(s), the field container, the field solver and the load balancer of a PIC mini-app
#ifndef IPPL_COMPAT_PIC_MANAGER_H
#define IPPL_COMPAT_PIC_MANAGER_H
#include <memory>
#include <stdexcept>
#include <vector>
#include "Decomposition/OrthogonalRecursiveBisection.h"
#include "Manager/BaseManager.h"
#include "Manager/FieldSolverBase.h"
namespace ippl {
template <typename T, unsigned Dim, class pc, class fc, class orb>
class PicManager : public BaseManager {
public:
    PicManager() = default;
    virtual ~PicManager() = default;
    virtual void par2grid() = 0;
    virtual void grid2par() = 0;
    std::shared_ptr<pc> getParticleContainer() { return pcontainer_m; }
    std::shared_ptr<pc> getParticleContainer(size_t i) { return pcontainers_m.at(i); }
    void setParticleContainer(std::shared_ptr<pc> p) {
        pcontainer_m = p;
        if (pcontainers_m.empty()) pcontainers_m.push_back(p);
        else pcontainers_m[0] = p;
    }
    size_t addParticleContainer(std::shared_ptr<pc> p) {
        pcontainers_m.push_back(p);
        if (pcontainers_m.size() == 1) pcontainer_m = p;
        return pcontainers_m.size() - 1;
    }
    size_t getNumParticleContainers() const { return pcontainers_m.size(); }
    const std::vector<std::shared_ptr<pc>>& getParticleContainers() const { return pcontainers_m; }
    std::shared_ptr<fc> getFieldContainer() { return fcontainer_m; }
    void setFieldContainer(std::shared_ptr<fc> f) { fcontainer_m = f; }
    std::shared_ptr<ippl::FieldSolverBase<T, Dim>> getFieldSolver() { return fsolver_m; }
    void setFieldSolver(std::shared_ptr<ippl::FieldSolverBase<T, Dim>> s) { fsolver_m = s; }
    std::shared_ptr<orb> getLoadBalancer() { return loadbalancer_m; }
    void setLoadBalancer(std::shared_ptr<orb> l) { loadbalancer_m = l; }

protected:
    std::shared_ptr<fc> fcontainer_m;
    std::shared_ptr<pc> pcontainer_m;
    std::vector<std::shared_ptr<pc>> pcontainers_m;
    std::shared_ptr<orb> loadbalancer_m;
    std::shared_ptr<ippl::FieldSolverBase<T, Dim>> fsolver_m;
};
}  // namespace ippl
#endif
