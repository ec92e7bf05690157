// compat/Manager/BaseManager.h -- ippl::BaseManager (src/Manager/BaseManager.h:13-75): the run loop of a mini-app.
// advance() is the one hook a manager must supply; the other three default to nothing.
#ifndef IPPL_COMPAT_BASE_MANAGER_H
#define IPPL_COMPAT_BASE_MANAGER_H
#include "Ippl.h"
namespace ippl {

class BaseManager {
public:
    virtual ~BaseManager() = default;

    virtual void advance() = 0;
    virtual void pre_run() {}
    virtual void pre_step() {}
    virtual void post_step() {}

    // `steps` time steps, each bracketed by its hooks
    void run(int steps) {
        int done = 0;
        while (done < steps) {
            pre_step(), advance(), post_step();
            ++done;
        }
    }
};

}  // namespace ippl
#endif
