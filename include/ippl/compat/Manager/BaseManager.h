// compat/Manager/BaseManager.h -- ippl::BaseManager (src/Manager/BaseManager.h:13-75): the run loop of a mini-app
#ifndef IPPL_COMPAT_BASE_MANAGER_H
#define IPPL_COMPAT_BASE_MANAGER_H
#include "Ippl.h"
namespace ippl {
class BaseManager {
public:
    BaseManager()          = default;
    virtual ~BaseManager() = default;
    virtual void pre_run() {}
    virtual void pre_step() {}
    virtual void post_step() {}
    virtual void advance() = 0;
    // nt times: pre_step, advance, post_step
    void run(int nt) {
        for (int it = 0; it < nt; ++it) {
            pre_step();
            advance();
            post_step();
        }
    }
};
}  // namespace ippl
#endif
