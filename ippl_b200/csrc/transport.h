// transport.h -- the message step of the multi-rank phases (comm.cu): NCCL send / recv groups between processes, device
// copies between the in-process ranks of an ipplb_loop.  Shared with fftdist.cu.
#pragma once
#include <cstddef>
#include <vector>

#include "common.cuh"

struct ipplb_loop {
    std::vector<ipplb_ctx*> ctx;
};

namespace ipplb {

struct Xfer {
    int peer;
    const void* sptr; size_t sbytes;
    void* rptr; size_t rbytes;
};

// one grouped exchange on ctx's stream (one Xfer per peer)
int nccl_exchange(ipplb_ctx* ctx, const std::vector<Xfer>& x);
// all ranks in one process: rank a's send to b is matched with b's receive from a; synchronises every rank's stream
int loop_exchange(ipplb_loop* L, const std::vector<std::vector<Xfer>>& all);

}  // namespace ipplb
