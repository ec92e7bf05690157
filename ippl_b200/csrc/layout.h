// layout.h -- host-side decomposition types shared by layout.cpp and comm.cu
#pragma once
#include <vector>

namespace ipplb {

struct IBox {
    int lo[3], hi[3];  // inclusive global cell indices
};

struct NeighborEntry {
    int comp;  // base-3 component: digit d = 0 lower, 1 upper, 2 parallel
    int peer;
    int send_lo[3], send_hi[3];  // my interior strip the peer's halo needs (local ghosted idx, hi excl.)
    int recv_lo[3], recv_hi[3];  // my halo strip the peer owns
};

int matching_component(int comp);

struct Layout {
    int ng[3]  = {0, 0, 0};
    int par[3] = {1, 1, 1};
    bool periodic = true;
    int nghost    = 1;
    std::vector<IBox> boxes;
    int init(const int ng[3], const int par[3], int nranks, int periodic, int nghost);
    std::vector<NeighborEntry> neighbors(int me) const;
    void regions(const double origin[3], const double h[3], double* out) const;
};

}  // namespace ipplb

struct ipplb_layout {
    ipplb::Layout L;
};
