// refshim.cpp -- TEST INFRASTRUCTURE.  Thin extern "C" entry points around the REAL reference
// headers (compiled in place from /root/reference/src, never copied) so tests can compare the
// oracle restatement with the reference's own arithmetic:
//   Interpolation/CIC.hpp            detail::scatterToField / gatherFromField
//   Particle/ParticleBC.h            detail::PeriodicBC
//   Index, NDIndex, Partitioner      detail::Partitioner<3>::split
//   FieldLayout/FieldLayout.hpp      findNeighbors / send+recv ranges / getMatchingIndex
// Kokkos, MPI and IpplTimings are replaced by the stand-ins in this directory.  Built only where
// /root/reference exists (this container); output goes to oracle/_ref/ (git-ignored).
#include <Kokkos_Core.hpp>

#include <array>
#include <cstddef>
#include <utility>
#include <vector>

#include "Utility/IpplException.h"
#include "Types/Vector.h"

#include "Index/NDIndex.h"
#include "Interpolation/CIC.hpp"
#include "Partition/Partitioner.h"
#include "Particle/ParticleBC.h"
#include "Region/NDRegion.h"

#include "FieldLayout/FieldLayout.h"

namespace {
    template <typename T>
    struct View3 {
        static constexpr unsigned rank = 3;
        using value_type               = T;
        T* p;
        long e0, e1;
        T& operator()(size_t i, size_t j, size_t k) const { return p[i + e0 * (j + e1 * k)]; }
    };
    struct PosView {
        using value_type = ippl::Vector<double, 3>;
        ippl::Vector<double, 3>* p;
        ippl::Vector<double, 3>& operator()(size_t i) const { return p[i]; }
    };
    struct BoxView {
        std::vector<ippl::NDIndex<3>>* v;
        ippl::NDIndex<3>& operator()(size_t i) const { return (*v)[i]; }
    };
}  // namespace

extern "C" {

// The index/weight lines are the call site of ParticleAttrib::scatter (ParticleAttrib.hpp:174-179)
// evaluated with the reference's own Vector expression templates; the deposit is the reference's
// detail::scatterToField.
void ref_scatter(long n, const double* x, const double* y, const double* z, const double* q,
                 const double* origin_, const double* h_, const int* first, int nghost, double* rho,
                 long e0, long e1) {
    using vector_type = ippl::Vector<double, 3>;
    vector_type origin = {origin_[0], origin_[1], origin_[2]};
    vector_type dx     = {h_[0], h_[1], h_[2]};
    const vector_type invdx = 1.0 / dx;
    ippl::Vector<int, 3> lfirst = {first[0], first[1], first[2]};
    View3<double> view{rho, e0, e1};
    for (long i = 0; i < n; ++i) {
        vector_type pp = {x[i], y[i], z[i]};
        vector_type l                = (pp - origin) * invdx + 0.5;
        ippl::Vector<int, 3> index   = l;
        ippl::Vector<double, 3> whi  = l - index;
        ippl::Vector<double, 3> wlo  = 1.0 - whi;
        ippl::Vector<size_t, 3> args = index - lfirst + nghost;
        const double& val            = q[i];
        ippl::detail::scatterToField(std::make_index_sequence<1 << 3>{}, view, wlo, whi, args, val);
    }
}

void ref_gather(long n, const double* x, const double* y, const double* z, const double* origin_,
                const double* h_, const int* first, int nghost, const double* efield, long e0,
                long e1, double* ex, double* ey, double* ez) {
    using vector_type = ippl::Vector<double, 3>;
    vector_type origin = {origin_[0], origin_[1], origin_[2]};
    vector_type dx     = {h_[0], h_[1], h_[2]};
    const vector_type invdx = 1.0 / dx;
    ippl::Vector<int, 3> lfirst = {first[0], first[1], first[2]};
    View3<vector_type> view{reinterpret_cast<vector_type*>(const_cast<double*>(efield)), e0, e1};
    static_assert(sizeof(vector_type) == 3 * sizeof(double));
    for (long i = 0; i < n; ++i) {
        vector_type pp = {x[i], y[i], z[i]};
        vector_type l                = (pp - origin) * invdx + 0.5;
        ippl::Vector<int, 3> index   = l;
        ippl::Vector<double, 3> whi  = l - index;
        ippl::Vector<double, 3> wlo  = 1.0 - whi;
        ippl::Vector<size_t, 3> args = index - lfirst + nghost;
        vector_type g = ippl::detail::gatherFromField(std::make_index_sequence<1 << 3>{}, view, wlo,
                                                      whi, args);
        ex[i] = g[0];
        ey[i] = g[1];
        ez[i] = g[2];
    }
}

void ref_periodic_bc(long n, double* x, double* y, double* z, const double* lo, const double* hi) {
    std::vector<ippl::Vector<double, 3>> R(n);
    for (long i = 0; i < n; ++i) R[i] = {x[i], y[i], z[i]};
    ippl::NDRegion<double, 3> nr;
    for (unsigned d = 0; d < 3; ++d) nr[d] = ippl::PRegion<double>(lo[d], hi[d]);
    PosView view{R.data()};
    for (unsigned d = 0; d < 3; ++d) {
        ippl::detail::PeriodicBC<double, 3, PosView> bc(view, nr, d, false);
        for (long i = 0; i < n; ++i) bc((size_t)i);
    }
    for (long i = 0; i < n; ++i) {
        x[i] = R[i][0];
        y[i] = R[i][1];
        z[i] = R[i][2];
    }
}

int ref_partition(const int* ng, const int* is_parallel, int nsplits, int* boxes_out) {
    ippl::Index i0(ng[0]), i1(ng[1]), i2(ng[2]);
    ippl::NDIndex<3> domain(i0, i1, i2);
    std::vector<ippl::NDIndex<3>> boxes(nsplits);
    BoxView view{&boxes};
    std::array<bool, 3> par = {is_parallel[0] != 0, is_parallel[1] != 0, is_parallel[2] != 0};
    ippl::detail::Partitioner<3> p;
    try {
        p.split(domain, view, par, nsplits);
    } catch (...) { return -1; }
    for (int r = 0; r < nsplits; ++r)
        for (int d = 0; d < 3; ++d) {
            boxes_out[r * 6 + d]     = boxes[r][d].first();
            boxes_out[r * 6 + 3 + d] = boxes[r][d].last();
        }
    return 0;
}

// FieldLayout<3> built as rank `my` of `nranks`: returns entries in component order, each
// comp, rank, send lo[3], send hi[3], recv lo[3], recv hi[3]; and the rank boxes.
int ref_neighbors(const int* ng, const int* is_parallel, int nranks, int periodic, int nghost, int my,
                  int* boxes_out, int* out, int max_entries) {
    refshim::g_rank = my;
    refshim::g_size = nranks;
    ippl::Index i0(ng[0]), i1(ng[1]), i2(ng[2]);
    ippl::NDIndex<3> domain(i0, i1, i2);
    std::array<bool, 3> par = {is_parallel[0] != 0, is_parallel[1] != 0, is_parallel[2] != 0};
    ippl::FieldLayout<3> fl(ippl::mpi::Communicator(), domain, par, periodic != 0, nghost);
    for (int r = 0; r < nranks; ++r)
        for (int d = 0; d < 3; ++d) {
            boxes_out[r * 6 + d]     = fl.getLocalNDIndex(r)[d].first();
            boxes_out[r * 6 + 3 + d] = fl.getLocalNDIndex(r)[d].last();
        }
    int cnt = 0;
    if (nranks < 2) return 0;
    const auto& nb = fl.getNeighbors();
    const auto& sr = fl.getNeighborsSendRange();
    const auto& rr = fl.getNeighborsRecvRange();
    for (size_t comp = 0; comp < nb.size(); ++comp)
        for (size_t i = 0; i < nb[comp].size(); ++i) {
            if (cnt < max_entries) {
                int* o = out + cnt * 14;
                o[0]   = (int)comp;
                o[1]   = nb[comp][i];
                for (int d = 0; d < 3; ++d) {
                    o[2 + d]  = (int)sr[comp][i].lo[d];
                    o[5 + d]  = (int)sr[comp][i].hi[d];
                    o[8 + d]  = (int)rr[comp][i].lo[d];
                    o[11 + d] = (int)rr[comp][i].hi[d];
                }
            }
            ++cnt;
        }
    return cnt;
}

// the same tables after FieldLayout::updateLayout(domains) with caller-supplied rank boxes ([nranks][6] lo, hi inclusive):
// what an ORB repartition leaves behind (FieldLayout.hpp:59-73 -> findNeighbors)
int ref_neighbors_boxes(const int* ng, int nranks, int periodic, int nghost, int my, const int* boxes, int* out,
                        int max_entries) {
    refshim::g_rank = my;
    refshim::g_size = nranks;
    ippl::Index i0(ng[0]), i1(ng[1]), i2(ng[2]);
    ippl::NDIndex<3> domain(i0, i1, i2);
    std::array<bool, 3> par = {true, true, true};
    ippl::FieldLayout<3> fl(ippl::mpi::Communicator(), domain, par, periodic != 0, nghost);
    std::vector<ippl::NDIndex<3>> doms(nranks);
    for (int r = 0; r < nranks; ++r)
        doms[r] = ippl::NDIndex<3>(ippl::Index(boxes[6 * r], boxes[6 * r + 3]), ippl::Index(boxes[6 * r + 1], boxes[6 * r + 4]),
                                   ippl::Index(boxes[6 * r + 2], boxes[6 * r + 5]));
    fl.updateLayout(doms);
    int cnt = 0;
    const auto& nb = fl.getNeighbors();
    const auto& sr = fl.getNeighborsSendRange();
    const auto& rr = fl.getNeighborsRecvRange();
    for (size_t comp = 0; comp < nb.size(); ++comp)
        for (size_t i = 0; i < nb[comp].size(); ++i) {
            if (cnt < max_entries) {
                int* o = out + cnt * 14;
                o[0]   = (int)comp;
                o[1]   = nb[comp][i];
                for (int d = 0; d < 3; ++d) {
                    o[2 + d]  = (int)sr[comp][i].lo[d];
                    o[5 + d]  = (int)sr[comp][i].hi[d];
                    o[8 + d]  = (int)rr[comp][i].lo[d];
                    o[11 + d] = (int)rr[comp][i].hi[d];
                }
            }
            ++cnt;
        }
    return cnt;
}

int ref_matching_index(int i) { return ippl::FieldLayout<3>::getMatchingIndex(i); }
}
