// layout.cpp -- host-only decomposition logic: rank boxes, neighbour components with send/recv
// ranges, physical regions.  Must agree bit-for-bit with the reference's FieldLayout / Partitioner /
// RegionLayout (src/FieldLayout/FieldLayout.hpp:76-362, src/Partition/Partitioner.hpp:15-123,
// src/Index/Index.hpp:162-191, src/Region/RegionLayout.hpp:68-98) because particle ownership and
// halo contents are defined by it; tests compare it with the oracle and the reference's golden tables.
#include "layout.h"

#include <algorithm>
#include <cstring>

#include "../../include/ippl_b200.h"

namespace ipplb {

void set_error(const char* fmt, ...);

namespace {

IBox grown(const IBox& b, int g) {
    IBox r = b;
    for (int d = 0; d < 3; ++d) {
        r.lo[d] -= g;
        r.hi[d] += g;
    }
    return r;
}
bool overlaps(const IBox& a, const IBox& b) {
    for (int d = 0; d < 3; ++d)
        if (a.lo[d] > b.hi[d] || a.hi[d] < b.lo[d]) return false;
    return true;
}
IBox common(const IBox& a, const IBox& b) {
    IBox r;
    for (int d = 0; d < 3; ++d) {
        r.lo[d] = std::max(a.lo[d], b.lo[d]);
        r.hi[d] = std::min(a.hi[d], b.hi[d]);
    }
    return r;
}
IBox shifted(const IBox& b, const int s[3], int sign) {
    IBox r = b;
    for (int d = 0; d < 3; ++d) {
        r.lo[d] += sign * s[d];
        r.hi[d] += sign * s[d];
    }
    return r;
}

// Recursive bisection.  Power-of-two rank counts: level-wise halving that cycles through the
// parallel dims, first cut = most significant bit of the rank id.  Otherwise: cut the longest
// parallel dim in the ratio floor(n/2) : n - floor(n/2), ranks [lo, lo+n/2) left.
void bisect_pow2(const IBox& box, int n, int rank0, unsigned d, const int par[3],
                 std::vector<IBox>& out) {
    if (n == 1) {
        out[rank0] = box;
        return;
    }
    while (!par[d]) d = (d + 1) % 3;
    const int len = box.hi[d] - box.lo[d] + 1;
    const int mid = box.lo[d] + len / 2 - 1;
    IBox l = box, r = box;
    l.hi[d] = mid;
    r.lo[d] = mid + 1;
    bisect_pow2(l, n / 2, rank0, (d + 1) % 3, par, out);
    bisect_pow2(r, n / 2, rank0 + n / 2, (d + 1) % 3, par, out);
}

void bisect_ratio(const IBox& box, int lo, int hi, const int par[3], std::vector<IBox>& out) {
    const int n = hi - lo;
    if (n <= 1) {
        out[lo] = box;
        return;
    }
    const int cut = lo + n / 2;
    double a      = cut - lo;
    a /= hi - lo;
    int d = -1;
    double longest = 0;
    for (int dd = 0; dd < 3; ++dd)
        if (par[dd]) {
            double len = box.hi[dd] - box.lo[dd] + 1;
            if (len > longest) {
                longest = len;
                d       = dd;
            }
        }
    const int len = box.hi[d] - box.lo[d] + 1;
    const int mid = box.lo[d] + static_cast<int>(len * a + 0.5) - 1;
    IBox l = box, r = box;
    l.hi[d] = mid;
    r.lo[d] = mid + 1;
    bisect_ratio(l, lo, cut, par, out);
    bisect_ratio(r, cut, hi, par, out);
}

}  // namespace

int Layout::init(const int ng_[3], const int par_[3], int nranks, int periodic_, int nghost_) {
    for (int d = 0; d < 3; ++d) {
        ng[d]  = ng_[d];
        par[d] = par_[d];
    }
    periodic = periodic_ != 0;
    nghost   = nghost_;
    boxes.assign(nranks, IBox());
    IBox dom;
    for (int d = 0; d < 3; ++d) {
        dom.lo[d] = 0;
        dom.hi[d] = ng[d] - 1;
    }
    if (nranks == 1) {
        boxes[0] = dom;
        return 0;
    }
    long totpar = 1;
    for (int d = 0; d < 3; ++d) totpar *= par[d] ? ng[d] : 1;
    if (totpar < nranks || !(par[0] || par[1] || par[2])) {
        set_error("layout: domain cannot be partitioned into %d local domains", nranks);
        return IPPLB_ERR_ARG;
    }
    if ((nranks & (nranks - 1)) == 0)
        bisect_pow2(dom, nranks, 0, 0, par, boxes);
    else
        bisect_ratio(dom, 0, nranks, par, boxes);
    return 0;
}

// All neighbour components of rank `me`, in component order, then peer-ascending discovery order.
std::vector<NeighborEntry> Layout::neighbors(int me) const {
    std::vector<NeighborEntry> out;
    const int nr = (int)boxes.size();
    if (nr < 2) return out;
    const IBox& mine = boxes[me];
    const IBox halo  = grown(mine, nghost);

    auto record = [&](const IBox& halo_img, const IBox& peer_img, const IBox& isect, int peer) {
        NeighborEntry e;
        e.peer = peer;
        // send: my interior cells the peer's halo needs; recv: my halo cells the peer owns
        const IBox s = common(grown(peer_img, nghost), mine);
        const IBox r = common(halo, peer_img);
        int code = 0, digit = 1;
        for (int d = 0; d < 3; ++d, digit *= 3) {
            e.send_lo[d] = s.lo[d] - mine.lo[d] + nghost;
            e.send_hi[d] = s.hi[d] - mine.lo[d] + nghost + 1;
            e.recv_lo[d] = r.lo[d] - mine.lo[d] + nghost;
            e.recv_hi[d] = r.hi[d] - mine.lo[d] + nghost + 1;
            const int ilen = isect.hi[d] - isect.lo[d] + 1;
            if (ilen == nghost)
                code += (halo_img.lo[d] != isect.lo[d]) ? digit : 0;  // 1 = upper face, 0 = lower
            else
                code += 2 * digit;  // parallel to this axis
        }
        e.comp = code;
        out.push_back(e);
    };

    for (int peer = 0; peer < nr; ++peer) {
        if (peer == me) continue;
        const IBox& theirs = boxes[peer];
        if (overlaps(halo, theirs)) record(halo, theirs, common(halo, theirs), peer);
        if (!periodic) continue;
        // periodic images: shift my halo by -period where I touch the upper domain face, +period where
        // I touch the lower one; combine dims in increasing order (up to all three)
        int shift[3] = {0, 0, 0};
        auto images  = [&](auto&& self, int d0, int depth) -> void {
            for (int d = d0; d < 3; ++d)
                for (int k = 0; k < 2; ++k) {
                    int off = 0;
                    if (k == 0 && mine.hi[d] == ng[d] - 1) off = -ng[d];
                    if (k == 1 && mine.lo[d] == 0) off = ng[d];
                    if (!off) continue;
                    shift[d]          = off;
                    const IBox himg   = shifted(halo, shift, +1);
                    if (overlaps(himg, theirs))
                        record(himg, shifted(theirs, shift, -1), common(himg, theirs), peer);
                    if (depth + 1 < 3) self(self, d + 1, depth + 1);
                    shift[d] = 0;
                }
        };
        images(images, 0, 0);
    }
    std::stable_sort(out.begin(), out.end(),
                     [](const NeighborEntry& a, const NeighborEntry& b) { return a.comp < b.comp; });
    return out;
}

int matching_component(int comp) {
    // swap lower<->upper in every digit, keep "parallel"
    int m = 0;
    for (int d = 0, w = 1; d < 3; ++d, w *= 3) {
        int digit = comp % 3;
        comp /= 3;
        m += (digit == 2 ? 2 : 1 - digit) * w;
    }
    return m;
}

void Layout::regions(const double origin[3], const double h[3], double* out) const {
    for (size_t r = 0; r < boxes.size(); ++r)
        for (int d = 0; d < 3; ++d) {
            out[r * 6 + d]     = boxes[r].lo[d] * h[d] + origin[d];
            out[r * 6 + 3 + d] = (boxes[r].hi[d] + 1) * h[d] + origin[d];
        }
}

}  // namespace ipplb

using ipplb::Layout;

extern "C" {

int ipplb_layout_create(ipplb_layout** out, const int ng[3], const int is_parallel[3], int nranks,
                        int periodic, int nghost) {
    if (!out || !ng || !is_parallel || nranks < 1 || nghost < 0) {
        ipplb::set_error("layout_create: bad arguments");
        return IPPLB_ERR_ARG;
    }
    ipplb_layout* l = new ipplb_layout();
    int rc          = l->L.init(ng, is_parallel, nranks, periodic, nghost);
    if (rc) {
        delete l;
        return rc;
    }
    *out = l;
    return IPPLB_OK;
}

int ipplb_layout_set_boxes(ipplb_layout* l, const int* boxes) {
    if (!l || !boxes) return IPPLB_ERR_ARG;
    for (size_t r = 0; r < l->L.boxes.size(); ++r)
        for (int d = 0; d < 3; ++d) {
            l->L.boxes[r].lo[d] = boxes[r * 6 + d];
            l->L.boxes[r].hi[d] = boxes[r * 6 + 3 + d];
        }
    return IPPLB_OK;
}

int ipplb_layout_destroy(ipplb_layout* l) {
    delete l;
    return IPPLB_OK;
}

int ipplb_layout_nranks(const ipplb_layout* l) { return l ? (int)l->L.boxes.size() : 0; }

int ipplb_layout_boxes(const ipplb_layout* l, int* o) {
    if (!l || !o) return IPPLB_ERR_ARG;
    for (size_t r = 0; r < l->L.boxes.size(); ++r)
        for (int d = 0; d < 3; ++d) {
            o[r * 6 + d]     = l->L.boxes[r].lo[d];
            o[r * 6 + 3 + d] = l->L.boxes[r].hi[d];
        }
    return IPPLB_OK;
}

int ipplb_layout_neighbors(const ipplb_layout* l, int rank, int* out, int max_entries) {
    if (!l || rank < 0 || rank >= (int)l->L.boxes.size()) return -1;
    auto v = l->L.neighbors(rank);
    for (int i = 0; i < (int)v.size() && i < max_entries; ++i) {
        int* o = out + i * 14;
        o[0]   = v[i].comp;
        o[1]   = v[i].peer;
        for (int d = 0; d < 3; ++d) {
            o[2 + d]  = v[i].send_lo[d];
            o[5 + d]  = v[i].send_hi[d];
            o[8 + d]  = v[i].recv_lo[d];
            o[11 + d] = v[i].recv_hi[d];
        }
    }
    return (int)v.size();
}

int ipplb_layout_regions(const ipplb_layout* l, const double origin[3], const double h[3],
                         double* regions_out) {
    if (!l || !regions_out) return IPPLB_ERR_ARG;
    l->L.regions(origin, h, regions_out);
    return IPPLB_OK;
}

int ipplb_layout_mesh(const ipplb_layout* l, int rank, const double origin[3], const double h[3],
                      ipplb_mesh* mesh) {
    if (!l || !mesh || rank < 0 || rank >= (int)l->L.boxes.size()) return IPPLB_ERR_ARG;
    for (int d = 0; d < 3; ++d) {
        mesh->ng[d]     = l->L.ng[d];
        mesh->first[d]  = l->L.boxes[rank].lo[d];
        mesh->nl[d]     = l->L.boxes[rank].hi[d] - l->L.boxes[rank].lo[d] + 1;
        mesh->origin[d] = origin[d];
        mesh->h[d]      = h[d];
    }
    mesh->nghost = l->L.nghost;
    return IPPLB_OK;
}
}
