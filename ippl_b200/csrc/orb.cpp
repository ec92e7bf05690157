// orb.cpp -- host state machine of the orthogonal recursive bisection (no CUDA: linked into the product library and, for
// tests, into the CPU mock of the C-ABI).  The device part -- the plane sums of the weight field -- is ipplb_orb_plane_sums
// in diag.cu.
#include <numeric>
#include <vector>

#include "../../include/ippl_b200.h"

namespace ipplb {
void set_error(const char* fmt, ...);
}
#define ORB_REQUIRE(cond, msg)     \
    do {                           \
        if (!(cond)) {             \
            ipplb::set_error(msg); \
            return IPPLB_ERR_ARG;  \
        }                          \
    } while (0)

// ---- ORB host state machine: OrthogonalRecursiveBisection.hpp:14-105 -------------------------------------------
struct ipplb_orb {
    struct Dom {
        int lo[3], hi[3];
        int length(int d) const { return hi[d] - lo[d] + 1; }
    };
    std::vector<Dom> domains;
    std::vector<int> procs;
    unsigned it  = 0;
    int maxprocs = 0;
    int axis     = 0;
};

// findMedian, OrthogonalRecursiveBisection.hpp:185-216 (unsigned loop bounds like the reference's w.size() - 1)
static int orb_find_median(const std::vector<double>& w) {
    if (w.size() == 4) return 1;
    const double tot  = std::accumulate(w.begin(), w.end(), 0.0);
    const double half = 0.5 * tot;
    double curr       = 0.0;
    for (unsigned int i = 0; i < w.size() - 1; i++) {
        curr += w[i];
        if (curr >= half) {
            if (i == 0) return 1;
            const double previous = curr - w[i];
            if ((curr + previous) <= tot && curr != half) {
                if (i == w.size() - 2) return (int)(i - 1);
                return (int)i;
            }
            return (i > 1) ? (int)(i - 1) : 1;
        }
    }
    return (int)(w.size() - 3);
}

extern "C" {

// ---- ORB -----------------------------------------------------------------------------------------------------
int ipplb_orb_begin(ipplb_orb** out, const int ng[3], int nranks) {
    ORB_REQUIRE(out && ng && nranks >= 1, "orb_begin: bad arguments");
    auto* o = new ipplb_orb;
    ipplb_orb::Dom d;
    for (int k = 0; k < 3; ++k) {
        d.lo[k] = 0;
        d.hi[k] = ng[k] - 1;
    }
    o->domains  = {d};
    o->procs    = {nranks};
    o->it       = 0;
    o->maxprocs = nranks;
    *out        = o;
    return IPPLB_OK;
}

int ipplb_orb_next(ipplb_orb* o, int dom_lo[3], int dom_hi[3], int* axis, int* pending) {
    ORB_REQUIRE(o && dom_lo && dom_hi && axis && pending, "orb_next: bad arguments");
    *pending = o->maxprocs > 1 ? 1 : 0;
    if (!*pending) return IPPLB_OK;
    const ipplb_orb::Dom& d = o->domains[o->it];
    // findCutAxis, :107-116: std::max_element over the axis lengths (first maximum wins)
    int best = 0;
    for (int k = 1; k < 3; ++k)
        if (d.length(best) < d.length(k)) best = k;
    o->axis = best;
    *axis   = best;
    for (int k = 0; k < 3; ++k) {
        dom_lo[k] = d.lo[k];
        dom_hi[k] = d.hi[k];
    }
    return IPPLB_OK;
}

int ipplb_orb_cut(ipplb_orb* o, const double* reduced, int n) {
    ORB_REQUIRE(o && reduced && o->maxprocs > 1, "orb_cut: no cut pending");
    ipplb_orb::Dom d = o->domains[o->it];
    const int ax     = o->axis;
    ORB_REQUIRE(n == d.length(ax), "orb_cut: weight vector length differs from the domain length along the axis");
    ORB_REQUIRE(n >= 3, "orb_cut: domain too thin to cut (findMedian needs at least 3 planes)");
    const int median = orb_find_median(std::vector<double>(reduced, reduced + n));
    // cutDomain, :218-232: NDIndex::split at global index median + first -> left [first, mid], right [mid + 1, last]
    // (Index::split(l, r, mid), src/Index/Index.hpp:171-181)
    const int mid       = median + d.lo[ax];
    ipplb_orb::Dom left = d, right = d;
    left.hi[ax]  = mid;
    right.lo[ax] = mid + 1;
    o->domains[o->it] = left;
    o->domains.insert(o->domains.begin() + o->it + 1, right);
    const int temp  = o->procs[o->it];
    o->procs[o->it] = temp / 2;
    o->procs.insert(o->procs.begin() + o->it + 1, temp - o->procs[o->it]);
    o->maxprocs = 0;
    for (unsigned i = 0; i < o->procs.size(); ++i) {
        if (o->procs[i] > o->maxprocs) {
            o->maxprocs = o->procs[i];
            o->it       = i;
        }
    }
    return IPPLB_OK;
}

int ipplb_orb_finish(ipplb_orb* o, int* boxes_out, int* ok) {
    ORB_REQUIRE(o && boxes_out && ok, "orb_finish: bad arguments");
    *ok = o->maxprocs > 1 ? 0 : 1;
    for (size_t r = 0; r < o->domains.size(); ++r) {
        for (int k = 0; k < 3; ++k) {
            boxes_out[r * 6 + k]     = o->domains[r].lo[k];
            boxes_out[r * 6 + 3 + k] = o->domains[r].hi[k];
            if (o->domains[r].length(k) == 1) *ok = 0;  // :93-99
        }
    }
    delete o;
    return IPPLB_OK;
}

// drops a state machine that will not be driven to ipplb_orb_finish (error paths)
int ipplb_orb_destroy(ipplb_orb* o) {
    delete o;
    return IPPLB_OK;
}

}  // extern "C"
