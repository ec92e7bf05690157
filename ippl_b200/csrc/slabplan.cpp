// slabplan.cpp -- host-side plan of the slab-decomposed FFT solve (see slabplan.h).  No CUDA.
#include "slabplan.h"

#include <algorithm>
#include <cstring>

#include "../../include/ippl_b200.h"

namespace ipplb {

void slab_range(int n, int p, int r, int& lo, int& hi) {
    const int base = n / p, rem = n % p;
    lo = r * base + std::min(r, rem);
    hi = lo + base + (r < rem ? 1 : 0);
}

namespace {

struct Span {
    int lo, hi;  // [lo, hi)
    int len() const { return hi > lo ? hi - lo : 0; }
};
Span cut(int a_lo, int a_hi, int b_lo, int b_hi) { return Span{std::max(a_lo, b_lo), std::min(a_hi, b_hi)}; }

SlabCopy copy(int sb, long so, long s0, long s1, long s2, int db, long d_o, long d0, long d1, long d2, int n0, int n1, int n2,
              int elem) {
    SlabCopy c;
    c.src_buf = sb; c.dst_buf = db; c.src_off = so; c.dst_off = d_o;
    c.ss[0] = s0; c.ss[1] = s1; c.ss[2] = s2;
    c.ds[0] = d0; c.ds[1] = d1; c.ds[2] = d2;
    c.n[0] = n0; c.n[1] = n1; c.n[2] = n2;
    c.elem = elem;
    return c;
}

}  // namespace

int SlabPlan::build(const Layout& L, int rank) {
    nranks = (int)L.boxes.size();
    me     = rank;
    if (rank < 0 || rank >= nranks) return IPPLB_ERR_ARG;
    for (int d = 0; d < 3; ++d) ng[d] = L.ng[d];
    nghost = L.nghost;
    nxh    = ng[0] / 2 + 1;
    const int nx = ng[0], ny = ng[1], nz = ng[2], g = nghost, P = nranks;
    slab_range(nz, P, me, zs, ze);
    slab_range(ny, P, me, ys, ye);
    const int nzl = ze - zs, nyl = ye - ys;
    const IBox& B = L.boxes[me];
    const int nl[3] = {B.hi[0] - B.lo[0] + 1, B.hi[1] - B.lo[1] + 1, B.hi[2] - B.lo[2] + 1};
    const long ex = nl[0] + 2 * g, ey = nl[1] + 2 * g, ez = nl[2] + 2 * g;
    const long SR = (long)nzl * ny * nx;    // reals per component in the z-slab
    const long S2 = (long)nzl * ny * nxh;   // complex per component in the z-slab
    const long SZ = (long)nz * nyl * nxh;   // complex per component in the y-slab
    size[SB_RHO]    = ex * ey * ez;
    size[SB_EF]     = 3 * ex * ey * ez;
    size[SB_REAL]   = 3 * SR;
    size[SB_SPEC2D] = 2 * 3 * S2;
    size[SB_SPECZ]  = 2 * 4 * SZ;
    long need_send = 0, need_recv = 0;
    for (auto& ph : phase) {
        ph.pre.clear();
        ph.post.clear();
        ph.msgs.clear();
    }

    // ---- P0: boxes -> z-slabs (rho, real) ----------------------------------------------------------------------------
    {
        SlabPhase& ph = phase[0];
        long so = 0, ro = 0;
        for (int d = 0; d < P; ++d) {
            int dz0, dz1;
            slab_range(nz, P, d, dz0, dz1);
            const Span s = cut(B.lo[2], B.hi[2] + 1, dz0, dz1);           // my box's planes inside d's slab
            const long scount = (long)nl[0] * nl[1] * s.len();
            const IBox& Bd = L.boxes[d];
            const int dl[2] = {Bd.hi[0] - Bd.lo[0] + 1, Bd.hi[1] - Bd.lo[1] + 1};
            const Span r = cut(Bd.lo[2], Bd.hi[2] + 1, zs, ze);          // d's box planes inside my slab
            const long rcount = (long)dl[0] * dl[1] * r.len();
            if (scount)
                ph.pre.push_back(copy(SB_RHO, g + ex * (g + ey * (long)(s.lo - B.lo[2] + g)), 1, ex, ex * ey, SB_SEND, so, 1, nl[0],
                                      (long)nl[0] * nl[1], nl[0], nl[1], s.len(), 1));
            if (rcount)
                ph.post.push_back(copy(SB_RECV, ro, 1, dl[0], (long)dl[0] * dl[1], SB_REAL,
                                       Bd.lo[0] + (long)nx * (Bd.lo[1] + (long)ny * (r.lo - zs)), 1, nx, (long)nx * ny, dl[0], dl[1],
                                       r.len(), 1));
            if (scount || rcount) ph.msgs.push_back(SlabMsg{d, so, scount, ro, rcount});
            so += scount;
            ro += rcount;
        }
        need_send = std::max(need_send, so);
        need_recv = std::max(need_recv, ro);
    }
    // ---- P1: z-slabs -> y-slabs (rho_hat after the 2-D transforms, complex) ---------------------------------------
    {
        SlabPhase& ph = phase[1];
        long so = 0, ro = 0;  // in complex elements
        for (int d = 0; d < P; ++d) {
            int dy0, dy1, dz0, dz1;
            slab_range(ny, P, d, dy0, dy1);
            slab_range(nz, P, d, dz0, dz1);
            const int dyl = dy1 - dy0, dzl = dz1 - dz0;
            const long scount = (long)nxh * dyl * nzl;   // my planes, d's rows
            const long rcount = (long)nxh * nyl * dzl;   // d's planes, my rows
            if (scount)
                ph.pre.push_back(copy(SB_SPEC2D, (long)nxh * dy0, 1, nxh, (long)nxh * ny, SB_SEND, so, 1, nxh, (long)nxh * dyl, nxh, dyl,
                                      nzl, 2));
            if (rcount)
                ph.post.push_back(copy(SB_RECV, ro, 1, nxh, (long)nxh * nyl, SB_SPECZ, (long)nxh * nyl * dz0, 1, nxh, (long)nxh * nyl, nxh,
                                       nyl, dzl, 2));
            if (scount || rcount) ph.msgs.push_back(SlabMsg{d, 2 * so, 2 * scount, 2 * ro, 2 * rcount});
            so += scount;
            ro += rcount;
        }
        need_send = std::max(need_send, 2 * so);
        need_recv = std::max(need_recv, 2 * ro);
    }
    // ---- P2: y-slabs -> z-slabs (three gradient spectra after the inverse z transforms, complex) -------------------
    {
        SlabPhase& ph = phase[2];
        long so = 0, ro = 0;  // complex elements
        for (int d = 0; d < P; ++d) {
            int dy0, dy1, dz0, dz1;
            slab_range(ny, P, d, dy0, dy1);
            slab_range(nz, P, d, dz0, dz1);
            const int dyl = dy1 - dy0, dzl = dz1 - dz0;
            const long sc = (long)nxh * nyl * dzl;   // per component: d's planes, my rows
            const long rc = (long)nxh * dyl * nzl;   // per component: my planes, d's rows
            for (int c = 0; c < 3; ++c) {
                if (sc)
                    ph.pre.push_back(copy(SB_SPECZ, (1 + c) * SZ + (long)nxh * nyl * dz0, 1, nxh, (long)nxh * nyl, SB_SEND, so + c * sc, 1,
                                          nxh, (long)nxh * nyl, nxh, nyl, dzl, 2));
                if (rc)
                    ph.post.push_back(copy(SB_RECV, ro + c * rc, 1, nxh, (long)nxh * dyl, SB_SPEC2D, c * S2 + (long)nxh * dy0, 1, nxh,
                                           (long)nxh * ny, nxh, dyl, nzl, 2));
            }
            if (sc || rc) ph.msgs.push_back(SlabMsg{d, 2 * so, 2 * 3 * sc, 2 * ro, 2 * 3 * rc});
            so += 3 * sc;
            ro += 3 * rc;
        }
        need_send = std::max(need_send, 2 * so);
        need_recv = std::max(need_recv, 2 * ro);
    }
    // ---- P3: z-slabs -> boxes (three real gradient components -> E interior, rho <- last component) ----------------
    {
        SlabPhase& ph = phase[3];
        long so = 0, ro = 0;
        for (int d = 0; d < P; ++d) {
            int dz0, dz1;
            slab_range(nz, P, d, dz0, dz1);
            const IBox& Bd = L.boxes[d];
            const int dl[2] = {Bd.hi[0] - Bd.lo[0] + 1, Bd.hi[1] - Bd.lo[1] + 1};
            const Span s = cut(Bd.lo[2], Bd.hi[2] + 1, zs, ze);          // d's box planes that I hold
            const long sc = (long)dl[0] * dl[1] * s.len();
            const Span r = cut(B.lo[2], B.hi[2] + 1, dz0, dz1);          // my box's planes that d holds
            const long rc = (long)nl[0] * nl[1] * r.len();
            for (int c = 0; c < 3; ++c) {
                if (sc)
                    ph.pre.push_back(copy(SB_REAL, c * SR + Bd.lo[0] + (long)nx * (Bd.lo[1] + (long)ny * (s.lo - zs)), 1, nx, (long)nx * ny,
                                          SB_SEND, so + c * sc, 1, dl[0], (long)dl[0] * dl[1], dl[0], dl[1], s.len(), 1));
                if (rc) {
                    const long cell = g + ex * (g + ey * (long)(r.lo - B.lo[2] + g));
                    ph.post.push_back(copy(SB_RECV, ro + c * rc, 1, nl[0], (long)nl[0] * nl[1], SB_EF, 3 * cell + c, 3, 3 * ex, 3 * ex * ey,
                                           nl[0], nl[1], r.len(), 1));
                    if (c == 2)   // the reference's inverse transform lands in rho's storage (FFTPeriodicPoissonSolver.hpp:153)
                        ph.post.push_back(copy(SB_RECV, ro + c * rc, 1, nl[0], (long)nl[0] * nl[1], SB_RHO, cell, 1, ex, ex * ey, nl[0],
                                               nl[1], r.len(), 1));
                }
            }
            if (sc || rc) ph.msgs.push_back(SlabMsg{d, so, 3 * sc, ro, 3 * rc});
            so += 3 * sc;
            ro += 3 * rc;
        }
        need_send = std::max(need_send, so);
        need_recv = std::max(need_recv, ro);
    }
    size[SB_SEND] = (need_send + 1) & ~1L;   // whole complex elements
    size[SB_RECV] = (need_recv + 1) & ~1L;
    return IPPLB_OK;
}

}  // namespace ipplb

using namespace ipplb;

extern "C" {

int ipplb_slabplan_create(const ipplb_layout* layout, int rank, ipplb_slabplan** out) {
    if (!layout || !out) return IPPLB_ERR_ARG;
    ipplb_slabplan* p = new ipplb_slabplan();
    const int rc      = p->P.build(layout->L, rank);
    if (rc != IPPLB_OK) {
        delete p;
        return rc;
    }
    *out = p;
    return IPPLB_OK;
}

int ipplb_slabplan_destroy(ipplb_slabplan* p) {
    delete p;
    return IPPLB_OK;
}

int ipplb_slabplan_info(const ipplb_slabplan* p, long info[16]) {
    if (!p || !info) return IPPLB_ERR_ARG;
    const SlabPlan& P = p->P;
    const long v[16]  = {P.nranks, P.me, P.ng[0], P.ng[1], P.ng[2], P.nxh, P.zs, P.ze, P.ys, P.ye, P.size[SB_REAL], P.size[SB_SPEC2D],
                         P.size[SB_SPECZ], P.size[SB_SEND], P.size[SB_RECV], P.nghost};
    std::memcpy(info, v, sizeof(v));
    return IPPLB_OK;
}

int ipplb_slabplan_rows(const ipplb_slabplan* p, int phase, int which, long* rows, int max_rows, int* nrows) {
    if (!p || phase < 0 || phase > 3 || which < 0 || which > 2 || !nrows) return IPPLB_ERR_ARG;
    const SlabPhase& ph = p->P.phase[phase];
    if (which == 1) {
        *nrows = (int)ph.msgs.size();
        for (int i = 0; rows && i < *nrows && i < max_rows; ++i) {
            const SlabMsg& m = ph.msgs[i];
            const long v[5]  = {m.peer, m.soff, m.scount, m.roff, m.rcount};
            std::memcpy(rows + 16 * i, v, sizeof(v));
        }
        return IPPLB_OK;
    }
    const std::vector<SlabCopy>& cs = which == 0 ? ph.pre : ph.post;
    *nrows = (int)cs.size();
    for (int i = 0; rows && i < *nrows && i < max_rows; ++i) {
        const SlabCopy& c = cs[i];
        const long v[14]  = {c.src_buf, c.dst_buf, c.src_off, c.dst_off, c.ss[0], c.ss[1], c.ss[2], c.ds[0], c.ds[1], c.ds[2],
                             c.n[0],    c.n[1],    c.n[2],    c.elem};
        std::memcpy(rows + 16 * i, v, sizeof(v));
    }
    return IPPLB_OK;
}

}  // extern "C"
