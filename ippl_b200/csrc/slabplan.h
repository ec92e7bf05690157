// slabplan.h -- host-side plan of the slab-decomposed periodic FFT solve (no CUDA in here: the tables are unit-tested on the
// CPU by executing them with numpy, tests/test_slabplan_cpu.py, and executed on the device by fftdist.cu).
//
// What it restates: the reference hands the field to heFFTe, which redistributes the FieldLayout boxes into the FFT's own
// pencil / slab decompositions and back (src/FFT/FFT.hpp:118-193 "heffte::fft3d_r2c" with the in / out boxes of the layout,
// src/PoissonSolvers/FFTPeriodicPoissonSolver.hpp:53-169).  Here the same data movement is four all-to-all phases between
// three decompositions of the periodic domain (x fastest everywhere):
//     boxes (FieldLayout)  --P0-->  z-slabs [nzl][ny][nx]    2-D real-to-complex transforms of whole (x, y) planes
//     z-slabs              --P1-->  y-slabs [nz][nyl][nxh]   1-D transforms along z, k-space multipliers, 1-D inverses
//     y-slabs              --P2-->  z-slabs (3 components)   2-D complex-to-real inverses
//     z-slabs              --P3-->  boxes   (3 components)   E interior (AoS-3) [+ rho <- last component, as the reference]
// Every phase is: local sub-box copies into one contiguous message per peer, one message exchange, local sub-box copies out.
#pragma once
#include <vector>

#include "layout.h"

namespace ipplb {

enum SlabBuf { SB_RHO = 0, SB_EF = 1, SB_REAL = 2, SB_SPEC2D = 3, SB_SPECZ = 4, SB_SEND = 5, SB_RECV = 6, SB_COUNT = 7 };

// one strided 3-D sub-box copy; offsets and strides in ELEMENTS of `elem` doubles (1: real, 2: complex)
struct SlabCopy {
    int src_buf, dst_buf;
    long src_off, dst_off;
    long ss[3], ds[3];
    int n[3];
    int elem;
};

// one message: `count` doubles from SEND + soff to the peer, `rcount` doubles from the peer into RECV + roff
struct SlabMsg {
    int peer;
    long soff, scount, roff, rcount;
};

struct SlabPhase {
    std::vector<SlabCopy> pre, post;
    std::vector<SlabMsg> msgs;
};

struct SlabPlan {
    int nranks = 1, me = 0;
    int ng[3] = {0, 0, 0}, nxh = 0, nghost = 1;
    int zs = 0, ze = 0, ys = 0, ye = 0;  // my z-slab [zs, ze) and y-slab [ys, ye)
    long size[SB_COUNT] = {0, 0, 0, 0, 0, 0, 0};  // doubles per buffer (RHO / EF: the ghosted field sizes, for checks)
    SlabPhase phase[4];
    int build(const Layout& L, int rank);
};

// even split of n planes over p ranks: rank r owns [lo, hi)
void slab_range(int n, int p, int r, int& lo, int& hi);

}  // namespace ipplb

struct ipplb_slabplan {
    ipplb::SlabPlan P;
};
