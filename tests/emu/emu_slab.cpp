// emu_slab.cpp -- TEST INFRASTRUCTURE: the source text of the two kernels of the slab-decomposed FFT solve
// (ippl_b200/csrc/fftdist.cu: slab_copy_kernel, kspace_slab_kernel, and the CopyDev / BufTable structs they take; cut out
// between their markers by tests/test_kernel_text_cpu.py and passed as STRUCT_TEXT / KERNEL_TEXT), compiled for the host and
// run thread by thread (neither kernel synchronises or uses warp intrinsics: a sequential sweep over the launch grid is an
// execution).  Built as a shared library; the test plugs it into the numpy executor of the plan (tests/test_slabplan_cpu.py)
// in place of numpy's copies and multipliers, with the launch geometry of fftdist.cu's run_copies / transforms.
#include <cmath>
#include <cstddef>

struct Dim3 { unsigned x = 1, y = 1, z = 1; };
static Dim3 threadIdx, blockIdx, blockDim, gridDim;
#define __global__
#define __launch_bounds__(n)
struct double2 { double x, y; };
typedef double2 cufftDoubleComplex;
static inline double2 make_cuDoubleComplex(double a, double b) { return double2{a, b}; }
enum { SB_COUNT = 7 };

#include STRUCT_TEXT
#include KERNEL_TEXT

extern "C" {

int emu_sizeof_copydev() { return (int)sizeof(CopyDev); }

// run_copies of fftdist.cu: grid (gx, n) with gx = ceil(biggest / 256) capped, 256 threads
void emu_slab_copies(const CopyDev* list, int n, double* const* bufs, long biggest, long cap) {
    BufTable B;
    for (int b = 0; b < SB_COUNT; ++b) B.p[b] = bufs[b];
    long gx = (biggest + 255) / 256;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    blockDim.x = 256; gridDim.x = (unsigned)gx; gridDim.y = (unsigned)n;
    for (unsigned by = 0; by < gridDim.y; ++by)
        for (unsigned bx = 0; bx < gridDim.x; ++bx)
            for (unsigned t = 0; t < 256; ++t) {
                blockIdx.x = bx; blockIdx.y = by; threadIdx.x = t;
                slab_copy_kernel(list, B);
            }
}

// transforms(step 1) of fftdist.cu: grid g = min(ceil(SZ / 256), cap), 256 threads
void emu_kspace(int nxh, int nyl, int nz, int ys, double inv_n, const double* kx, const double* ky, const double* kz,
                const double* rh, double* g0, double* g1, double* g2, long cap) {
    const long SZ = (long)nxh * nyl * nz;
    long g = (SZ + 255) / 256;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    blockDim.x = 256; gridDim.x = (unsigned)g; gridDim.y = 1;
    for (unsigned bx = 0; bx < gridDim.x; ++bx)
        for (unsigned t = 0; t < 256; ++t) {
            blockIdx.x = bx; blockIdx.y = 0; threadIdx.x = t;
            kspace_slab_kernel(nxh, nyl, nz, ys, inv_n, kx, ky, kz, (const cufftDoubleComplex*)rh, (cufftDoubleComplex*)g0,
                               (cufftDoubleComplex*)g1, (cufftDoubleComplex*)g2);
        }
}

}  // extern "C"
