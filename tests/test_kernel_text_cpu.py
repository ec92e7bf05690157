"""The SOURCE TEXT of the two kernel variants that were written without GPU access, executed on the CPU:
  * bins_move2_kernel (ippl_b200/csrc/bins.cu, ipplb_bins_build variant 2), cut out between its markers and run by a
    lock-step warp emulator (tests/emu/emu_bins_move2.cpp: the 32 lanes of a warp are 32 host threads that meet at every warp
    intrinsic): every particle lands once in the bucket of its tile with all six attributes, cursors end at the totals, what
    exceeds a bucket is written nowhere -- random, tile-ordered, single-tile and sub-warp inputs;
  * gather_point3_vec (ippl_b200/csrc/push.cuh, ipplb_ctx_set_gather_variant(2)): push.cuh itself compiled for the host
    (tests/emu/emu_gather_v2.cpp) and compared bit for bit with gather_point<3>, the kernel pinned to the oracle on the GPU;
    every load is checked for alignment and against the field's byte range.
This is a check of the kernels' logic, not of their behaviour on a GPU: tests/test_zz_variants_gpu.py is that."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")


def _build_and_run(tmp_path, src, flags, token):
    exe = str(tmp_path / "emu")
    cc = subprocess.run(["g++", "-O1", *flags, os.path.join(EMU, src), "-o", exe], capture_output=True, text=True, timeout=300)
    assert cc.returncode == 0, cc.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and token in out.stdout, out.stdout[-3000:] + out.stderr[-1000:]
    return out.stdout


def _kernel_text(tmp_path, mutate=None):
    src = open(os.path.join(ROOT, "ippl_b200", "csrc", "bins.cu")).read()
    m = re.search(r"// \[host-emulation begin: bins_move2_kernel\].*?\n(__global__.*?)// \[host-emulation end: bins_move2_kernel\]", src, re.S)
    assert m, "markers around bins_move2_kernel not found in bins.cu"
    text = m.group(1)
    if mutate:
        assert mutate[0] in text
        text = text.replace(*mutate)
    path = tmp_path / "bins_move2.inc"
    path.write_text(text)
    return ["-std=c++20", "-pthread", f'-DKERNEL_TEXT="{path}"']


def test_bins_move2_kernel_text_under_a_lockstep_warp_emulator(tmp_path):
    log = _build_and_run(tmp_path, "emu_bins_move2.cpp", _kernel_text(tmp_path), "EMU_BINS_MOVE2_OK")
    assert log.count(" ok") == 5 and "FAILED" not in log


def test_the_emulator_notices_a_wrong_lane_rank(tmp_path):
    """the same harness on a deliberately broken copy of the kernel text (lane rank counted inclusively): it must fail"""
    flags = _kernel_text(tmp_path, mutate=("(1u << lane) - 1u", "(2u << lane) - 1u"))
    exe = str(tmp_path / "emu")
    cc = subprocess.run(["g++", "-O1", *flags, os.path.join(EMU, "emu_bins_move2.cpp"), "-o", exe], capture_output=True, text=True, timeout=300)
    assert cc.returncode == 0, cc.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "EMU_BINS_MOVE2_FAILED" in out.stdout


def test_gather_variant2_device_functions_on_the_host(tmp_path):
    log = _build_and_run(tmp_path, "emu_gather_v2.cpp",
                         ["-std=c++17", "-ffp-contract=off", "-w", "-I/usr/local/cuda/include", f"-I{ROOT}"], "EMU_GATHER_V2_OK")
    assert log.count(": ok") == 5 and "FAILED" not in log


# ---- the slab-decomposed FFT solve: its two kernels' text inside the numpy executor of the plan --------------------------------
def _slab_emulation_library(tmp_path):
    import ctypes as C
    src = open(os.path.join(ROOT, "ippl_b200", "csrc", "fftdist.cu")).read()
    structs = re.search(r"// \[host-emulation begin: slab structs\].*?\n(.*?)// \[host-emulation end: slab structs\]", src, re.S)
    kernels = re.search(r"// \[host-emulation begin: slab kernels\]\n(.*?)// \[host-emulation end: slab kernels\]", src, re.S)
    assert structs and kernels, "markers not found in fftdist.cu"
    (tmp_path / "structs.inc").write_text(structs.group(1))
    (tmp_path / "kernels.inc").write_text(kernels.group(1))
    lib = str(tmp_path / "libemu_slab.so")
    cc = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", f'-DSTRUCT_TEXT="{tmp_path / "structs.inc"}"',
                         f'-DKERNEL_TEXT="{tmp_path / "kernels.inc"}"', os.path.join(EMU, "emu_slab.cpp"), "-o", lib],
                        capture_output=True, text=True, timeout=300)
    assert cc.returncode == 0, cc.stderr[-3000:]
    return C.CDLL(lib)


def test_slab_kernels_text_in_the_numpy_executor_of_the_plan(tmp_path):
    """tests/test_slabplan_cpu.py's execution of the plan for all ranks, with the device executor's own kernels (source text
    of slab_copy_kernel and kspace_slab_kernel, launch geometry of run_copies / transforms) doing the copies and the k-space
    step; numpy only transforms (cuFFT conventions: unnormalised, half spectrum in x).  <= 1e-12 of the whole-domain solve."""
    import ctypes as C

    import numpy as np

    import ippl_b200 as ib
    import oracle
    import test_slabplan_cpu as T
    emu = _slab_emulation_library(tmp_path)

    class CopyDev(C.Structure):
        _fields_ = [("src_off", C.c_long), ("dst_off", C.c_long), ("ss", C.c_long * 3), ("ds", C.c_long * 3), ("n", C.c_int * 3),
                    ("src_buf", C.c_int), ("dst_buf", C.c_int), ("elem", C.c_int)]
    assert emu.emu_sizeof_copydev() == C.sizeof(CopyDev)
    BUFS = ib.SlabPlan.BUFS

    def k_tables(ng, origin, h):   # poisson_k_tables (poisson.cu), FFTPeriodicPoissonSolver.hpp:66-70, 127-137
        out = []
        for d in range(3):
            N = ng[d]
            i = np.arange(N // 2 + 1 if d == 0 else N)
            Len = (origin[d] + N * h[d]) - origin[d]
            out.append(np.ascontiguousarray((i != N // 2) * 2 * np.pi / Len * (i - (i > N // 2) * N)))
        return out

    class Rank(T.Rank):
        cap = 4096

        def copies(self, phase, which):
            rows = self.plan.rows(phase, which)
            if not rows:
                return
            arr = (CopyDev * len(rows))()
            for a, c in zip(arr, rows):
                a.src_off, a.dst_off, a.elem = c["src_off"], c["dst_off"], c["elem"]
                a.src_buf, a.dst_buf = BUFS.index(c["src"]), BUFS.index(c["dst"])
                for k in range(3):
                    a.ss[k], a.ds[k], a.n[k] = c["ss"][k], c["ds"][k], c["n"][k]
            ptrs = (C.c_void_p * len(BUFS))(*[self.buf[b].ctypes.data for b in BUFS])
            biggest = max(c["n"][0] * c["n"][1] * c["n"][2] for c in rows)
            emu.emu_slab_copies(arr, len(rows), ptrs, C.c_long(biggest), C.c_long(self.cap))

    def step1(rk, origin, h):
        p = rk.plan
        nx, ny, nz = p.ng
        nxh, nyl = p.nxh, p.ye - p.ys
        if not nyl:
            return
        SZ = nz * nyl * nxh
        sz = rk.buf["specz"]
        sz[:SZ] = np.fft.fft(sz[:SZ].reshape(nz, nyl, nxh), axis=0).ravel()               # cufftExecZ2Z forward, in place
        kx, ky, kz = k_tables(p.ng, origin, h)
        at = lambda k: C.c_void_p(sz.ctypes.data + 16 * SZ * k)                              # noqa: E731
        emu.emu_kspace(nxh, nyl, nz, p.ys, C.c_double(1.0 / (nx * ny * nz)), kx.ctypes.data_as(C.c_void_p),
                       ky.ctypes.data_as(C.c_void_p), kz.ctypes.data_as(C.c_void_p), at(0), at(1), at(2), at(3), C.c_long(rk.cap))
        for c in range(3):                                                                 # cufftExecZ2Z inverse: unnormalised
            sz[(1 + c) * SZ:(2 + c) * SZ] = (np.fft.ifft(sz[(1 + c) * SZ:(2 + c) * SZ].reshape(nz, nyl, nxh), axis=0) * nz).ravel()

    for ng, world, kind, cap in (((16, 12, 10), 2, "default", 4096), ((24, 16, 16), 8, "orb", 3), ((9, 7, 11), 5, "default", 1),
                                 ((10, 6, 5), 8, "default", 4096)):
        origin, h = (0.0, 0.5, -1.0), (0.3, 0.25, 0.4)
        layout = ib.Layout(ng, world)
        if kind == "orb":
            layout.set_boxes(T.orb_like(ng, world))
        rng = np.random.default_rng(7)
        rho_g = rng.normal(size=(ng[2], ng[1], ng[0]))
        rho_g -= rho_g.mean()
        N, nxh = ng[0] * ng[1] * ng[2], ng[0] // 2 + 1
        rhat = np.fft.rfftn(rho_g) / N
        want = np.stack([np.fft.irfftn(rhat * np.broadcast_to(M, rho_g.shape)[:, :, :nxh], s=rho_g.shape, axes=(0, 1, 2)) * N
                         for M in oracle.poisson_kspace_multipliers(ng, origin, h)], axis=-1)
        Rank.cap = cap          # small caps force the kernels' grid-stride loops
        ranks = [Rank(layout, r, origin, h) for r in range(world)]
        g = ranks[0].plan.nghost
        for rk in ranks:
            f = rk.first
            rk.buf["rho"].reshape(rk.ext[2], rk.ext[1], rk.ext[0])[g:-g, g:-g, g:-g] = \
                rho_g[f[2]:f[2] + rk.nl[2], f[1]:f[1] + rk.nl[1], f[0]:f[0] + rk.nl[0]]
        for phase in range(4):
            for rk in ranks:
                rk.copies(phase, 0)
            T.exchange(ranks, phase)
            for rk in ranks:
                rk.copies(phase, 2)
            if phase < 3:
                for rk in ranks:
                    if phase == 1:
                        step1(rk, origin, h)
                    else:
                        T.transforms(rk, phase, origin, h)
        scale = np.abs(want).max()
        for r, rk in enumerate(ranks):
            f = rk.first
            ef = rk.buf["ef"].reshape(rk.ext[2], rk.ext[1], rk.ext[0], 3)
            got = ef[g:-g, g:-g, g:-g]
            ref = want[f[2]:f[2] + rk.nl[2], f[1]:f[1] + rk.nl[1], f[0]:f[0] + rk.nl[0]]
            assert np.isfinite(got).all(), f"{ng} x{world} rank {r}: E interior not fully written"
            assert np.max(np.abs(got - ref)) <= 1e-12 * scale, (ng, world, r, np.max(np.abs(got - ref)) / scale)
            halo = ef.copy()
            halo[g:-g, g:-g, g:-g] = np.nan
            assert np.isnan(halo).all(), f"rank {r}: the solve wrote into E's ghost layers"
            rho = rk.buf["rho"].reshape(rk.ext[2], rk.ext[1], rk.ext[0])
            assert np.array_equal(rho[g:-g, g:-g, g:-g], got[..., 2])
        for rk in ranks:
            rk.plan.close()
        layout.close()


def test_kokkos_shim_reduction_kernels_under_a_lockstep_block_emulator(tmp_path):
    """for_kernel / reduce1_kernel / reduce2_kernel / block_join / atomic_join of include/ippl/KokkosShim.cuh -- the device
    code behind the unchanged reference drivers' Kokkos::parallel_for / parallel_reduce, which the shim's own host-emulation
    mode replaces by host loops -- cut out of the header and run as 256 host threads per block (tests/emu/emu_shim_reduce.cpp)"""
    src = open(os.path.join(ROOT, "include", "ippl", "KokkosShim.cuh")).read()
    red = re.search(r"// \[host-emulation begin: shim reducers\].*?\n(.*?)// \[host-emulation end: shim reducers\]", src, re.S)
    ker = re.search(r"// \[host-emulation begin: shim kernels\]\n(.*?)// \[host-emulation end: shim kernels\]", src, re.S)
    assert red and ker, "markers not found in KokkosShim.cuh"
    (tmp_path / "reducers.inc").write_text(red.group(1))
    (tmp_path / "kernels.inc").write_text(ker.group(1))
    _build_and_run(tmp_path, "emu_shim_reduce.cpp", ["-std=c++20", "-pthread", f'-DREDUCER_TEXT="{tmp_path / "reducers.inc"}"',
                                                     f'-DKERNEL_TEXT="{tmp_path / "kernels.inc"}"'], "EMU_SHIM_REDUCE_OK")


def test_arrivals_kernel_text_under_a_lockstep_block_emulator(tmp_path):
    """arrivals_p2p_kernel (ippl_b200/csrc/comm.cu), the multi-GPU step's kernel that drops the inbox records into their
    buckets: GPU-verified in round 2, then given a check that refuses records outside the rank's box without GPU access.
    Its text, with the product's own cic.cuh / bins.h, as 256 host threads per block (tests/emu/emu_arrivals.cpp): valid
    records land once (bucket or tail), records on the box's upper faces are accepted, zeroed / non-finite / foreign
    records are refused and flagged, a full tail is flagged, status words and deposited charge add up.  (This run found
    that a NaN coordinate converts to cell 0 and passed the first version of the check on ranks whose box starts there.)"""
    src = open(os.path.join(ROOT, "ippl_b200", "csrc", "comm.cu")).read()
    m = re.search(r"// \[host-emulation begin: arrivals_p2p_kernel\].*?\n(struct ArriveArgs.*?)// \[host-emulation end: arrivals_p2p_kernel\]", src, re.S)
    assert m, "markers not found in comm.cu"
    (tmp_path / "arrivals.inc").write_text(m.group(1))
    log = _build_and_run(tmp_path, "emu_arrivals.cpp", ["-std=c++20", "-w", "-pthread", "-ffp-contract=off", "-I/usr/local/cuda/include",
                                                         f"-I{ROOT}", f'-DKERNEL_TEXT="{tmp_path / "arrivals.inc"}"'], "EMU_ARRIVALS_OK")
    assert log.count(": ok") == 3 and "FAILED" not in log
