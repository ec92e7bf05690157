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
