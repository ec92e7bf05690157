"""The slab-decomposed FFT Poisson solve on a GPU (in-process ranks; tests/slab_fft_check.py in its own process).

STATUS: the plan behind it is verified on the CPU (tests/test_slabplan_cpu.py executes the same tables with numpy for all
ranks); the device executor (ippl_b200/csrc/fftdist.cu: copy-list kernel, batched cuFFT plans, k-space kernel) was written
after this round's GPU budget was spent and has not run yet.  xfail(strict=False) until it has passed once; the file
sorts last and the check runs in its own process."""
import os
import subprocess
import sys

import pytest

from util import first_run_timeout

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="device executor of the slab FFT not yet executed on a GPU")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_fft_solve_on_in_process_ranks():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "slab_fft_check.py")], capture_output=True, text=True,
                         timeout=first_run_timeout(150))
    assert out.returncode == 0 and "SLAB_FFT_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
