"""First executions, each in its OWN pytest process: the two kernel variants written after this round's GPU budget was spent
-- ipplb_bins_build variant 2 (tests/variant_build_check.py) and the gather kernels' variant 2
(tests/variant_gather_check.py) -- and the multi-rank parity tests on a layout whose z cut is off the middle
(tests/loop_orb_z_check.py: the geometry of the PenningTrap run).  Why separate processes: a CUDA
fault in code that has never run poisons the CUDA context of the process it happens in, and must not reach the rest of
the suite or the other checks.  xfail(strict=False) until each has passed once on a GPU (a pass shows
as XPASS); both variants are opt-in, the defaults are the verified kernels."""
import os
import subprocess
import sys

import pytest

from util import first_run_timeout

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="kernel variant not yet executed on a GPU (written without GPU access)")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("check", ["variant_build_check.py", "variant_gather_check.py", "loop_orb_z_check.py"])
def test_kernel_variant_in_its_own_process(check):
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", check), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"],
                         cwd=ROOT, capture_output=True, text=True, timeout=first_run_timeout(150))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1500:]
    assert " passed" in out.stdout and "failed" not in out.stdout, out.stdout[-1500:]
