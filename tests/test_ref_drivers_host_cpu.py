"""The reference's UNCHANGED alpine drivers (/root/reference/demos/alpine/{LandauDamping,PenningTrap,BumponTailInstability}.cpp
with their own *Manager.h, AlpineManager.h, FieldContainer / FieldSolver / LoadBalancer / ParticleContainer headers) run on
the CPU: compiled by the host compiler against include/ippl/compat + include/ippl/KokkosShim.cuh in host-emulation mode and
linked to oracle/mock -- a CPU stand-in for the C-ABI built on the oracle (TEST INFRASTRUCTURE; the product has no CPU path).

What this pins, without a GPU: the host logic between the drivers and the C-ABI -- the managers' call sequences through the
facade, the functor-shaped samplers, the reductions and the CSV dumps -- against the reference's own known-answer file
(demos/alpine/validation/FieldLandau_valid_result.csv, tolerance 0.4 as in validation/CMakeLists.txt:23-26) and the physical
anchors tests/test_y_facade.py uses.  What it cannot pin: the CUDA launch layer of the shim and the kernels behind the
C-ABI (tests/test_zz_reference_drivers.py runs the nvcc-built drivers on a GPU)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK = os.path.join(ROOT, "oracle", "_build", "mock")


@pytest.fixture(scope="module")
def drivers():
    if not os.path.isdir("/root/reference/demos/alpine"):
        pytest.skip("needs the reference tree")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "mock"], stderr=subprocess.DEVNULL)
    return MOCK


def _run(drivers, tmp_path, exe, grid, np_, nt, csv):
    d = tmp_path / exe
    (d / "data").mkdir(parents=True)
    cmd = [os.path.join(drivers, f"ref_{exe}_host"), str(grid), str(grid), str(grid), str(np_), str(nt), "FFT", "0.01", "LeapFrog",
           "--overallocate", "2.0", "--info", "0"]
    out = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    text = (d / "data" / csv).read_text().splitlines()
    assert all(line.strip() for line in text), "blank lines in the CSV (Inform::flush must not emit a message)"
    assert os.path.exists(d / "timing.dat")
    return np.loadtxt(d / "data" / csv, skiprows=1), out.stdout


def test_unchanged_landau_driver_reproduces_the_reference_known_answer(drivers, tmp_path):
    golden = np.loadtxt(os.path.join(ROOT, "tests", "golden", "FieldLandau_valid_result.csv"), skiprows=1)
    got, log = _run(drivers, tmp_path, "LandauDamping", 16, 10000000, 25, "FieldLandau_1_manager.csv")
    assert got.shape == golden.shape == (26, 3)
    assert np.allclose(got[:, 0], golden[:, 0], atol=1e-12)
    assert np.max(np.abs(got[:, 1:] - golden[:, 1:])) <= 0.4
    assert got[-1, 1] < 0.7 * got[0, 1]
    for timer in ("particlesCreation", "pushVelocity", "pushPosition", "update", "solve", "dumpData"):   # the drivers' own timers
        assert timer in log


def test_unchanged_bumpontail_driver(drivers, tmp_path):
    got, _ = _run(drivers, tmp_path, "BumponTailInstability", 16, 2000000, 6, "FieldBumponTail_1_manager.csv")
    assert got.shape == (7, 3) and np.isfinite(got).all()
    k, delta = 0.21, 0.01
    theory = 0.5 * (delta / k) ** 2 * (2 * np.pi / k) ** 3
    assert 0.8 * theory <= got[0, 1] <= 1.6 * theory, (got[0, 1], theory)


def test_unchanged_penningtrap_driver(drivers, tmp_path):
    n = 1000000
    got, _ = _run(drivers, tmp_path, "PenningTrap", 16, n, 6, "ParticleField_1_manager.csv")
    cols = [1, 2, 3, 5, 6, 7]   # column 4 is rhoNorm_m, which the reference never assigns (AlpineManager.h:71)
    assert got.shape == (7, 8) and np.isfinite(got[:, cols]).all() and (got[:, cols] > 0).all()
    assert abs(got[0, 2] / (1.5 * n) - 1.0) <= 5e-3           # v ~ N(0, 1) per component
    h3 = (20.0 / 16) ** 3
    assert np.allclose(got[:, 1], 0.5 * h3 * (got[:, 5] ** 2 + got[:, 6] ** 2 + got[:, 7] ** 2), rtol=1e-8)
    assert np.allclose(got[:, 3], got[:, 1] + got[:, 2], rtol=1e-9)
    # total energy is conserved to a few 1e-3 over the first steps of the trap
    assert abs(got[-1, 3] / got[0, 3] - 1.0) <= 2e-2


def test_reference_lambdas_harness_on_the_mock(drivers):
    """demos/ref_lambdas.cu in host-emulation mode: the eight lambda bodies cut out of the reference drivers against the mock's
    (= the oracle's) implementation of the same C-ABI calls; Kick1 / Kick2 bit for bit.  Pins the harness the GPU run uses."""
    out = subprocess.run([os.path.join(drivers, "ref_lambdas_host")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count(": ok") >= 7 and "all reference lambdas agree" in out.stdout
    assert "Kick1: 0 of" in out.stdout and "Kick2: 0 of" in out.stdout


def _run_any(drivers, tmp_path, exe, tag, grid, np_, nt, csv):
    d = tmp_path / tag
    (d / "data").mkdir(parents=True)
    cmd = [os.path.join(drivers, exe), str(grid), str(grid), str(grid), str(np_), str(nt), "FFT", "0.01", "LeapFrog", "--overallocate", "2.0",
           "--info", "0"]
    out = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return np.loadtxt(d / "data" / csv, skiprows=1)


@pytest.mark.parametrize("app,csv,cols", [("LandauDamping", "FieldLandau_1_manager.csv", [0, 1, 2]),
                                           ("PenningTrap", "ParticleField_1_manager.csv", [0, 1, 2, 3, 5, 6, 7]),
                                           ("BumponTailInstability", "FieldBumponTail_1_manager.csv", [0, 1, 2])])
def test_restated_drivers_equal_the_unchanged_reference_drivers(drivers, tmp_path, app, csv, cols):
    """demos/<app>.cpp (the repo's restatement on demos/Alpine.h, C-ABI samplers and reductions) and the reference's unchanged
    <app>.cpp (its own managers, functor-shaped samplers and Kokkos lambdas through the compat layer) on the same backend
    write the same CSV: same particles, same call sequence, same arithmetic."""
    ref = _run_any(drivers, tmp_path, f"ref_{app}_host", "ref", 16, 500000, 5, csv)
    own = _run_any(drivers, tmp_path, f"demo_{app}_host", "own", 16, 500000, 5, csv)
    assert ref.shape == own.shape
    assert np.max(np.abs(ref[:, cols] - own[:, cols]) / np.maximum(np.abs(own[:, cols]), 1e-300)) <= 1e-12


def test_api_check_demo_on_the_mock(drivers):
    out = subprocess.run([os.path.join(drivers, "demo_api_check_host")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "scatter(policy, hash)" in out.stdout and "ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("app,csv,cols,fused_steps", [("LandauDamping", "FieldLandau_1_manager.csv", [0, 1, 2], 8),
                                                       ("BumponTailInstability", "FieldBumponTail_1_manager.csv", [0, 1, 2], 8),
                                                       ("PenningTrap", "ParticleField_1_manager.csv", [0, 1, 2, 3, 5, 6, 7], 0)])
def test_lazy_fusion_runs_the_unchanged_drivers_on_the_fused_step(drivers, tmp_path, app, csv, cols, fused_steps):
    """IPPL_B200_FUSE=1: the facade records the unchanged driver's gather / kick / kick / drift / update() and executes them
    together with the scatter as ONE ipplb_bins_step (here: the mock's emulation of it; the host logic is what is under test).
    LandauDamping and BumponTail fuse every step and never materialise; PenningTrap's kicks are driver lambdas over getView(),
    so every step materialises and nothing fuses.  Either way the CSV is the one of the plain run."""
    def go(fuse):
        d = tmp_path / f"fuse{fuse}"
        (d / "data").mkdir(parents=True)
        cmd = [os.path.join(drivers, f"ref_{app}_host"), "16", "16", "16", "300000", "8", "FFT", "0.01", "LeapFrog", "--overallocate", "2.0",
               "--info", "0"]
        out = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=600, env=dict(os.environ, IPPL_B200_FUSE=str(fuse)))
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        return np.loadtxt(d / "data" / csv, skiprows=1)[:, cols], out.stdout
    plain, log0 = go(0)
    fused, log1 = go(1)
    assert "ippl_b200 fusion" not in log0
    assert f"ippl_b200 fusion: {fused_steps} fused steps, {0 if fused_steps else 9} materialisations" in log1, log1[-500:]
    assert np.max(np.abs(plain - fused) / np.maximum(np.abs(plain), 1e-300)) <= 1e-12


def test_fusion_engine_materialises_correctly_when_the_driver_peeks(drivers):
    """demos/fusion_check.cpp on the mock: peeks at the particles after the closing kick, between drift and update and right
    after a fused scatter; same particles and field-energy history as the plain run, 8 of 12 steps fused, 6 materialisations."""
    out = subprocess.run([os.path.join(drivers, "demo_fusion_check_host")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "fusion_check: ok" in out.stdout, out.stdout + out.stderr
    assert "8 fused steps, 6 materialisations" in out.stdout


def _run_ranks(drivers, tmp_path, exe, tag, world, grid, np_, nt, csv, threads=4, fuse=None):
    """`world` processes of one driver on the mock: ippl::initialize reads RANK / WORLD_SIZE like under torchrun; the mock's
    communicator is a directory (IPPLB_MOCK_DIR)"""
    d = tmp_path / tag
    (d / "data").mkdir(parents=True)
    (d / "comm").mkdir()
    cmd = [os.path.join(drivers, exe), str(grid), str(grid), str(grid), str(np_), str(nt), "FFT", "0.01", "LeapFrog", "--overallocate", "2.0",
           "--info", "0"]
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), IPPLB_NCCL_ID_FILE=str(d / "comm" / "id"),
                   IPPLB_MOCK_DIR=str(d / "comm"), OMP_NUM_THREADS=str(threads))
        if fuse is not None:
            env["IPPL_B200_FUSE"] = str(fuse)
        procs.append(subprocess.Popen(cmd, cwd=d, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        out, _ = p.communicate(timeout=900)
        logs.append(out)
    assert all(p.returncode == 0 for p in procs), "\n".join(l[-1500:] for l in logs)
    return np.loadtxt(d / "data" / csv, skiprows=1), logs[0]


def test_unchanged_landau_driver_on_two_ranks_reproduces_the_known_answer(drivers, tmp_path):
    """The reference generated its known-answer file with 2 ranks (demos/alpine/validation/CMakeLists.txt): the unchanged driver
    on 2 ranks of the mock -- its own LoadBalancer doing the first ORB repartition, the facade's two-phase migrate, halo
    exchanges, the replicated solve, the reductions of the dumps and of IpplTimings -- stays within the reference's tolerance."""
    golden = np.loadtxt(os.path.join(ROOT, "tests", "golden", "FieldLandau_valid_result.csv"), skiprows=1)
    got, log = _run_ranks(drivers, tmp_path, "ref_LandauDamping_host", "landau2", 2, 16, 10000000, 25, "FieldLandau_2_manager.csv", threads=8)
    assert got.shape == golden.shape
    assert np.max(np.abs(got[:, 1:] - golden[:, 1:])) <= 0.4
    assert "Timing results for 2 rank(s)" in log and "loadBalance" in log and "Could not repartition" not in log


def test_unchanged_penning_driver_on_four_ranks_equals_the_restated_one(drivers, tmp_path):
    """PenningTrap on 4 ranks: the blob makes the ORB repartition non-trivial (first repartition on the analytic density, one
    more during the run).  The reference's unchanged driver (its LoadBalancer.hpp, ORB through the compat layer) and the
    repo's restated driver write the same CSV."""
    kw = dict(world=4, grid=16, np_=400000, nt=6, csv="ParticleField_4_manager.csv")
    ref, log = _run_ranks(drivers, tmp_path, "ref_PenningTrap_host", "pt_ref", **kw)
    own, _ = _run_ranks(drivers, tmp_path, "demo_PenningTrap_host", "pt_own", **kw)
    cols = [0, 1, 2, 3, 5, 6, 7]
    assert ref.shape == own.shape == (7, 8)
    assert np.max(np.abs(ref[:, cols] - own[:, cols]) / np.maximum(np.abs(own[:, cols]), 1e-300)) <= 1e-12
    assert abs(ref[0, 2] / (1.5 * 400000) - 1.0) <= 1e-2
    assert "Could not repartition" not in log


@pytest.mark.parametrize("app,csv,world,expect", [("LandauDamping", "FieldLandau_2_manager.csv", 2, "8 fused steps, 0 materialisations"),
                                                   ("BumponTailInstability", "FieldBumponTail_4_manager.csv", 4, "fused steps")])
def test_lazy_fusion_on_several_ranks(drivers, tmp_path, app, csv, world, expect):
    """IPPL_B200_FUSE=1 on several ranks: the recorded update() includes the migration, so the fused step is followed by
    ipplb_bins_migrate and the container's count follows; a mid-run ORB repartition (BumponTail on 4 ranks) materialises once
    and fusion resumes (the charge is still one value on every rank).  Same CSV as the plain run."""
    kw = dict(world=world, grid=16, np_=400000, nt=8, csv=csv)
    plain, _ = _run_ranks(drivers, tmp_path, f"ref_{app}_host", "plain", fuse=0, **kw)
    fused, log = _run_ranks(drivers, tmp_path, f"ref_{app}_host", "fused", fuse=1, **kw)
    assert expect in log, log[-600:]
    import re
    m = re.search(r"ippl_b200 fusion: (\d+) fused steps, (\d+) materialisations", log)
    assert m and int(m.group(1)) >= 6 and int(m.group(2)) <= 2, log[-600:]
    assert np.max(np.abs(plain - fused) / np.maximum(np.abs(plain), 1e-300)) <= 1e-12
