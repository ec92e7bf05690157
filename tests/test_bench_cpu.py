"""bench.py's host side without a GPU: the reference arm (oracle port on the host cores) runs and reports the same
metric / unit / config object as our arm would; workloads of BASELINE.json's configs; the traffic extract bench.py reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


@pytest.mark.parametrize("gpus", [1, 8])
def test_reference_arm_line(gpus):
    import bench
    from ippl_b200 import app
    lines = _run(["--impl", "reference", "--gpus", str(gpus), "--log2-particles", "16", "--steps", "1", "--warmup", "1"],
                 env={"OMP_NUM_THREADS": "1"})   # what torchrun exports: the arm must not inherit it
    d = json.loads(lines[-1])
    assert d["impl"] == "reference" and d["unit"] == "particles/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same metric and the same config object as our arm prints for this N
    w = app.workload("landau", gpus, 16)
    assert d["metric"] == w["metric"] and d["config"] == bench.config_dict(w, gpus, True) and d["n_gpus"] == gpus


def test_reference_arm_other_ranks_exit_quietly():
    assert _run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1"}) == []


def test_workloads_follow_baseline_configs():
    from ippl_b200 import app
    w = app.workload("landau", 1)
    assert w["ng"] == (128, 128, 128) and w["n_local"] == 1 << 27 and abs(w["dt"] - 0.5 * w["h"][0]) < 1e-15
    assert app.workload("landau", 8)["ng"] == (256, 256, 256) and app.workload("landau", 4)["ng"] == (256, 256, 128)
    w = app.workload("bumpontail", 8)
    assert w["ng"] == (512, 512, 512) and w["n_local"] == 1 << 29 and len(w["vel"]) == 2
    w = app.workload("penning", 8)
    assert w["ng"] == (256, 256, 256) and w["n_local"] * 8 == 1 << 30 and w["push"] == "penning"
    assert abs(w["dt"] - 0.5 * 20.0 / 2048) < 1e-18


def test_traffic_extract_is_what_bench_reports():
    import bench
    t, src = bench.profiled_traffic("fused_step3_kernel")
    raw = json.load(open(os.path.join(ROOT, "profiles", "r2_fused_traffic.json")))
    assert t == raw["dram_bytes_read"] + raw["dram_bytes_write"] and 1.2e10 < t < 1.4e10 and src["kernel"] == raw["kernel"]
    assert bench.profiled_traffic("some_other_kernel") == (None, None)
