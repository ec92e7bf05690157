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


# ---- the secondary-measurement leg (bench.extras_leg): process plumbing, no GPU ---------------------------------------
_CHILD = r'''
import json, os, sys, datetime
mode = sys.argv[1]
rank = int(os.environ.get("RANK", "0"))
if mode == "job":      # a multi-rank sub-job: its own rendezvous on the shifted port, rank 0 prints the line
    import torch, torch.distributed as dist
    assert "TORCHELASTIC_USE_AGENT_STORE" not in os.environ and os.environ["IPPLB_PG_TIMEOUT_S"] == "90"
    dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=60))
    t = torch.tensor([rank + 1.0])
    dist.all_reduce(t)
    if rank == 0:
        print("noise before the line")
        print(json.dumps({"metric": "m", "value": float(t[0]), "unit": "particles/s", "n_gpus": dist.get_world_size(),
                          "ms_per_step": 1.0, "roofline": {"kernel": "k", "frac": 0.5, "ms_per_launch": 0.9, "junk": 1},
                          "config": {"workload": "w"}, "clocks": {"dropped": True}}))
    dist.destroy_process_group()
elif mode == "late_crash":      # the measurement is complete and printed, then the tear-down fails
    print(json.dumps({"metric": "m", "value": 7.0}), flush=True)
    sys.exit(5)
elif mode == "crash":
    sys.stderr.write("boom\n")
    sys.exit(3)
elif mode == "hang":
    import time
    time.sleep(60)
elif mode == "variant":
    print(json.dumps({"part": sys.argv[2], "rows": [{"v2_same_bits": True}]}))
elif mode == "micro":
    print(json.dumps({"rows": [{"ppc": 8, "order": "random", "n": 10, "gather_gpps": 20.0 + rank, "scatter_atomic_gpps": 30.0 - rank}],
                      "bins_build": []}))
'''

_PARENT = r'''
import json, os, sys, datetime
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
import bench
child = sys.argv[2]
world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
dist.init_process_group("gloo", timeout=datetime.timedelta(seconds=60))
jobs = [("good", [sys.executable, child, "job"], 60), ("bad", [sys.executable, child, "crash"], 60),
        ("stuck", [sys.executable, child, "hang"], 2)]
ex = bench.extras_leg(None, world, rank, rank, dist, jobs=jobs, micro_cmd=[sys.executable, child, "micro"], micro_limit=30,
                      variant_cmds=[("va", [sys.executable, child, "variant", "a"], 30), ("vb", [sys.executable, child, "crash"], 30)])
if rank == 0:
    print("RESULT " + json.dumps(ex))
dist.destroy_process_group()
'''


def test_extras_leg_sub_jobs_under_torchrun(tmp_path):
    """two torchrun workers (gloo): each starts the sub-jobs; a multi-rank sub-job gets its own rendezvous, a crashing one
    and a hanging one are reported as errors, the per-rank microbench lines are merged through files"""
    (tmp_path / "child.py").write_text(_CHILD)
    (tmp_path / "parent.py").write_text(_PARENT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29871", str(tmp_path / "parent.py"), ROOT, str(tmp_path / "child.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    ex = json.loads(line[len("RESULT "):])
    assert ex["good"]["value"] == 3.0 and ex["good"]["n_gpus"] == 2 and ex["good"]["config"] == {"workload": "w"}
    assert ex["good"]["roofline"] == {"kernel": "k", "frac": 0.5, "achieved": None, "peak": None, "unit": None, "ms_per_launch": 0.9}
    assert "clocks" not in ex["good"]
    assert "rc 3" in ex["bad"]["error"] and "boom" in ex["bad"]["error"]
    assert "no result within 2" in ex["stuck"]["error"] and "killed" in ex["stuck"]["error"]
    rep = ex["micro"]["replicas"]
    assert rep["n"] == 2 and rep["failed"] == 0
    assert rep["min_over_ranks_gpps"] == {"ppc8_random_gather_gpps": 20.0, "ppc8_random_scatter_atomic_gpps": 29.0}
    assert ex["micro"]["rows"][0]["gather_gpps"] == 20.0
    assert ex["va"] == {"part": "a", "rows": [{"v2_same_bits": True}]} and "rc 3" in ex["vb"]["error"]


def test_extras_leg_single_process(tmp_path):
    import bench
    (tmp_path / "child.py").write_text(_CHILD)
    child = str(tmp_path / "child.py")
    ex = bench.extras_leg(None, 1, 0, 0, None, jobs=[("bad", [sys.executable, child, "crash"], 30), ("late", [sys.executable, child, "late_crash"], 30)],
                          micro_cmd=[sys.executable, child, "micro"], micro_limit=30,
                          variant_cmds=[("va", [sys.executable, child, "variant", "a"], 30)])
    assert "rc 3" in ex["bad"]["error"] and ex["micro"]["rows"][0]["ppc"] == 8 and "replicas" not in ex["micro"]
    assert ex["late"]["value"] == 7.0 and ex["late"]["exit_code_after_the_line"] == 5      # a complete measurement is kept
    assert ex["va"]["part"] == "a"


def test_extras_leg_budget(tmp_path):
    """all sub-runs together stay inside one budget: what does not fit is not started"""
    import bench
    (tmp_path / "child.py").write_text(_CHILD)
    child = str(tmp_path / "child.py")
    ex = bench.extras_leg(None, 1, 0, 0, None, jobs=[("stuck", [sys.executable, child, "hang"], 60)],
                          micro_cmd=[sys.executable, child, "micro"], micro_limit=30, variant_cmds=[], budget_s=3.0, min_left_s=1.0)
    assert "no result within" in ex["stuck"]["error"] and "not started" in ex["micro"]["error"] and ex["seconds"] < 10


def test_main_line_on_stand_ins(monkeypatch, capsys):
    """bench.main() from argument parsing to the printed line, with the GPU side (context, mini-app, CUDA events) replaced by
    stand-ins: the line carries what the contract asks for, `config` is the workload object, the findings sit under `run`,
    the secondary measurements are attached after everything else and a failure in them cannot lose the line"""
    import types

    import torch

    import bench
    import ippl_b200 as ib
    from ippl_b200 import app

    class Ev:
        def __init__(self, enable_timing=True):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 40.0     # ms for the whole timed region

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    for name in ("synchronize", "set_device", "empty_cache"):
        monkeypatch.setattr(torch.cuda, name, lambda *a, **k: None)

    class Ctx:
        device = "cpu"
        launches = 0

        def __init__(self, local):
            pass

        def close(self):
            pass

    class Bins:
        def set_timing(self, on):
            pass

        def kernel_ms(self):
            return [3.9, 4.1]

    class Mini:
        def __init__(self, ctx, w, rank, world, mode=2, fft="replicated", dist=None):
            self.w, self.bins, self.solve_ms, self.n_mine = w, Bins(), 0.2, w["n_local"]
            self.mesh = types.SimpleNamespace(nl=w["ng"], cells=(w["ng"][0] + 2) ** 3)
            self.steps = 0

        def initialise(self):
            pass

        def step(self, first=False):
            self.steps += 1

        def status(self):
            return {"n_local": self.n_mine, "n_tail": 5, "n_exit": 0, "flags": 0}

        def phase_ms(self):
            return {"rho_zero": 0.01}

        def extra_config(self):
            return {"found": 1}

        def e2e(self, steps, barrier):
            return {"ms_per_step": 5.0, "steps": steps, "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 2, "what": "w"}

        def e2e_streamed(self, barrier):
            return {"value": 1.0}

        def close(self):
            self.closed = True

    monkeypatch.setattr(ib, "Context", Ctx)
    monkeypatch.setattr(app, "MiniApp", Mini)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1.0, "sm_max_mhz": 1.0, "reasons": []})
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)

    def line(extras):
        monkeypatch.setattr(bench, "extras_leg", extras)
        monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "10", "--warmup", "3", "--no-cpu"])
        bench.main()
        return json.loads(capsys.readouterr().out.strip().splitlines()[-1])

    d = line(lambda *a, **k: {"stub": True})
    w = app.workload("landau", 1)
    assert d["config"] == bench.config_dict(w, 1, True) and d["run"] == {"tail_fraction": 5 / w["n_local"], "found": 1}
    assert d["ms_per_step"] == 4.0 and d["value"] == w["n_local"] / 4.0e-3 and d["n_gpus"] == 1 and d["steps"] == 10 and d["warmup"] == 3
    assert d["roofline"]["ms_per_launch"] == 4.0 and d["roofline"]["ms_per_launch"] <= d["ms_per_step"] and 0 < d["roofline"]["frac"] < 1
    assert d["roofline"]["traffic"] and d["e2e"]["value"] == w["n_local"] / 5.0e-3 and d["e2e"]["unit"] == "particles/s"
    assert d["extras"] == {"stub": True} and d["dtype"] == "f64" and d["higher_is_better"] is True and d["vs_baseline"] is None

    def broken(*a, **k):
        raise RuntimeError("secondary measurement fell over")
    d = line(broken)
    assert "fell over" in d["extras"]["error"] and d["ms_per_step"] == 4.0
    monkeypatch.setenv("IPPLB_BENCH_EXTRAS", "0")
    assert "extras" not in line(broken)


def test_first_solve_restores_rho_and_checks_the_slab_solve(monkeypatch):
    """MiniApp._first_solve on stand-ins: a solve leaves the last gradient component in rho, so the timed solve must start
    from the restored rho; with --fft slab the slab result is compared with the replicated one on the same rho and a
    disagreeing slab field is reported and not used for the steps"""
    import types

    import torch

    import ippl_b200 as ib
    from ippl_b200 import app

    class Ev:
        def __init__(self, enable_timing=True):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 1.5

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    seen = []

    class Solver:
        def __init__(self, ctx, mesh, layout=None, origin=None, h=None, slab=False):
            self.slab = slab

        def solve(self, rho, ef):
            seen.append(("slab" if self.slab else "replicated", float(rho.sum())))
            ef.copy_(torch.cat([rho, 2 * rho, 3 * rho]) * (Solver.slab_factor if self.slab else 1.0))
            rho.copy_(ef[-len(rho):])          # the solve clobbers rho with the last component

        def close(self):
            pass

    monkeypatch.setattr(ib, "Poisson", Solver)
    ctx = types.SimpleNamespace(device="cpu", scatter=lambda *a, **k: None, halo_exchange=lambda *a, **k: None,
                                halo_accumulate_periodic=lambda *a, **k: None, field_density=lambda *a, **k: None)
    w = app.workload("landau", 2, 10)

    def make(fft, world):
        m = app.MiniApp(ctx, w, 0, world, fft=fft)
        m.parts = types.SimpleNamespace(arr={k: torch.zeros(4) for k in "xyz"})
        m.n_mine, m.n_total, m.q, m.mesh, m.layout = 4, 8, -1.0, None, None
        m.rho, m.ef = torch.arange(1.0, 6.0, dtype=torch.float64), torch.zeros(15, dtype=torch.float64)
        return m

    Solver.slab_factor = 1.0
    m = make("replicated", 1)
    m._first_solve()
    assert [s[1] for s in seen] == [15.0, 15.0] and m.solve_ms == 1.5          # both solves saw the same rho
    assert torch.equal(m.ef[:5], torch.arange(1.0, 6.0, dtype=torch.float64)) and not hasattr(m, "solve_check")
    seen.clear()
    m = make("slab", 2)
    m._first_solve()
    assert [s for s in seen] == [("slab", 15.0), ("slab", 15.0), ("replicated", 15.0), ("replicated", 15.0)]
    assert m.solve_check["rel_l2_slab_vs_replicated"] == 0.0 and m.solve_check["field_used_by_the_steps"] == "slab solve"
    assert m.extra_config()["field_solve_check"] == m.solve_check
    Solver.slab_factor = 1.001                                                  # a slab solve that is off by 0.1 %
    m = make("slab", 2)
    m._first_solve()
    assert abs(m.solve_check["rel_l2_slab_vs_replicated"] - 1e-3) < 1e-9 and "rejected" in m.solve_check["field_used_by_the_steps"]
    assert torch.equal(m.ef[:5], torch.arange(1.0, 6.0, dtype=torch.float64))  # the steps get the replicated field


def test_main_line_on_two_ranks_with_stand_ins(tmp_path):
    """bench.main() under torchrun with 2 workers (tests/bench_main_stand_in.py: gloo for NCCL, stand-ins for the GPU side):
    one line from rank 0, whole-job value from the slowest rank's time, the parity object, the end-to-end maximum, the
    secondary measurements merged over the ranks"""
    from ippl_b200 import app
    (tmp_path / "child.py").write_text(_CHILD)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29921", os.path.join(ROOT, "tests", "bench_main_stand_in.py"), str(tmp_path / "child.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    w = app.workload("landau", 2)
    assert d["n_gpus"] == 2 and d["ms_per_step"] == 25.0 / 6 and d["value"] == 2 * w["n_local"] / (25.0 / 6 * 1e-3)
    assert d["parity"] == {"counts_exact": True, "ranks": 2} and d["config"]["workload"].startswith("alpine LandauDamping 256x128x128")
    assert d["run"]["field_solve"] == "replicated" and d["roofline"]["traffic"] is None
    assert d["e2e"]["ms_per_step"] == 6.0 and d["e2e"]["value"] == 2 * w["n_local"] / 6.0e-3 and "e2e_particles_streamed" not in d
    ex = d["extras"]
    assert ex["job"]["value"] == 3.0 and ex["job"]["n_gpus"] == 2 and ex["micro"]["replicas"]["n"] == 2 and ex["micro"]["replicas"]["failed"] == 0
