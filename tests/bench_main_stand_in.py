"""Runner for tests/test_bench_cpu.py::test_main_line_on_two_ranks_with_stand_ins (not collected by name): bench.main() under
torchrun on the CPU -- the real control flow of the multi-rank arm (process group, the oracle-checked parity step's place,
max-over-ranks reductions, the particle-conservation check, the secondary-measurement leg with its file rendezvous, the
printed line), with gloo in place of NCCL and stand-ins for everything that needs a GPU."""
import datetime
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
child = sys.argv[1]
sys.argv = ["bench.py", "--gpus", os.environ["WORLD_SIZE"], "--steps", "6", "--warmup", "3", "--no-cpu"]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
import ippl_b200 as ib  # noqa: E402
from ippl_b200 import app  # noqa: E402


class Ev:
    def __init__(self, enable_timing=True):
        pass

    def record(self):
        pass

    def elapsed_time(self, other):
        return 24.0 + int(os.environ["RANK"])     # rank 1 is the slower one: the line must carry the maximum


torch.cuda.Event = Ev
for name in ("synchronize", "set_device", "empty_cache"):
    setattr(torch.cuda, name, lambda *a, **k: None)
real_init = dist.init_process_group
dist.init_process_group = lambda backend, device_id=None, timeout=None: real_init("gloo", timeout=timeout or datetime.timedelta(seconds=60))


class Ctx:
    device = "cpu"
    launches = 0

    def __init__(self, local):
        pass

    def comm_init(self, rank, world, uid):
        assert uid == b"stand-in id"

    def close(self):
        pass


class Bins:
    def set_timing(self, on):
        pass

    def kernel_ms(self):
        return [3.9, 4.1]


class Mini:
    def __init__(self, ctx, w, rank, world, mode=2, fft="replicated", dist=None):
        assert dist is not None and world == 2
        self.w, self.bins, self.solve_ms, self.n_mine = w, Bins(), 0.2, w["n_local"] + (5 if rank == 0 else -5)
        self.mesh = types.SimpleNamespace(nl=(w["ng"][0] // 2, w["ng"][1], w["ng"][2]), cells=10)

    def initialise(self):
        pass

    def step(self, first=False):
        pass

    def status(self):
        return {"n_local": self.n_mine, "n_tail": 3, "n_exit": 0, "flags": 8}     # 8 = informational bit only

    def phase_ms(self):
        return {"halo_fill_E": 0.05}

    def extra_config(self):
        return {"field_solve": "replicated", "migration": "stand-in"}

    def e2e(self, steps, barrier):
        return {"ms_per_step": 5.0 + int(os.environ["RANK"]), "steps": steps, "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 2, "what": "w"}

    def close(self):
        pass


ib.Context, ib.nccl_unique_id, app.MiniApp = Ctx, (lambda: b"stand-in id"), Mini
sys.modules["mgpu_parity"] = types.SimpleNamespace(multi_rank_step=lambda ctx, d, rank, world: {"counts_exact": True, "ranks": world})
bench.ClockSampler.start = lambda self: None
bench.ClockSampler.stop = lambda self: {"sm_mhz": 1.0, "sm_max_mhz": 1.0, "reasons": []}
real_extras = bench.extras_leg
bench.extras_leg = lambda args, world, rank, local, d: real_extras(args, world, rank, local, d, jobs=[("job", [sys.executable, child, "job"], 60)],
                                                                   micro_cmd=[sys.executable, child, "micro"], micro_limit=30, variant_cmds=[])
bench.main()
