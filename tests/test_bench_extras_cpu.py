"""scripts/bench_extras.py (the microbench process bench.py starts after its headline numbers) driven on the CPU against
a stand-in for the ippl_b200 bindings: checks the script's own logic -- the sweep, the variant comparison, the JSON line --
not any kernel (those are the `-m gpu` tests)."""
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_bindings(torch, calls):
    ib = types.ModuleType("ippl_b200")

    class Mesh:
        @staticmethod
        def make(ng, origin, h):
            m = Mesh()
            m.ng, m.cells = ng, (ng[0] + 2) * (ng[1] + 2) * (ng[2] + 2)
            return m

    class Particles:
        NAMES = ("x", "y", "z", "px", "py", "pz")

        def __init__(self, cap, dev, q=0.0):
            self.arr = {k: torch.zeros(cap, dtype=torch.float64) for k in self.NAMES}
            self.n, self.q_scalar = 0, q

    class Context:
        def __init__(self, dev):
            self.device = "cpu"

        def field(self, mesh, ncomp=1):
            return torch.zeros(mesh.cells * ncomp, dtype=torch.float64)

        def offsets_buffer(self, mesh):
            return torch.zeros(8, dtype=torch.int32)

        def sort_by_cell(self, mesh, src, dst, off):
            for k in Particles.NAMES:
                dst.arr[k].copy_(src.arr[k])
            dst.n = src.n

        def field_fill(self, f, v):
            f.fill_(v)

        def close(self):
            calls.append("close")

        def set_gather_variant(self, v):
            calls.append(f"gather_variant{v}")

        def gather(self, mesh, x, y, z, ef, out, add=False):
            for o in out:
                o.zero_()
            calls.append("gather")

        def __getattr__(self, name):   # halo_fill_periodic, scatter, gather, gather_push, scatter_sorted
            def op(*a, **k):
                calls.append(name)
            return op

    class Bins:
        def __init__(self, ctx, mesh, cap):
            self.variant, self.n = 1, 0

        def set_build_variant(self, v):
            self.variant = v
            calls.append(f"variant{v}")

        def build(self, src, dst):
            order = torch.arange(src.n) if self.variant == 1 else torch.arange(src.n - 1, -1, -1)   # same set, other order
            for k in Particles.NAMES:
                dst.arr[k][:src.n] = src.arr[k][:src.n][order]
            dst.n = self.n = src.n

        def status(self):
            return self.n, 0, 0, 0

        def compact(self, cur, out):
            for k in Particles.NAMES:
                out.arr[k][:self.n] = cur.arr[k][:self.n]
            return self.n

        def step(self, push, cur, nxt, ef, rho):
            rho += 1.0
            cur.arr, nxt.arr = nxt.arr, cur.arr
            cur.arr, nxt.arr = nxt.arr, cur.arr   # (the stand-in leaves the particles where they are)

        def tables(self):
            return tuple(np.arange(4, dtype=np.int32) for _ in range(3))

        def close(self):
            pass

    ib.Mesh, ib.Particles, ib.Context, ib.Bins = Mesh, Particles, Context, Bins
    ib.leapfrog_push = lambda dt: object()
    return ib


def _run_part(monkeypatch, capsys, part):
    import torch
    calls = []
    monkeypatch.setitem(sys.modules, "ippl_b200", _fake_bindings(torch, calls))

    class Ev:
        def __init__(self, enable_timing=True):
            pass

        def record(self):
            pass

        def elapsed_time(self, other):
            return 2.0

    monkeypatch.setattr(torch.cuda, "Event", Ev)
    for name in ("synchronize", "set_device", "empty_cache"):
        monkeypatch.setattr(torch.cuda, name, lambda *a, **k: None)
    spec = importlib.util.spec_from_file_location("bench_extras", os.path.join(ROOT, "scripts", "bench_extras.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(sys, "argv", ["bench_extras.py", "--part", part, "--device", "0", "--grid", "6", "--ppc", "1", "3", "--reps", "3"])
    mod.main()
    d = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert d["part"] == part
    return d, calls


def test_bench_extras_verified_part(monkeypatch, capsys):
    d, calls = _run_part(monkeypatch, capsys, "verified")
    rows = d["rows"]
    assert [(r["ppc"], r["order"]) for r in rows] == [(1, "sorted"), (1, "random"), (1, "bucketed"), (3, "sorted"), (3, "random"), (3, "bucketed")]
    assert rows[0]["scatter_sorted_gpps"] == 216 / 2.0 / 1e6 and "scatter_sorted_gpps" not in rows[1]
    assert rows[2]["fused_step_gpps"] > 0 and rows[5]["bins_build_ms"] == 2.0 and rows[5]["n"] == 3 * 216
    assert {"gather", "gather_push", "scatter", "scatter_sorted", "close"} <= set(calls)
    assert not any(c.startswith(("variant", "gather_variant")) for c in calls)     # no variant is touched by this part


def test_bench_extras_gather_variants_part(monkeypatch, capsys):
    d, calls = _run_part(monkeypatch, capsys, "gather_variants")
    rows = d["rows"]
    assert [(r["ppc"], r["order"]) for r in rows] == [(1, "sorted"), (1, "random"), (3, "sorted"), (3, "random")]
    for r in rows:
        assert r["v2_same_bits"] is True and r["gather_speedup"] == 1.0 and r["gather_push_v2_gpps"] == r["gather_push_v1_gpps"] > 0
    seq = [c for c in calls if c.startswith("gather_variant")]
    assert seq[:3] == ["gather_variant1", "gather_variant2", "gather_variant1"] and seq[-1] == "gather_variant1"


def test_bench_extras_build_variants_part(monkeypatch, capsys):
    d, calls = _run_part(monkeypatch, capsys, "build_variants")
    rows = d["rows"]
    assert [r["ppc"] for r in rows] == [1, 3]
    b = rows[1]
    assert b["n"] == 3 * 216 and b["build_v1_ms"] == b["build_v2_ms"] == 2.0 and b["build_speedup"] == 1.0
    assert b["v2_same_tables"] and b["v2_same_particles"] and b["v2_step_rho_rel_l2"] == 0.0
    assert [c for c in calls if c.startswith("variant")] == ["variant1", "variant2"] * 2
