#!/usr/bin/env python
"""bench.py -- particles/s per PIC step (scatter + push + gather) of the alpine mini-apps' hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config landau|bumpontail|penning]   our arm (CUDA, sm_100a, C-ABI)
  python bench.py --impl reference [...]                                                     the reference algorithm on the host cores

Default workload (BASELINE.json configs[1], the one the metric is quoted on): LandauDamping, 128^3 cells and 2^27 fp64
particles PER GPU (weak scaling: the mesh doubles along x, y, z as N = 2, 4, 8; FieldLayout decomposition), CIC, LeapFrog.
A "step" is one pass of the owned path:
  N = 1: rho = 0 -> ONE fused kernel (gather E + kick + kick + drift + periodic BC + re-bucketing + charge deposit; the
         periodic fillHalo(E) / accumulateHalo(rho) are folded into it: ghost nodes alias the opposite interior layer);
  N > 1: fillHalo(E) over NVLink -> rho = 0 -> the fused kernel with the ownership test -> migration of the leavers ->
         accumulateHalo(rho).
--config bumpontail = BASELINE.json configs[3] (512^3 mesh held fixed, 2^29 particles per GPU), --config penning =
configs[2] (256^3 mesh, 2^30 particles over the GPUs, ORB repartition before the timed region, Boris-type push).
The FFT field solve is a non-owned stage: it is run once before the timed region to produce a self-consistent E and is
timed separately (`solve_ms`).  Inputs (6.4 GB of particles per GPU) are far larger than the 126 MB L2: no flush needed.
The mini-app driver itself (initial condition, ORB, step, parity check, end-to-end variants) is ippl_b200/app.py.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "particles/s"
BYTES_PER_PARTICLE_STEP = 120  # gather+push 96 B + scatter 24 B (uniform scalar charge; 32 with a q array)
BYTES_PER_CELL_STEP = 40       # rho zero + rho write + E read


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, read from the committed summary of
    an `ncu --set full` capture of this workload (profiles/r2_fused_traffic.json, written by scripts/ncu_traffic.py from
    the .ncu-rep); None when there is no capture for this kernel."""
    path = os.path.join(ROOT, "profiles", "r2_fused_traffic.json")
    if not os.path.exists(path):
        return None, None
    t = json.load(open(path))
    if kernel_substr not in t.get("kernel", ""):
        return None, None
    return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"]), {k: t.get(k) for k in ("kernel", "commit", "command", "file")}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def config_dict(w, world, bins):
    ng, n_local = w["ng"], w["n_local"]
    n_total = n_local * world
    return {"workload": f"{w['label']} {ng[0]}x{ng[1]}x{ng[2]} mesh, {n_local} particles/GPU fp64, CIC, "
                        f"{'LeapFrog' if w['push'] == 'leapfrog' else 'Boris-type kicks (PenningTrap LeapFrogStep)'}",
            "particles_total": n_total, "ppc": n_total / (ng[0] * ng[1] * ng[2]),
            "decomposition": f"FieldLayout {world} rank(s)" + (", ORB repartition before the timed region"
                                                              if w["name"] == "penning" and world > 1 else ""),
            "l2": f"inputs ({48 * n_local / 1e9:.1f} GB/GPU) exceed L2; no flush needed",
            "sort": "none: single-pass fused step on per-tile buckets" if bins else "counting sort by cell every step",
            "solve": "excluded (non-owned cuFFT stage)", "charge": "uniform scalar q (24 B/particle scatter)"}


# ---- reference arm: the reference algorithm on the host cores --------------------------------------------------------
def cpu_run(w, n_sample, steps, warmup, threads=None):
    """The reference algorithm (oracle port: OpenMP, atomic scatter like Kokkos-OpenMP, three-kernel leapfrog, per-
    dimension periodic wrap) on a sample of the workload: same mesh, same distributions, `n_sample` particles."""
    import oracle
    if threads:
        oracle.set_num_threads(threads)   # torchrun exports OMP_NUM_THREADS=1: set the thread count explicitly
    ng, h, L = w["ng"], w["h"], w["Lg"]
    m = oracle.Mesh.make(ng, (0, 0, 0), h)
    kinds, par = w["dist"]
    R = []
    for d in range(3):
        if kinds[d] == 1:
            R.append(oracle.sample_landau(n_sample, par[2 * d], par[2 * d + 1], 0.0, L[d], 42 + d))
        elif kinds[d] == 0:
            R.append(oracle.sample_landau(n_sample, 0.0, 1.0, 0.0, L[d], 42 + d))      # alpha = 0: uniform
        else:
            R.append(np.clip(par[2 * d] + par[2 * d + 1] * oracle.sample_normal(n_sample, 42 + d), 1e-9, L[d]))
    P = []
    for d in range(3):
        p = oracle.sample_normal(n_sample, 52 + d)
        lo = 0
        for i, (mu, sd, frac) in enumerate(w["vel"]):
            hi = n_sample if i == len(w["vel"]) - 1 else lo + int(frac * n_sample)
            p[lo:hi] = mu[d] + sd[d] * p[lo:hi]
            lo = hi
        P.append(p)
    E = [np.zeros(n_sample) for _ in range(3)]
    cells = m.ext[0] * m.ext[1] * m.ext[2]
    ef = 0.05 * np.sin(np.arange(cells * 3) * 1e-3)   # smooth synthetic E (the solve is a non-owned stage)
    rho = oracle.field_zeros(m)
    dt, q = w["dt"], -(L[0] * L[1] * L[2]) / n_sample
    for _ in range(warmup):
        oracle.pic_step_nosolve(m, R, P, E, q, dt, ef, rho)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.pic_step_nosolve(m, R, P, E, q, dt, ef, rho)
    sec = (time.perf_counter() - t0) / steps
    return n_sample / sec, sec, oracle.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ippl_b200 import app
    world = args.gpus
    w = app.workload(args.config, world, args.log2_particles)
    # N = 1: the whole workload (2^27 particles).  N > 1: the N-GPU mesh with one GPU's share of the particles (a bounded
    # sample: the host's throughput per particle does not depend on how many there are), all host cores either way.
    n_sample = min(w["n_local"], 1 << 27)
    steps, warmup = args.steps, max(1, args.warmup)     # exactly what was asked for (1.5 s per step on 16 cores at C2)
    cores = os.cpu_count() or 1
    value, sec, used = cpu_run(w, n_sample, steps, warmup, threads=cores)
    whole = world == 1 and n_sample == w["n_local"]
    sample = (f"{'the whole workload: ' if whole else 'bounded sample: '}{n_sample} of the {w['n_local'] * world} particles on the full "
              f"{w['ng'][0]}x{w['ng'][1]}x{w['ng'][2]} mesh, {w['label']} initial condition, {steps} steps after {warmup} warm-up; "
              f"OpenMP restatement of the reference algorithm (atomic scatter, three-pass leapfrog), solve excluded")
    print(json.dumps({
        "impl": "reference", "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (initial condition sampled on the host)",
        "config": config_dict(w, world, True),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---- secondary measurements (after the headline numbers are final; each in its own process) -------------------------
def _run_sub(cmd, env, limit_s):
    """One secondary measurement as its own process group: a fault in it (or a hang: killed at the limit) cannot touch the
    headline run.  Returns (last JSON line of its stdout parsed | None, error text | None)."""
    import signal
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT, start_new_session=True)
    try:
        so, se = p.communicate(timeout=limit_s)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)
        except ProcessLookupError:
            pass
        so, se = p.communicate()
        return None, f"no result within {limit_s:.0f} s (process group killed); stderr tail: {se[-300:]}"
    line = next((ln for ln in reversed(so.splitlines()) if ln.startswith("{")), None)
    if line is None:
        return None, f"rc {p.returncode}, no JSON line; stderr tail: {se[-400:]}"
    try:
        d = json.loads(line)
    except ValueError as e:
        return None, f"rc {p.returncode}, unparsable JSON line: {e}"
    if p.returncode != 0:   # the line is printed when the measurement is complete: a non-zero exit after it is the tear-down's
        d["exit_code_after_the_line"] = p.returncode
        d["stderr_tail"] = se[-300:]
    return d, None


def _sub_env(port_shift):
    """Environment of a multi-rank sub-job started from inside a torchrun worker: same RANK / LOCAL_RANK / WORLD_SIZE, its
    own rendezvous port, and rank 0 hosts the store itself (torchrun's agent store lives on the parent's port)."""
    e = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC")}
    e["MASTER_PORT"] = str(int(e.get("MASTER_PORT", "29500")) + port_shift)
    e["IPPLB_PG_TIMEOUT_S"] = "90"
    return e


def _brief(d):
    b = {k: d[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "kernels_ms", "solve_ms",
                           "parity", "gpu_launches", "step_roofline", "run", "exit_code_after_the_line", "stderr_tail") if k in d}
    r = d.get("roofline") or {}
    b["roofline"] = {k: r.get(k) for k in ("kernel", "frac", "achieved", "peak", "unit", "ms_per_launch")}
    b["config"] = d.get("config")
    return b


def extras_leg(args, world, rank, local, dist, jobs=None, micro_cmd=None, micro_limit=240, variant_cmds=None, budget_s=600.0, min_left_s=15.0):
    """BASELINE.json's other configurations, measured in the same driver run as the headline (configs[1]) line:
    configs[3] (BumponTail 512^3, 2^29 particles per GPU) at every N, configs[2] (PenningTrap 256^3, 2^30 particles, ORB)
    at N = 8, configs[4] (scatter / gather microbench, sorted against random order; at N > 1 one replica per GPU, all at
    once), and on rank 0 the A/B runs of the two kernel variants that were written without GPU access -- each as a separate
    process (group) with its own time limit, after the timed region, the roofline and the end-to-end figures of the
    headline line are final.  A failure is recorded as {"error": ...} and changes nothing else."""
    import tempfile
    ex = {"what": "secondary measurements, each in its own process after the headline numbers were final"}
    t0 = time.perf_counter()

    def run_sub(cmd, env, limit):
        # every sub-run has its own limit, and all of them together `budget_s` (the driver's limit for one bench run is fixed)
        left = budget_s - (time.perf_counter() - t0)
        if left < min_left_s:
            return None, f"not started: the {budget_s:.0f} s budget of the secondary measurements was spent"
        return _run_sub(cmd, env, min(float(limit), left))
    box = None
    if world > 1:
        # the ranks meet again through files in a directory named by rank 0 (agreed on while they are still in step):
        # no collective of the headline run's process group is left waiting while the sub-jobs run
        token = [f"{os.getpid()}_{time.time_ns()}" if rank == 0 else None]
        dist.broadcast_object_list(token, src=0)
        box = os.path.join(tempfile.gettempdir(), "ipplb_extras_" + token[0])
        os.makedirs(box, exist_ok=True)
    me = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", str(world), "--steps", "5", "--warmup", "3",
          "--no-e2e", "--no-cpu", "--no-extras"]
    if jobs is None:    # (name, command, time limit in seconds); the tests pass their own
        jobs = [("c4_bumpontail", me + ["--config", "bumpontail"], 180)]
        if world == 8:
            jobs.insert(0, ("c3_penning", me + ["--config", "penning"], 180))
        if world == 1:  # the same workload on the unfused kernels (gather+push, counting sort by cell, sorted scatter): what the fused step replaces
            jobs.append(("c2_unfused_mode1", me + ["--mode", "1"], 120))
        if world > 1:   # the slab-decomposed FFT solve on the 512^3 mesh, checked against and timed next to the replicated one
            jobs.append(("fft_slab_512", me + ["--config", "bumpontail", "--fft", "slab", "--log2-particles", "26"], 150))
    micro = [sys.executable, os.path.join(ROOT, "scripts", "bench_extras.py"), "--device", str(local), "--part"]
    if micro_cmd is None:
        micro_cmd = micro + ["verified"]
    if variant_cmds is None:    # kernel variants that have never run: one process each, in the one-GPU run only
        variant_cmds = [] if world > 1 else [("micro_gather_variants", micro + ["gather_variants"], 150),
                                             ("micro_build_variants", micro + ["build_variants"], 150)]
    for i, (name, cmd, limit) in enumerate(jobs):
        d, err = run_sub(cmd, _sub_env(11 + i) if world > 1 else dict(os.environ), limit)
        if rank == 0:       # only rank 0 of a sub-job prints a line
            ex[name] = _brief(d) if d else {"error": err}
    d, err = run_sub(micro_cmd, dict(os.environ), micro_limit)
    micro = d if d else {"error": err}
    if world > 1:
        mine = os.path.join(box, f"micro_{rank}.json")
        with open(mine + ".tmp", "w") as f:
            json.dump(micro, f)
        os.replace(mine + ".tmp", mine)
    if rank == 0:
        for name, cmd, limit in variant_cmds:
            d, err = run_sub(cmd, dict(os.environ), limit)
            ex[name] = d if d else {"error": err}
    if world > 1:
        # every rank leaves this wait at the same moment: rank 0's marker appears when its variant runs are over
        with open(os.path.join(box, f"done_{rank}"), "w") as f:
            f.write("1")
        files = [os.path.join(box, f"done_{r}") for r in range(world)]
        deadline = min(time.perf_counter() + micro_limit + sum(v[2] for v in variant_cmds), t0 + budget_s) + 30
        while time.perf_counter() < deadline and not all(os.path.exists(f) for f in files):
            time.sleep(0.2)
        if rank == 0:
            files = [os.path.join(box, f"micro_{r}.json") for r in range(world)]
            every = [json.load(open(f)) if os.path.exists(f) else None for f in files]
            worst = {}
            for m in every:
                for r in (m or {}).get("rows", []):
                    for k, v in r.items():
                        if k.endswith("_gpps"):
                            key = f"ppc{r.get('ppc')}_{r.get('order')}_{k}"
                            worst[key] = min(worst.get(key, v), v)
            micro = dict(micro, replicas={"n": world, "failed": sum(1 for m in every if not m or "error" in m),
                                          "min_over_ranks_gpps": worst,
                                          "what": "the same one-GPU sweep on every GPU of the box at the same time"})
    ex["micro"] = micro
    ex["seconds"] = time.perf_counter() - t0
    return ex


# ---- our arm ---------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="landau", choices=["landau", "bumpontail", "penning"])
    ap.add_argument("--log2-particles", type=int, default=None, help="particles per GPU = 2^k (default: the config's own)")
    ap.add_argument("--mode", type=int, default=2, help="2: fused single-pass step (default); 1: push + counting sort + sorted scatter")
    ap.add_argument("--fft", default="replicated", choices=["replicated", "slab"],
                    help="multi-GPU field solve (not part of the timed step; reported as solve_ms): replicated cuFFT solve or the slab-decomposed one")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (other BASELINE configs, microbench)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import ippl_b200 as ib
    from ippl_b200 import app

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    ctx = ib.Context(local)
    dev = ctx.device
    if world > 1:
        # a rank that dies must not leave its peers waiting for the whole time limit: collectives give up after 3 minutes
        dist.init_process_group("nccl", device_id=dev,
                                timeout=datetime.timedelta(seconds=int(os.environ.get("IPPLB_PG_TIMEOUT_S", "180"))))
        uid = [ib.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])

    run = app.MiniApp(ctx, app.workload(args.config, world, args.log2_particles), rank, world, mode=args.mode, fft=args.fft,
                      dist=dist if world > 1 else None)
    w = run.w
    n_local, n_total = w["n_local"], w["n_local"] * world
    parity = None
    if world > 1:    # one small oracle-checked multi-rank step on the same communicator, before the workload is set up
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import mgpu_parity
        parity = mgpu_parity.multi_rank_step(ctx, dist, rank, world)
    run.initialise()                # particles sampled on the device, [ORB], first solve, bucketing
    mesh, bins = run.mesh, run.bins

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run.step(first=True)
    for _ in range(args.warmup - 1):
        run.step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    if bins is not None:
        bins.set_timing(True)
    l0 = ctx.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        run.step()
    t1.record()
    barrier()
    launches = ctx.launches - l0
    clocks = sampler.stop() if rank == 0 else None
    ms = t0.elapsed_time(t1)
    kernel_ms = bins.kernel_ms() if bins is not None else []
    if bins is not None:
        bins.set_timing(False)
    st = run.status()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        c = torch.tensor([st["n_local"], st["flags"] & 7], device=dev, dtype=torch.int64)
        dist.all_reduce(c)
        assert int(c[0]) == n_total and int(c[1]) == 0, f"particles lost in migration: {c.tolist()} of {n_total}"
    else:
        assert st["n_local"] == run.n_mine and (st["flags"] & 7) == 0 and st["n_exit"] == 0, f"fused store lost particles: {st}"
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel: algorithmic bytes per launch / its CUDA-event time INSIDE the timed region --
    peak, peak_src = peaks()
    ncell_int = mesh.nl[0] * mesh.nl[1] * mesh.nl[2]
    kern = dict(run.phase_ms())     # per-phase CUDA-event times of one more step (outside the timed region)
    if bins is not None:
        rk = "fused_step3_kernel"
        ms_launch = float(np.mean(kernel_ms))
        alg = float(BYTES_PER_PARTICLE_STEP) * st["n_local"] + 32.0 * ncell_int
        traffic, traffic_src = (profiled_traffic(rk) if (world == 1 and args.config == "landau" and n_local == 1 << 27)
                                else (None, None))
        kern[rk] = ms_launch
        how = f"mean of {len(kernel_ms)} CUDA-event pairs around the kernel alone, inside the timed region"
    else:
        rk = "gather_push_kernel"
        ms_launch = kern["gather_push"]
        alg = 96.0 * st["n_local"] + 24.0 * mesh.cells
        traffic, traffic_src, how = None, None, "CUDA events around one launch"
    achieved = alg / (ms_launch * 1e-3) / 1e9
    step_bytes = BYTES_PER_PARTICLE_STEP * n_local + BYTES_PER_CELL_STEP * ncell_int
    step_achieved = step_bytes / (ms_per_step * 1e-3) / 1e9
    cfg = config_dict(w, world, bins is not None)     # the workload: the same object in both arms' lines
    run_info = {"tail_fraction": st["n_tail"] / max(st["n_local"], 1)}   # what this run found: kept out of `config`
    run_info.update(run.extra_config())

    out = {
        "metric": w["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic (initial condition sampled on the device)",
        "config": cfg, "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": rk, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg, "ms_per_launch": ms_launch, "ms_per_launch_how": how},
        "step_roofline": {"bytes_per_particle": BYTES_PER_PARTICLE_STEP, "achieved": step_achieved, "peak": peak,
                          "unit": "GB/s", "frac": step_achieved / peak},
        "kernels_ms": kern, "solve_ms": run.solve_ms, "run": run_info,
    }
    if parity is not None:
        out["parity"] = parity

    if rank == 0 and not args.no_cpu and world == 1:
        v, sec, cores = cpu_run(w, 1 << 24, 3, 1, threads=os.cpu_count())
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"2^24 of the {n_local} particles on the same {w['ng'][0]}^3 mesh, same initial condition, 3 steps "
                                         "after 1 warm-up; OpenMP restatement of the reference algorithm, solve excluded "
                                         "(`--impl reference` times the whole workload)"}

    # ---- e2e: the same step through the C-ABI with HOST buffers, copies inside the timed region ----------------------
    if not args.no_e2e and bins is not None:
        e2e = run.e2e(steps=max(4, args.steps), barrier=barrier)     # as many steps as the device-resident arm
        if world > 1:
            t = torch.tensor([e2e["ms_per_step"]], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e["ms_per_step"] = float(t[0])
        e2e["value"] = n_total / (e2e["ms_per_step"] * 1e-3)
        e2e["unit"] = UNIT
        out["e2e"] = e2e
        if world == 1 and args.config == "landau":
            out["e2e_particles_streamed"] = run.e2e_streamed(barrier)

    run.close()
    # ---- secondary measurements: only on the headline workload, only after everything above is final ----------------
    if (not args.no_extras and os.environ.get("IPPLB_BENCH_EXTRAS", "1") != "0" and args.config == "landau"
            and args.log2_particles is None and args.mode == 2):
        del run, bins
        torch.cuda.empty_cache()
        try:
            extras = extras_leg(args, world, rank, local, dist)
        except Exception as e:  # noqa: BLE001 -- never lose the headline line to a secondary measurement
            extras = {"error": f"{type(e).__name__}: {e}"[:400]}
        out["extras"] = extras
    if rank == 0:
        print(json.dumps(out), flush=True)      # on its way before any tear-down: a stuck destroy must not cost the line
    if world > 1:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
