#!/bin/bash
# one 8-GPU box: parity over NCCL, the Landau scaling points, C3 (PenningTrap + ORB) and C4 (BumponTail 512^3)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --nproc-per-node 8 --master-port 29521 tests/mgpu_parity.py 2>&1 | grep -v "^\*\|OMP_NUM" | tail -3
for n in 8 4; do
  $TR --nproc-per-node $n --master-port 2953$n bench.py --gpus $n 2>gpurun_out/bench_r2_n$n.err | tail -1 > gpurun_out/bench_r2_n$n.json
  tail -3 gpurun_out/bench_r2_n$n.err | grep -v "^\*\|OMP_NUM"
done
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --config penning 2>gpurun_out/bench_r2_penning_n8.err | tail -1 > gpurun_out/bench_r2_penning_n8.json
tail -3 gpurun_out/bench_r2_penning_n8.err | grep -v "^\*\|OMP_NUM"
$TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --config bumpontail 2>gpurun_out/bench_r2_bumpontail_n8.err | tail -1 > gpurun_out/bench_r2_bumpontail_n8.json
tail -3 gpurun_out/bench_r2_bumpontail_n8.err | grep -v "^\*\|OMP_NUM"
python - <<P
import json
for f in ("n8", "n4", "penning_n8", "bumpontail_n8"):
    try:
        d = json.load(open(f"gpurun_out/bench_r2_{f}.json"))
        print(f, "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "kernel %.4f" % d["roofline"]["ms_per_launch"], "frac %.3f" % d["roofline"]["frac"],
              "e2e %.4g" % d["e2e"]["value"], d["kernels_ms"], d.get("run", d["config"]).get("tail_fraction"), d.get("run", d["config"]).get("orb"), d.get("run", d["config"]).get("migrated_fraction_last_step"), d.get("parity", {}).get("rho_rel_l2"))
    except Exception as e:
        print(f, "failed", e)
P
