#!/bin/bash
# GPU-side experiment driver: parity of the fused step under a CTA-shape / kernel-generation knob, then the
# bench line per knob.  Usage: scripts/cfg_sweep.sh "<test cfgs>" "<bench cfgs>" [ncu cfg]
mkdir -p gpurun_out
for c in $1; do
  echo "== tests cfg $c"
  IPPLB_FUSED_VAR=$c timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_step or landau_energy or smoke" 2>&1 | tail -4
done
for c in $2; do
  echo "== bench cfg $c"
  IPPLB_FUSED_VAR=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/sweep_cfg$c.json 2> gpurun_out/sweep_cfg$c.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/sweep_cfg$c.json").read().strip().splitlines()[-1])
    print("cfg $c ms/step", round(d["ms_per_step"], 4), "fused ms", round(d["kernels_ms"]["fused_step"], 4), "frac", round(d["roofline"]["frac"], 4), d["clocks"])
except Exception as e:
    print("cfg $c failed", e); print(open("gpurun_out/sweep_cfg$c.err").read()[-2000:])
P
done
if [ -n "$3" ]; then
  IPPLB_FUSED_VAR=$3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_step -s 4 -c 1 -o gpurun_out/ncu_cfg$3 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_cfg$3.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
