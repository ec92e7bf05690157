#!/bin/bash
# First GPU execution of everything that was written after the round-2 GPU budget was spent.  Every step runs under its OWN
# timeout (a hung multi-rank job must not eat the rest of the budget -- the lesson of profiles/r2_summary.md) and logs to
# gpurun_out/first_*.log.  One GPU is enough for steps 1-4; steps 5-6 need 2 / 8 GPUs and are skipped otherwise.
#   gpurun --timeout 1500 -- bash scripts/gpu_first_run.sh            (1 GPU)
#   gpurun --gpus 8 --timeout 1500 -- bash scripts/gpu_first_run.sh   (8 GPUs)
mkdir -p gpurun_out
NG=$(python -c "import torch; print(torch.cuda.device_count())")
run() {   # run <seconds> <log name> <command...>
    local t=$1 name=$2; shift 2
    timeout "$t" "$@" > "gpurun_out/first_$name.log" 2>&1
    echo "$name: rc=$? ($(tail -1 gpurun_out/first_$name.log | cut -c1-160))"
}
# 1. the validated suite first (kernel parity, loop ranks, facade drivers): must stay green
run 900 suite python -m pytest tests -q -m gpu -x --deselect tests/test_zz_reference_drivers.py --deselect tests/test_zz_slab_fft_gpu.py --deselect tests/test_zz_variants_gpu.py
# 2. the reference's lambdas / unchanged drivers / lazy fusion, one by one
run 120 ref_lambdas demos/ref_lambdas
run 300 ref_drivers python -m pytest tests/test_zz_reference_drivers.py -q -rA
run 120 fusion_check demos/fusion_check
# 2b. the two opt-in kernel variants (bucket build v2, gather v2), each in its own pytest process
run 400 variants python -m pytest tests/test_zz_variants_gpu.py -q -rA
# 3. slab-decomposed FFT on in-process ranks
run 300 slab_fft python tests/slab_fft_check.py
# 4. the headline bench (unchanged kernel; checks nothing regressed)
run 600 bench_n1 python bench.py
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$NG" -ge 2 ]; then
    # 5. two ranks: parity over NCCL, the slab solve over NCCL (bench --fft slab reports solve_ms), facade drivers
    run 300 parity_n2 $TR --nproc-per-node 2 --master-port 29601 tests/mgpu_parity.py
    run 400 bench_n2_slab $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --fft slab --no-e2e
fi
if [ "$NG" -ge 8 ]; then
    # 6. C3 (PenningTrap 256^3, 2^30 particles, ORB) -- the run that faulted in round 2 (unequal inbox segments, fixed blind) -- and C4
    run 600 bench_penning_n8 $TR --nproc-per-node 8 --master-port 29603 bench.py --gpus 8 --config penning
    run 600 bench_bumpontail_n8 $TR --nproc-per-node 8 --master-port 29604 bench.py --gpus 8 --config bumpontail
    run 600 bench_bumpontail_n8_slab $TR --nproc-per-node 8 --master-port 29605 bench.py --gpus 8 --config bumpontail --fft slab --no-e2e
fi
