#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench, then the ncu launch list of ONE steady-state step and a
# `--set full` capture of the tensor-core conv / wgrad kernels of that step.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
set -o pipefail
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
echo "bench_full exit=$?"
tail -c 600 gpurun_out/bench_full.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_conv_tc|k_wgrad_tc' -c 60 \
    -f -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
ls -la gpurun_out | tail -20
