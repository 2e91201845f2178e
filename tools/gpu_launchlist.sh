#!/bin/bash
# ncu launch list (device time per launch) of ONE steady-state train step -> gpurun_out/launches_step.csv
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range --no-graph $BENCH_ARGS"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit=$?"
