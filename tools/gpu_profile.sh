#!/bin/bash
# Runs on the GPU box (under gpurun): plain bench, then the ncu launch list of ONE steady-state step and a
# `--set full` capture of the tensor-core conv / wgrad kernels of that step.  Outputs land in gpurun_out/
# (<= 64 MiB in total: the big capture stays in /tmp, only its CSV pages and two single-launch reports travel).
mkdir -p gpurun_out
# (ncu cannot replay the kernel nodes of the captured step graph reliably: the profiled command keeps train_step eager,
#  which launches exactly the same kernels; the benchmark line itself uses the graph)
if [ -z "$SKIP_BENCH" ]; then
  python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.log 2> gpurun_out/bench_full.err
  echo "bench_full exit=$?"
  tail -c 600 gpurun_out/bench_full.err
fi
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range --no-graph"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit=$?"
ncu --set full --clock-control none --profile-from-start off -k regex:'k_conv_tc|k_wgrad_tc' -c 140 \
    -f -o /tmp/prof_all $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_tc_raw.csv 2> /dev/null
python tools/pick_top_launch.py gpurun_out/prof_tc_raw.csv > gpurun_out/top_launch.txt
cat gpurun_out/top_launch.txt
read -r SKIP_CONV SKIP_WGRAD < <(tail -n 1 gpurun_out/top_launch.txt)
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_conv_tc|k_wgrad_tc' -s $SKIP_CONV -c 1 \
    -f -o gpurun_out/prof_conv_top $CMD > gpurun_out/ncu_top1.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_conv_tc|k_wgrad_tc' -s $SKIP_WGRAD -c 1 \
    -f -o gpurun_out/prof_wgrad_top $CMD > gpurun_out/ncu_top2.log 2>&1
echo "top captures exit=$?"
du -sh gpurun_out
ls -la gpurun_out | tail -20
