#!/bin/bash
# Runs on the GPU box (under gpurun): the ncu launch list of ONE steady-state eager step, then `--set full` single-launch
# captures (with source) of the longest k_conv_tc and the longest k_wgrad_tc launch of that step.  Outputs in gpurun_out/.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --profile-range --no-graph"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_step.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit=$?"
read -r SKIP_CONV SKIP_WGRAD < <(python tools/pick_top_from_list.py gpurun_out/launches_step.csv)
echo "top conv #$SKIP_CONV, top wgrad #$SKIP_WGRAD (indices among the tensor-core launches)"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_conv_tc|k_wgrad_tc' -s $SKIP_CONV -c 1 \
    -f -o gpurun_out/prof_conv_top $CMD > gpurun_out/ncu_top1.log 2>&1
echo "conv capture exit=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'k_conv_tc|k_wgrad_tc' -s $SKIP_WGRAD -c 1 \
    -f -o gpurun_out/prof_wgrad_top $CMD > gpurun_out/ncu_top2.log 2>&1
echo "wgrad capture exit=$?"
ncu -i gpurun_out/prof_conv_top.ncu-rep --page details > gpurun_out/conv_tc_top_details.txt 2>/dev/null
ncu -i gpurun_out/prof_wgrad_top.ncu-rep --page details > gpurun_out/wgrad_tc_top_details.txt 2>/dev/null
ls -la gpurun_out | tail -12
