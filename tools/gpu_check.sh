#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests in separate processes so that a faulting kernel cannot take the
# other results down with it.  Logs land in gpurun_out/.  Usage: bash tools/gpu_check.sh [group ...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, pytest args...
  local name=$1; shift
  timeout 900 python -m pytest "$@" -m gpu -q -p no:cacheprovider --tb=short --maxfail=10 > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" >> gpurun_out/summary.txt
  tail -n 30 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
groups="$@"
[ -z "$groups" ] && groups="simt tc convT models"
for g in $groups; do
  case $g in
    simt)   run simt   tests/test_kernels_gpu.py -k "not tc and not transpose" ;;
    tc)     run tc     tests/test_kernels_gpu.py -k "tc" ;;
    convT)  run convT  tests/test_kernels_gpu.py -k "transpose" ;;
    models) run models tests/test_models_gpu.py ;;
    *)      run $g tests -k "$g" ;;
  esac
done
cat gpurun_out/summary.txt
