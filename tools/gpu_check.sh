#!/bin/bash
# Runs on the GPU box (under gpurun): kernel-level parity tests in separate processes so that a faulting
# tensor-core kernel cannot take the other results down with it.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, pytest args...
  local name=$1; shift
  timeout 600 python -m pytest "$@" -m gpu -q -p no:cacheprovider --tb=short --maxfail=8 > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" >> gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log
}
: > gpurun_out/summary.txt
run simt   tests/test_kernels_gpu.py -k "not tc and not transpose"
run tc_bf16 tests/test_kernels_gpu.py -k "tc and bf16"
run tc_tf32 tests/test_kernels_gpu.py -k "tc and tf32"
run convT  tests/test_kernels_gpu.py -k "transpose"
cat gpurun_out/summary.txt
