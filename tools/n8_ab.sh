#!/bin/bash
# N = 8 (or $NG): where the data-parallel overhead goes -- step time with single exchanges switched off (diagnostics: wrong results)
run() {
  name="$1"; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus ${NG:-8} --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$name', round(d['ms_per_step'], 3), 'ms/step', round(d['value'], 1), 'img/s')"
}
run diag_no_small_exchanges SGAN_DIAG_LOCAL_SMALL=1
run diag_no_g_bucket SGAN_DIAG_SKIP_G_BUCKET=1
run nccl_buckets SGAN_NO_CE_ALLREDUCE=1
