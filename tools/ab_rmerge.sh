run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'])"; }
run merged_r_backward
SGAN_NO_MERGED_R_BWD=1 run separate_r_backward
