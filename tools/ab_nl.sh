run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'])"; }
run nl_tensor_ops
SGAN_NO_NL_TC=1 run nl_ffma
