run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'])"; }
run fused_shortcut
SGAN_NO_FUSED_SHORTCUT=1 run separate_shortcut
run fused_shortcut_again
