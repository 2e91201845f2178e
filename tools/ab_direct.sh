run() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['ms_per_step'], d['value'])"; }
run kmajor_only
SGAN_DIRECT_NMAJOR=1 run both_direct
SGAN_NO_DIRECT=1 run packed
run kmajor_only_again
