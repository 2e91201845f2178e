run() { timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$2', d['ms_per_step'], d['value'])"; }
run 29511 default
NCCL_MAX_CTAS=4 run 29512 max_ctas4
NCCL_MAX_CTAS=2 run 29513 max_ctas2
SGAN_DP_SYNC_ALLREDUCE=1 run 29514 sync
SGAN_NO_PEER=1 run 29515 nopeer
