run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$1 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', d['ms_per_step'], d['value'])"; }
run 1 base
NCCL_MAX_CTAS=8 run 2 maxctas8
GH_SM_BUDGET=132 NCCL_MAX_CTAS=16 run 3 budget132_ctas16
GH_SM_BUDGET=140 NCCL_MAX_CTAS=8 run 4 budget140_ctas8
