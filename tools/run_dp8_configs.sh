mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 8 --warmup 3 --no-roofline --no-parity "$@" > gpurun_out/r2j_dp8_$tag.log 2>&1; grep "^{" gpurun_out/r2j_dp8_$tag.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$tag', d['config']['name'], d['value'], d['unit'], d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'mfu', d['mfu']['of_nominal_2250'], d.get('dp_check',{}).get('ranks_identical'), d['clocks']['sm_mhz'], 'host', d['host_enqueue_ms_per_step'])
"; tail -2 gpurun_out/r2j_dp8_$tag.log | cut -c1-200 | grep -v "^{"; }
run img_ga1 --config img336_stage1
run img_ga2 --config img336_stage1 --grad-accum 2
run siglip --config siglip384_stage2_all
run use2 --config use2frames336_stage1
run sliding --config sliding336_stage1
