#!/bin/bash
# Round-2 same-box A/B of this session's work (profiles/r02_same_box_ab.txt):
#   old = libgenhancer_b200_ab.so built from commit 15a565a (first-form attention backward, general epilogue for aux_out, 64-wide
#         split-K tiles) + GH_LORA_FUSED=0 GH_LORA_ONES=0 (dropout kernel + skinny GEMMs, column-sum bias gradients)
#   new = the product library, defaults
# arms alternate inside ONE gpurun call; bench.py --steps 8 --warmup 3, no baselines.
cd "$(dirname "$0")/.."
run() {  # label, config, env...
  local label=$1 cfg=$2; shift 2
  env "$@" timeout 400 python bench.py --config $cfg --steps 8 --warmup 3 --no-cpu-baseline --no-library-baseline --no-parity 2>/dev/null | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$label', '$cfg', d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['achieved'])"
}
OLD="GH_LIB_PATH=genhancer_b200/libgenhancer_b200_ab.so GH_LORA_FUSED=0 GH_LORA_ONES=0"
for rep in 1 2; do
  run old img336_stage1 $OLD
  run new img336_stage1 GH_X=1
  run old siglip384_stage2_all $OLD
  run new siglip384_stage2_all GH_X=1
done
run old use2frames336_stage1 $OLD
run new use2frames336_stage1 GH_X=1
run old sliding336_stage1 $OLD
run new sliding336_stage1 GH_X=1
