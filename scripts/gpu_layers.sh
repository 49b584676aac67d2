#!/bin/bash
# one-layer timings (profiles/run_group.py) for the evidence set.  usage: bash scripts/gpu_layers.sh <tag>
OUT=gpurun_out/${1:-layers}; mkdir -p $OUT
for cfg in "1 f32 352x640 480" "4 f32 352x640 480" "8 f32 352x640 480" "1 bf16 352x640 480" "1 f32 512x1408 48" "1 bf16 512x1408 48" "1 f32 256x704 480"; do
  n=$(echo $cfg | tr ' ' '_'); timeout 300 python profiles/run_group.py $cfg > $OUT/layer_$n.json 2> $OUT/layer_$n.err; echo "layer $cfg rc=$?"
  python -c "
import json,sys; l=json.load(open('$OUT/layer_$n.json'))['layer']; print(l['fwd_grouped_us'], l['bwd_grouped_us'], l['bwd_grouped_stage_us'])"
done
