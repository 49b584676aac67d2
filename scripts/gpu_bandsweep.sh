#!/bin/bash
# usage: scripts/gpu_bandsweep.sh TAG  -- band-target / CTA-size sweep of the backward chain on one stage-2 layer
set -u
OUT=gpurun_out/${1:-bands}; mkdir -p $OUT
for th in 512 256; do
for bt in 512 768 1024 1536 2048; do
for bs in 1 4; do
HIPAD_DFA_SORT_THREADS=$th HIPAD_DFA_BAND_TARGET=$bt timeout 300 python profiles/run_group.py $bs f32 > $OUT/t${th}_bt${bt}_bs$bs.json 2> $OUT/t${th}_bt${bt}_bs$bs.err
python -c "import json;d=json.load(open('$OUT/t${th}_bt${bt}_bs$bs.json'));print('threads',$th,'target',$bt,'bs',$bs,{k:d[k]['bwd_group1_us'] for k in ('det','map','plan','ego')}, d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
done
done
done
