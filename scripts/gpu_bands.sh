#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2g}; mkdir -p $OUT
for B in 1 2 4 6; do
  HIPAD_DFA_BANDS=$B timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_b$B.json 2> $OUT/group_b$B.err; echo "bands=$B rc=$?"; python -c "import json;d=json.load(open('$OUT/group_b$B.json'));print({k:d[k]['bwd_group1_us'] for k in ('det','map','plan','ego')}, d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
done
HIPAD_DFA_BANDS=2 timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_b2_bs4.json 2> $OUT/group_b2_bs4.err; python -c "import json;d=json.load(open('$OUT/group_b2_bs4.json'));print('bs4 b2',{k:d[k]['bwd_group1_us'] for k in ('det','map','plan','ego')}, d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_bs4.json 2> $OUT/group_bs4.err; python -c "import json;d=json.load(open('$OUT/group_bs4.json'));print('bs4 default',{k:d[k]['bwd_group1_us'] for k in ('det','map','plan','ego')}, d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
