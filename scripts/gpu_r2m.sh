#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2m}; mkdir -p $OUT
HIPAD_DFA_REDUCE_LIGHT=1 timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt; tail -3 $OUT/pytest.log
for V in 0 1; do
HIPAD_DFA_REDUCE_LIGHT=$V timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_l$V.json 2> $OUT/group_l$V.err; echo "light=$V rc=$?"; python -c "import json;d=json.load(open('$OUT/group_l$V.json'));print({k:(d[k]['fwd_group1_us'],d[k]['bwd_group1_us']) for k in ('det','map','plan','ego')}, d['layer']['fwd_grouped_us'], d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
HIPAD_DFA_REDUCE_LIGHT=$V timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_bs4_l$V.json 2> $OUT/group_bs4_l$V.err; python -c "import json;d=json.load(open('$OUT/group_bs4_l$V.json'));print('bs4', d['layer']['fwd_grouped_us'], d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
done
