#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2l}; mkdir -p $OUT
HIPAD_DFA_GROUP_CTAS=8 timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt; tail -3 $OUT/pytest.log
for V in 6 8; do
HIPAD_DFA_GROUP_CTAS=$V timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_c$V.json 2> $OUT/group_c$V.err; echo "ctas=$V rc=$?"; python -c "import json;d=json.load(open('$OUT/group_c$V.json'));print({k:(d[k]['fwd_group1_us'],d[k]['bwd_group1_us']) for k in ('det','map','plan','ego')}, d['layer'])"
HIPAD_DFA_GROUP_CTAS=$V timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_bs4_c$V.json 2> $OUT/group_bs4_c$V.err; python -c "import json;d=json.load(open('$OUT/group_bs4_c$V.json'));print('bs4', d['layer']['fwd_grouped_us'], d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
done
