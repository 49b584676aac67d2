#!/bin/bash
set -u
OUT=gpurun_out/${1:-r2h}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -x -q -p timeout --timeout 300 > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt; tail -3 $OUT/pytest.log
for T in 256 512; do
  HIPAD_DFA_SORT_THREADS=$T timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_t$T.json 2> $OUT/group_t$T.err; echo "sort threads=$T rc=$?"; python -c "import json;d=json.load(open('$OUT/group_t$T.json'));print({k:d[k]['bwd_group1_us'] for k in ('det','map','plan','ego')}, d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
done
HIPAD_DFA_SORT_THREADS=256 HIPAD_DFA_BANDS=32 timeout 300 python profiles/run_group.py 1 f32 > $OUT/group_t256_b32.json 2> $OUT/group_t256_b32.err; python -c "import json;d=json.load(open('$OUT/group_t256_b32.json'));print('b32',{k:d[k]['bwd_group1_us'] for k in ('det','map','plan','ego')}, d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
for T in 256 512; do
  HIPAD_DFA_SORT_THREADS=$T timeout 300 python profiles/run_group.py 4 f32 > $OUT/group_bs4_t$T.json 2> $OUT/group_bs4_t$T.err; python -c "import json;d=json.load(open('$OUT/group_bs4_t$T.json'));print('bs4 t$T',{k:d[k]['bwd_group1_us'] for k in ('det','map','plan','ego')}, d['layer']['bwd_grouped_us'], d['layer']['bwd_grouped_stage_us'])"
done
