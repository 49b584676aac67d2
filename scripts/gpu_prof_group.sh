#!/bin/bash
# ncu --set full of the grouped layer's kernels (second repetition); keeps CSV exports + summary
set -u
TAG=${1:-r2prof}; OUT=gpurun_out/$TAG; mkdir -p $OUT
# per repetition: fwd memset? + dfa_group_kernel(fwd), then bwd: dfa_group_kernel(bwd), vis_compact, band_sort, classify, reduce
python profiles/prof_group.py 3 1 > $OUT/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dfa_group --launch-skip 2 -c 2 -f -o /tmp/full_group \
    python profiles/prof_group.py 3 1 > $OUT/ncu_full.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/full_group.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/full_group_raw.csv.gz
ncu -i /tmp/full_group.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/full_group_source.csv.gz
python profiles/summarize_ncu.py /tmp/full_group.ncu-rep > $OUT/full_group_summary.txt 2>&1
cat $OUT/full_group_summary.txt
du -sh $OUT
