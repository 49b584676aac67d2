#!/bin/bash
# short refresh of the evidence that depends on the kernels (after a kernel change): smoke, bench lines, ncu capture of one grouped layer.
# usage: bash scripts/gpu_r2_refresh.sh <tag>
set -u
TAG=${1:-r02}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt
timeout 900 python bench.py > $OUT/bench_f32_bs1.json 2> $OUT/bench_f32_bs1.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --dtype bf16 --skip-cpu --skip-decoder > $OUT/bench_bf16_bs1.json 2> $OUT/bench_bf16.err; echo "bench bf16 rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --bs 4 --skip-cpu --skip-decoder --steps 10 > $OUT/bench_f32_bs4.json 2> $OUT/bench_bs4.err; echo "bench bs4 rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --streams 1 --skip-cpu --skip-e2e --skip-decoder > $OUT/bench_f32_bs1_serial.json 2> $OUT/bench_serial.err; echo "bench serial rc=$?" | tee -a $OUT/status.txt
timeout 300 python profiles/run_group.py 1 f32 352x640 480 > $OUT/layer_1_f32_352x640_480.json 2> /dev/null; echo "layer rc=$?" | tee -a $OUT/status.txt
python profiles/prof_group.py 3 1 > $OUT/plain_prof.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dfa_ --launch-skip 6 -c 6 -f -o /tmp/full_group \
    python profiles/prof_group.py 3 1 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/status.txt
ncu -i /tmp/full_group.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/full_layer_raw.csv.gz
ncu -i /tmp/full_group.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/full_layer_source.csv.gz
python profiles/summarize_ncu.py /tmp/full_group.ncu-rep > $OUT/full_layer_summary.txt 2>&1
cp profiles/r02/l2_gather_peak.json $OUT/l2_gather_peak.json
python profiles/gather_roofline.py $OUT/full_layer_raw.csv.gz $OUT/l2_gather_peak.json $OUT/gather_roofline.json > /dev/null 2>> $OUT/ncu_full.log
cat $OUT/status.txt
