#!/bin/bash
# bench.py runs + launch list.  usage: bash scripts/gpu_bench.sh <tag>
set -u
TAG=${1:-bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py > $OUT/bench_f32.json 2> $OUT/bench_f32.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --streams 1 --skip-cpu --skip-e2e > $OUT/bench_f32_serial.json 2> $OUT/bench_f32_serial.err; echo "bench serial rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --dtype bf16 --skip-cpu > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "bench bf16 rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --bs 4 --skip-cpu --steps 10 > $OUT/bench_f32_bs4.json 2> $OUT/bench_f32_bs4.err; echo "bench bs4 rc=$?" | tee -a $OUT/status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dfa_ -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --no-graph --streams 1 > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/status.txt
python - <<PY
import json
for f in ("bench_f32","bench_f32_serial","bench_bf16","bench_f32_bs4"):
    try:
        d=json.load(open("$OUT/%s.json"%f)); print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["kernel_avg_us"])
    except Exception as e: print(f, "failed", e)
PY
for f in $OUT/*.err; do tail -2 $f; done
