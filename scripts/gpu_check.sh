#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (N=1), launch list and a full ncu capture of one call of each kind.
# usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -p timeout --timeout 60 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py > $OUT/bench_f32.json 2> $OUT/bench_f32.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --dtype bf16 --skip-cpu > $OUT/bench_bf16.json 2> $OUT/bench_bf16.err; echo "bench bf16 rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --bs 4 --skip-cpu --steps 10 > $OUT/bench_f32_bs4.json 2> $OUT/bench_f32_bs4.err; echo "bench bs4 rc=$?" | tee -a $OUT/status.txt
for k in det map plan; do timeout 120 python profiles/run_kernels.py $k 4 1 >> $OUT/kernels.txt 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --no-graph > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/status.txt
for k in det map plan; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:dfa_ -c 8 -f -o $OUT/full_$k \
      python profiles/run_kernels.py $k 2 1 > $OUT/ncu_full_$k.log 2>&1; echo "ncu full $k rc=$?" | tee -a $OUT/status.txt
done
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench_f32.json; cat $OUT/kernels.txt
