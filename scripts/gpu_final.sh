#!/bin/bash
# end-of-round evidence run: GPU tests, smoke, bench set, sweep, launch list
set -u
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -p timeout --timeout 60 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt; tail -2 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt; tail -1 $OUT/smoke.log
bash scripts/gpu_bench.sh ${1:-final}
timeout 400 python profiles/run_sweep.py 1 > $OUT/sweep_bs1.md 2> $OUT/sweep.err; echo "sweep rc=$?" | tee -a $OUT/status.txt; head -8 $OUT/sweep_bs1.md
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2>> $OUT/sweep.err; echo "ref arm rc=$?" | tee -a $OUT/status.txt
