#!/bin/bash
# round-2 first call: decoder harness on the GPU (parity vs reference CUDA op, decoder timing), baseline bench line
set -u
OUT=gpurun_out/${1:-r2a}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_decoder_harness.py -m gpu -x -q -p timeout --timeout 600 > $OUT/pytest_harness.log 2>&1; echo "harness pytest rc=$?" | tee -a $OUT/status.txt; tail -15 $OUT/pytest_harness.log
timeout 600 python harness/bench_decoder.py --frames 30 > $OUT/decoder.json 2> $OUT/decoder.err; echo "decoder bench rc=$?" | tee -a $OUT/status.txt; cat $OUT/decoder.json; tail -5 $OUT/decoder.err
timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/status.txt; cut -c1-600 $OUT/bench.json
