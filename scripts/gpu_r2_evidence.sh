#!/bin/bash
# round-2 evidence: full GPU test suite, smoke, bench set, sweeps, launch list, ncu captures.  usage: bash scripts/gpu_r2_evidence.sh <tag>
set -u
TAG=${1:-r02}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -p timeout --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/status.txt; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/status.txt; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench_f32_bs1.json 2> $OUT/bench_f32_bs1.err; echo "bench rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --dtype bf16 --skip-cpu --skip-decoder > $OUT/bench_bf16_bs1.json 2> $OUT/bench_bf16.err; echo "bench bf16 rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --bs 4 --skip-cpu --skip-decoder --steps 10 > $OUT/bench_f32_bs4.json 2> $OUT/bench_bs4.err; echo "bench bs4 rc=$?" | tee -a $OUT/status.txt
timeout 600 python bench.py --streams 1 --skip-cpu --skip-e2e --skip-decoder > $OUT/bench_f32_bs1_serial.json 2> $OUT/bench_serial.err; echo "bench serial rc=$?" | tee -a $OUT/status.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_ref.err; echo "ref arm rc=$?" | tee -a $OUT/status.txt
for cfg in "1 f32 352x640 480" "4 f32 352x640 480" "8 f32 352x640 480" "1 bf16 352x640 480" "1 f32 512x1408 48" "1 bf16 512x1408 48" "1 f32 256x704 480"; do
  n=$(echo $cfg | tr ' ' '_'); timeout 300 python profiles/run_group.py $cfg > $OUT/layer_$n.json 2> $OUT/layer_$n.err; echo "layer $cfg rc=$?" | tee -a $OUT/status.txt
done
timeout 600 python profiles/run_sweep.py 1 > $OUT/sweep_configs3_bs1.md 2> $OUT/sweep1.err; echo "sweep bs1 rc=$?" | tee -a $OUT/status.txt
timeout 600 python profiles/run_sweep.py 4 > $OUT/sweep_configs3_bs4.md 2> $OUT/sweep4.err; echo "sweep bs4 rc=$?" | tee -a $OUT/status.txt
timeout 600 python profiles/run_sweep.py 8 > $OUT/sweep_configs3_bs8.md 2> $OUT/sweep8.err; echo "sweep bs8 rc=$?" | tee -a $OUT/status.txt
timeout 600 python harness/parity_report.py > $OUT/decoder_parity.json 2> $OUT/decoder_parity.err; echo "decoder parity rc=$?" | tee -a $OUT/status.txt
python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --skip-decoder --skip-ref-op --no-graph --streams 1 > $OUT/plain_for_ncu.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dfa_ -c 600 --csv --log-file $OUT/ncu_launch_list_bench_steps2.csv \
    python bench.py --steps 2 --warmup 3 --skip-e2e --skip-cpu --skip-decoder --skip-ref-op --no-graph --streams 1 > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?" | tee -a $OUT/status.txt
python profiles/prof_group.py 3 1 > $OUT/plain_prof.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dfa_ --launch-skip 6 -c 6 -f -o /tmp/full_group \
    python profiles/prof_group.py 3 1 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $OUT/status.txt
ncu -i /tmp/full_group.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/full_layer_raw.csv.gz
ncu -i /tmp/full_group.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/full_layer_source.csv.gz
python profiles/summarize_ncu.py /tmp/full_group.ncu-rep > $OUT/full_layer_summary.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/l2_gather_peak profiles/micro/l2_gather_peak.cu && /tmp/l2_gather_peak > $OUT/l2_gather_peak.json; echo "l2 peak rc=$?" | tee -a $OUT/status.txt
python profiles/gather_roofline.py $OUT/full_layer_raw.csv.gz $OUT/l2_gather_peak.json $OUT/gather_roofline.json > /dev/null 2>> $OUT/ncu_full.log
python profiles/prof_percall.py 2 > $OUT/plain_percall.log 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:dfa_ --launch-skip 25 -c 25 -f -o /tmp/full_percall \
    python profiles/prof_percall.py 2 > $OUT/ncu_percall.log 2>&1; echo "ncu percall rc=$?" | tee -a $OUT/status.txt
ncu -i /tmp/full_percall.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/full_percall_raw.csv.gz
python profiles/summarize_ncu.py /tmp/full_percall.ncu-rep > $OUT/full_percall_summary.txt 2>&1
python profiles/ncu_traffic.py $OUT/full_percall_raw.csv.gz $OUT/dominant_kernel_traffic.json > /dev/null 2>> $OUT/ncu_percall.log
cat $OUT/status.txt; du -sh $OUT
