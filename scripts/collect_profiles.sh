#!/bin/bash
# copy the judged evidence of a gpurun evidence call into profiles/r02/ (tracked).  usage: bash scripts/collect_profiles.sh <gpurun_out tag>
set -eu
SRC=gpurun_out/$1; DST=profiles/r02; mkdir -p $DST
for f in bench_f32_bs1.json bench_bf16_bs1.json bench_f32_bs4.json bench_f32_bs1_serial.json bench_reference_arm.json \
         pytest_gpu.log smoke.log sweep_configs3_bs1.md sweep_configs3_bs4.md sweep_configs3_bs8.md decoder_parity.json \
         ncu_launch_list_bench_steps2.csv full_layer_summary.txt full_layer_raw.csv.gz full_layer_source.csv.gz \
         full_percall_summary.txt full_percall_raw.csv.gz dominant_kernel_traffic.json l2_gather_peak.json gather_roofline.json; do
  [ -f $SRC/$f ] && cp $SRC/$f $DST/$f || echo "missing $f"
done
for f in $SRC/layer_*.json; do cp $f $DST/; done
[ -f $SRC/dominant_kernel_traffic.json ] && cp $SRC/dominant_kernel_traffic.json profiles/dominant_kernel_traffic.json
ls $DST | wc -l
