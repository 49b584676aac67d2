#!/bin/bash
# A/B of compile-time variants on the GPU box: rebuild the library with each flag set, time one stage-2 layer.
# usage: scripts/exp_variants.sh <outdir> "<name>|<nvcc flags>|<env assignments>" ...
out=gpurun_out/$1; shift
mkdir -p $out
for spec in "$@"; do
  name=${spec%%|*}; rest=${spec#*|}; flags=${rest%%|*}; envs=${rest#*|}
  export HIPAD_DFA_NVCC_EXTRA="$flags"
  python hip-ad_b200/build.py > $out/build_$name.log 2>&1 || { echo "$name build failed"; tail -5 $out/build_$name.log; continue; }
  for bs in ${EXP_BS:-1}; do
    env $envs python profiles/run_group.py $bs ${EXP_DTYPE:-f32} > $out/${name}_bs$bs.json 2> $out/${name}_bs$bs.err
    python - $out/${name}_bs$bs.json $name $bs <<'P'
import json,sys
d=json.load(open(sys.argv[1])); l=d["layer"]
print(sys.argv[2], "bs", sys.argv[3], "fwd", l["fwd_grouped_us"], "bwd", l["bwd_grouped_us"], "acc", l["bwd_grouped_accumulate_us"], l["bwd_grouped_stage_us"],
      {k:(d[k]["fwd_group1_us"],d[k]["bwd_group1_us"]) for k in ("det","map","plan","ego")})
P
  done
done
