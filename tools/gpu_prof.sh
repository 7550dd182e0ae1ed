#!/bin/bash
# ncu --set full captures of named kernels through the device-resident sweeps (one launch each, after warm-up); raw and
# source pages are exported as CSV on the box (the reports themselves exceed what gpurun brings back).
#   gpurun --timeout 900 -- 'bash tools/gpu_prof.sh <tag> "name=kernel_regex:solver:points:mode:extra_flags ..."'
#   e.g.  "eigen_eval=k_linear_eigen:linear_eigen:10000000:f64:--eval  ls_100M=k_linear_ls:linear_LS:100000000:f64:"
#   solver "mv8m" = masked 8-view multi-view sweep; "+" inside extra_flags stands for a blank.
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for SPEC in ${2:-iter_eval=k_iterative_ls:iterative_LS:10000000:f64:--eval}; do
  NAME=${SPEC%%=*}; REST=${SPEC#*=}
  IFS=: read -r K S N MODE FLAGS <<< "$REST"
  FLAGS=${FLAGS//+/ }            # "+" stands for a blank inside a spec: --eval+--rig+forward
  if [ "$S" = mv8m ]; then
    CMD="python tools/sweep_multiview.py --views 8 --visible 0.7 --iters 2 --points $N"
  else
    VAR=0; [ "$MODE" = f32 ] && VAR=-1
    CMD="python tools/sweep_kernels.py --points $N --solvers $S --modes $MODE --variants $VAR --ppts 4 --iters 2 $FLAGS"
  fi
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:^$K\$" -s 3 -c 1 -f -o $OUT/full_$NAME $CMD > $OUT/full_$NAME.log 2>&1
  echo "$NAME ($K) rc=$?"
  ncu -i $OUT/full_$NAME.ncu-rep --page raw --csv > $OUT/full_$NAME.raw.csv 2>/dev/null
  ncu -i $OUT/full_$NAME.ncu-rep --page source --csv > $OUT/full_$NAME.source.csv 2>/dev/null
  rm -f $OUT/full_$NAME.ncu-rep
done
ls -la $OUT | head -30
