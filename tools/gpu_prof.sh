#!/bin/bash
# ncu --set full captures of named kernels through the device-resident sweep (one launch each, after warm-up); raw and
# source pages are exported as CSV on the box (the reports themselves exceed what gpurun brings back).
#   gpurun --timeout 900 -- 'bash tools/gpu_prof.sh <tag> "k_iterative_ls:iterative_LS k_polynomial:polynomial"'
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for KS in ${2:-k_iterative_ls:iterative_LS}; do
  K=${KS%%:*}; S=${KS##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:^$K\$" -s 1 -c 1 -f -o $OUT/full_$K \
      python tools/sweep_kernels.py --points ${N:-10000000} --solvers $S --modes f64 --variants 0 --ppts 4 --iters 2 ${SWEEP_FLAGS:-} > $OUT/full_$K.log 2>&1
  echo "$K rc=$?"
  ncu -i $OUT/full_$K.ncu-rep --page raw --csv > $OUT/full_$K.raw.csv 2>/dev/null
  ncu -i $OUT/full_$K.ncu-rep --page source --csv > $OUT/full_$K.source.csv 2>/dev/null
  rm -f $OUT/full_$K.ncu-rep
done
ls -la $OUT
