#!/bin/bash
# Quick GPU iteration: parity tests + device-resident kernel sweep.   gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag>'
set -u
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/sweep_pixel.py > $OUT/sweep_pixel.jsonl 2> $OUT/sweep_pixel.err; cat $OUT/sweep_pixel.jsonl; tail -3 $OUT/sweep_pixel.err
for N in 10000000 100000000; do
  timeout 600 python tools/sweep_kernels.py --points $N --solvers linear_LS,iterative_LS,linear_eigen,polynomial --modes ${MODES:-f64} --variants 0 --ppts 4 --rig ${RIG:-rotating} > $OUT/sweep_$N.jsonl 2> $OUT/sweep_$N.err
  cut -c1-200 $OUT/sweep_$N.jsonl; tail -3 $OUT/sweep_$N.err
done
