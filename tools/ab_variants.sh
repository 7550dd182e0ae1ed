#!/bin/bash
# A/B of compile-time variants of libtriangl_cuda on the GPU box: the default build and every variants/*.so run the same
# device-resident kernel sweep.   gpurun --timeout 900 -- 'bash tools/ab_variants.sh <tag> [solvers] [points]'
set -u
TAG=${1:-ab}
SOLVERS=${2:-iterative_LS,linear_eigen,polynomial}
N=${3:-10000000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for LIB in default variants/*.so; do
  NAME=$(basename $LIB .so)
  if [ "$LIB" = default ]; then unset TRGL_CUDA_LIB; else export TRGL_CUDA_LIB=$PWD/$LIB; fi
  timeout 300 python tools/sweep_kernels.py --points $N --solvers $SOLVERS --modes f64 --variants 0 --ppts 4 --rig ${RIG:-rotating} ${SWEEP_FLAGS:-} \
      > $OUT/sweep_$NAME.jsonl 2> $OUT/sweep_$NAME.err
  echo "== $NAME"; cut -c1-260 $OUT/sweep_$NAME.jsonl; tail -2 $OUT/sweep_$NAME.err
done
unset TRGL_CUDA_LIB
