#!/bin/bash
# bench.py at N GPUs of one box (default line: compute-only value + gather + scene sub-records).
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_scale.sh <tag> N [bench flags]'
set -u
TAG=${1:-scale}; N=${2:-8}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$N" = 1 ]; then
  timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > $OUT/bench_n1.json 2> $OUT/bench_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
      bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
fi
echo "rc=$?"; tail -2 $OUT/bench_n$N.err | cut -c1-300
