#!/bin/bash
# Weak-scaling run on one 8-GPU box: bench.py at N = 1, 2, 4, 8 (no data-path collective), then the two gather variants
# at N = 8.   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scale.sh <tag>'
set -u
TAG=${1:-scale}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() {   # N, extra flags, file tag
    local N=$1; shift; local NAME=$1; shift
    if [ "$N" = "1" ]; then
        timeout 400 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
    else
        timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
            bench.py --gpus $N --steps 5 --warmup 3 "$@" > $OUT/$NAME.json 2> $OUT/$NAME.err
    fi
    python - $OUT/$NAME.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], d["n_gpus"], "GPUs", round(d["value"] / 1e9, 2), "G pts/s", round(d["ms_per_step"], 3), "ms/step",
          "e2e", round(d["e2e"]["value"] / 1e9, 3), {k: round(v["kernel_ms"], 3) for k, v in d["per_solver"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for N in 1 2 4 8; do run $N n$N; done
run 8 n8_gather --gather
run 8 n8_p2p --p2p-gather
