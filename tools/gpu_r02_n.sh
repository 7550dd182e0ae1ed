#!/bin/bash
# Multi-GPU call of round 2: real-peer PeerGather test + the default bench line at N GPUs (gather / scene sub-records).
#   gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_r02_n.sh r02n2 N'
set -u
TAG=${1:-r02n}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest peer gather"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "peer_gather or mirrors" > $OUT/pytest_peer.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest_peer.log | cut -c1-300
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 ${BENCH_FLAGS:-} > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?"; tail -5 $OUT/bench_n$N.err | cut -c1-300
python tools/show_bench.py $OUT/bench_n$N.json | head -260
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-600
