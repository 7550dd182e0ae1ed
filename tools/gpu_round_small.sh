#!/bin/bash
# Reduced evidence round (when the GPU budget is short): GPU tests, bench 10 M / 100 M, linear_LS sweeps, launch list and
# ncu --set full of linear_LS only.   gpurun --timeout 900 -- 'bash tools/gpu_round_small.sh r01h'
set -u
TAG=${1:-small}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
echo "== bench 10M"; timeout 300 python bench.py > $OUT/bench_10M.json 2> $OUT/bench_10M.err; cut -c1-300 $OUT/bench_10M.json
echo "== bench 100M"; timeout 300 python bench.py --points 100000000 --steps 5 --no-cpu-baseline > $OUT/bench_100M.json 2> $OUT/bench_100M.err; cut -c1-300 $OUT/bench_100M.json
echo "== sweeps"; for R in rotating forward; do timeout 200 python tools/sweep_kernels.py --points 10000000 --solvers linear_LS,iterative_LS,linear_eigen,polynomial --modes f64 --variants 0 --ppts 4 --rig $R | sed "s/^{/{\"rig\": \"$R\", /" >> $OUT/sweep_rigs_10M.jsonl; done; cut -c1-150 $OUT/sweep_rigs_10M.jsonl
timeout 200 python tools/sweep_kernels.py --points 100000000 --solvers linear_LS --modes f64 --variants 0 --ppts 4 > $OUT/sweep_100M.jsonl; cat $OUT/sweep_100M.jsonl
echo "== ncu launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
for K in k_linear_ls; do
    timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^$K\$" -s 1 -c 1 -f -o $OUT/full_$K python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/full_$K.log 2>&1
    ncu -i $OUT/full_$K.ncu-rep --page raw --csv > $OUT/full_$K.raw.csv 2>/dev/null; ncu -i $OUT/full_$K.ncu-rep --page source --csv > $OUT/full_$K.source.csv 2>/dev/null; rm -f $OUT/full_$K.ncu-rep
done
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^k_linear_ls\$" -s 1 -c 1 -f -o $OUT/full_k_linear_ls_100M python tools/sweep_kernels.py --points 100000000 --solvers linear_LS --modes f64 --variants 0 --ppts 4 --iters 2 > $OUT/full_k_linear_ls_100M.log 2>&1
ncu -i $OUT/full_k_linear_ls_100M.ncu-rep --page raw --csv > $OUT/full_k_linear_ls_100M.raw.csv 2>/dev/null; ncu -i $OUT/full_k_linear_ls_100M.ncu-rep --page source --csv > $OUT/full_k_linear_ls_100M.source.csv 2>/dev/null; rm -f $OUT/full_k_linear_ls_100M.ncu-rep
ls $OUT | head -30
