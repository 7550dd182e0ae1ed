#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (10 M and 100 M points), ncu launch list and ncu --set full
# captures of the four solver kernels.  Everything lands in gpurun_out/ (scratch); tools/summarise_ncu.py
# turns the captures into the committed summaries under profiles/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01b'
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/nvidia_smi.csv 2>&1

echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -6 $OUT/smoke.log
echo "== bench 10M"; timeout 600 python bench.py > $OUT/bench_10M.json 2> $OUT/bench_10M.err; echo "bench rc=$?"; cut -c1-600 $OUT/bench_10M.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cut -c1-400 $OUT/bench_reference.json
echo "== bench 10M, evaluation of linear_LS fused as well"; timeout 300 python bench.py --no-cpu-baseline --fuse-ls-eval > $OUT/bench_10M_fuse_ls.json 2>> $OUT/bench_10M.err; cut -c1-200 $OUT/bench_10M_fuse_ls.json
for V in variants/*.so; do [ -f $V ] || continue; echo "== bench 10M with $V"; TRGL_CUDA_LIB=$PWD/$V timeout 300 python bench.py --no-cpu-baseline > $OUT/bench_10M_$(basename $V .so).json 2>> $OUT/bench_10M.err; done
echo "== bench 100M"; timeout 600 python bench.py --points 100000000 --steps 5 --no-cpu-baseline > $OUT/bench_100M.json 2> $OUT/bench_100M.err; cut -c1-400 $OUT/bench_100M.json
echo "== sweep 100M"; timeout 600 python tools/sweep_kernels.py --points 100000000 --solvers linear_LS,iterative_LS,linear_eigen,polynomial --modes f64,f32 --variants 0,8 --ppts 4 > $OUT/sweep_100M.jsonl 2> $OUT/sweep_100M.err; cat $OUT/sweep_100M.jsonl
echo "== sweep 10M per rig"; for R in rotating translating forward general; do timeout 300 python tools/sweep_kernels.py --points 10000000 --solvers linear_LS,iterative_LS,linear_eigen,polynomial --modes f64 --variants 0 --ppts 4 --rig $R | sed "s/^{/{\"rig\": \"$R\", /" >> $OUT/sweep_rigs_10M.jsonl; done; cut -c1-150 $OUT/sweep_rigs_10M.jsonl
echo "== multi-view"; timeout 300 python tools/sweep_multiview.py --views 2,4,8,16 > $OUT/sweep_multiview.jsonl 2> $OUT/sweep_multiview.err; timeout 300 python tools/sweep_multiview.py --views 8 --visible 0.7 >> $OUT/sweep_multiview.jsonl; cut -c1-200 $OUT/sweep_multiview.jsonl

echo "== ncu launch list (bench --steps 2 --warmup 1)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/launches_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/launches.csv)"

# gpurun brings back at most 64 MiB: export the raw / source pages as CSV on the box and drop the 20 MB reports
export_rep() {
    ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null
    ncu -i $1.ncu-rep --page source --csv > $1.source.csv 2>/dev/null
    rm -f $1.ncu-rep
}
echo "== ncu --set full, the four solver kernels at 10 M points (one launch each, after warm-up)"
for K in k_linear_ls k_iterative_ls k_linear_eigen k_polynomial; do
    # anchored: k_linear_eigen must not match its follow-up kernel k_linear_eigen_general, k_linear_ls not its ring / tma variants
    timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^$K\$" -s 1 -c 1 -f -o $OUT/full_$K \
        python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/full_$K.log 2>&1
    echo "$K rc=$?"
    export_rep $OUT/full_$K
done
echo "== ncu --set full, linear_LS at 100 M points"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^k_linear_ls\$" -s 1 -c 1 -f -o $OUT/full_k_linear_ls_100M \
    python tools/sweep_kernels.py --points 100000000 --solvers linear_LS --modes f64 --variants 0 --ppts 4 --iters 2 > $OUT/full_k_linear_ls_100M.log 2>&1
echo "100M rc=$?"
export_rep $OUT/full_k_linear_ls_100M
ls -la $OUT
