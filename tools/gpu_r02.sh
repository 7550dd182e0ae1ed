#!/bin/bash
# Round-2 GPU call: parity tests, smoke, the default bench line (N = 1), reference arm, SLAM latency.
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02.sh r02a [phases]'      phases: tests,smoke,bench,ref,slam,rigs,mv,launches,ncu,rare,probe (default: tests,bench,ref,slam)
set -u
TAG=${1:-r02}
PH=${2:-tests,bench,ref,slam}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $OUT/nvidia_smi.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
has() { [[ ",$PH," == *",$1,"* ]]; }
if has tests; then echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -60 $OUT/pytest_gpu.log | cut -c1-400; fi
if has smoke; then echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -8 $OUT/smoke.log; fi
if has bench; then echo "== bench default"; timeout 900 python bench.py > $OUT/bench_10M.json 2> $OUT/bench_10M.err; echo "bench rc=$?"; tail -5 $OUT/bench_10M.err; python tools/show_bench.py $OUT/bench_10M.json; fi
if has ref; then echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cut -c1-500 $OUT/bench_reference.json; fi
if has slam; then echo "== slam"; timeout 600 python bench.py --workload slam --steps 200 --warmup 20 > $OUT/slam_latency.json 2> $OUT/slam.err; echo "slam rc=$?"; tail -3 $OUT/slam.err; python tools/show_bench.py $OUT/slam_latency.json; fi
if has rigs; then echo "== sweep 10M per rig (with the evaluation epilogue)"; rm -f $OUT/sweep_rigs_10M.jsonl; for R in rotating translating forward general; do timeout 300 python tools/sweep_kernels.py --points 10000000 --solvers linear_LS,iterative_LS,linear_eigen,polynomial --modes f64 --variants 0 --ppts 4 --eval --rig $R | sed "s/^{/{\"rig\": \"$R\", /" >> $OUT/sweep_rigs_10M.jsonl; done; cut -c1-160 $OUT/sweep_rigs_10M.jsonl; fi
if has mv; then echo "== multi-view"; timeout 300 python tools/sweep_multiview.py --views 2,3,4,6,8,12,16 > $OUT/sweep_multiview.jsonl 2> $OUT/sweep_multiview.err; timeout 300 python tools/sweep_multiview.py --views 4,8,16 --visible 0.7 >> $OUT/sweep_multiview.jsonl; cut -c1-200 $OUT/sweep_multiview.jsonl; fi
if has launches; then
  echo "== ncu launch list (bench --steps 2 --warmup 1)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-100m > $OUT/launches_bench.log 2>&1
  echo "launch list rc=$? lines=$(wc -l < $OUT/launches.csv)"
fi
if has ncu; then
  bash tools/gpu_prof.sh $TAG "ls=k_linear_ls:linear_LS:10000000:f64: ls_100M=k_linear_ls:linear_LS:100000000:f64: ls_eval=k_linear_ls:linear_LS:10000000:f64:--eval iter_eval=k_iterative_ls:iterative_LS:10000000:f64:--eval eigen_eval=k_linear_eigen:linear_eigen:10000000:f64:--eval poly_eval=k_polynomial:polynomial:10000000:f64:--eval ls_f32_100M=k_linear_ls_f32x4:linear_LS:100000000:f32: mv8_masked=k_multiview_ls:mv8m:10000000:: polyg_forward=k_polynomial_general:polynomial:10000000:f64:--eval+--rig+forward eigeng_forward=k_linear_eigen_general:linear_eigen:10000000:f64:--eval+--rig+forward"
fi
if has rare; then echo "== rare paths of polynomial"; timeout 300 python tools/exp/count_rare_paths.py > $OUT/rare_paths.txt 2>&1; cat $OUT/rare_paths.txt | cut -c1-300; fi
if has probe; then
  echo "== FP64 instruction-kind probe"
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/fp64_ops_probe tools/exp/fp64_ops_probe.cu > /dev/null 2>&1 && /tmp/fp64_ops_probe > $OUT/fp64_ops_probe.txt 2>&1
  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dfma_probe tools/exp/dfma_probe.cu > /dev/null 2>&1 && /tmp/dfma_probe >> $OUT/fp64_ops_probe.txt 2>&1
  cat $OUT/fp64_ops_probe.txt
fi
ls -la $OUT | head -60
