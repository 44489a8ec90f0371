#!/bin/bash
# epilogue v4 (compile-time feature sets): tests, microbench, bench with the per-op breakdown, launch list, ncu of fc1-train GEMM
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py epi ln sdpa > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-graph > gpurun_out/ncu.log 2>&1; echo "ncu_list rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 120 -c 1 -f -o gpurun_out/gemm_fc1train_full python tools/microbench.py epi > gpurun_out/ncu_epi.log 2>&1; echo "ncu_fc1 rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 95 -c 1 -f -o gpurun_out/gemm_fco_full python tools/microbench.py epi > gpurun_out/ncu_epi2.log 2>&1; echo "ncu_fco rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log | cut -c1-3000; tail -5 gpurun_out/bench.err; cat gpurun_out/microbench.jsonl
