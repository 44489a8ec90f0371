#!/bin/bash
# GEMM epilogue v2, attention fwd/bwd staging, LN templating, M4C decode, CUDA-graph bench
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py gemm sdpa ln > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2> gpurun_out/bench_eager.err; echo "bench_eager rc=$?" >> gpurun_out/summary.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-e2e > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 12 -c 3 -f -o gpurun_out/gemm_full python tools/microbench.py gemm > gpurun_out/ncu_full.log 2>&1; echo "ncu_full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_eager.log; tail -3 gpurun_out/microbench.err
