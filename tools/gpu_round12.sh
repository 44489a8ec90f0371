#!/bin/bash
# fused bias gradients (GEMM colsum epilogue, analytic dbk, dbv from dctx), tc5 dispatch heuristics
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log | cut -c1-1200; tail -5 gpurun_out/bench.err
