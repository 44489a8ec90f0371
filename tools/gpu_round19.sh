#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py tests/test_gpu_parity.py -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py epi > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log; grep -v meta gpurun_out/microbench.jsonl | cut -c1-200; cut -c1-300 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
