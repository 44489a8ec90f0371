#!/bin/bash
# tests (incl. the new tensor-core attention backward), kernel microbenchmarks, bench, launch list
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py gemm sdpa ln > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-e2e > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; tail -3 gpurun_out/microbench.err
