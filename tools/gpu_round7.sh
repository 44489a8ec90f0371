#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py epi > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --workload m4c_decode --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_m4c.log 2> gpurun_out/bench_m4c.err; echo "bench_m4c rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_m4c.log; tail -5 gpurun_out/bench_m4c.err; cat gpurun_out/microbench.jsonl
