#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_attention_tc5.py tests/test_gpu_attention_core.py -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/tc5.log 2>&1; echo "tc5 tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py sdpa > gpurun_out/microbench_sdpa.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --workload m4c_decode --no-cpu-baseline > gpurun_out/bench_m4c.log 2> gpurun_out/bench_m4c.err; echo "m4c rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/tc5.log | cut -c1-250; tail -4 gpurun_out/pytest_gpu.log; grep -v meta gpurun_out/microbench_sdpa.jsonl | cut -c1-330; cut -c1-300 gpurun_out/bench_m4c.log
