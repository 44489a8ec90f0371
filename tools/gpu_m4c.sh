#!/bin/bash
# M4C iteration: attention / M4C tests + the decode bench
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_attention_tc5.py tests/test_gpu_attention_core.py tests/test_gpu_m4c.py -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/tc5.log 2>&1; echo "tc5+m4c tests rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --workload m4c_decode --no-cpu-baseline > gpurun_out/bench_m4c.log 2> gpurun_out/bench_m4c.err; echo "m4c rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/tc5.log | cut -c1-250; cut -c1-300 gpurun_out/bench_m4c.log; tail -3 gpurun_out/bench_m4c.err
