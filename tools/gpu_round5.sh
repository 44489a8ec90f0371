#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gemm_tcgen05.py tests/test_gpu_parity.py -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py gemm > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --workload m4c_decode --steps 5 --warmup 3 > gpurun_out/bench_m4c.log 2> gpurun_out/bench_m4c.err; echo "bench_m4c rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 35 -c 2 -f -o gpurun_out/gemm_full python tools/microbench.py gemm > gpurun_out/ncu_full.log 2>&1; echo "ncu_full rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; cat gpurun_out/bench_m4c.log; tail -5 gpurun_out/bench_m4c.err
