#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 100 -c 1 -f -o gpurun_out/gemm_fco2_full python tools/microbench.py epi > gpurun_out/ncu_fco.log 2>&1; echo "ncu_fco rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --workload m4c_decode --no-cpu-baseline > gpurun_out/bench_m4c.log 2> gpurun_out/bench_m4c.err; echo "m4c rc=$?" >> gpurun_out/summary.txt
timeout 900 python tools/microbench.py mha > gpurun_out/microbench_mha.jsonl 2> gpurun_out/microbench.err; echo "mha rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cut -c1-500 gpurun_out/bench_m4c.log; tail -3 gpurun_out/bench_m4c.err
