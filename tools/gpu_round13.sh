#!/bin/bash
# CTA-pair GEMM (cta_group::2): dedicated tests first, then the suite + bench with pairs on, and the bench with pairs off
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_gemm_tcgen05.py -q -m gpu --timeout 300 -p no:cacheprovider -x -k "cta_pair" > gpurun_out/pair.log 2>&1; PAIR=$?; echo "pair tests rc=$PAIR" >> gpurun_out/summary.txt
[ $PAIR -ne 0 ] && export OVQA_GEMM_2CTA=0
echo "env 2CTA=$OVQA_GEMM_2CTA" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py gemm epi > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
OVQA_GEMM_2CTA=0 timeout 600 python tools/microbench.py gemm epi > gpurun_out/microbench_1cta.jsonl 2>> gpurun_out/microbench.err; echo "microbench1 rc=$?" >> gpurun_out/summary.txt
OVQA_GEMM_2CTA=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1cta.log 2> gpurun_out/bench_1cta.err; echo "bench1 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -25 gpurun_out/pair.log; tail -5 gpurun_out/pytest_gpu.log; cut -c1-400 gpurun_out/bench.log; cut -c1-400 gpurun_out/bench_1cta.log; tail -5 gpurun_out/bench.err
