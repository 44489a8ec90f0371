#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
OVQA_DEBUG=1 timeout 600 python tools/microbench.py gemm > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 330 -c 1 -f -o gpurun_out/gemm_pair_full python tools/microbench.py gemm > gpurun_out/ncu_pair.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_attention_tc5.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/tc5.log 2>&1; echo "tc5 rc=$?" >> gpurun_out/summary.txt
timeout 300 python tools/microbench.py sdpa > gpurun_out/microbench_sdpa.jsonl 2>> gpurun_out/microbench.err; echo "sdpa rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -i "co-resident" gpurun_out/microbench.err | sort | uniq -c; cat gpurun_out/microbench.jsonl | cut -c1-260; tail -3 gpurun_out/tc5.log
