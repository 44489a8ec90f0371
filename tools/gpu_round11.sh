#!/bin/bash
# ncu of the tcgen05 attention kernels, wgrad split-K sweep
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_attention_tc5.py tests/test_gpu_attention_core.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/tc5.log 2>&1; echo "tc5 tests rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py wgrad > gpurun_out/microbench_wgrad.jsonl 2> gpurun_out/microbench.err; echo "wgrad rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdpa_fwd_tc5 --launch-skip 5 -c 1 -f -o gpurun_out/tc5_fwd_full python tools/microbench.py sdpa > gpurun_out/ncu_tc5f.log 2>&1; echo "ncu_fwd rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdpa_bwd_tc5 --launch-skip 5 -c 1 -f -o gpurun_out/tc5_bwd_full python tools/microbench.py sdpa > gpurun_out/ncu_tc5b.log 2>&1; echo "ncu_bwd rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/tc5.log; cat gpurun_out/microbench_wgrad.jsonl
