#!/bin/bash
# tcgen05 attention core (fwd, then fwd+bwd), LN-bwd / colsum rewrites, bench
mkdir -p gpurun_out; : > gpurun_out/summary.txt
OVQA_SDPA_TC5_BWD=0 timeout 300 python -m pytest tests/test_gpu_attention_tc5.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/tc5_fwd.log 2>&1; FWD=$?; echo "tc5_fwd rc=$FWD" >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_attention_tc5.py -q -m gpu --timeout 200 -p no:cacheprovider -x > gpurun_out/tc5_bwd.log 2>&1; BWD=$?; echo "tc5_bwd rc=$BWD" >> gpurun_out/summary.txt
[ $FWD -ne 0 ] && export OVQA_SDPA_TC5=0
[ $BWD -ne 0 ] && export OVQA_SDPA_TC5_BWD=0
echo "env TC5=$OVQA_SDPA_TC5 TC5_BWD=$OVQA_SDPA_TC5_BWD" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py sdpa ln > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
OVQA_SDPA_TC5=0 OVQA_SDPA_TC5_BWD=0 timeout 600 python tools/microbench.py sdpa > gpurun_out/microbench_mma.jsonl 2> gpurun_out/microbench_mma.err; echo "microbench_mma rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -15 gpurun_out/tc5_fwd.log; tail -15 gpurun_out/tc5_bwd.log; tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log | cut -c1-1500; tail -5 gpurun_out/bench.err
