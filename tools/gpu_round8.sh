#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gemm_tcgen05.py tests/test_gpu_parity.py -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python tools/microbench.py epi > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
# full captures: the heaviest fused epilogue (2nd epi case = fc1 train-style), attention fwd/bwd at 100x100, LayerNorm bwd
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 30 -c 1 -f -o gpurun_out/gemm_epi_full python tools/microbench.py epi > gpurun_out/ncu_epi.log 2>&1; echo "ncu_epi rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sdpa_fwd_mma|sdpa_bwd_mma" --launch-skip 8 -c 2 -f -o gpurun_out/sdpa_full python tools/microbench.py sdpa > gpurun_out/ncu_sdpa.log 2>&1; echo "ncu_sdpa rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd" --launch-skip 5 -c 1 -f -o gpurun_out/ln_full python tools/microbench.py ln > gpurun_out/ncu_ln.log 2>&1; echo "ncu_ln rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; cat gpurun_out/microbench.jsonl
