#!/bin/bash
# first GPU round trip: diagnose the tcgen05 GEMM, then the parity suite with the CUDA-core GEMM twin, then for real
mkdir -p gpurun_out; : > gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 300 python tests/diag_tcgen05.py > gpurun_out/diag.log 2>&1; echo "diag rc=$?" >> gpurun_out/summary.txt
OVQA_BF16_GEMM=simt timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/parity_simt.log 2>&1; echo "parity_simt rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_gemm_tcgen05.py tests/test_gpu_parity.py -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/parity_tc.log 2>&1; echo "parity_tc rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/diag.log; tail -15 gpurun_out/parity_simt.log; tail -15 gpurun_out/parity_tc.log; tail -5 gpurun_out/smoke.log
