#!/bin/bash
# `ncu --set full` of the fused-epilogue GEMMs (second launch of each case); the command first exits 0 without ncu
mkdir -p gpurun_out
CASES="${CASES:-fc_o fc1 plain}"
N=$(echo $CASES | wc -w)
timeout 200 python tools/ncu_gemm_cases.py $CASES > gpurun_out/ncu_gemm_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_gemm_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c $((2 * N)) -f -o gpurun_out/${TAG:-r02p}_gemm_epi python tools/ncu_gemm_cases.py $CASES > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gemm.log
