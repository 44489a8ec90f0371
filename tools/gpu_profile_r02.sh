#!/bin/bash
# round-2 profiling run on one GPU: launch lists of the three workloads (eager, serialised: SHARES are meaningful) and `ncu --set full` of
# the top kernels.  Every ncu command runs only after the same command line exited 0 without ncu (B200_PROFILING.md).
mkdir -p gpurun_out; : > gpurun_out/prof_summary.txt
run() { echo "== $*" >> gpurun_out/prof_summary.txt; }
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-graph --no-cpu-baseline"
for wl in mcan_train m4c_decode xmt_train; do
  timeout 300 $B --workload $wl > gpurun_out/plain_$wl.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_$wl.csv $B --workload $wl > gpurun_out/ncu_$wl.log 2>&1
  echo "launch list $wl rc=$?" >> gpurun_out/prof_summary.txt
done
# full captures: wide-head attention forward of the M4C decode, tensor-memory attention fwd + bwd and the fc_o-class GEMM of the MCAN step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdpa_fwd_tc5w_kernel --launch-skip 20 -c 1 -f -o gpurun_out/r02_tc5w_full $B --workload m4c_decode > gpurun_out/ncu_tc5w.log 2>&1; echo "ncu tc5w rc=$?" >> gpurun_out/prof_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sdpa_fwd_tc5_kernel|sdpa_bwd_tc5_kernel" --launch-skip 40 -c 2 -f -o gpurun_out/r02_tc5_full $B --workload mcan_train > gpurun_out/ncu_tc5.log 2>&1; echo "ncu tc5 rc=$?" >> gpurun_out/prof_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"geo_bias|attr_pool" --launch-skip 12 -c 4 -f -o gpurun_out/r02_geo_pool_full $B --workload xmt_train > gpurun_out/ncu_geo.log 2>&1; echo "ncu geo/pool rc=$?" >> gpurun_out/prof_summary.txt
cat gpurun_out/prof_summary.txt
