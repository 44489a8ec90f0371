#!/bin/bash
# round-2 `ncu --set full` captures (one launch each; every profiled command first exits 0 without ncu)
mkdir -p gpurun_out; : > gpurun_out/ncu_summary.txt
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-graph --no-cpu-baseline"
cap() {  # name, kernel regex, launch-skip, count, workload
  timeout 200 $B --workload $5 > gpurun_out/plain_$1.log 2>&1 &&
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip $3 -c $4 -f -o gpurun_out/r02_$1 $B --workload $5 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1 rc=$?" >> gpurun_out/ncu_summary.txt
}
cap tc5w_full "sdpa_fwd_tc5w_kernel" 4 1 m4c_decode
cap tc5_full "sdpa_fwd_tc5_kernel|sdpa_bwd_tc5_kernel" 30 2 mcan_train
cap geo_pool_full "geo_bias|attr_pool" 6 4 xmt_train
cat gpurun_out/ncu_summary.txt
