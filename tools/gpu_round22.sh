#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_m4c.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-graph --workload m4c_decode --no-cpu-baseline > gpurun_out/ncu_m4c.log 2>&1; echo "ncu_m4c rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/ncu_m4c.log | cut -c1-300
