#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 2 --warmup 2 --no-e2e > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-e2e > gpurun_out/ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -12 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
