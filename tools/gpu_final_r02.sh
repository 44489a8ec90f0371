#!/bin/bash
# end-of-round-2 validation on one GPU: full GPU suite, smoke(), the default bench line (all four workloads), the reference arm, kernel
# microbenchmarks, eager launch lists of the three workloads (ncu, serialised: SHARES are meaningful), the kineto step audit
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
S=$(date +%s); timeout 1200 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
S=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02s_bench_reference.json 2> gpurun_out/r02s_bench_reference.err; echo "bench_ref rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
timeout 900 python tools/microbench.py gemm epi sdpa ln mha > gpurun_out/r02s_microbench.jsonl 2> gpurun_out/r02s_microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
timeout 120 python tools/lstm_probe.py 256 30 512 > gpurun_out/r02s_lstm_probe.json 2>&1; echo "lstm probe rc=$?" >> gpurun_out/summary.txt
timeout 200 python tools/step_audit.py > gpurun_out/r02s_step_audit.json 2> gpurun_out/r02s_step_audit.err; echo "audit rc=$?" >> gpurun_out/summary.txt
B="python bench.py --steps 2 --warmup 2 --no-e2e --no-graph --no-cpu-baseline"
for wl in mcan_train m4c_decode xmt_train; do
  timeout 300 $B --workload $wl > gpurun_out/plain_$wl.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02s_launches_$wl.csv $B --workload $wl > gpurun_out/ncu_$wl.log 2>&1
  echo "launch list $wl rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cut -c1-300 gpurun_out/r02s_bench.json; cut -c1-400 gpurun_out/r02s_bench_reference.json
