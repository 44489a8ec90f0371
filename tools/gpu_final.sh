#!/bin/bash
# end-of-round validation on one GPU: full GPU suite, smoke(), default bench (both workloads), launch list, ncu of the top kernels
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
S=$(date +%s); timeout 1200 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
S=$(date +%s); timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
timeout 900 python bench.py --workload m4c_decode > gpurun_out/bench_m4c.log 2> gpurun_out/bench_m4c.err; echo "m4c rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 2 --no-e2e --no-graph > gpurun_out/ncu.log 2>&1; echo "ncu_list rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 120 -c 1 -f -o gpurun_out/gemm_fc1train_full python tools/microbench.py epi > gpurun_out/ncu_epi.log 2>&1; echo "ncu_fc1 rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel --launch-skip 52 -c 1 -f -o gpurun_out/gemm_fc2_full python tools/microbench.py gemm > gpurun_out/ncu_gemm.log 2>&1; echo "ncu_fc2 rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sdpa_fwd_tc5_kernel|sdpa_bwd_tc5_kernel" --launch-skip 8 -c 2 -f -o gpurun_out/tc5_full python tools/microbench.py sdpa > gpurun_out/ncu_tc5.log 2>&1; echo "ncu_tc5 rc=$?" >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layernorm_bwd" --launch-skip 5 -c 1 -f -o gpurun_out/ln_full python tools/microbench.py ln > gpurun_out/ncu_ln.log 2>&1; echo "ncu_ln rc=$?" >> gpurun_out/summary.txt
timeout 900 python tools/microbench.py gemm epi sdpa ln mha decode beam > gpurun_out/microbench.jsonl 2> gpurun_out/microbench.err; echo "microbench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.log; cut -c1-400 gpurun_out/bench_m4c.log; cut -c1-300 gpurun_out/bench_ref.log
