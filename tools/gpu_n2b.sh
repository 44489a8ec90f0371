#!/bin/bash
# 2 GPUs: full GPU test suite, then the driver's launch line for N=2 (ours, captured DDP graph), exit must be clean and prompt
mkdir -p gpurun_out; : > gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpus.txt
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --workload m4c_decode > gpurun_out/bench_m4c_n2.log 2> gpurun_out/bench_m4c_n2.err; echo "bench_m4c_n2 rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2> gpurun_out/bench_ref_n2.err; echo "bench_ref_n2 rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log; cut -c1-600 gpurun_out/bench_n2.log; tail -4 gpurun_out/bench_n2.err; cut -c1-400 gpurun_out/bench_m4c_n2.log; cut -c1-600 gpurun_out/bench_ref_n2.log
