#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out; : > gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpus.txt
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench_n$N rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
S=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 5 --warmup 3 --workload m4c_decode > gpurun_out/bench_m4c_n$N.log 2> gpurun_out/bench_m4c_n$N.err; echo "bench_m4c_n$N rc=$? wall=$(( $(date +%s) - S ))s" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; cut -c1-500 gpurun_out/bench_n$N.log; tail -4 gpurun_out/bench_n$N.err | cut -c1-300; cut -c1-300 gpurun_out/bench_m4c_n$N.log
