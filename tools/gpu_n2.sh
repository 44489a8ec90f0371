#!/bin/bash
# 2-GPU: DDP step captured in a CUDA graph (NCCL all-reduce inside the capture), eager fallback, and M4C decode partition
mkdir -p gpurun_out; : > gpurun_out/summary.txt
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -p no:cacheprovider -x -k "layernorm" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$?" >> gpurun_out/summary.txt
OVQA_DDP_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_eager.log 2> gpurun_out/bench_n2_eager.err; echo "bench_n2_eager rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --workload m4c_decode > gpurun_out/bench_m4c_n2.log 2> gpurun_out/bench_m4c_n2.err; echo "bench_m4c_n2 rc=$?" >> gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench_n1 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_n2.log; tail -8 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2_eager.log; tail -3 gpurun_out/bench_n2_eager.err; cat gpurun_out/bench_m4c_n2.log; tail -3 gpurun_out/bench_m4c_n2.err; cat gpurun_out/bench.log
