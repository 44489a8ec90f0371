#!/bin/bash
# N = 2 data-parallel MCAN step with SMs reserved for the collective during backward
mkdir -p gpurun_out
for R in ${RESERVES:-0 8 16 32}; do
  OVQA_DDP_SM_RESERVE=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload mcan_train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/n2_reserve_$R.json 2> gpurun_out/n2_reserve_$R.err
  echo "reserve $R rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/n2_reserve_$R.json").read().strip().splitlines()[-1])
    print("reserve $R", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), d.get("step_breakdown",{}).get("allreduce"))
except Exception as e:
    print("parse failed", e)
PY
done
