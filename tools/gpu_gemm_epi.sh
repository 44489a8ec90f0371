#!/bin/bash
# GEMM epilogue iteration: the tcgen05 GEMM tests, then the fused-epilogue microbench with the TMA-fed epilogue on and off
mkdir -p gpurun_out; : > gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gemm_tcgen05.py -x -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc=$?" >> gpurun_out/summary.txt
tail -15 gpurun_out/pytest_gemm.log | cut -c1-300
timeout 300 python tools/microbench.py epi gemm > gpurun_out/mb_epi_tma.jsonl 2> gpurun_out/mb_epi_tma.err; echo "mb tma rc=$?" >> gpurun_out/summary.txt
OVQA_GEMM_TMA_EPILOGUE=0 timeout 300 python tools/microbench.py epi gemm > gpurun_out/mb_epi_old.jsonl 2> gpurun_out/mb_epi_old.err; echo "mb old rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
python - <<'PY'
import json
def rd(p):
    d={}
    for l in open(p):
        try: r=json.loads(l)
        except Exception: continue
        if r.get("kind") in ("gemm_epi","gemm"): d[r["tag"]]=r
    return d
a,b=rd("gpurun_out/mb_epi_tma.jsonl"),rd("gpurun_out/mb_epi_old.jsonl")
for t in a:
    r=a[t]; o=b.get(t,{})
    print(f"{t:32s} tma {r['us']:7.1f} us  old {o.get('us',0):7.1f} us  floor hbm {r.get('hbm_floor_us',0):5.1f} tensor {r.get('tensor_floor_us',0):5.1f} cublas {r.get('cublas_us',0):6.1f}")
PY
tail -3 gpurun_out/mb_epi_tma.err
