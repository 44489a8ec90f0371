"""Stand-alone diagnosis of the tcgen05 GEMM against its CUDA-core twin (run on the GPU box; writes a report).
Each configuration runs in THIS process; a device trap aborts the script, so the riskiest forms come last."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvivqa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
report = []


def run(tag, m, n, k, a_op="N", b_op="N", out_dtype=torch.float32, split_k=1, bias=False, gelu=False, structured=False):
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    if structured:   # small integers: exact in bf16 and in fp32 accumulation -> any mismatch is a layout bug
        A = torch.randint(-3, 4, (m, k), generator=g).float()
        B = torch.randint(-3, 4, (n, k), generator=g).float()
    else:
        A, B = torch.randn((m, k), generator=g), torch.randn((n, k), generator=g)
    a = (A if a_op == "N" else A.t().contiguous()).to(dev).bfloat16()
    b = (B if b_op == "N" else B.t().contiguous()).to(dev).bfloat16()
    bs = torch.randn(n, generator=g).to(dev) if bias else None
    ref = ops.gemm_bf16_simt(a, b, a_op=a_op, b_op=b_op, out_dtype=torch.float32, bias=bs, gelu=gelu)
    out = ops.gemm(a, b, a_op=a_op, b_op=b_op, out_dtype=out_dtype, bias=bs, gelu=gelu, split_k=split_k)
    torch.cuda.synchronize()
    err = (out.float() - ref).abs()
    scale = ref.abs().max().item() + 1e-9
    tol = (2e-2 if out_dtype == torch.bfloat16 else 1e-3) * scale
    bad = err > tol
    rec = dict(tag=tag, m=m, n=n, k=k, a_op=a_op, b_op=b_op, split_k=split_k, max_err=err.max().item(), scale=scale,
               frac_bad=bad.float().mean().item(), ok=bool(not bad.any()))
    if bad.any():
        rows = bad.any(1).nonzero().flatten()[:16].tolist()
        cols = bad.any(0).nonzero().flatten()[:16].tolist()
        rec.update(bad_rows=rows, bad_cols=cols, sample_out=out.float()[:2, :8].tolist(), sample_ref=ref[:2, :8].tolist(),
                   bad_by_rowblock=[bad[i:i + 32].float().mean().item() for i in range(0, min(m, 256), 32)],
                   bad_by_colblock=[bad[:, j:j + 32].float().mean().item() for j in range(0, min(n, 512), 32)])
    report.append(rec)
    print(json.dumps(rec), flush=True)
    return rec["ok"]


cases = [
    ("nn_tiny_structured", 128, 128, 64, "N", "N", dict(structured=True)),
    ("nn_k128_structured", 128, 128, 128, "N", "N", dict(structured=True)),
    ("nn_basic", 256, 256, 512, "N", "N", {}),
    ("nn_ragged", 300, 200, 328, "N", "N", {}),
    ("nn_bias_gelu_bf16out", 512, 2048, 512, "N", "N", dict(bias=True, gelu=True, out_dtype=torch.bfloat16)),
    ("nn_many_tiles", 4096, 512, 512, "N", "N", {}),
    ("nn_wide256", 40000, 2048, 512, "N", "N", dict(out_dtype=torch.bfloat16)),
    ("nt_structured", 128, 128, 64, "N", "T", dict(structured=True)),
    ("nt_dgrad", 600, 512, 2048, "N", "T", {}),
    ("tt_structured", 128, 128, 64, "T", "T", dict(structured=True)),
    ("tt_wgrad", 512, 2048, 3000, "T", "T", {}),
    ("tt_wgrad_splitk", 512, 512, 6000, "T", "T", dict(split_k=8)),
    ("tn_structured", 128, 128, 64, "T", "N", dict(structured=True)),
]
ok_all = True
for tag, m, n, k, ao, bo, kw in cases:
    try:
        ok_all &= run(tag, m, n, k, ao, bo, **kw)
    except Exception as e:  # noqa: BLE001
        report.append(dict(tag=tag, error=repr(e)))
        print("EXC", tag, repr(e), flush=True)
        ok_all = False
        break
os.makedirs("gpurun_out", exist_ok=True)
json.dump(report, open("gpurun_out/diag_tcgen05.json", "w"), indent=1)
print("ALL OK" if ok_all else "FAILURES")
sys.exit(0 if ok_all else 1)
