#!/usr/bin/env python
"""What runs in one eager MCAN training step besides the library's GEMM / attention / LayerNorm kernels (one GPU).

    python tools/step_audit.py > gpurun_out/step_audit.json

Kineto (torch.profiler) records three eager steps; for the last one this prints
  * every CUDA kernel name with launches / total us (the same table as the ncu launch list, but at eager back-to-back timing),
  * for the kernels of interest (fills, fp32 SIMT GEMM, pooling head, copies): the CPU op (with input shapes) that launched them,
    so that a zero-fill or a cast can be traced to the line of functional.py / harness that asks for it.
There is no nsys on this image; this is the substitute.
"""
import collections
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openvivqa_b200 as ovqa  # noqa: E402
from openvivqa_b200 import functional as Fn, harness as H  # noqa: E402

WATCH = ("FillFunctor", "gemm_simt", "attr_pool", "direct_copy", "CUDAFunctor_add", "elementwise_kernel", "Memset", "Memcpy")


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ovqa.set_compute_dtype("bf16")
    H.bench_library_math().__enter__()
    wl = bench.MCAN
    torch.manual_seed(0)
    model = H.MCAN(H.mcan_cfg(wl["d_model"], wl["d_feature"], wl["layers"], wl["h"], wl["d_ff"], dropout=wl["dropout"], compute_dtype="bf16"),
                   H.StubVocab(wl["vocab"], wl["answers"])).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.98), fused=True)
    feats, toks, ans = (t.to(dev) for t in bench.synth_batches(1, wl, seed=1, feature_dtype=torch.bfloat16)[0])

    def step():
        Fn.dropout_stream.advance_on_device(dev)
        loss = F.nll_loss(H.bench_forward(model, H.Batch(region_features=feats, question_tokens=toks)).float(), ans)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    marks = []
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU], record_shapes=True,
                                with_stack=True) as prof:
        for _ in range(3):
            step()
            torch.cuda.synchronize()
    evs = list(prof.events())
    cuda = [e for e in evs if e.device_type == torch.autograd.DeviceType.CUDA and e.name]
    cuda.sort(key=lambda e: e.time_range.start)
    # the last step = the last third of the kernels (every step launches the same sequence)
    n = len(cuda) // 3
    last = cuda[2 * n:]
    by = collections.OrderedDict()
    for e in last:
        k = by.setdefault(e.name[:110], [0, 0.0])
        k[0] += 1
        k[1] += e.time_range.elapsed_us()
    table = sorted(([name, c, round(us, 1)] for name, (c, us) in by.items()), key=lambda r: -r[2])
    span = last[-1].time_range.end - last[0].time_range.start
    busy = sum(e.time_range.elapsed_us() for e in last)
    # who launched the watched kernels: correlate through the CPU op that encloses the launch (kineto links them by correlation id)
    cpu_ops = [e for e in evs if e.device_type == torch.autograd.DeviceType.CPU]
    who = collections.Counter()
    for op in cpu_ops:
        for k in getattr(op, "kernels", []) or []:
            if any(w in k.name for w in WATCH):
                shapes = getattr(op, "input_shapes", None)
                stack = [s for s in (getattr(op, "stack", None) or []) if "openvivqa_b200" in s or "bench.py" in s or "harness" in s][:2]
                who[(k.name[:60], op.name, json.dumps(shapes)[:120], " | ".join(s.split("/")[-1][:70] for s in stack))] += 1
    out = {"kernels_last_step": len(last), "span_us": round(span, 1), "busy_us": round(busy, 1), "table": table[:70],
           "watched": [[*k, c] for k, c in who.most_common(80)]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
