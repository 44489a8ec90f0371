#!/usr/bin/env python
"""Timeline of one data-parallel MCAN training step (torchrun, N >= 2): where the gradient all-reduce runs relative to backward.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_timeline.py > gpurun_out/ddp_timeline.json

Kineto (torch.profiler) records every CUDA kernel of three eager steps; rank 0 prints, for the last step: the span of the step, every NCCL
kernel (start offset, duration), the compute kernels that overlap an NCCL kernel and how much longer they run than the same kernels of a
step region without NCCL, and the gap between the last compute kernel of backward and the optimiser's first kernel (the exposed tail).
There is no nsys on this image; this is the substitute.
"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import openvivqa_b200 as ovqa  # noqa: E402
from openvivqa_b200 import dist as D, functional as Fn, harness as H  # noqa: E402


def main():
    rank, local, world = D.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ovqa.set_compute_dtype("bf16")
    H.bench_library_math().__enter__()
    wl = bench.MCAN
    torch.manual_seed(0)
    model = H.MCAN(H.mcan_cfg(wl["d_model"], wl["d_feature"], wl["layers"], wl["h"], wl["d_ff"], dropout=wl["dropout"], compute_dtype="bf16"),
                   H.StubVocab(wl["vocab"], wl["answers"])).to(dev).train()
    net = D.data_parallel(model, device_ids=[local], bucket_cap_mb=int(os.environ.get("OVQA_DDP_BUCKET_MB", "25")),
                          grad_dtype=os.environ.get("OVQA_DDP_GRAD_DTYPE", "bf16"))
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.98), fused=True)
    feats, toks, ans = (t.to(dev) for t in bench.synth_batches(1, wl, seed=1 + rank, feature_dtype=torch.bfloat16)[0])

    def step():
        Fn.dropout_stream.advance_on_device(dev)
        loss = F.nll_loss(H.bench_forward(net, H.Batch(region_features=feats, question_tokens=toks)).float(), ans)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(12):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
            torch.cuda.synchronize()
            dist.barrier()
    if rank == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name and not e.name.startswith("Memcpy") and not e.name.startswith("Memset")]
        ev.sort(key=lambda e: e.time_range.start)
        # split into steps at the advance_u64 launches (first kernel of every step)
        starts = [i for i, e in enumerate(ev) if "advance_u64" in e.name]
        last = ev[starts[-1]:]
        t0 = last[0].time_range.start
        is_nccl = lambda e: "nccl" in e.name.lower()
        is_adam = lambda e: "FusedAdam" in e.name or "multi_tensor_apply" in e.name
        nccl = [e for e in last if is_nccl(e) and e.time_range.elapsed_us() > 5]
        adam = [e for e in last if is_adam(e)]
        comp = [e for e in last if not is_nccl(e)]
        first_adam = min(e.time_range.start for e in adam) if adam else None
        bwd_compute_before_adam = [e for e in comp if first_adam is None or e.time_range.end <= first_adam]
        last_compute_end = max(e.time_range.end for e in bwd_compute_before_adam if not is_adam(e))
        # compute kernels overlapping any NCCL kernel, grouped by name: mean duration with / without overlap
        def overlaps(e):
            return any(e.time_range.start < n.time_range.end and n.time_range.start < e.time_range.end for n in nccl)
        by = {}
        for e in comp:
            k = e.name[:60]
            d = by.setdefault(k, [[], []])
            d[0 if overlaps(e) else 1].append(e.time_range.elapsed_us())
        slow = []
        for k, (a, b) in by.items():
            if len(a) >= 1 and len(b) >= 1:
                slow.append(dict(kernel=k, n_overlapped=len(a), n_clear=len(b), us_overlapped=round(sum(a) / len(a), 1), us_clear=round(sum(b) / len(b), 1)))
        slow.sort(key=lambda r: -(r["us_overlapped"] - r["us_clear"]) * r["n_overlapped"])
        out = dict(world=world, step_span_us=round(last[-1].time_range.end - t0, 1),
                   compute_kernel_us=round(sum(e.time_range.elapsed_us() for e in comp), 1),
                   nccl=[dict(name=e.name[:50], start_us=round(e.time_range.start - t0, 1), us=round(e.time_range.elapsed_us(), 1)) for e in nccl],
                   nccl_total_us=round(sum(e.time_range.elapsed_us() for e in nccl), 1),
                   last_backward_kernel_end_us=round(last_compute_end - t0, 1),
                   first_adam_start_us=round(first_adam - t0, 1) if first_adam else None,
                   exposed_tail_us=round(first_adam - last_compute_end, 1) if first_adam else None,
                   extra_us_of_overlapped_kernels=round(sum((r["us_overlapped"] - r["us_clear"]) * r["n_overlapped"] for r in slow), 1),
                   most_slowed=slow[:12])
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
