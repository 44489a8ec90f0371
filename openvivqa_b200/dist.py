"""Host-side multi-GPU plumbing of the path: one process per GPU, batch data-parallel.

The path shards by independent samples: training adds ONE collective (the gradient all-reduce, NCCL over NVLink,
bucketed and overlapped with backward by ``DistributedDataParallel``); inference (M4C decode) is a pure batch
partition with no data-path collective.  Everything here is backend-agnostic (``nccl`` on GPUs, ``gloo`` in the CPU
tests).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process per GPU); (0, 0, 1) when absent."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(n_samples: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of a global batch owned by ``rank``; sizes differ by at most one, every sample is
    owned exactly once (inference partitions a ragged last batch with this)."""
    base, rem = divmod(n_samples, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_max(value: float, device=None) -> float:
    """Max over ranks of a host scalar (timings are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def freeze_unused(model: torch.nn.Module) -> List[str]:
    """Parameters a model declares as never used in ``forward`` (``unused_parameters()``: the dead cross-attention
    branch of CrossModalityEncoderLayer, encoders.py:39-66) get ``requires_grad=False`` so that DDP's reducer does not
    wait for gradients that never arrive.  Returns their names."""
    unused = set()
    for m in model.modules():
        fn = getattr(m, "unused_parameters", None)
        if callable(fn):
            unused.update(id(p) for p in fn())
    names = []
    for n, p in model.named_parameters():
        if id(p) in unused:
            p.requires_grad_(False)
            names.append(n)
    return names


def data_parallel(model: torch.nn.Module, device_ids: Iterable[int] | None = None, bucket_cap_mb: int = 25, grad_dtype: str = "fp32",
                  sm_reserve: int = 0):
    """Wrap for data-parallel training: gradient all-reduce in ``bucket_cap_mb`` buckets launched from autograd hooks
    (overlaps the rest of backward); gradients are bucket views (no extra copy).  No-op for world size 1.

    ``grad_dtype="bf16"``: every bucket crosses NVLink as bf16 (cast, all-reduce, cast back and average — the stock
    ``bf16_compress_hook``): half the bytes, and half the time the collective's CTAs compete with the backward kernels for SMs and HBM
    ("bf16 grads halve it").  Parameters and optimiser state stay fp32.

    ``sm_reserve``: SMs the library's persistent kernels leave free DURING BACKWARD for the collective's CTAs (set when the gradient of
    the wrapped model's output arrives, restored when the backward pass ends — the grid sizes are baked into a captured CUDA graph).
    Without it a one-CTA-per-SM kernel launched while the all-reduce holds SMs runs its displaced CTAs as a second wave."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    freeze_unused(model)
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=list(device_ids) if device_ids is not None else None,
                                                    gradient_as_bucket_view=True, bucket_cap_mb=bucket_cap_mb)
    net._ovqa_grad_dtype = "fp32"
    if sm_reserve > 0 and next(model.parameters()).is_cuda:
        _reserve_sms_during_backward(net, int(sm_reserve))
    if grad_dtype == "bf16" and next(model.parameters()).is_cuda:
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        net.register_comm_hook(None, default_hooks.bf16_compress_hook)
        net._ovqa_grad_dtype = "bf16"
    return net


def _reserve_sms_during_backward(net: torch.nn.Module, sms: int) -> None:
    from . import ops

    def on_backward_start(_grad):
        prev = ops.set_sm_reserve(sms)
        torch.autograd.Variable._execution_engine.queue_callback(lambda: ops.set_sm_reserve(prev))

    def attach(_mod, _inp, out):
        for t in (out if isinstance(out, (tuple, list)) else (out,)):
            if isinstance(t, torch.Tensor) and t.requires_grad:
                t.register_hook(on_backward_start)
                break

    net.register_forward_hook(attach)


def grad_dtype_of(net) -> str:
    return getattr(net, "_ovqa_grad_dtype", "fp32")
