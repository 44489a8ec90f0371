"""Deterministic synthetic parameters and inputs shared by the golden generator, the tests and bench.py.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Weights are drawn on the CPU generator from a seed that
depends only on (seed, parameter name), so the GPU box regenerates bit-identical tensors from the
``(name, shape)`` lists stored in ``tests/golden/*.pt`` without needing ``/root/reference``.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def synth_params(spec: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Xavier-uniform matrices, small non-zero biases, LayerNorm gains near 1 (so every affine path is exercised)."""
    out = {}
    for name, shape in spec:
        shape = tuple(shape)
        g = _gen(seed, name)
        leaf = name.rsplit(".", 2)
        owner = leaf[-2].lower() if len(leaf) >= 2 else ""
        is_ln = any(t in owner for t in ("layer_norm", "layernorm"))
        if len(shape) >= 2:
            fan_out, fan_in = shape[0], math.prod(shape[1:])
            a = math.sqrt(6.0 / (fan_in + fan_out))
            if "embedding" in owner:                       # lookup tables (nn.Embedding)
                t = torch.randn(shape, generator=g) * 0.5
            elif owner == "lstm":                          # nn.LSTM default: U(+-1/sqrt(H))
                t = (torch.rand(shape, generator=g) * 2 - 1) / math.sqrt(shape[1])
            else:
                t = (torch.rand(shape, generator=g) * 2 - 1) * a
        elif is_ln and leaf[-1] == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_ln:
            t = 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        out[name] = t.float()
    return out


def spec_of(state_dict) -> list:
    return [(k, tuple(v.shape)) for k, v in state_dict.items() if v.dtype.is_floating_point]


def padded_features(B: int, L: int, D: int, seed: int, min_valid: int = 1, relu_like: bool = True,
                    all_pad_first: bool = False) -> torch.Tensor:
    """|N(0,1)| features with a random number of trailing all-zero (= padding) rows per sample."""
    g = _gen(seed, f"feat{B}x{L}x{D}")
    x = torch.randn((B, L, D), generator=g)
    if relu_like:
        x = x.abs()
    n = torch.randint(min_valid, L + 1, (B,), generator=g)
    if all_pad_first:
        n[0] = 0
    x[torch.arange(L).unsqueeze(0) >= n.unsqueeze(1)] = 0.0
    return x


def padded_tokens(B: int, L: int, vocab: int, seed: int, min_len: int = 1) -> torch.Tensor:
    g = _gen(seed, f"tok{B}x{L}x{vocab}")
    t = torch.randint(4, vocab, (B, L), generator=g)
    n = torch.randint(min_len, L + 1, (B,), generator=g)
    t[torch.arange(L).unsqueeze(0) >= n.unsqueeze(1)] = 0
    return t


def boxes(B: int, N: int, seed: int, scale: float = 600.0) -> torch.Tensor:
    g = _gen(seed, f"box{B}x{N}")
    a = torch.rand((B, N, 2, 2), generator=g) * scale
    lo, hi = a.min(dim=2).values, a.max(dim=2).values
    return torch.cat([lo, hi], dim=-1)  # x_min, y_min, x_max, y_max
