#!/usr/bin/env python
"""A handful of the training step's fused-epilogue GEMMs, launched twice each (for an `ncu -k regex:gemm_tcgen05 --launch-skip ...` capture).

    python tools/ncu_gemm_cases.py [case ...]     cases: fc_o fc1 plain dpre wgrad
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvivqa_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = 25600
which = sys.argv[1:] or ["fc_o", "fc1", "plain"]
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
for w in which:
    if w == "fc_o":
        a, b = torch.randn((M, 512), device=dev).bfloat16(), torch.randn((512, 512), device=dev).bfloat16()
        kw = dict(bias=torch.randn(512, device=dev), dropout_p=0.1, seed=3, offset=7, residual=torch.randn((M, 512), device=dev), out=torch.empty((M, 512), device=dev))
    elif w == "fc1":
        a, b = torch.randn((M, 512), device=dev).bfloat16(), torch.randn((2048, 512), device=dev).bfloat16()
        kw = dict(bias=torch.randn(2048, device=dev), dropout_p=0.1, seed=3, offset=7, gelu=True, save_dgelu=True,
                  aux_out=torch.empty((M, 2048), device=dev, dtype=torch.bfloat16), out=torch.empty((M, 2048), device=dev, dtype=torch.bfloat16))
    elif w == "plain":
        a, b = torch.randn((M, 512), device=dev).bfloat16(), torch.randn((512, 512), device=dev).bfloat16()
        kw = dict(out=torch.empty((M, 512), device=dev, dtype=torch.bfloat16))
    elif w == "dpre":
        a, b = torch.randn((M, 512), device=dev).bfloat16(), torch.randn((512, 2048), device=dev).bfloat16()
        kw = dict(b_op="T", mul_aux=torch.randn((M, 2048), device=dev).bfloat16(), out=torch.empty((M, 2048), device=dev, dtype=torch.bfloat16))
    elif w == "wgrad":
        a, b = torch.randn((M, 512), device=dev).bfloat16(), torch.randn((M, 512), device=dev).bfloat16()
        kw = dict(a_op="T", b_op="T", split_k=0, out=torch.zeros((512, 512), device=dev))
    for _ in range(2):
        flush.zero_()
        ops.gemm(a, b, **kw)
    torch.cuda.synchronize()
print("ok")
