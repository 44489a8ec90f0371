#!/usr/bin/env python
"""Timing probe for the LSTM recurrence kernels (B x L x H = 256 x 30 x 512 by default): us per launch, forward and backward."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openvivqa_b200 import ops

def main():
    B, L, H = (int(a) for a in (sys.argv[1:4] or (256, 30, 512)))
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    xg = torch.randn((B, L, 4 * H), generator=g).to(dev)
    w = ((torch.rand((4 * H, H), generator=g) * 2 - 1) / H ** 0.5).to(torch.bfloat16).to(dev)
    b = torch.zeros(4 * H, device=dev)
    dy = torch.randn((B, L, H), generator=g).to(dev)
    y, y16, hp, gt, c = ops.lstm_fwd(xg, w, b, save=True)
    def t(fn, n=20):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / n
    print(json.dumps(dict(dbg=os.environ.get("OVQA_LSTM_DBG", "0"), B=B, L=L, H=H, fwd_us=round(t(lambda: ops.lstm_fwd(xg, w, b, save=True)), 1),
                          fwd_nosave_us=round(t(lambda: ops.lstm_fwd(xg, w, b, save=False)), 1),
                          bwd_us=round(t(lambda: ops.lstm_bwd(dy, w, gt, c)), 1))))
main()
