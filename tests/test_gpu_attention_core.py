"""GPU: the bf16 tensor-core attention core (forward + backward, attention_mma.cu) against an fp32 torch restatement
of attentions.py:53-57 on the same bf16-rounded inputs, over the shape classes of SURVEY.md §8a and every mask rank."""
import math

import pytest
import torch

from openvivqa_b200 import ops

pytestmark = pytest.mark.gpu

SHAPES = [  # (B, h, Lq, Lk, d)
    (3, 8, 30, 30, 64), (2, 8, 100, 100, 64), (2, 8, 100, 30, 64), (2, 8, 30, 100, 64), (2, 8, 30, 600, 64),
    (2, 8, 182, 182, 96), (3, 4, 17, 45, 32), (2, 1, 36, 36, 64), (1, 8, 1, 50, 64)]


def _mask(kind, B, Lq, Lk, dev, g):
    if kind == "none":
        return None
    valid = torch.randint(max(1, Lk // 3), Lk + 1, (B,), generator=g)
    pad = (torch.arange(Lk)[None, :] >= valid[:, None]).float() * -10e4          # models/utils.py:56
    if kind == "padding":
        return pad.view(B, 1, 1, Lk).to(dev)
    if kind == "all_masked_row":                                                  # a sample with zero valid keys
        pad[0] = -10e4
        return pad.view(B, 1, 1, Lk).to(dev)
    if kind == "full":                                                            # (B,1,Lq,Lk): padding OR causal-like band
        band = (torch.arange(Lk)[None, :] > (torch.arange(Lq)[:, None] * Lk) // max(Lq, 1)).float() * -10e4
        return torch.minimum(pad.view(B, 1, 1, Lk), band.view(1, 1, Lq, Lk)).contiguous().to(dev)
    if kind == "shared":                                                          # (1,1,Lq,Lk) broadcast over the batch
        return ((torch.arange(Lk)[None, :] > (torch.arange(Lq)[:, None] * Lk) // max(Lq, 1)).float() * -10e4).view(1, 1, Lq, Lk).to(dev)
    raise AssertionError(kind)


def _reference(q, k, v, mask, h, d):
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    qh = q.view(B, Lq, h, d).permute(0, 2, 1, 3)
    kh = k.view(B, Lk, h, d).permute(0, 2, 3, 1)
    vh = v.view(B, Lk, h, d).permute(0, 2, 1, 3)
    att = torch.matmul(qh, kh) / math.sqrt(d)
    if mask is not None:
        att = att + mask
    att = torch.softmax(att, dim=-1)
    return torch.matmul(att, vh).permute(0, 2, 1, 3).reshape(B, Lq, h * d)


@pytest.mark.parametrize("kind", ["none", "padding", "full", "shared", "all_masked_row"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_bf16_core_forward_backward(shape, kind):
    B, h, Lq, Lk, d = shape
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(B * 1000 + Lq * 10 + Lk)
    fused = Lq == Lk                                       # self-attention layout: one (B, L, 3*h*d) buffer, strided views
    if fused:
        buf = torch.randn((B, Lq, 3 * h * d), generator=g).to(dev).bfloat16()
        q, k, v = buf[..., :h * d], buf[..., h * d:2 * h * d], buf[..., 2 * h * d:]
    else:
        q = torch.randn((B, Lq, h * d), generator=g).to(dev).bfloat16()
        k = torch.randn((B, Lk, h * d), generator=g).to(dev).bfloat16()
        v = torch.randn((B, Lk, h * d), generator=g).to(dev).bfloat16()
    dout = torch.randn((B, Lq, h * d), generator=g).to(dev).bfloat16()
    mask = _mask(kind, B, Lq, Lk, dev, g)

    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref = _reference(qr, kr, vr, mask, h, d)
    ref.backward(dout.float())

    out, lse, _ = ops.sdpa_fwd(q, k, v, h=h, d_k=d, d_v=d, mask=mask, need_lse=True)
    assert (out.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    if fused:
        dbuf = torch.full_like(buf, float("nan"))
        dq, dk, dv = dbuf[..., :h * d], dbuf[..., h * d:2 * h * d], dbuf[..., 2 * h * d:]
    else:
        dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    ops.sdpa_bwd(q, k, v, out, dout, lse, dq, dk, dv, h=h, d_k=d, d_v=d, mask=mask)
    for name, got, want in (("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
        assert torch.isfinite(got.float()).all(), name
        err = (got.float() - want).abs().max().item()
        assert err <= 3e-2 * max(want.abs().max().item(), 1e-3), f"{name}: {err:.3e} vs scale {want.abs().max().item():.3e}"
