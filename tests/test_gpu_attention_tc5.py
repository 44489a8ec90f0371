"""GPU: the tcgen05 / TMEM / TMA attention core (attention_tc5.cu: head dim 64, Lq, Lk <= 128, per-key mask) against an fp32
torch restatement of attentions.py:53-57 on the same bf16-rounded inputs.  Shapes chosen so that persistent CTAs wrap their
load / TMEM / P-buffer stages many times (items >> #SM), tiles are full (128), minimal (1), and ragged (non-multiples of 16)."""
import math

import pytest
import torch

from openvivqa_b200 import ops

pytestmark = pytest.mark.gpu

SHAPES = [  # (B, h, Lq, Lk)
    (256, 8, 100, 100), (64, 8, 128, 128), (130, 8, 100, 30), (300, 2, 30, 100), (77, 8, 30, 30), (9, 8, 1, 1), (5, 3, 17, 16), (5, 3, 16, 17),
    (3, 8, 113, 127), (2, 1, 128, 1),
    # wide heads / rows longer than one 128-row tile (sdpa_fwd_tc5w_kernel: M4C's 96-wide heads over 182 positions, 128-wide heads)
    (128, 8, 182, 182, 96), (16, 8, 62, 182, 96), (5, 2, 100, 256, 96), (3, 4, 300, 200, 128), (2, 3, 17, 33, 128), (40, 8, 182, 182, 96)]


def _ref(q, k, v, mask, h, d):
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    qh = q.view(B, Lq, h, d).permute(0, 2, 1, 3)
    kh = k.view(B, Lk, h, d).permute(0, 2, 3, 1)
    vh = v.view(B, Lk, h, d).permute(0, 2, 1, 3)
    att = torch.matmul(qh, kh) / math.sqrt(d)
    if mask is not None:
        att = att + mask
    att = torch.softmax(att, dim=-1)
    return torch.matmul(att, vh).permute(0, 2, 1, 3).reshape(B, Lq, h * d), att


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tc5_core_forward_backward(shape, masked, monkeypatch):
    monkeypatch.setenv("OVQA_SDPA_TC5", "1")            # also the shapes the dispatcher would give to the warp-tile core
    monkeypatch.setenv("OVQA_SDPA_TC5_BWD", "1")
    B, h, Lq, Lk = shape[:4]
    d = shape[4] if len(shape) > 4 else 64
    if masked and Lk == 1:
        pytest.skip("single fully-masked key: the exact gradients are 0 and only the -1e5 rounding noise of exp(s - lse) is left")
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + Lq * 10 + Lk)
    fused = Lq == Lk
    if fused:
        buf = torch.randn((B, Lq, 3 * h * d), generator=g).to(dev).bfloat16()
        q, k, v = buf[..., :h * d], buf[..., h * d:2 * h * d], buf[..., 2 * h * d:]
    else:
        q = torch.randn((B, Lq, h * d), generator=g).to(dev).bfloat16()
        k = torch.randn((B, Lk, h * d), generator=g).to(dev).bfloat16()
        v = torch.randn((B, Lk, h * d), generator=g).to(dev).bfloat16()
    dout = torch.randn((B, Lq, h * d), generator=g).to(dev).bfloat16()
    mask = None
    if masked:
        valid = torch.randint(max(1, Lk // 3), Lk + 1, (B,), generator=g)
        mask = ((torch.arange(Lk)[None, :] >= valid[:, None]).float() * -10e4).view(B, 1, 1, Lk).to(dev)
        mask[0] = -10e4                                                       # one sample with every key masked
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref, att = _ref(qr, kr, vr, mask, h, d)
    ref.backward(dout.float())
    out, lse, _ = ops.sdpa_fwd(q, k, v, h=h, d_k=d, d_v=d, mask=mask, need_lse=True)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    assert (out.float() - ref).abs().max().item() <= 2e-2 * max(1.0, ref.abs().max().item())
    # lse against the reference scores
    s = torch.matmul(q.float().view(B, Lq, h, d).permute(0, 2, 1, 3), k.float().view(B, Lk, h, d).permute(0, 2, 3, 1)) / math.sqrt(d)
    if mask is not None:
        s = s + mask
    want_lse = torch.logsumexp(s, dim=-1)
    assert (lse - want_lse).abs().max().item() <= 2e-2 + 1e-6 * want_lse.abs().max().item()
    if fused:
        dbuf = torch.full_like(buf, float("nan"))
        dq, dk, dv = dbuf[..., :h * d], dbuf[..., h * d:2 * h * d], dbuf[..., 2 * h * d:]
    else:
        dq, dk, dv = (torch.full_like(t, float("nan")) for t in (q, k, v))
    ops.sdpa_bwd(q, k, v, out, dout, lse, dq, dk, dv, h=h, d_k=d, d_v=d, mask=mask)
    torch.cuda.synchronize()
    for name, got, want in (("dq", dq, qr.grad), ("dk", dk, kr.grad), ("dv", dv, vr.grad)):
        assert torch.isfinite(got.float()).all(), name
        err = (got.float() - want).abs().max().item()
        # a sample with EVERY key masked has scores ~ -1e5, where one fp32 ulp (0.0078) is a 0.8 % error of exp(s - lse): the
        # reference's own softmax carries the same noise; it only shows where the exact gradient is ~0 (Lk == 1)
        floor = 0.5 * dout.float().abs().max().item() if masked else 1e-3
        assert err <= 3e-2 * max(want.abs().max().item(), floor), f"{name}: {err:.3e} vs scale {want.abs().max().item():.3e}"


@pytest.mark.parametrize("Lq_rows", [None, 62])
def test_prefix_lm_mask_structured_vs_dense(Lq_rows):
    """M4C's (B,1,L,L) mask (mmf_m4c.py:310-343) given as per-key row + causal tail (ovqa_sdpa_fwd_prefix_lm) against the dense tensor
    through the generic entry point, both against an fp32 torch reference; also the row-pruned last layer (queries = the last 62 rows)."""
    from openvivqa_b200.harness.m4c import MMF_M4C
    dev = torch.device("cuda:0")
    B, h, L, T, d = 9, 8, 182, 12, 96
    g = torch.Generator().manual_seed(3)
    valid = [torch.randint(5, n + 1, (B,), generator=g) for n in (20, 100, 50)]
    pads = [((torch.arange(n)[None, :] >= v[:, None]).float() * -10e4).view(B, 1, 1, n).to(dev) for n, v in zip((20, 100, 50), valid)]
    dense = MMF_M4C.mmt_mask(pads[0], pads[1], pads[2], T)
    assert getattr(dense, "_ovqa_prefix_lm", None) is not None and dense.shape == (B, 1, L, L)
    x = torch.randn((B, L, 3 * h * d), generator=g).to(dev).bfloat16()
    q, k, v = x[..., :h * d], x[..., h * d:2 * h * d], x[..., 2 * h * d:]
    mask = dense
    if Lq_rows is not None:
        q = q[:, -Lq_rows:].contiguous()
        mask = dense[:, :, -Lq_rows:]
        mask._ovqa_prefix_lm = dense._ovqa_prefix_lm
    ref, _ = _ref(q.float(), k.float(), v.float(), mask, h, d)
    out_s, lse_s, _ = ops.sdpa_fwd(q, k, v, h=h, d_k=d, d_v=d, mask=mask, need_lse=True)                 # structured (hint attached)
    plain = mask.clone()                                                                                  # same values, no hint -> dense path
    out_d, lse_d, _ = ops.sdpa_fwd(q, k, v, h=h, d_k=d, d_v=d, mask=plain, need_lse=True)
    torch.cuda.synchronize()
    scale = max(1.0, ref.abs().max().item())
    assert (out_s.float() - ref).abs().max().item() <= 2e-2 * scale
    assert (out_d.float() - ref).abs().max().item() <= 2e-2 * scale
    assert (lse_s - lse_d).abs().max().item() <= 2e-2
