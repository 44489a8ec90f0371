"""GPU parity of the cluster-resident LSTM recurrence (csrc/lstm.cu; SURVEY §8f rank 4, text_embeddings.py:234,244), through the C ABI.

* kernel level: ``ovqa_lstm_fwd`` / ``ovqa_lstm_bwd`` against the step-by-step restatement in tests/cpu_ops.py, which rounds the same
  operands to bf16 (h_{t-1}, W_hh, the saved gates) — what is left is accumulation order and the exp-based activations;
* function level: ``Fn.lstm`` (input projection GEMM + recurrence + the gradient GEMMs) against autograd of the fp32 oracle
  (oracle/restated.py::lstm_forward) within the bf16 path's tolerance."""
import pytest
import torch

import cpu_ops as E
from openvivqa_b200 import functional as Fn
from openvivqa_b200 import ops
from oracle import restated as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _inputs(B, L, H, seed):
    g = torch.Generator().manual_seed(seed)
    xg = torch.randn((B, L, 4 * H), generator=g) * 1.5
    w = (torch.rand((4 * H, H), generator=g) * 2 - 1) / H ** 0.5 * 2.0        # a little hotter than nn.LSTM's default: the recurrence matters
    b = (torch.rand((4 * H,), generator=g) * 2 - 1) / H ** 0.5
    return xg, w.to(torch.bfloat16), b


@pytest.mark.parametrize("B,L,H", [(5, 7, 128), (40, 30, 256), (32, 30, 512), (70, 13, 512), (1, 1, 512), (256, 30, 512)])
def test_lstm_recurrence_kernels_against_the_restatement(dev, B, L, H):
    xg, w16, b = _inputs(B, L, H, 11 + B)
    y_r, y16_r, hp_r, g_r, c_r = E.lstm_fwd(xg, w16, b, save=True)
    y, y16, hp, gt, c = ops.lstm_fwd(xg.to(dev), w16.to(dev), b.to(dev), save=True)
    # h feeds back through a bf16 rounding: a last-bit difference in step t can flip a rounding in step t+1, so the bound is a few bf16 ulps of h
    assert float((y.cpu() - y_r).abs().max()) <= 2e-2, float((y.cpu() - y_r).abs().max())
    assert _rel(y, y_r) <= 4e-3, _rel(y, y_r)
    assert _rel(c, c_r) <= 4e-3, _rel(c, c_r)
    assert _rel(gt, g_r) <= 6e-3, _rel(gt, g_r)
    assert torch.equal(y16.cpu(), y.cpu().to(torch.bfloat16))
    assert torch.equal(hp[:, 0].cpu(), torch.zeros(B, H, dtype=torch.bfloat16))
    if L > 1:
        assert torch.equal(hp[:, 1:].cpu(), y16[:, :-1].cpu())
    # inference form: nothing saved, same y
    y2 = ops.lstm_fwd(xg.to(dev), w16.to(dev), b.to(dev), save=False)[0]
    assert torch.equal(y2, y)
    # backward on the tensors the forward kernel saved (so that the comparison does not compound the forward's rounding)
    dy = torch.randn((B, L, H), generator=torch.Generator().manual_seed(5))
    dg_r = E.lstm_bwd(dy, w16, gt.cpu(), c.cpu())
    dg = ops.lstm_bwd(dy.to(dev), w16.to(dev), gt, c)
    assert _rel(dg, dg_r) <= 1e-2, _rel(dg, dg_r)
    # the last step has no recurrent term: exact up to the activations
    assert _rel(dg[:, -1], dg_r[:, -1]) <= 5e-3


def test_lstm_rejects_unsupported_hidden_sizes(dev):
    assert ops.lstm_supported(512) and ops.lstm_supported(128) and not ops.lstm_supported(96)
    xg = torch.zeros((2, 3, 4 * 96), device=dev)
    with pytest.raises(Exception):
        ops.lstm_fwd(xg, torch.zeros((4 * 96, 96), device=dev, dtype=torch.bfloat16), None)


@pytest.mark.parametrize("B,L,In,H", [(9, 12, 64, 128), (64, 30, 512, 512)])
def test_lstm_function_against_oracle_autograd(dev, B, L, In, H):
    torch.manual_seed(3)
    ref = torch.nn.LSTM(In, H, batch_first=True)
    x = torch.randn(B, L, In)
    xr = x.clone().requires_grad_(True)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in ref.named_parameters()}
    yr = R.lstm_forward(p, xr)
    w = torch.randn(yr.shape, generator=torch.Generator().manual_seed(7))
    (yr * w).sum().backward()

    m = torch.nn.LSTM(In, H, batch_first=True).to(dev)
    m.load_state_dict(ref.state_dict())
    xi = x.clone().to(dev).requires_grad_(True)
    assert Fn.lstm_supported(m, xi, "bf16") and not Fn.lstm_supported(m, xi, "fp32")
    y = Fn.lstm(m, xi)
    (y * w.to(dev)).sum().backward()
    assert float((y.cpu() - yr).abs().max()) <= 2e-2, float((y.cpu() - yr).abs().max())
    assert _rel(xi.grad, xr.grad) <= 3e-2, _rel(xi.grad, xr.grad)
    for k, v in m.named_parameters():
        assert v.grad is not None, k
        assert _rel(v.grad, p[k].grad) <= 3e-2, (k, _rel(v.grad, p[k].grad))
    # eval / no_grad: same output, nothing saved
    with torch.no_grad():
        y0 = Fn.lstm(m, xi)
    assert torch.equal(y0, y.detach())
