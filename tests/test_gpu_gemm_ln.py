"""GPU parity of the fused projection + bias + dropout + residual + LayerNorm kernel (csrc/gemm_ln.cu; attentions.py:58,331 and
positionwise_feed_forward.py:25-26) against the two-kernel form it replaces (ovqa_gemm with the same epilogue, then ovqa_layernorm_fwd)
and against plain PyTorch fp32 on the bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from openvivqa_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _case(M, K, seed, dev):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn((M, K), generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn((512, K), generator=g) / K ** 0.5).to(torch.bfloat16).to(dev)
    bias = torch.randn((512,), generator=g).to(dev)
    res = torch.randn((M, 512), generator=g).to(dev) * 1.5 + 0.3
    gamma = (1.0 + 0.2 * torch.randn((512,), generator=g)).to(dev)
    beta = (0.1 * torch.randn((512,), generator=g)).to(dev)
    return a, w, bias, res, gamma, beta


@pytest.mark.parametrize("M,K", [(128, 64), (200, 512), (1, 512), (333, 1536), (3200, 2048), (7680, 512), (25600, 512)])
@pytest.mark.parametrize("p", [0.0, 0.1])
def test_gemm_ln_matches_the_two_kernel_form(dev, M, K, p):
    a, w, bias, res, gamma, beta = _case(M, K, 3 + M + K, dev)
    assert ops.gemm_ln_supported(a, w)
    seed, off = 1234567, 4242
    s1, y1, y161, mean1, rstd1 = ops.gemm_ln(a, w, bias, res, gamma, beta, eps=1e-5, dropout_p=p, seed=seed, offset=off, save=True, want_bf16=True)
    s0 = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, dropout_p=p, seed=seed, offset=off, residual=res)
    y0, y160, _, mean0, rstd0 = ops.layernorm_fwd(s0, gamma, beta, eps=1e-5, want_bf16=True, save_stats=True)
    # the same dropout mask (backward regenerates it from (seed, offset)) and the same pre-LayerNorm sum
    assert float((s1 - s0).abs().max()) <= 2e-5 * max(1.0, float(s0.abs().max())), float((s1 - s0).abs().max())
    if p > 0:
        assert torch.equal((s1 - res) == 0, (s0 - res) == 0) or float((((s1 - res) == 0) != ((s0 - res) == 0)).float().mean()) < 1e-6
    assert float((mean1 - mean0).abs().max()) <= 1e-5 * max(1.0, float(mean0.abs().max()))
    assert float(((rstd1 - rstd0) / rstd0).abs().max()) <= 1e-4
    assert float((y1 - y0).abs().max()) <= 2e-4 * max(1.0, float(y0.abs().max())), float((y1 - y0).abs().max())
    assert float((y161.float() - y1).abs().max()) <= 2 ** -8 * float(y1.abs().max()) + 1e-6
    # plain PyTorch on the same bf16-rounded operands (no dropout)
    if p == 0:
        ref = F.layer_norm(a.float() @ w.float().t() + bias + res, (512,), gamma, beta, 1e-5)
        assert float((y1 - ref).abs().max()) <= 2e-3, float((y1 - ref).abs().max())
    # inference form: nothing saved, same y
    _, y2, y162, m2, r2 = ops.gemm_ln(a, w, bias, res, gamma, beta, eps=1e-5, dropout_p=p, seed=seed, offset=off, save=False, want_bf16=False)
    assert m2 is None and r2 is None and y162 is None and torch.equal(y2, y1)


def test_gemm_ln_rejects_other_widths(dev):
    a = torch.zeros((8, 64), device=dev, dtype=torch.bfloat16)
    w = torch.zeros((256, 64), device=dev, dtype=torch.bfloat16)
    assert not ops.gemm_ln_supported(a, w)
