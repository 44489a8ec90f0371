"""GPU: the tcgen05/TMEM/TMA GEMM against its CUDA-core twin and an fp64 reference, every operand form, both tile
widths, fused epilogues, split-K."""
import pytest
import torch
import torch.nn.functional as F

from openvivqa_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _mk(m, n, k, a_op, b_op, dev, seed=0):
    g = torch.Generator().manual_seed(seed + m + 3 * n + 7 * k)
    A, B = torch.randn((m, k), generator=g), torch.randn((n, k), generator=g)
    a = (A if a_op == "N" else A.t().contiguous()).to(dev).bfloat16()
    b = (B if b_op == "N" else B.t().contiguous()).to(dev).bfloat16()
    ref = (A.bfloat16().double() @ B.bfloat16().double().t()).float()
    return a, b, ref


@pytest.mark.parametrize("a_op,b_op", [("N", "N"), ("N", "T"), ("T", "T"), ("T", "N")])
@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (300, 200, 328), (1000, 512, 512), (257, 1024, 2048), (40000, 2048, 512)])
def test_forms_and_shapes(dev, a_op, b_op, m, n, k):
    if a_op == "T":
        m = (m + 7) // 8 * 8       # TMA: the contiguous dimension's pitch must be a multiple of 16 bytes
    if b_op == "T":
        n = (n + 7) // 8 * 8
    a, b, ref = _mk(m, n, k, a_op, b_op, dev)
    out = ops.gemm(a, b, a_op=a_op, b_op=b_op, out_dtype=torch.float32)
    err = (out.cpu() - ref).abs().max().item()
    assert err <= 1e-3 * ref.abs().max().item(), (a_op, b_op, m, n, k, err)
    twin = ops.gemm_bf16_simt(a, b, a_op=a_op, b_op=b_op)
    assert (out - twin).abs().max().item() <= 1e-3 * ref.abs().max().item()


def test_fused_epilogues(dev):
    m, n, k = 700, 768, 512
    a, b, ref = _mk(m, n, k, "N", "N", dev)
    g = torch.Generator().manual_seed(1)
    bias, res = torch.randn(n, generator=g), torch.randn((m, n), generator=g)
    pre = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    out = ops.gemm(a, b, bias=bias.to(dev), gelu=True, aux_out=pre, out_dtype=torch.bfloat16)
    want = F.gelu(ref + bias)
    assert (out.float().cpu() - want).abs().max() <= 2e-2 * want.abs().max()
    assert (pre.float().cpu() - (ref + bias)).abs().max() <= 2e-2 * ref.abs().max()
    c2 = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    out = ops.gemm(a, b, bias=bias.to(dev), residual=res.to(dev), out_dtype=torch.float32, c2=c2)
    want = ref + bias + res
    assert (out.cpu() - want).abs().max() <= 1e-3 * want.abs().max()
    assert (c2.float().cpu() - want).abs().max() <= 2e-2 * want.abs().max()
    # dgelu epilogue and dropout mask are the same functions as on the fp32 path
    keep = ops.dropout_keep_mask(m * n, 0.1, 9, 5, dev).view(m, n)
    out = ops.gemm(a, b, out_dtype=torch.float32, dropout_p=0.1, seed=9, offset=5)
    assert (out.cpu() - ref * keep.cpu() / 0.9).abs().max() <= 1e-3 * ref.abs().max()


@pytest.mark.parametrize("split_k", [0, 2, 7, 37])
def test_split_k_wgrad(dev, split_k):
    m, n, k = 512, 768, 9000
    a, b, ref = _mk(m, n, k, "T", "T", dev)
    out = ops.gemm(a, b, a_op="T", b_op="T", out_dtype=torch.float32, split_k=split_k)
    assert (out.cpu() - ref).abs().max() <= 2e-3 * ref.abs().max()


def test_unaligned_operands_take_the_cuda_core_twin(dev):
    """TMA needs 16-byte row pitches: operands that do not have them (K = 36 -> 72-byte rows; a column slice starting at an odd offset)
    are computed by the CUDA-core twin — same operand forms and epilogue — instead of failing."""
    a = torch.randn((64, 36), device=dev).bfloat16()      # K=36 -> row pitch 72 B, not a multiple of 16
    b = torch.randn((40, 36), device=dev).bfloat16()
    bias = torch.randn(40, device=dev)
    out = ops.gemm(a, b, bias=bias, out_dtype=torch.float32)
    ref = a.float() @ b.float().t() + bias
    assert (out - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()
    wide = torch.randn((64, 72), device=dev).bfloat16()
    sl = wide[:, 3:67]                                     # 16-byte pitch but a base address off by 6 bytes
    out2 = ops.gemm(sl, sl, out_dtype=torch.float32)
    ref2 = sl.float() @ sl.float().t()
    assert (out2 - ref2).abs().max().item() <= 1e-3 * ref2.abs().max().item()


@pytest.mark.parametrize("n,odt", [(520, torch.float32), (520, torch.bfloat16), (516, torch.float32), (250, torch.bfloat16), (2048, torch.bfloat16)])
def test_epilogue_vector_and_ragged_paths(dev, n, odt):
    """The transposed 8-wide epilogue: full tiles, a tile with a ragged 8-column tail (n % 8 != 0), rows not a multiple
    of 128, with bias + GELU + residual + bf16 copy in one pass, against the CUDA-core twin."""
    m, k = 300, 192
    a, b, ref = _mk(m, n, k, "N", "N", dev)
    bias = torch.randn(n, device=dev)
    res = torch.randn((m, n), device=dev)
    pre = torch.empty((m, n), device=dev, dtype=odt)
    c2 = torch.empty((m, n), device=dev, dtype=torch.bfloat16) if n % 8 == 0 else None
    out = ops.gemm(a, b, out_dtype=odt, bias=bias, gelu=True, aux_out=pre, residual=res, c2=c2)
    want_pre = ref.to(dev) + bias
    want = torch.nn.functional.gelu(want_pre) + res
    tol = 2e-2 if odt == torch.bfloat16 else 2e-3
    assert (pre.float() - want_pre).abs().max() <= tol * want_pre.abs().max()
    assert (out.float() - want).abs().max() <= tol * want.abs().max()
    if c2 is not None:
        assert (c2.float() - want).abs().max() <= 2e-2 * want.abs().max()
    # backward-style epilogue: multiply by GELU'(saved pre-activation)
    d = ops.gemm(a, b, out_dtype=odt, dgelu_aux=pre)
    xs = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(xs).sum().backward()
    want_d = ref.to(dev) * xs.grad
    assert (d.float() - want_d).abs().max() <= tol * want_d.abs().max()
    # the training pair: forward saves GELU'(pre) itself, backward multiplies by it
    dsave = torch.empty((m, n), device=dev, dtype=odt)
    out2 = ops.gemm(a, b, out_dtype=odt, bias=bias, gelu=True, aux_out=dsave, save_dgelu=True, residual=res)
    assert torch.equal(out2, out)
    xs2 = want_pre.clone().requires_grad_(True)
    torch.nn.functional.gelu(xs2).sum().backward()
    assert (dsave.float() - xs2.grad).abs().max() <= (1e-2 if odt == torch.bfloat16 else 1e-5)
    d2 = ops.gemm(a, b, out_dtype=odt, mul_aux=dsave)
    want_d2 = ref.to(dev) * dsave.float()
    assert (d2.float() - want_d2).abs().max() <= tol * want_d2.abs().max()


def test_fast_gelu_matches_erf_gelu(dev):
    """normal_cdf_fast (Abramowitz-Stegun 7.1.26) vs the exact erf GELU over the whole useful range."""
    x = torch.linspace(-12, 12, 128 * 256, device=dev).view(128, 256)
    eye = torch.eye(256, device=dev).bfloat16()
    xb = x.bfloat16()
    out = ops.gemm(xb, eye, out_dtype=torch.float32, gelu=True)           # y = gelu(x_bf16 * I)
    want = torch.nn.functional.gelu(xb.float())
    assert (out - want).abs().max().item() <= 1e-6


@pytest.mark.parametrize("m,n,k", [(300, 512, 192), (1000, 2048, 512), (257, 768, 2048), (300, 520, 192), (77, 72, 64), (5000, 1000, 512)])
def test_specialised_epilogues(dev, m, n, k):
    """The compile-time feature sets of gemm_tcgen05.cu (fc1 train / eval, dpre, fc_o / fc2 train / eval, dx + residual)
    against torch: erf GELU (the tanh-fit epilogue must stay within bf16 rounding of it), the same Philox keep mask as every
    other kernel, and the saved derivative = GELU' x dropout scale."""
    a, b, ref = _mk(m, n, k, "N", "N", dev)
    ref = ref.to(dev)
    g = torch.Generator().manual_seed(2)
    bias, res = torch.randn(n, generator=g).to(dev), torch.randn((m, n), generator=g).to(dev)
    pre = ref + bias
    p, seed, off = 0.1, 11, 3
    keep = ops.dropout_keep_mask(m * n, p, seed, off, dev).view(m, n).float()
    assert 0.88 < keep.mean().item() < 0.92
    xs = pre.clone().requires_grad_(True)
    F.gelu(xs).sum().backward()
    dgelu, gelu = xs.grad, F.gelu(pre)
    tol = 2e-2
    # fc1, training: bias + GELU + dropout -> bf16, saved derivative
    dsave = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    y = ops.gemm(a, b, bias=bias, gelu=True, aux_out=dsave, save_dgelu=True, dropout_p=p, seed=seed, offset=off, out_dtype=torch.bfloat16)
    assert (y.float() - gelu * keep / (1 - p)).abs().max() <= tol * gelu.abs().max()
    assert (dsave.float() - dgelu * keep / (1 - p)).abs().max() <= tol
    assert (y[keep == 0] == 0).all() and (dsave[keep == 0] == 0).all()
    # fc1, training with p = 0, and eval
    y0 = ops.gemm(a, b, bias=bias, gelu=True, aux_out=dsave, save_dgelu=True, out_dtype=torch.bfloat16)
    y1 = ops.gemm(a, b, bias=bias, gelu=True, out_dtype=torch.bfloat16)
    assert torch.equal(y0, y1)
    assert (y1.float() - gelu).abs().max() <= tol * gelu.abs().max()
    assert (y1.float() - gelu).abs().max() <= 3e-3 * torch.maximum(gelu.abs(), torch.ones_like(gelu)).max()
    assert (dsave.float() - dgelu).abs().max() <= 1e-2
    # dpre = (dy W2) * saved derivative  (N,T operand form, bf16 out)
    bt = b.t().contiguous()                                   # stored [k, n]: the "T" form of the same B
    d = ops.gemm(a, bt, b_op="T", out_dtype=torch.bfloat16, mul_aux=dsave)
    want = ref * dsave.float()
    assert (d.float() - want).abs().max() <= tol * want.abs().max()
    # the same with the fused column sums (bias gradient of the consuming layer), and the plain dgrad with column sums
    cs = torch.zeros(n, device=dev)
    d_cs = ops.gemm(a, bt, b_op="T", out_dtype=torch.bfloat16, mul_aux=dsave, colsum=cs)
    assert torch.equal(d_cs, d)
    assert (cs - want.sum(0)).abs().max() <= 2e-3 * want.abs().sum(0).max()
    cs = torch.zeros(n, device=dev)
    plain = ops.gemm(a, bt, b_op="T", out_dtype=torch.bfloat16, colsum=cs)
    assert (plain.float() - ref).abs().max() <= tol * ref.abs().max()
    assert (cs - ref.sum(0)).abs().max() <= 2e-3 * ref.abs().sum(0).max()
    cs = torch.zeros(n, device=dev)                            # run-time feature path (fp32 out, bias + column sums), accumulating twice
    for _ in range(2):
        s32 = ops.gemm(a, b, bias=bias, out_dtype=torch.float32, colsum=cs)
    assert (cs - 2 * pre.sum(0)).abs().max() <= 2e-3 * pre.abs().sum(0).max()
    assert (s32 - pre).abs().max() <= 1e-3 * pre.abs().max()
    # fc_o / fc2: bias + dropout + residual -> fp32 (training), bias + residual (eval)
    s = ops.gemm(a, b, bias=bias, dropout_p=p, seed=seed, offset=off, residual=res, out_dtype=torch.float32)
    want = pre * keep / (1 - p) + res
    assert (s - want).abs().max() <= 1e-3 * want.abs().max()
    s = ops.gemm(a, b, bias=bias, residual=res, out_dtype=torch.float32)
    assert (s - (pre + res)).abs().max() <= 1e-3 * (pre + res).abs().max()
    # dx = dy W + the other branch's gradient  (N,T form, fp32 out)
    s = ops.gemm(a, bt, b_op="T", residual=res, out_dtype=torch.float32)
    assert (s - (ref + res)).abs().max() <= 1e-3 * (ref + res).abs().max()
    # ... accumulated in place (result written over the residual it read), and the saved derivative multiplied in place
    acc = res.clone()
    ops.gemm(a, bt, b_op="T", residual=acc, out=acc)
    assert torch.equal(acc, s)
    dd = dsave.clone()
    ops.gemm(a, bt, b_op="T", mul_aux=dd, out=dd)
    assert torch.equal(dd, d)


def test_tanhfit_gelu_matches_erf_gelu(dev):
    """The specialised bf16 epilogue's GELU (tanh form fitted to the erf definition, tanh.approx) against F.gelu on a
    dense grid: error well below the bf16 rounding of the result."""
    x = torch.linspace(-12, 12, 128 * 256, device=dev).view(128, 256)
    eye = torch.eye(256, device=dev).bfloat16()
    xb = x.bfloat16()
    zero = torch.zeros(256, device=dev)
    dsave = torch.empty((128, 256), device=dev, dtype=torch.bfloat16)
    out = ops.gemm(xb, eye, bias=zero, out_dtype=torch.bfloat16, gelu=True, aux_out=dsave, save_dgelu=True)
    xs = xb.float().requires_grad_(True)
    want = F.gelu(xs)
    want.sum().backward()
    err = (out.float() - want.detach()).abs()
    assert (err <= 2.0 ** -8 * want.detach().abs() + 1.5e-3).all(), err.max().item()
    assert (dsave.float() - xs.grad).abs().max().item() <= 6e-3


@pytest.mark.parametrize("a_op,b_op", [("N", "N"), ("N", "T"), ("T", "T"), ("T", "N")])
@pytest.mark.parametrize("m,n,k", [(40000, 512, 512), (20008, 768, 256), (38000 + 77, 256, 1024), (40000, 2048, 512), (30000, 1100, 128)])
def test_cta_pair_matches_single_cta(dev, monkeypatch, a_op, b_op, m, n, k):
    """cta_group::2 (a CTA pair per 256-row tile, B split across the pair) against the single-CTA kernel: same K order, so
    the results must be bit-identical, ragged last pair tiles included."""
    if a_op == "T":
        m = (m + 7) // 8 * 8
    if b_op == "T":
        n = (n + 7) // 8 * 8       # TMA: the contiguous dimension's pitch must be a multiple of 16 bytes
    a, b, ref = _mk(m, n, k, a_op, b_op, dev)
    bias = torch.randn(n, device=dev)
    res = torch.randn((m, n), device=dev)
    outs = {}
    for pair in ("0", "1"):
        monkeypatch.setenv("OVQA_GEMM_2CTA", pair)
        o1 = ops.gemm(a, b, a_op=a_op, b_op=b_op, out_dtype=torch.float32)
        o2 = ops.gemm(a, b, a_op=a_op, b_op=b_op, out_dtype=torch.bfloat16, bias=bias)
        o3 = ops.gemm(a, b, a_op=a_op, b_op=b_op, out_dtype=torch.float32, bias=bias, residual=res, dropout_p=0.1, seed=3, offset=7)
        torch.cuda.synchronize()
        outs[pair] = (o1, o2, o3)
    assert (outs["1"][0].cpu() - ref).abs().max() <= 1e-3 * ref.abs().max()
    for x, y in zip(outs["0"], outs["1"]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("split_k", [0, 3, 11])
def test_cta_pair_split_k(dev, monkeypatch, split_k):
    m, n, k = 1536, 512, 20000
    a, b, ref = _mk(m, n, k, "T", "T", dev)
    for pair in ("0", "1"):
        monkeypatch.setenv("OVQA_GEMM_2CTA", pair)
        out = ops.gemm(a, b, a_op="T", b_op="T", out_dtype=torch.float32, split_k=split_k)
        assert (out.cpu() - ref).abs().max() <= 2e-3 * ref.abs().max(), pair
