"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs and
against the reference-generated golden vectors.

Tolerances (BASELINE.json north_star): fp32 path <= 1e-4 relative (max|a-b| / max|ref|) on attention outputs and
logits; bf16 path <= 2e-2 absolute on LayerNorm-scale outputs / logits.
"""
import math

import pytest
import torch
import torch.nn.functional as F

import openvivqa_b200 as ovqa
from openvivqa_b200 import functional as Fn
from openvivqa_b200 import harness as H
from openvivqa_b200 import ops
from oracle import cases as C
from oracle import restated as R
from oracle import synth as S

pytestmark = pytest.mark.gpu

FP32_REL = 1e-4
BF16_ABS = 2e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    H.strict_fp32_library_math()      # the non-hot-path torch ops (cuDNN LSTM, embedding Linear) must not use TF32 here
    return torch.device("cuda:0")


def rel_err(a, ref):
    a, ref = a.detach().float().cpu(), ref.detach().float().cpu()
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def abs_err(a, ref):
    return (a.detach().float().cpu() - ref.detach().float().cpu()).abs().max().item()


def check(a, ref, mode, what="", bf16_abs=BF16_ABS):
    """fp32: 1e-4 relative.  bf16: 2e-2 absolute on logits; activations of LayerNorm scale (|x| up to ~4-5 after
    several stacked blocks) get the same bound relative to a unit-scale logit, i.e. scaled by max|ref|/2."""
    if mode == "fp32":
        e = rel_err(a, ref)
        assert e <= FP32_REL, f"{what}: fp32 rel err {e:.3e} > {FP32_REL}"
    else:
        e = abs_err(a, ref)
        assert e <= bf16_abs * max(1.0, ref.abs().max().item() / 2), f"{what}: bf16 abs err {e:.3e}"
    return e


def grad_check(named_grads, ref_grads, tol):
    """Compare parameter gradients; the error of each tensor is measured against the largest gradient of its
    kind (weights / biases) because some are analytically ZERO (fc_k.bias: softmax is shift invariant) and only
    hold rounding noise in the oracle."""
    scale = {}
    for n, g in ref_grads.items():
        if g is not None:
            kind = n.rsplit(".", 1)[-1]
            scale[kind] = max(scale.get(kind, 0.0), g.abs().max().item())
    worst = (0.0, None)
    for n, g in named_grads:
        assert g is not None, n
        r = ref_grads[n]
        kind = n.rsplit(".", 1)[-1]
        e = abs_err(g, r) / max(r.abs().max().item(), 0.05 * scale[kind], 1e-12)
        worst = max(worst, (e, n))
        assert e <= tol, f"{n}: grad err {e:.3e} > {tol}"
    return worst


def load(module, params, dev):
    module.load_state_dict(params, strict=True)
    return module.to(dev).eval()


# ------------------------------------------------------------------------------------------------ GEMM (fp32 FFMA path)
@pytest.mark.parametrize("a_op,b_op", [("N", "N"), ("N", "T"), ("T", "T"), ("T", "N")])
def test_gemm_fp32_operand_forms(dev, a_op, b_op):
    g = torch.Generator().manual_seed(3)
    m, n, k = 77, 130, 203
    A = torch.randn((m, k), generator=g)
    B = torch.randn((n, k), generator=g)
    ref = A.double() @ B.double().t()
    a = (A if a_op == "N" else A.t().contiguous()).to(dev)
    b = (B if b_op == "N" else B.t().contiguous()).to(dev)
    out = ops.gemm(a, b, a_op=a_op, b_op=b_op)
    assert rel_err(out, ref.float()) < 1e-5


def test_gemm_fp32_epilogue(dev):
    g = torch.Generator().manual_seed(4)
    m, n, k = 50, 96, 64
    A, B = torch.randn((m, k), generator=g), torch.randn((n, k), generator=g)
    bias, res = torch.randn(n, generator=g), torch.randn((m, n), generator=g)
    pre = A @ B.t() + bias
    aux = torch.empty((m, n), device=dev)
    out = ops.gemm(A.to(dev), B.to(dev), bias=bias.to(dev), gelu=True, aux_out=aux, residual=res.to(dev))
    assert rel_err(aux, pre) < 1e-5
    assert rel_err(out, F.gelu(pre) + res) < 1e-5
    dy = torch.randn((m, n), generator=g)
    # dgelu epilogue: (dy W) * gelu'(aux)
    W = torch.randn((n, 40), generator=g)          # stored [k_red=n, n_out=40]
    pre2 = torch.randn((m, 40), generator=g)
    x = pre2.clone().requires_grad_(True)
    F.gelu(x).backward(dy @ W)
    out2 = ops.gemm(dy.to(dev), W.to(dev), b_op="T", dgelu_aux=pre2.to(dev))
    assert rel_err(out2, x.grad) < 1e-5
    # the training pair: forward saves GELU'(pre) in aux_out, backward multiplies by the saved tensor
    dsave = torch.empty((m, n), device=dev)
    out_s = ops.gemm(A.to(dev), B.to(dev), bias=bias.to(dev), gelu=True, aux_out=dsave, save_dgelu=True, residual=res.to(dev))
    xp = pre.clone().requires_grad_(True)
    F.gelu(xp).sum().backward()
    assert torch.equal(out_s, out) and rel_err(dsave, xp.grad) < 1e-5
    gsave = torch.randn((m, 40), generator=g)
    out_m = ops.gemm(dy.to(dev), W.to(dev), b_op="T", mul_aux=gsave.to(dev))
    assert rel_err(out_m, (dy @ W) * gsave) < 1e-5
    # split-k accumulates into fp32
    out3 = ops.gemm(A.t().contiguous().to(dev), B.t().contiguous().to(dev), a_op="T", b_op="T", split_k=3)
    assert rel_err(out3, A @ B.t()) < 1e-5


# ------------------------------------------------------------------------------------------------ attention blocks
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(C.ATT_CASES))
def test_attention_cases_vs_golden_and_oracle(dev, golden, name, mode):
    B, Lq, Lk, D, h, dk, dv, _, _ = C.ATT_CASES[name]
    gold = golden("attention")[name]
    x = C.att_inputs(name)
    params = S.synth_params(C.att_spec(D, h, dk, dv), 0)
    cfg = H.attention_cfg(D, h, dk, dv, compute_dtype=mode)
    mha = load(ovqa.MultiHeadAttention(cfg), params, dev)
    q = x["queries"].to(dev)
    k = q if x["keys"] is x["queries"] else x["keys"].to(dev)
    mask = None if x["mask"] is None else x["mask"].to(dev)
    with torch.no_grad():
        y = mha(q, k, k, mask)
        core_out, att = mha.attention(q, k, k, mask)
    check(y, gold["mha_out"], mode, f"{name} MHA")
    check(core_out, gold["core_out"], mode, f"{name} core out")
    assert att.shape == gold["att"].shape
    if mode == "fp32":
        assert abs_err(att, gold["att"]) <= 1e-5
    else:
        assert abs_err(att, gold["att"]) <= 2e-2
    # rows of the probabilities sum to one, fully-masked rows stay finite
    assert torch.isfinite(att).all()
    assert abs_err(att.sum(-1), torch.ones(att.shape[:-1])) < 1e-4
    # and the restated oracle agrees on the same inputs (not only the stored vector)
    ref = R.mha(params, x["queries"], x["keys"], x["values"], x["mask"], h=h, d_k=dk, d_v=dv)
    check(y, ref, mode, f"{name} MHA vs oracle")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_mha_aoa(dev, golden, mode):
    B, Lq, Lk, D, h, dk, dv, _, _ = C.ATT_CASES["guided_100x30"]
    x = C.att_inputs("guided_100x30")
    params = S.synth_params(C.att_spec(D, h, dk, dv, aoa=True), 0)
    mha = load(ovqa.MultiHeadAttention(H.attention_cfg(D, h, dk, dv, use_aoa=True, compute_dtype=mode)), params, dev)
    with torch.no_grad():
        y = mha(x["queries"].to(dev), x["keys"].to(dev), x["keys"].to(dev), x["mask"].to(dev))
    check(y, golden("attention")["guided_100x30+aoa"]["mha_out"], mode, "AoA")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name,D,dff,B,L", [("pwff_512", 512, 2048, 2, 30), ("pwff_768", 768, 3072, 1, 19)])
def test_pwff(dev, golden, name, D, dff, B, L, mode):
    params = S.synth_params(C.pwff_spec(D, dff), 0)
    m = load(ovqa.PositionWiseFeedForward(H.attention_cfg(D, 8, d_ff=dff, compute_dtype=mode)), params, dev)
    x = torch.randn((B, L, D), generator=S._gen(7, name))
    with torch.no_grad():
        y = m(x.to(dev))
    check(y, golden("attention")[name]["out"], mode, name)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", list(C.GEO_CASES))
def test_geometry_core(dev, golden, name, mode):
    B, N, D, h, trig = C.GEO_CASES[name]
    gold = golden("geometry")[name]
    x = C.geo_inputs(name)
    params = S.synth_params(C.att_spec(D, h, D // h, D // h, geometry_dg=x["d_g"]), 0)
    cfg = H.attention_cfg(D, h, arch="AugmentedGeometryScaledDotProductAttention", trig=trig, compute_dtype=mode)
    mha = load(ovqa.MultiHeadAttention(cfg), params, dev)
    xd, boxes = x["x"].to(dev), x["boxes"].to(dev)
    with torch.no_grad():
        out, mn = mha.attention(xd, xd, xd, boxes)            # the reference's only executable call (mask=None)
    # sin/cos arguments reach ~700: the reference's own fp32 embedding is only good to ~1e-4 (tests/test_oracle_golden)
    tol = 5e-4 if mode == "fp32" else 3e-2
    assert abs_err(mn, gold["att"]) <= tol
    assert rel_err(out, gold["core_out"]) <= (5e-4 if mode == "fp32" else 5e-2)
    # through MHA with a padding mask (documented deviation: the reference crashes here) vs the restated oracle
    feats = S.padded_features(B, N, D, 77, min_valid=N // 2, relu_like=False)
    mask = R.padding_mask(feats, 0)
    ref = R.mha(params, feats, feats, feats, mask, h=h, d_k=D // h, d_v=D // h, boxes=x["boxes"], d_g=x["d_g"], trig=trig)
    fd = feats.to(dev)
    with torch.no_grad():
        y = mha(fd, fd, fd, mask.to(dev), boxes=boxes)
    # (this reference is computed on the host at test time: its sin/cos of arguments ~700 differ between host CPUs at the 1e-4 level —
    # 4.9e-4 on one box, 5.6e-4 on another for the same GPU output — hence the margin over the golden-vector bound above)
    assert rel_err(y, ref) <= (1e-3 if mode == "fp32" else 5e-2)
    # materialised bias kernel agrees with the oracle's log-bias
    geo = ops.GeoParams(boxes, *[t.to(dev) for t in (torch.cat([params[f"attention.fc_gs.{i}.weight"] for i in range(h)]),
                                                       torch.cat([params[f"attention.fc_gs.{i}.bias"] for i in range(h)]))], trig)
    bias = ops.box_geometry_bias(geo, h)
    refb = R.geometry_log_bias(R.sub(params, "attention"), x["boxes"], h=h, d_g=x["d_g"], trig=trig)
    close = (bias.cpu() - refb).abs() <= 2e-3 + 2e-3 * refb.abs()
    assert close.float().mean() > 0.999     # the -13.8 clamp floor flips on isolated near-zero entries


# ------------------------------------------------------------------------------------------------ encoders + models
def _enc_inputs():
    D = 128
    vis = S.padded_features(2, 36, D, 31, min_valid=10, relu_like=False)
    txt = S.padded_features(2, 14, D, 32, min_valid=4, relu_like=False)
    return vis, R.padding_mask(vis, 0), txt, R.padding_mask(txt, 0)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_encoders(dev, golden, mode):
    g = golden("encoders")
    vis, vm, txt, tm = [t.to(dev) for t in _enc_inputs()]
    x = H.xmt_cfg("cross", 128, 256, 2, 4, 256, compute_dtype=mode)
    a = H.attention_cfg(128, 4, d_ff=256, compute_dtype=mode)
    with torch.no_grad():
        m = load(H.Encoder(H.NS(D_MODEL=128, LAYERS=2, SELF_ATTENTION=a)), S.synth_params(g["Encoder"]["spec"], 0), dev)
        check(m(txt, tm), g["Encoder"]["out"], mode, "Encoder")
        m = load(H.GuidedAttentionEncoder(H.NS(D_MODEL=128, LAYERS=2, SELF_ATTENTION=a, GUIDED_ATTENTION=a)),
                 S.synth_params(g["GuidedAttentionEncoder"]["spec"], 0), dev)
        check(m(vis, vm, txt, tm), g["GuidedAttentionEncoder"]["out"], mode, "GuidedAttentionEncoder")
        m = load(H.CoAttentionEncoder(H.xmt_cfg("co", 128, 256, 2, 4, 256, compute_dtype=mode).ENCODER),
                 S.synth_params(g["CoAttentionEncoder"]["spec"], 0), dev)
        v, l = m(vis, vm, txt, tm)
        check(v, g["CoAttentionEncoder"]["out_v"], mode, "CoAtt v"); check(l, g["CoAttentionEncoder"]["out_l"], mode, "CoAtt l")
        m = load(H.CrossModalityEncoder(x.ENCODER), S.synth_params(g["CrossModalityEncoder"]["spec"], 0), dev)
        v, l = m(vis, vm, txt, tm)
        check(v, g["CrossModalityEncoder"]["out_v"], mode, "XMT v"); check(l, g["CrossModalityEncoder"]["out_l"], mode, "XMT l")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["small", "full"])
def test_mcan_logits(dev, golden, name, mode):
    c = C.MCAN_SMALL if name == "small" else C.MCAN_FULL
    g = golden("mcan")[name]
    m = H.MCAN(H.mcan_cfg(c["d_model"], c["d_feature"], c["layers"], c["h"], c["d_ff"], compute_dtype=mode),
               H.StubVocab(c["vocab"], c["answers"]))
    m = load(m, S.synth_params(g["spec"], 0), dev)
    with torch.no_grad():
        logp = m(H.Batch(C.mcan_inputs(c)).to(dev))
    e = check(logp, g["logp"], mode, f"MCAN {name}")
    agree = (logp.argmax(-1).cpu() == g["logp"].argmax(-1)).float().mean().item()
    print(f"MCAN {name} {mode}: err {e:.3e}, argmax agreement {agree:.3f}")
    if mode == "fp32":
        assert agree == 1.0


def test_mcan_bf16_argmax_agreement_at_batch(dev):
    """>= 99.9% argmax agreement where the oracle's own top-2 margin exceeds twice the bf16 tolerance; the raw
    agreement and the margin distribution are printed (random-init margins are tiny)."""
    c = dict(C.MCAN_SMALL, B=256)
    spec_model = H.MCAN(H.mcan_cfg(c["d_model"], c["d_feature"], c["layers"], c["h"], c["d_ff"]), H.StubVocab(c["vocab"], c["answers"]))
    params = S.synth_params([(k, tuple(v.shape)) for k, v in spec_model.state_dict().items()], 0)
    x = C.mcan_inputs(c, seed=9)
    ref = R.mcan_forward(params, x["region_features"], x["question_tokens"], d_model=c["d_model"], h=c["h"], layers=c["layers"])
    m = H.MCAN(H.mcan_cfg(c["d_model"], c["d_feature"], c["layers"], c["h"], c["d_ff"], compute_dtype="bf16"), H.StubVocab(c["vocab"], c["answers"]))
    m = load(m, params, dev)
    with torch.no_grad():
        logp = m(H.Batch(x).to(dev)).cpu()
    top2 = ref.topk(2, dim=-1).values
    margin = top2[:, 0] - top2[:, 1]
    agree = logp.argmax(-1) == ref.argmax(-1)
    decided = margin > 2 * BF16_ABS
    print(f"bf16 argmax agreement raw {agree.float().mean():.4f}; margin median {margin.median():.4f} min {margin.min():.2e}; "
          f"decided rows {int(decided.sum())}/{len(margin)} agreement {agree[decided].float().mean():.4f}")
    assert abs_err(logp, ref) <= BF16_ABS
    assert agree[decided].float().mean() >= 0.999


# ------------------------------------------------------------------------------------------------ gradients (dropout 0)
def _oracle_grads(fn, params, inputs):
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xs = [x.clone().requires_grad_(True) for x in inputs]
    out = fn(p, *xs)
    w = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    (out * w).sum().backward()
    return out, w, {k: v.grad for k, v in p.items()}, [x.grad for x in xs]


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("kind", ["self", "cross", "causal"])
def test_mha_gradients(dev, kind, mode):
    D, h, dk = 128, 4, 32
    B, Lq, Lk = 3, (21 if kind != "cross" else 37), 21
    params = S.synth_params(C.att_spec(D, h, dk, dk), 2)
    keys = S.padded_features(B, Lk, D, 41, min_valid=5, relu_like=False)
    queries = keys if kind != "cross" else torch.randn((B, Lq, D), generator=S._gen(1, "gq"))
    mask = R.padding_mask(keys, 0)
    if kind == "causal":
        mask = R.self_attention_masks(mask, R.sequential_mask(Lk))
    if kind == "cross":
        f = lambda p, q, k: R.mha(p, q, k, k, mask, h=h, d_k=dk, d_v=dk)
        out, w, gp, gx = _oracle_grads(f, params, [queries, keys])
    else:
        f = lambda p, q: R.mha(p, q, q, q, mask, h=h, d_k=dk, d_v=dk)
        out, w, gp, gx = _oracle_grads(f, params, [queries])
    m = ovqa.MultiHeadAttention(H.attention_cfg(D, h, dk, dk, dropout=0.0, compute_dtype=mode))
    m.load_state_dict(params)
    m = m.to(dev).train()
    qd = queries.to(dev).requires_grad_(True)
    kd = qd if kind != "cross" else keys.to(dev).requires_grad_(True)
    y = m(qd, kd, kd, mask.to(dev))
    (y * w.to(dev)).sum().backward()
    tol = 2e-4 if mode == "fp32" else 4e-2
    assert rel_err(y, out) <= (FP32_REL if mode == "fp32" else 3e-2)
    assert rel_err(qd.grad, gx[0]) <= tol, "dqueries"
    if kind == "cross":
        assert rel_err(kd.grad, gx[1]) <= tol, "dkeys"
    grad_check([(n, p.grad) for n, p in m.named_parameters()], gp, tol)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_pwff_gradients(dev, mode):
    D, dff = 128, 320
    params = S.synth_params(C.pwff_spec(D, dff), 3)
    x = torch.randn((3, 17, D), generator=S._gen(2, "px"))
    out, w, gp, gx = _oracle_grads(lambda p, t: R.pwff(p, t), params, [x])
    m = ovqa.PositionWiseFeedForward(H.attention_cfg(D, 4, d_ff=dff, dropout=0.0, compute_dtype=mode))
    m.load_state_dict(params)
    m = m.to(dev).train()
    xd = x.to(dev).requires_grad_(True)
    y = m(xd)
    (y * w.to(dev)).sum().backward()
    tol = 2e-4 if mode == "fp32" else 4e-2
    assert rel_err(xd.grad, gx[0]) <= tol
    grad_check([(n, p.grad) for n, p in m.named_parameters()], gp, tol)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_feature_embedding_front_end(dev, mode):
    """FeatureEmbedding (vision_embeddings.py:10-25) against the restated oracle: output, padding mask of the all-zero
    rows, gradients of proj.weight / proj.bias; and the fused dropout of the bf16 path."""
    B, N, K, D = 5, 37, 256, 128
    feats = S.padded_features(B, N, K, 41, min_valid=5, relu_like=True)
    params = {"proj.weight": torch.randn((D, K), generator=S._gen(7, "fw")) * 0.05, "proj.bias": torch.randn((D,), generator=S._gen(7, "fb")) * 0.1}
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref, ref_mask = R.feature_embedding(p, feats)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (ref * w).sum().backward()
    m = ovqa.FeatureEmbedding(H.NS(D_FEATURE=K, D_MODEL=D, DROPOUT=0.0, COMPUTE_DTYPE=mode))
    m.load_state_dict(params)
    m = m.to(dev).train()
    y, mask = m(feats.to(dev))
    assert mask.shape == ref_mask.shape and torch.equal(mask.cpu(), ref_mask.to(mask.dtype))
    check(y, ref, mode, "FeatureEmbedding")
    (y * w.to(dev)).sum().backward()
    tol = 2e-4 if mode == "fp32" else 4e-2
    grad_check([(n, q.grad) for n, q in m.named_parameters()], {k: v.grad for k, v in p.items()}, tol)
    if mode == "bf16":                                   # fused dropout: zero pattern with the right rate, kept values scaled by 1 / (1 - p)
        m.dropout.p = 0.25
        yd, _ = m(feats.to(dev))
        live = y.detach().abs() > 1e-3
        kept = (yd.detach() != 0) & live
        frac = kept.sum().item() / live.sum().item()
        assert 0.70 < frac < 0.80, frac
        assert (yd.detach()[kept] - y.detach()[kept] / 0.75).abs().max().item() <= 2e-2 * y.detach().abs().max().item()


def test_geometry_gradients(dev):
    D, h = 128, 4
    B, N = 2, 13
    params = S.synth_params(C.att_spec(D, h, D // h, D // h, geometry_dg=D // h), 4)
    x = torch.randn((B, N, D), generator=S._gen(3, "gx"))
    boxes = S.boxes(B, N, 5)
    f = lambda p, t: R.mha(p, t, t, t, None, h=h, d_k=D // h, d_v=D // h, boxes=boxes, d_g=D // h, trig=True)
    out, w, gp, gx = _oracle_grads(f, params, [x])
    m = ovqa.MultiHeadAttention(H.attention_cfg(D, h, dropout=0.0, arch="AugmentedGeometryScaledDotProductAttention"))
    m.load_state_dict(params)
    m = m.to(dev).train()
    xd = x.to(dev).requires_grad_(True)
    y = m(xd, xd, xd, None, boxes=boxes.to(dev))
    (y * w.to(dev)).sum().backward()
    assert rel_err(y, out) <= 5e-4
    assert rel_err(xd.grad, gx[0]) <= 2e-3
    grad_check([(n, p.grad) for n, p in m.named_parameters()], gp, 5e-3)


def test_mcan_training_step_gradients_fp32(dev, golden):
    """Whole-model check: one fwd+NLL+bwd of MCAN-small (dropout 0) against the oracle's autograd."""
    c = C.MCAN_SMALL
    params = S.synth_params(golden("mcan")["small"]["spec"], 0)
    x = C.mcan_inputs(c)
    ans = torch.arange(c["B"]) % c["answers"]
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss_ref = F.nll_loss(R.mcan_forward(p, x["region_features"], x["question_tokens"], d_model=c["d_model"], h=c["h"], layers=c["layers"]), ans)
    loss_ref.backward()
    m = H.MCAN(H.mcan_cfg(c["d_model"], c["d_feature"], c["layers"], c["h"], c["d_ff"], dropout=0.0, compute_dtype="fp32"),
               H.StubVocab(c["vocab"], c["answers"]))
    m.load_state_dict(params)
    m = m.to(dev).train()
    loss = F.nll_loss(m(H.Batch(x).to(dev)), ans.to(dev))
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * abs(loss_ref.item())
    worst = grad_check([(n, q.grad) for n, q in m.named_parameters()], {n: t.grad for n, t in p.items()}, 1e-3)
    print("worst grad rel err", worst)


# ------------------------------------------------------------------------------------------------ dropout
def test_dropout_mask_is_consistent_between_kernels(dev):
    p, seed, off = 0.1, 1234, 77
    m, n, k = 64, 128, 32
    keep = ops.dropout_keep_mask(m * n, p, seed, off, dev).view(m, n)
    rate = 1 - keep.mean().item()
    assert abs(rate - p) < 0.02
    A = torch.randn((m, k), device=dev)
    Bm = torch.randn((n, k), device=dev)
    plain = ops.gemm(A, Bm)
    dropped = ops.gemm(A, Bm, dropout_p=p, seed=seed, offset=off)
    assert rel_err(dropped, plain * keep / (1 - p)) < 1e-6
    x = torch.randn((m, n), device=dev)
    assert rel_err(ops.dropout_apply(x, p, seed, off), x * keep / (1 - p)) < 1e-6
    y16 = ops.dropout_apply(x, p, seed, off, torch.bfloat16)
    assert abs_err(y16, (x * keep / (1 - p)).bfloat16()) == 0.0
    # a different offset gives a different mask
    assert (ops.dropout_keep_mask(m * n, p, seed, off + 10000, dev).view(m, n) != keep).any()
    # device-side offset base (CUDA-graph replays): offset + *base is what the kernels use, in every kernel family
    base = torch.tensor([10000], dtype=torch.int64, device=dev)
    keep2 = ops.dropout_keep_mask(m * n, p, seed, off + 10000, dev).view(m, n)
    assert torch.equal(ops.dropout_keep_mask(m * n, p, seed, off, dev, offset_base=base).view(m, n), keep2)
    assert rel_err(ops.gemm(A, Bm, dropout_p=p, seed=seed, offset=off, offset_base=base), plain * keep2 / (1 - p)) < 1e-6
    assert rel_err(ops.gemm(A.bfloat16(), Bm.bfloat16(), out_dtype=torch.float32, dropout_p=p, seed=seed, offset=off, offset_base=base),
                   ops.gemm(A.bfloat16(), Bm.bfloat16(), out_dtype=torch.float32) * keep2 / (1 - p)) < 1e-6
    assert rel_err(ops.dropout_apply(x, p, seed, off, offset_base=base), x * keep2 / (1 - p)) < 1e-6
    ops.advance_u64(base, 5)
    assert base.item() == 10005


def test_cuda_graph_replays_draw_fresh_dropout_masks(dev):
    """A captured training step freezes (seed, offset) in its launch parameters; advancing the device-side offset base
    inside the graph makes every replay use new masks, and backward still matches its own forward."""
    torch.manual_seed(0)
    m = ovqa.MultiHeadAttention(H.attention_cfg(128, 4, dropout=0.5, compute_dtype="bf16")).to(dev).train()
    x = torch.randn((4, 9, 128), device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            m(x, x, x, None).square().sum().backward()
    torch.cuda.current_stream().wait_stream(side)
    m.zero_grad(set_to_none=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        Fn.dropout_stream.advance_on_device(dev)
        y = m(x, x, x, None)
        y.square().sum().backward()
    outs = []
    for _ in range(3):
        g.replay()
        outs.append(y.clone())
    torch.cuda.synchronize()
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("block", ["mha", "pwff"])
def test_training_dropout_backward_matches_forward(dev, block):
    """With p=0.1 the backward must use exactly the forward masks: directional derivative check in fp32."""
    D = 128
    torch.manual_seed(0)
    if block == "mha":
        m = ovqa.MultiHeadAttention(H.attention_cfg(D, 4, dropout=0.1, compute_dtype="fp32")).to(dev).train()
        fwd = lambda t: m(t, t, t, None)
    else:
        m = ovqa.PositionWiseFeedForward(H.attention_cfg(D, 4, d_ff=256, dropout=0.1, compute_dtype="fp32")).to(dev).train()
        fwd = lambda t: m(t)
    x = torch.randn((2, 9, D), device=dev)
    w = torch.randn((2, 9, D), device=dev)
    d = torch.randn((2, 9, D), device=dev)

    def run(t):
        Fn.dropout_stream._seed_of = None        # restart the counter stream: identical masks on every evaluation
        return (fwd(t) * w).sum()

    xg = x.clone().requires_grad_(True)
    run(xg).backward()
    analytic = (xg.grad * d).sum().item()
    eps = 1e-2
    with torch.no_grad():
        numeric = ((run(x + eps * d) - run(x - eps * d)) / (2 * eps)).item()
    assert abs(analytic - numeric) <= 2e-2 * max(1.0, abs(numeric)), (analytic, numeric)
    # and dropout really is active in train mode, inactive in eval
    with torch.no_grad():
        Fn.dropout_stream._seed_of = None
        a = fwd(x)
        b = fwd(x)           # stream advanced -> different mask
        assert (a != b).any()
        m.eval()
        assert torch.equal(fwd(x), fwd(x))


# ------------------------------------------------------------------------------------------------ LayerNorm kernels
@pytest.mark.parametrize("D", [96, 512, 768])
def test_layernorm_fwd_bwd(dev, D):
    g = torch.Generator().manual_seed(D)
    x, r = torch.randn((37, D), generator=g) * 3 + 1, torch.randn((37, D), generator=g)
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    xs = (x + r).clone().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xs, (D,), gm, bt, 1e-5)
    dy = torch.randn((37, D), generator=g)
    ref.backward(dy)
    y, y16, s, mean, rstd = ops.layernorm_fwd(x.to(dev), gamma.to(dev), beta.to(dev), residual=r.to(dev), want_bf16=True,
                                              save_stats=True, save_sum=True)
    assert rel_err(y, ref) < 1e-5 and abs_err(y16, ref.detach().bfloat16()) <= 0.04 and rel_err(s, x + r) < 1e-6
    dx, _, dg, db = ops.layernorm_bwd(dy.to(dev), s, mean, rstd, gamma.to(dev))
    assert rel_err(dx, xs.grad) < 1e-4 and rel_err(dg, gm.grad) < 1e-4 and rel_err(db, bt.grad) < 1e-4
    # fused bias gradient of the producing GEMM: column sums of the dropout-masked dx (and the bf16 copy carries the same mask)
    p, seed, off = 0.25, 77, 1000
    keep = ops.dropout_keep_mask(37 * D, p, seed, off, dev).view(37, D)
    dx2, dx16, dg2, db2, dbias = ops.layernorm_bwd(dy.to(dev), s, mean, rstd, gamma.to(dev), want_bf16=True, dropout_p=p, seed=seed, offset=off,
                                                   want_dbias=True)
    masked = dx.to(dev) * keep / (1 - p)
    assert torch.equal(dx2, dx) and rel_err(dg2, dg) < 1e-6 and rel_err(db2, db) < 1e-6
    assert rel_err(dbias, masked.sum(0)) < 1e-4 and abs_err(dx16, masked.bfloat16()) <= 0.02 * masked.abs().max().item()
    # eps of the BERT layers
    y2, *_ = ops.layernorm_fwd(x.to(dev), gamma.to(dev), beta.to(dev), eps=1e-12)
    assert rel_err(y2, F.layer_norm(x, (D,), gamma, beta, 1e-12)) < 1e-5


def test_layernorm_bwd_many_rows_persistent(dev):
    """Enough rows that every warp of the persistent backward loops (cp.async double buffer) and many CTAs meet in the
    column-sum atomics."""
    rows, D = 9001, 512
    g = torch.Generator().manual_seed(5)
    x = (torch.randn((rows, D), generator=g) * 2 + 0.5).to(dev)
    gamma, beta, dy = torch.randn(D, generator=g).to(dev), torch.randn(D, generator=g).to(dev), torch.randn((rows, D), generator=g).to(dev)
    xs, gm, bt = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xs, (D,), gm, bt, 1e-5).backward(dy)
    _, _, _, mean, rstd = ops.layernorm_fwd(x, gamma, beta, save_stats=True)
    dx, dx16, dg, db, dbias = ops.layernorm_bwd(dy, x, mean, rstd, gamma, want_bf16=True, want_dbias=True)
    assert rel_err(dx, xs.grad) < 1e-4 and rel_err(dg, gm.grad) < 1e-4 and rel_err(db, bt.grad) < 1e-4
    assert rel_err(dbias, xs.grad.sum(0)) < 1e-3 and abs_err(dx16, xs.grad.bfloat16()) <= 0.02 * xs.grad.abs().max().item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("m,n", [(1, 8), (37, 96), (5000, 1536), (25600, 512), (333, 100)])
def test_colsum(dev, dtype, m, n):
    g = torch.Generator().manual_seed(m + n)
    x = torch.randn((m, n), generator=g).to(dev).to(dtype)
    want = x.double().sum(0)
    tol = 1e-5 * (x.double().abs().sum(0).max().item() + 1.0)
    out = ops.colsum(x)
    assert (out.double() - want).abs().max().item() <= tol
    out2 = ops.colsum(x, out=out.clone(), accumulate=True)
    assert (out2.double() - 2 * want).abs().max().item() <= 2 * tol
    if n % 16 == 0:                                   # a column slice of a wider buffer (row pitch != n), as the fused QKV gradients are
        xs = x[:, n // 2:]
        assert (ops.colsum(xs).double() - want[n // 2:]).abs().max().item() <= tol


# ------------------------------------------------------------------------------------------------ M4C kernels
def test_m4c_pointer_prev_pred_and_greedy_kernels(dev, golden):
    c = C.M4C_SMALL
    g = golden("m4c")
    p = S.synth_params(g["spec"], 0)
    x = C.m4c_inputs(c)
    enc = R.m4c_encode_inputs(p, x)
    V, Dh, T = c["vocab"], c["hidden"], c["T"]
    prev = g["step"]["prev_inds"]
    pp = R.sub(p, "mmt.prev_pred_embeddings")
    d = lambda t: t.to(dev).contiguous()
    out = ops.prev_pred_embed(d(p["classifier.weight"]), d(enc[3]), d(prev), d(pp["position_embeddings.weight"]),
                              d(pp["token_type_embeddings.weight"]),
                              (d(pp["ans_layer_norm.weight"]), d(pp["ans_layer_norm.bias"])),
                              (d(pp["ocr_layer_norm.weight"]), d(pp["ocr_layer_norm.bias"])),
                              (d(pp["emb_layer_norm.weight"]), d(pp["emb_layer_norm.bias"])), 1e-12)
    assert rel_err(out, g["parts"]["dec_emb"]) <= FP32_REL
    # pointer network module (projections through ovqa_gemm + fused score kernel), written in place after V columns
    ptr = ovqa.OcrPtrNet(Dh, Dh)
    ptr.load_state_dict(R.sub(p, "ocr_ptr_net"))
    ptr = ptr.to(dev).eval()
    scores = torch.zeros((c["B"], T, V + c["Nocr"]), device=dev)
    with torch.no_grad():
        ptr(d(g["step"]["mmt_dec_output"]), d(g["step"]["mmt_ocr_output"]), d(enc[4]), scores_out=scores, col_offset=V)
    ref = g["parts"]["ptr_scores"]
    got = scores[..., V:].cpu()
    valid = ref > -1e4
    assert (got[valid] - ref[valid]).abs().max() <= 1e-4 * ref[valid].abs().max()
    assert (got[~valid] < -9e4).all()
    assert (scores[..., :V] == 0).all()
    # greedy step against the oracle's loop body on the golden step scores
    sc = g["step"]["scores"]
    prev_inds = prev.clone()
    last = torch.zeros(c["B"])
    arg = sc.argmax(-1)
    prev_ref = prev_inds.clone(); prev_ref[:, 1:] = arg[:, :-1]
    step = 2
    last_ref = torch.where(last == 2, last, arg[:, step].float())
    pi, li = d(prev_inds), torch.zeros(c["B"], dtype=torch.long, device=dev)
    am = torch.empty((c["B"], T), dtype=torch.long, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.greedy_step(d(sc), pi, li, am, flag, step, 2)
    assert torch.equal(am.cpu(), arg) and torch.equal(pi.cpu(), prev_ref) and torch.equal(li.cpu().float(), last_ref)
    assert flag.item() == int((last_ref == 2).all())
    # ties resolve to the first index, like torch.argmax
    tie = torch.zeros((1, 1, 40), device=dev); tie[0, 0, 7] = 1; tie[0, 0, 33] = 1
    am1 = torch.empty((1, 1), dtype=torch.long, device=dev)
    ops.greedy_step(tie, torch.zeros((1, 1), dtype=torch.long, device=dev), torch.zeros(1, dtype=torch.long, device=dev), am1, flag, 0, 2)
    assert am1.item() == 7


# ------------------------------------------------------------------------------------------------ host-side caches
def test_weight_cache_follows_optimizer_updates(dev):
    """fc_q/k/v stay separate parameters; their fused bf16 copy must refresh after every (fused) optimiser step."""
    torch.manual_seed(0)
    m = ovqa.MultiHeadAttention(H.attention_cfg(128, 4, dropout=0.0, compute_dtype="bf16")).to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-1, fused=True)
    x = torch.randn((2, 9, 128), device=dev)
    y0 = m(x, x, x, None)
    y0.sum().backward()
    opt.step()
    with torch.no_grad():
        y1 = m(x, x, x, None)
        ref = ovqa.MultiHeadAttention(H.attention_cfg(128, 4, dropout=0.0, compute_dtype="bf16")).to(dev)
        ref.load_state_dict(m.state_dict())
        y2 = ref(x, x, x, None)
    assert (y1 - y0).abs().max() > 1e-3          # the step changed the output ...
    assert torch.equal(y1, y2)                   # ... exactly as a freshly built module with the new weights sees it


# ------------------------------------------------------------------------------------------------ stateful decode
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_decoder_layer_stepwise_with_kv_cache(dev, golden, mode):
    """DecoderLayer (decoders.py:13-26) over the whole sequence and one token at a time inside ``statefulness()``: the inference
    path keeps PROJECTED keys / values (self-attention: appended per step and re-gathered with the beams; encoder attention:
    projected once) and must reproduce the reference, which re-projects everything at every step."""
    c, g = C.DEC_CASE, golden("decoder")
    att = lambda st: H.attention_cfg(c["D"], c["h"], d_ff=c["d_ff"], stateful=st, compute_dtype=mode)
    layer = load(H.DecoderLayer(H.NS(SELF_ATTENTION=att(True), ENC_ATTENTION=att(False))), S.synth_params(g["spec"], 0), dev)
    x = C.dec_inputs()
    d = lambda t: t.to(dev)
    with torch.no_grad():
        check(layer(d(x["x"]), d(x["enc"]), d(x["enc"]), d(x["full_mask"]), d(x["enc_mask"])), g["full"], mode, "decoder layer, whole sequence")
    per_step, perm = C.dec_stepwise_inputs(x)

    def run(grad):
        outs = []
        with torch.set_grad_enabled(grad), layer.statefulness(c["B"]):
            running_mask, enc_prev = torch.zeros((c["B"], 1, 1, 0), device=dev), None
            for t, (xt, mt, enc, enc_mask) in enumerate(per_step):
                running_mask = torch.cat([running_mask, d(mt)], -1)
                if enc_prev is None or not torch.equal(enc_prev[0].cpu(), enc):       # (the same tensor object across steps, as beam search passes it)
                    enc_prev = (d(enc), d(enc_mask))
                outs.append(layer(d(xt), enc_prev[0], enc_prev[0], running_mask, enc_prev[1]))
                if t == c["regather_at"]:
                    layer.apply_to_states(lambda s_: s_[d(perm)])
                    running_mask = running_mask[d(perm)]
            sa = layer.self_attn
            if grad:      # reference bookkeeping: raw keys accumulate, nothing cached
                assert sa.running_keys.shape == (c["B"], c["T"], c["D"]) and sa._kv_len == 0 and layer.enc_attn._enc_kv is None
            else:         # inference: projected rows accumulate, encoder keys projected once per tensor
                assert sa._kv_len == c["T"] and sa.running_kv_proj.shape[0] == c["B"] and sa.running_keys.shape[1:] == (1, 1)
                assert layer.enc_attn._enc_kv is not None
        assert layer.self_attn.running_kv_proj.shape == (1, 2 * c["D"]) and layer.self_attn._kv_len == 0 and layer.enc_attn._enc_kv is None
        assert layer.self_attn.running_keys.shape == (0, c["D"])
        return torch.cat(outs, 1)

    cached, recomputed = run(False), run(True)
    check(cached, g["steps"], mode, "stepwise decode, projected K/V cache")
    check(recomputed, g["steps"], mode, "stepwise decode, reference bookkeeping")
    # the two paths run the same kernels on the same rows; only the GEMM batch differs
    assert abs_err(cached, recomputed) <= (1e-5 if mode == "fp32" else 2e-2)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("D,h", [(128, 4), (192, 2)])
def test_mha_decode_step_call(dev, mode, D, h):
    """ovqa_mha_decode_step on its own: multi-row prefixes, single rows, appended rows from another source, per-query masks — each
    step against the oracle block over the concatenated RAW inputs (what the reference computes, attentions.py:321-331)."""
    dk = D // h
    p = S.synth_params(C.att_spec(D, h, dk, dk), 3)
    P = {k: v.to(dev) for k, v in p.items()}
    W = dict(wq=P["attention.fc_q.weight"], bq=P["attention.fc_q.bias"], wk=P["attention.fc_k.weight"], bk=P["attention.fc_k.bias"],
             wv=P["attention.fc_v.weight"], bv=P["attention.fc_v.bias"], wo=P["attention.fc_o.weight"], bo=P["attention.fc_o.bias"],
             gamma=P["layer_norm.weight"], beta=P["layer_norm.bias"], mode=mode)
    plan = Fn.mha_decode_plan(**W)
    g = S._gen(5, f"decode{D}")
    B, cap = 3, 16
    cache = torch.full((B, cap, 2 * D), float("nan"), device=dev, dtype=torch.float32 if mode == "fp32" else torch.bfloat16)
    raw, length = [], 0
    schedule = [("self", 3), ("self", 1), ("cross", 2), ("self", 2), ("none", 1)]
    for kind, n in schedule:
        q = torch.randn((B, n if kind == "self" else 2, D), generator=g)
        src = q if kind == "self" else (torch.randn((B, n, D), generator=g) if kind == "cross" else None)
        if src is not None:
            raw.append(src)
        keys = torch.cat(raw, 1)
        Lk = keys.shape[1]
        mask = (torch.rand((B, 1, q.shape[1], Lk), generator=g) < 0.25).float() * -10e4
        mask[..., 0] = 0.0
        ref = R.mha(p, q, keys, keys, mask, h=h, d_k=dk, d_v=dk)
        y = Fn.mha_decode_block(plan, q.to(dev), None if src is None else src.to(dev), kind == "self", cache, length, mask.to(dev),
                                h=h, d_k=dk, d_v=dk)
        length = Lk
        check(y, ref, mode, f"decode step {kind}/{n} at length {Lk}")
    assert torch.isfinite(cache[:, :length].float()).all() and torch.isnan(cache[:, length:].float()).all()     # filed exactly the valid rows


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_pwff_inference_call_matches_the_block(dev, mode):
    torch.manual_seed(0)
    m = ovqa.PositionWiseFeedForward(H.attention_cfg(128, 4, d_ff=256, dropout=0.1, compute_dtype=mode)).to(dev).eval()
    x = torch.randn((3, 5, 128), device=dev)
    with torch.no_grad():
        y_call = m(x)                                    # one library call
    y_block = m(x)                                       # grad mode: the autograd block (same kernels)
    assert abs_err(y_call, y_block) <= 1e-6
    m.train()
    with torch.no_grad():
        assert abs_err(m(x), y_block) > 1e-3             # training-mode dropout still applies under no_grad


def test_encoder_projection_follows_beam_regather(dev, monkeypatch):
    """Beam search re-gathers the model's encoder features AND the attention blocks' states with one function per step
    (beam_search.py:19-33, containers.py:26-31).  The projected encoder keys must be re-gathered, not re-projected — and must be
    re-projected whenever the keys are NOT provably that function's output."""
    torch.manual_seed(0)
    layer = H.DecoderLayer(H.NS(SELF_ATTENTION=H.attention_cfg(128, 4, d_ff=256, stateful=True, compute_dtype="bf16"),
                                ENC_ATTENTION=H.attention_cfg(128, 4, d_ff=256, compute_dtype="bf16"))).to(dev).eval()
    calls = []
    real = Fn.project_kv_fused
    monkeypatch.setattr(Fn, "project_kv_fused", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    B, beam, N = 2, 3, 9
    enc = torch.randn((B, N, 128), device=dev)
    xs = [torch.randn((B * beam, 1, 128), device=dev) for _ in range(4)]

    class Model(ovqa.StatefulModule):                      # the owner of ``encoder_features``, as base_transformer.py:17
        def __init__(self):
            super().__init__()
            self.layer = layer
            self.register_state("encoder_features", None)

    model = Model()
    with torch.no_grad():
        with model.statefulness(B):
            model.encoder_features = enc
            expand = lambda s: s.repeat_interleave(beam, 0)                         # step 0: every sample fans out into its beams
            perm = torch.tensor([1, 0, 2, 5, 4, 3], device=dev)
            outs, keys_seen = [], []
            x0 = torch.randn((B, 1, 128), device=dev)
            outs.append(layer(x0, model.encoder_features, model.encoder_features, None, None))
            model.apply_to_states(expand)
            for t, x in enumerate(xs):
                keys_seen.append(model.encoder_features)
                outs.append(layer(x, model.encoder_features, model.encoder_features, None, None))
                model.apply_to_states(lambda s: s[perm])
            assert len(calls) == 1                                                    # projected once for the whole search
            # keys that did NOT come out of the sweep (same shape, other contents): projected again
            other = torch.randn_like(model.encoder_features)
            y_other = layer(xs[0], other, other, None, None)
            assert len(calls) == 2
        # reference bookkeeping on the same inputs (grad mode: no caches) gives the same numbers
        ref_layer_out = []
    with layer.statefulness(B):
        with torch.enable_grad():
            r0 = layer(x0, enc, enc, None, None)
            layer.apply_to_states(expand)
            e = expand(enc)
            for t, x in enumerate(xs):
                ref_layer_out.append(layer(x, e, e, None, None))
                layer.apply_to_states(lambda s: s[perm])
                e = e[perm]
    assert abs_err(outs[0], r0) <= BF16_ABS
    for a, b in zip(outs[1:], ref_layer_out):
        assert abs_err(a, b) <= BF16_ABS


def test_encoder_kv_cache_follows_in_place_updates(dev):
    """The once-per-session projection of the encoder keys is keyed on the tensor's identity AND version."""
    torch.manual_seed(0)
    m = ovqa.MultiHeadAttention(H.attention_cfg(128, 4, compute_dtype="bf16")).to(dev).eval()
    q, enc = torch.randn((2, 1, 128), device=dev), torch.randn((2, 9, 128), device=dev)
    with torch.no_grad():
        ref0 = m(q, enc, enc, None)
        with m.statefulness(2):
            y0 = m(q, enc, enc, None)
            tag0 = m._enc_kv[0]
            y0b = m(q, enc, enc, None)
            assert m._enc_kv[0] == tag0                        # second step: cache hit
            enc.mul_(-1.0)                                     # same storage, new contents
            y1 = m(q, enc, enc, None)
            assert m._enc_kv[0] != tag0
        ref1 = m(q, enc, enc, None)
    # (the one-call decode path scores with fp32 probabilities, the block's tensor-core kernel rounds them to bf16)
    assert torch.equal(y0, y0b) and abs_err(y0, ref0) <= BF16_ABS and abs_err(y1, ref1) <= BF16_ABS
    assert abs_err(y1, y0) > 10 * BF16_ABS


# ------------------------------------------------------------------------------------------------ beam selection
@pytest.mark.parametrize("name", list(C.BEAM_CASES))
def test_beam_select_matches_reference(dev, golden, name):
    """BeamSearch.select (beam_search.py:35-38): identical candidate indices and log-probabilities (index work: bit-exact)."""
    b_s, cur, V, beam = C.BEAM_CASES[name]
    idx, logprob = ovqa.beam_select(C.beam_inputs(name).to(dev), b_s, beam)
    g = golden("beam")[name]
    assert torch.equal(idx.cpu(), g["idx"]) and torch.equal(logprob.cpu(), g["logprob"])


@pytest.mark.parametrize("rows,n,k", [(1, 5, 5), (7, 1000, 1), (64, 50000, 5), (3, 200000, 16), (300, 33, 8)])
def test_topk_rows_against_sort(dev, rows, n, k):
    torch.manual_seed(rows * 1000 + k)
    x = torch.randn((rows, n), device=dev)
    x[0, n // 2] = float("nan")                          # NaN sorts first, as torch.sort(descending=True)
    x[-1, : min(n, 3)] = 7.5                             # ties: ascending index
    val, idx = ops.topk_rows(x, k)
    ref_val, ref_idx = torch.sort(x, dim=-1, descending=True, stable=True)
    assert torch.equal(idx, ref_idx[:, :k])
    assert torch.equal(torch.nan_to_num(val, nan=123.0), torch.nan_to_num(ref_val[:, :k], nan=123.0))
    # a strided view (row pitch larger than n)
    big = torch.randn((rows, n + 8), device=dev)
    v2, i2 = ops.topk_rows(big[:, :n], k)
    assert torch.equal(i2, torch.sort(big[:, :n], dim=-1, descending=True, stable=True)[1][:, :k])


# ------------------------------------------------------------------------------------------------ beam search end to end
def _captioner(mode, dev, g):
    c = C.GEN_CASE
    att = lambda st: H.attention_cfg(c["D"], c["h"], d_ff=c["d_ff"], stateful=st, compute_dtype=mode)
    cfg = H.NS(D_MODEL=c["D"], LAYERS=c["layers"], ATTENTION=H.NS(SELF_ATTENTION=att(True), ENC_ATTENTION=att(False)))
    m = H.Captioner(cfg, H.StubVocab(c["vocab"], max_answer_length=c["T"]))
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items() if v.dtype.is_floating_point] == g["spec"]
    m.load_state_dict(S.synth_params(g["spec"], 0), strict=False)
    m.eos_idx = c["eos"]
    return m.to(dev).eval()


@pytest.mark.parametrize("beam", [1, C.GEN_CASE["beam"]])
def test_beam_search_token_sequences_match_the_reference(dev, golden, beam, monkeypatch):
    """The reference's beam search over a stateful decoder (base_transformer.py:30-54, beam_search.py, decoders.py:28-76; golden
    vectors made by running it on CPU) against the same search through the decode path: projected-K/V caches, one-call blocks,
    top-k selection, provenance-tracked re-gathers.  fp32: IDENTICAL token sequences, log-probabilities to 1e-4."""
    g, c = golden("generative"), C.GEN_CASE
    x = {k: v.to(dev) for k, v in C.gen_inputs().items()}
    m = _captioner("fp32", dev, g)
    calls = []
    real = Fn.project_kv_fused
    monkeypatch.setattr(Fn, "project_kv_fused", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    with torch.no_grad():
        tokens, logp = m.beam_search(x["enc"], x["enc_mask"], c["B"], beam)
    assert torch.equal(tokens.cpu(), g[f"beam{beam}"]["tokens"])
    assert rel_err(logp, g[f"beam{beam}"]["logp"]) <= FP32_REL
    assert len(calls) == c["layers"]                       # the encoder output was projected once per layer for the whole search
    # with gradients enabled (self-critical training) the reference's bookkeeping runs instead: same sequences
    tokens_g, logp_g = m.beam_search(x["enc"], x["enc_mask"], c["B"], beam)
    assert torch.equal(tokens_g.cpu(), g[f"beam{beam}"]["tokens"]) and logp_g.requires_grad
    assert rel_err(logp_g, g[f"beam{beam}"]["logp"]) <= FP32_REL
    # bf16: same machinery; the sequences may legitimately differ where two continuations are within bf16 noise, the scores may not drift
    mb = _captioner("bf16", dev, g)
    with torch.no_grad():
        tb, lb = mb.beam_search(x["enc"], x["enc_mask"], c["B"], beam)
    same = (tb.cpu() == g[f"beam{beam}"]["tokens"]).all(-1)
    assert tb.shape == (c["B"], c["T"]) and torch.isfinite(lb).all()
    if same.any():
        assert abs_err(lb.cpu()[same], g[f"beam{beam}"]["logp"][same]) <= 5e-2


def test_decoder_whole_sequence_scoring(dev, golden):
    """decoders.py:47-76 as the training-time scorer (teacher forcing, causal + padding mask) — the stateless twin of the search."""
    g, c = golden("generative"), C.GEN_CASE
    x = {k: v.to(dev) for k, v in C.gen_inputs().items()}
    toks = torch.cat([torch.full((c["B"], 1), 1, dtype=torch.long), g[f"beam{c['beam']}"]["tokens"][:, :-1]], 1).to(dev)
    for mode in ("fp32", "bf16"):
        m = _captioner(mode, dev, g)
        with torch.no_grad():
            check(m.decoder(toks, x["enc"], x["enc_mask"]), g["teacher_forced"], mode, f"decoder scoring {mode}", bf16_abs=5e-2)


# ------------------------------------------------------------------------------------------------ empty inputs
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_empty_inputs(dev, dt):
    """A batch of zero rows goes through every wrapper and returns correctly shaped empty results (what torch's own ops do)."""
    a, w = torch.empty((0, 64), device=dev, dtype=dt), torch.randn((32, 64), device=dev).to(dt)      # w: Linear layout (n, k)
    assert ops.gemm(a, w).shape == (0, 32)
    assert ops.gemm(w, a).shape == (32, 0)
    dw = ops.gemm(torch.empty((0, 32), device=dev, dtype=dt), a, a_op="T", b_op="T", out_dtype=torch.float32)   # weight gradient over zero rows
    assert dw.shape == (32, 64) and float(dw.abs().sum()) == 0.0
    q = torch.empty((0, 5, 64), device=dev, dtype=dt)
    out, lse, _ = ops.sdpa_fwd(q, q, q, h=4, d_k=16, d_v=16, need_lse=True)
    assert out.shape == (0, 5, 64) and lse.shape == (0, 4, 5)
    x = torch.empty((0, 7, 64), device=dev)
    g, b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    y, y16, _, mean, rstd = ops.layernorm_fwd(x, g, b, want_bf16=True, save_stats=True)
    assert y.shape == x.shape and y16.shape == x.shape and mean.numel() == 0
    dx, _, dg, db = ops.layernorm_bwd(x, x, mean, rstd, g)
    assert dx.shape == x.shape and float(dg.abs().sum()) == 0.0 and float(db.abs().sum()) == 0.0
    assert float(ops.colsum(torch.empty((0, 64), device=dev)).abs().sum()) == 0.0
    assert ops.to_bf16(torch.empty((0, 8), device=dev)).shape == (0, 8)
    yb, m = ops.cast_rows_padmask(torch.empty((0, 64), device=dev), -1e4)
    assert yb.shape == (0, 64) and m.shape == (0,)
    # the module level: an empty batch through a training-mode block, forward and backward
    torch.manual_seed(0)
    mha = ovqa.MultiHeadAttention(H.attention_cfg(64, 4, dropout=0.1, compute_dtype="bf16" if dt == torch.bfloat16 else "fp32")).to(dev).train()
    xe = torch.empty((0, 5, 64), device=dev, requires_grad=True)
    ye = mha(xe, xe, xe, None)
    assert ye.shape == (0, 5, 64)
    ye.sum().backward()
    assert xe.grad.shape == xe.shape
    for p in mha.parameters():
        assert p.grad is None or float(p.grad.abs().sum()) == 0.0


# ------------------------------------------------------------------------------------------------ geometry bias as a tensor (bf16 route)
@pytest.mark.parametrize("name", list(C.GEO_CASES))
def test_geo_bias_kernels_against_the_oracle(dev, name):
    """ovqa_geo_bias_fwd against attentions.py:113-119,131-132 (restated), ovqa_geo_bias_bwd against autograd of the same."""
    from openvivqa_b200 import ops
    B, N, D, h, trig = C.GEO_CASES[name]
    x = C.geo_inputs(name)
    dg = x["d_g"]
    params = S.synth_params([(f"fc_gs.{i}.{t}", (1, dg) if t == "weight" else (1,)) for i in range(h) for t in ("weight", "bias")], 3)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref = R.geometry_log_bias(pr, x["boxes"], h=h, d_g=dg, trig=trig)
    gsel = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    # d log(raw) / d raw = 1 / raw reaches 1e6 next to the clamp, where raw itself is only known to ~1e-5 (sin / cos at hundreds of radians):
    # the loss is taken on exp(bias) = clamp(relu(raw), 1e-6), i.e. the kernel is handed dbias = gsel * exp(bias), which keeps it conditioned
    (ref.exp() * gsel).sum().backward()
    w = torch.cat([params[f"fc_gs.{i}.weight"] for i in range(h)], 0).to(dev)
    b = torch.cat([params[f"fc_gs.{i}.bias"] for i in range(h)], 0).to(dev)
    geo = ops.GeoParams(x["boxes"].to(dev), w, b, trig)
    key_mask = torch.zeros((B, 1, 1, N), device=dev)
    key_mask[:, :, :, -3:] = -10e4
    out = ops.geo_bias(geo, h)
    # compared as clamp(relu(raw), 1e-6) = exp(bias): the log amplifies the last-bit differences of sin / cos at arguments of several
    # hundred radians wherever the 64-term sum nearly cancels (raw ~ 1e-3 -> 2e-2 on the log)
    torch.testing.assert_close(out.cpu().exp(), ref.detach().exp(), rtol=1e-3, atol=5e-5)
    assert ((out.cpu() - ref.detach()).abs() > 1e-3).float().mean().item() < 1e-3
    masked = ops.geo_bias(geo, h, key_mask=key_mask).cpu()
    torch.testing.assert_close(masked[..., :-3], out.cpu()[..., :-3], rtol=0, atol=0)
    assert (masked[..., -3:] < -9e4).all()
    dw, db = torch.zeros((h, dg), device=dev), torch.zeros((h,), device=dev)
    ops.geo_bias_bwd(geo, h, (gsel * ref.detach().exp()).to(dev).contiguous(), dw, db)
    dw_ref = torch.cat([pr[f"fc_gs.{i}.weight"].grad for i in range(h)], 0)
    db_ref = torch.cat([pr[f"fc_gs.{i}.bias"].grad for i in range(h)], 0)
    assert rel_err(dw, dw_ref) <= 2e-3 and rel_err(db, db_ref) <= 2e-3, (rel_err(dw, dw_ref), rel_err(db, db_ref))


def test_geometry_gradients_bf16_route(dev):
    """The MHA block with the geometry core in bf16 mode (bias tensor + tensor-core attention kernels + dS -> fc_g gradients)."""
    D, h = 128, 4
    B, N = 3, 37
    params = S.synth_params(C.att_spec(D, h, D // h, D // h, geometry_dg=D // h), 4)
    x = S.padded_features(B, N, D, 3, min_valid=9, relu_like=False)
    mask = R.padding_mask(x, 0)
    boxes = S.boxes(B, N, 5)
    f = lambda p, t: R.mha(p, t, t, t, mask, h=h, d_k=D // h, d_v=D // h, boxes=boxes, d_g=D // h, trig=True)
    out, w, gp, gx = _oracle_grads(f, params, [x])
    m = ovqa.MultiHeadAttention(H.attention_cfg(D, h, dropout=0.0, arch="AugmentedGeometryScaledDotProductAttention", compute_dtype="bf16"))
    m.load_state_dict(params)
    m = m.to(dev).train()
    xd = x.to(dev).requires_grad_(True)
    y = m(xd, xd, xd, mask.to(dev), boxes=boxes.to(dev))
    (y * w.to(dev)).sum().backward()
    check(y, out, "bf16", "geometry block bf16")
    named = dict(m.named_parameters())
    for i in range(h):
        g, gr = named[f"attention.fc_gs.{i}.weight"].grad.cpu(), gp[f"attention.fc_gs.{i}.weight"]
        assert float((g - gr).norm() / gr.norm().clamp_min(1e-9)) <= 6e-2, (i, g, gr)
    g, gr = named["attention.fc_q.weight"].grad.cpu(), gp["attention.fc_q.weight"]
    assert float((g - gr).norm() / gr.norm()) <= 6e-2


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("kind", ["self", "kv_same", "three"])
def test_one_call_inference_block_matches_the_per_op_block(dev, mode, kind):
    """ovqa_mha_block_infer (eval mode, no autograd: one library call) against the per-op block it replaces — same kernels, same bits."""
    from openvivqa_b200 import functional as Fn
    D, h, B, Lq, Lk = 128, 4, 5, 21, 37
    m = ovqa.MultiHeadAttention(H.attention_cfg(D, h, dropout=0.1, compute_dtype=mode))
    m.load_state_dict(S.synth_params(C.att_spec(D, h, D // h, D // h), 2))
    m = m.to(dev).eval()
    g = S._gen(4, "onecall")
    q = torch.randn((B, Lq, D), generator=g).to(dev)
    k = q if kind == "self" else torch.randn((B, Lk, D), generator=g).to(dev)
    v = k if kind != "three" else torch.randn((B, Lk, D), generator=g).to(dev)
    mask = torch.zeros((B, 1, 1, k.shape[1]), device=dev)
    mask[:, :, :, -5:] = -10e4
    with torch.no_grad():
        assert Fn.ONE_CALL_INFERENCE
        y1 = m(q, k, v, mask)
        Fn.ONE_CALL_INFERENCE = False
        try:
            y0 = m(q, k, v, mask)
        finally:
            Fn.ONE_CALL_INFERENCE = True
    assert torch.equal(y1, y0)
    if mode == "bf16":
        assert torch.equal(y1._ovqa_bf16[1], y0._ovqa_bf16[1])
    pw = ovqa.PositionWiseFeedForward(H.attention_cfg(D, h, d_ff=256, dropout=0.1, compute_dtype=mode)).to(dev).eval()
    with torch.no_grad():
        z1 = Fn.pwff_block(y1, w1=pw.fc1.weight, b1=pw.fc1.bias, w2=pw.fc2.weight, b2=pw.fc2.bias, gamma=pw.layer_norm.weight, beta=pw.layer_norm.bias, mode=mode)
        Fn.ONE_CALL_INFERENCE = False
        try:
            z0 = Fn.pwff_block(y0, w1=pw.fc1.weight, b1=pw.fc1.bias, w2=pw.fc2.weight, b2=pw.fc2.bias, gamma=pw.layer_norm.weight, beta=pw.layer_norm.bias, mode=mode)
        finally:
            Fn.ONE_CALL_INFERENCE = True
    assert torch.equal(z1, z0)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("p_drop", [0.0, 0.1])
def test_attribute_reduce_pooling_head(dev, mode, p_drop):
    """SURVEY 8f rank 3 (models/mcan.py:21-25,70-77): ReLU + dropout GEMM epilogue and the pooling kernels, forward and backward, against
    autograd of the reference composition (with dropout: the composition applies the library's own keep mask)."""
    from openvivqa_b200 import functional as Fn, ops
    B, N, D = 7, 100, 512
    g = torch.Generator().manual_seed(3)
    params = S.synth_params([("fc1.weight", (D, D)), ("fc1.bias", (D,)), ("fc2.weight", (1, D)), ("fc2.bias", (1,))], 2)
    x = torch.randn((B, N, D), generator=g)
    torch.manual_seed(0)
    Fn.dropout_stream.offset = 0
    pi = {k: v.clone().to(dev).requires_grad_(True) for k, v in params.items()}
    xi = x.clone().to(dev).requires_grad_(True)
    y = Fn.attr_reduce_pool(xi, w1=pi["fc1.weight"], b1=pi["fc1.bias"], w2=pi["fc2.weight"], b2=pi["fc2.bias"], mode=mode, dropout=p_drop, training=True)
    # (the device-side offset base may have been advanced by an earlier CUDA-graph test of this process: read the mask with it)
    keep = ops.dropout_keep_mask(B * N * D, p_drop, Fn.dropout_stream.seed, 0, dev, offset_base=Fn.dropout_stream.base(dev)).view(B, N, D).cpu() \
        if p_drop > 0 else torch.ones(B, N, D)
    ks = keep / (1.0 - p_drop)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xr = x.clone().requires_grad_(True)
    hid = torch.relu(torch.nn.functional.linear(xr, pr["fc1.weight"], pr["fc1.bias"])) * ks
    att = torch.softmax(torch.nn.functional.linear(hid, pr["fc2.weight"], pr["fc2.bias"]), dim=1)
    ref = (xr * att).sum(dim=1)
    w = torch.randn(ref.shape, generator=g)
    (ref * w).sum().backward()
    (y * w.to(dev)).sum().backward()
    check(y, ref, mode, "pooled")
    tol = 2e-3 if mode == "fp32" else 6e-2
    assert rel_err(xi.grad, xr.grad) <= tol, rel_err(xi.grad, xr.grad)
    for k in params:
        if k == "fc2.bias":          # softmax gradients sum to zero over the sequence: the true gradient is 0, both sides hold rounding noise
            assert float(pi[k].grad.abs().max()) <= 1e-3 * float(w.abs().sum()) / N
            continue
        e = float((pi[k].grad.cpu() - pr[k].grad).norm() / pr[k].grad.norm().clamp_min(1e-12))
        assert e <= tol, (k, e)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_probability_dropout_in_the_attention_core(dev, mode):
    """Dropout on the attention probabilities (HF BertSelfAttention, the layers MMF_M4C trains through): forward and backward of the core
    against autograd of softmax(.) * keep-scale @ v, with the keep-scales read back from the library's own counter stream."""
    from openvivqa_b200 import functional as Fn, ops
    B, h, Lq, Lk, d, p = 3, 4, 37, 53, 32, 0.2
    g = torch.Generator().manual_seed(8)
    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    q, k, v = (torch.randn((B, L, h * d), generator=g).to(dt) for L in (Lq, Lk, Lk))
    mask = torch.zeros((B, 1, 1, Lk))
    mask[:, :, :, -7:] = -10e4
    seed, off = 1234, 77
    pd = (p, seed, off, None)
    qd, kd, vd = q.to(dev), k.to(dev), v.to(dev)
    out, lse, _ = ops.sdpa_fwd(qd, kd, vd, h=h, d_k=d, d_v=d, mask=mask.to(dev), need_lse=True, prob_dropout=pd)
    keep = ops.dropout_keep_mask(B * h * Lq * Lk, p, seed, off, dev).view(B, h, Lq, Lk).cpu() / (1.0 - p)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    sp = lambda t, L: t.view(B, L, h, d).permute(0, 2, 1, 3)
    att = torch.softmax(sp(qr, Lq) @ sp(kr, Lk).transpose(-1, -2) / math.sqrt(d) + mask, -1) * keep
    ref = (att @ sp(vr, Lk)).permute(0, 2, 1, 3).reshape(B, Lq, h * d)
    w = torch.randn(ref.shape, generator=g)
    (ref * w).sum().backward()
    check(out, ref, mode, "core with probability dropout")
    dq, dk, dv = torch.empty_like(qd), torch.empty_like(kd), torch.empty_like(vd)
    ops.sdpa_bwd(qd, kd, vd, out, w.to(dev).to(dt), lse, dq, dk, dv, h=h, d_k=d, d_v=d, mask=mask.to(dev), prob_dropout=pd)
    tol = 1e-3 if mode == "fp32" else 4e-2
    for got, want, name in ((dq, qr.grad, "dq"), (dk, kr.grad, "dk"), (dv, vr.grad, "dv")):
        e = float((got.float().cpu() - want).norm() / want.norm())
        assert e <= tol, (name, e)


def test_feature_embedding_accepts_bf16_staged_features(dev):
    """Region features staged as bf16 (pinned bf16 feed): same outputs, mask and gradients as the fp32 features they were rounded from."""
    from openvivqa_b200 import functional as Fn
    B, N, K, D = 3, 40, 256, 128
    feats = S.padded_features(B, N, K, 3, min_valid=7).to(torch.bfloat16).float()        # exactly representable in bf16
    m = ovqa.FeatureEmbedding(H.NS(D_FEATURE=K, D_MODEL=D, DROPOUT=0.0, COMPUTE_DTYPE="bf16")).to(dev).train()
    y32, m32 = m(feats.to(dev))
    y32.square().sum().backward()
    g32 = m.proj.weight.grad.clone()
    m.zero_grad()
    y16, m16 = m(feats.to(dev).to(torch.bfloat16))
    y16.square().sum().backward()
    assert torch.equal(y16, y32) and torch.equal(m16, m32)
    assert torch.equal(m.proj.weight.grad, g32)
    assert (m16 != 0).any() and (m16 == 0).any()
