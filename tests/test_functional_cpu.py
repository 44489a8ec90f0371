"""CPU: the hand-written backward passes and the module-level autograd plumbing of the drop-ins, checked against autograd of the
restated oracle.  Every library call is replaced by its plain-torch restatement (tests/cpu_ops.py), so what is under test is the HOST
logic: which tensors are saved, which gradients are returned for which inputs, the analytic bias-gradient shortcuts, the differentiable
probabilities of a direct core call (mmf_lorra.py:118-146) and the pointer network's gradients (mmf_m4c.py:379-396)."""
import math

import pytest
import torch

import openvivqa_b200 as ovqa
from openvivqa_b200 import functional as Fn, harness as H
from oracle import cases as C, restated as R, synth as S


def _grads(params):
    return {k: v.grad for k, v in params.items()}


def _close(a, b, tol=2e-4):
    torch.testing.assert_close(a, b, rtol=tol, atol=tol * max(1.0, float(b.abs().max())))


@pytest.mark.parametrize("case", ["sa_text_30", "guided_100x30", "micro_30x100", "dec_causal_13", "lorra_h1"])
def test_mha_block_backward_matches_oracle_autograd(cpu_ops, case):
    B, Lq, Lk, D, h, dk, dv, _, self_att = C.ATT_CASES[case]
    x = C.att_inputs(case)
    params = S.synth_params(C.att_spec(D, h, dk, dv), 3)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    qr = x["queries"].clone().requires_grad_(True)
    kr = qr if self_att else x["keys"].clone().requires_grad_(True)
    ref = R.mha(pr, qr, kr, kr, x["mask"], h=h, d_k=dk, d_v=dv)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (ref * w).sum().backward()

    m = ovqa.MultiHeadAttention(H.attention_cfg(D, h, dk, dv, dropout=0.0, compute_dtype="fp32")).train()
    m.load_state_dict(params, strict=True)
    q = x["queries"].clone().requires_grad_(True)
    k = q if self_att else x["keys"].clone().requires_grad_(True)
    y = m(q, k, k, x["mask"])
    (y * w).sum().backward()
    # a sample whose keys are ALL masked computes softmax over scores quantised by the fp32 add of -1e5 (steps of 2^-7): its
    # probabilities depend on the last bit of q.k, so only that case gets a loose bound
    tol = 1e-2 if case == "micro_30x100" else 2e-4
    _close(y.detach(), ref.detach(), 1e-5 if tol < 1e-3 else 1e-3)
    _close(q.grad, qr.grad, tol)
    if not self_att:
        _close(k.grad, kr.grad, tol)
    for name, p in m.named_parameters():
        assert p.grad is not None, name
        _close(p.grad, pr[name].grad, tol)


def test_pwff_block_backward_matches_oracle_autograd(cpu_ops):
    D, dff = 64, 128
    params = S.synth_params(C.pwff_spec(D, dff), 4)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    x = torch.randn((3, 7, D), generator=torch.Generator().manual_seed(1))
    xr = x.clone().requires_grad_(True)
    ref = R.pwff(pr, xr)
    ref.square().sum().backward()
    m = ovqa.PositionWiseFeedForward(H.attention_cfg(D, 4, d_ff=dff, dropout=0.0, compute_dtype="fp32")).train()
    m.load_state_dict(params, strict=True)
    xi = x.clone().requires_grad_(True)
    y = m(xi)
    y.square().sum().backward()
    _close(y.detach(), ref.detach(), 1e-5)
    _close(xi.grad, xr.grad)
    for name, p in m.named_parameters():
        _close(p.grad, pr[name].grad)


def test_dropout_masks_are_consistent_between_forward_and_backward(cpu_ops):
    """With p > 0 the backward pass must re-create the forward's masks from (seed, offset): the block's gradient has to match autograd
    through an explicit composition that applies the SAME keep-scales."""
    torch.manual_seed(0)
    D, h = 64, 4
    m = ovqa.MultiHeadAttention(H.attention_cfg(D, h, dropout=0.3, compute_dtype="fp32")).train()
    x = torch.randn(2, 9, D)
    q = x.clone().requires_grad_(True)
    Fn.dropout_stream.offset = 0
    y = m(q, q, q, None)
    y.sum().backward()
    # explicit composition: the mask of the fc_o output is keep-scale(seed, offset 0) over (B*L, D)
    seed = Fn.dropout_stream.seed
    ks = cpu_ops._keep_scale(2 * 9 * D, 0.3, seed, 0).view(2, 9, D)
    p = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    xr = x.clone().requires_grad_(True)
    out, _ = R.sdpa_core(R.sub(p, "attention"), xr, xr, xr, None, h=h, d_k=D // h, d_v=D // h)
    ref = torch.nn.functional.layer_norm(xr + out * ks, (D,), p["layer_norm.weight"], p["layer_norm.bias"])
    ref.sum().backward()
    _close(y.detach(), ref.detach(), 1e-5)
    _close(q.grad, xr.grad)
    _close(m.attention.fc_o.weight.grad, p["attention.fc_o.weight"].grad)
    _close(m.attention.fc_o.bias.grad, p["attention.fc_o.bias"].grad)


@pytest.mark.parametrize("use", ["att_only", "both"])
def test_direct_core_probabilities_are_differentiable(cpu_ops, use):
    """LoRRA's call pattern (mmf_lorra.py:127-146): ``_, att = core(q, k, v, mask)`` and features built FROM att."""
    B, Lq, Lk, D, h, dk, dv, _, _ = C.ATT_CASES["lorra_h1"]
    x = C.att_inputs("lorra_h1")
    params = {k[len("attention."):]: v for k, v in S.synth_params(C.att_spec(D, h, dk, dv), 3).items() if k.startswith("attention.")}
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    qr, kr = x["queries"].clone().requires_grad_(True), x["keys"].clone().requires_grad_(True)
    w_att = torch.randn((B, h, Lq, Lk), generator=torch.Generator().manual_seed(9))
    out_r, att_r = R.sdpa_core(pr, qr, kr, kr, x["mask"], h=h, d_k=dk, d_v=dv)
    loss_r = (att_r * w_att).sum() + (out_r.square().sum() if use == "both" else 0.0)
    loss_r.backward()

    core = ovqa.ScaledDotProductAttention(H.attention_cfg(D, h, dk, dv, compute_dtype="fp32"))
    core.load_state_dict(params, strict=True)
    q, k = x["queries"].clone().requires_grad_(True), x["keys"].clone().requires_grad_(True)
    out, att = core(q, k, k, x["mask"])
    assert att.requires_grad
    loss = (att * w_att).sum() + (out.square().sum() if use == "both" else 0.0)
    loss.backward()
    _close(att.detach(), att_r.detach(), 1e-5)
    _close(q.grad, qr.grad)
    _close(k.grad, kr.grad)
    for name in ("fc_q", "fc_k") + (("fc_v", "fc_o") if use == "both" else ()):
        _close(getattr(core, name).weight.grad, pr[name + ".weight"].grad)
    if use == "att_only":                       # the value / output projections do not influence the probabilities
        assert core.fc_o.weight.grad is None or float(core.fc_o.weight.grad.abs().max()) == 0.0


def test_direct_geometry_core_trains_its_geometry_weights(cpu_ops):
    name = "geo_raw4"
    B, N, D, h, trig = C.GEO_CASES[name]
    x = C.geo_inputs(name)
    spec = [(k, s) for k, s in C.att_spec(D, h, D // h, D // h, geometry_dg=x["d_g"]) if k.startswith("attention.")]
    params = {k[len("attention."):]: v for k, v in S.synth_params(spec, 0).items()}
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xr = x["x"].clone().requires_grad_(True)
    out_r, att_r = R.geometry_core(pr, xr, xr, xr, x["boxes"], None, h=h, d_k=D // h, d_v=D // h, d_g=x["d_g"], trig=trig)
    (out_r.square().sum() + att_r.square().sum()).backward()
    core = ovqa.AugmentedGeometryScaledDotProductAttention(H.attention_cfg(D, h, arch="AugmentedGeometryScaledDotProductAttention", trig=trig,
                                                                           compute_dtype="fp32"))
    core.load_state_dict(params, strict=True)
    xi = x["x"].clone().requires_grad_(True)
    out, att = core(xi, xi, xi, x["boxes"])
    (out.square().sum() + att.square().sum()).backward()
    _close(out.detach(), out_r.detach(), 1e-5)
    _close(xi.grad, xr.grad)
    for i in range(h):
        _close(core.fc_gs[i].weight.grad, pr[f"fc_gs.{i}.weight"].grad)
        _close(core.fc_gs[i].bias.grad, pr[f"fc_gs.{i}.bias"].grad)


def test_ocr_ptr_net_is_differentiable(cpu_ops):
    """mmf_m4c.py:379-396: the pointer scores train the query / key projections AND the MMT outputs they read."""
    D, qk, B, T, L = 48, 32, 3, 5, 7
    g = torch.Generator().manual_seed(2)
    params = S.synth_params([("query.weight", (qk, D)), ("query.bias", (qk,)), ("key.weight", (qk, D)), ("key.bias", (qk,))], 1)
    dec, ocr = torch.randn((B, T, D), generator=g), torch.randn((B, L, D), generator=g)
    mask = (torch.rand((B, 1, 1, L), generator=g) < 0.3).float() * R.MASK_VALUE     # (B,1,1,Locr) as generate_padding_mask builds it
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    dr, orr = dec.clone().requires_grad_(True), ocr.clone().requires_grad_(True)
    ref = R.ocr_ptr_net(pr, dr, orr, mask, query_key_size=qk)
    w = torch.randn(ref.shape, generator=g)
    (ref * w).sum().backward()
    net = ovqa.OcrPtrNet(D, qk)
    net.compute_dtype = "fp32"
    net.load_state_dict(params, strict=True)
    d, o = dec.clone().requires_grad_(True), ocr.clone().requires_grad_(True)
    s = net(d, o, mask)
    assert s.requires_grad
    (s * w).sum().backward()
    _close(s.detach(), ref.detach(), 1e-5)
    _close(d.grad, dr.grad)
    _close(o.grad, orr.grad)
    for name, p in net.named_parameters():
        _close(p.grad, pr[name].grad)
    with pytest.raises(ValueError):                              # the in-place score buffer is an inference-only fast path
        net(d, o, mask, scores_out=torch.zeros(B, T, 20), col_offset=4)
    with torch.no_grad():                                        # ... which still writes in place under no_grad
        buf = torch.zeros(B, T, 20)
        net(dec, ocr, mask, scores_out=buf, col_offset=4)
    _close(buf[..., 4:4 + L], ref.detach(), 1e-5)
    # 2-D query (a single decoding step): squeezed result, as the reference
    s1 = net(dec[:, 0], ocr, mask)
    assert s1.shape == (B, L)


def test_dynamic_pointer_network_gradients_skip_masked_keys(cpu_ops):
    D, B, T, L = 32, 2, 4, 6
    g = torch.Generator().manual_seed(3)
    params = S.synth_params([("fc_q.weight", (D, D)), ("fc_q.bias", (D,)), ("fc_k.weight", (D, D)), ("fc_k.bias", (D,))], 1)
    dec, ocr = torch.randn((B, T, D), generator=g), torch.randn((B, L, D), generator=g)
    pad = torch.zeros((B, L), dtype=torch.bool)
    pad[0, -2:] = True
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    orr = ocr.clone().requires_grad_(True)
    ref = R.dynamic_pointer_network(pr, dec, orr, pad, d_model=D)
    torch.where(torch.isinf(ref), torch.zeros_like(ref), ref).square().sum().backward()
    net = ovqa.DynamicPointerNetwork(H.NS(D_MODEL=D, COMPUTE_DTYPE="fp32"))
    net.load_state_dict(params, strict=True)
    o = ocr.clone().requires_grad_(True)
    s = net(dec, o, pad)
    assert torch.isinf(s[0, :, -2:]).all()
    torch.where(torch.isinf(s), torch.zeros_like(s), s).square().sum().backward()
    _close(o.grad, orr.grad)
    _close(net.fc_k.weight.grad, pr["fc_k.weight"].grad)


def test_beam_select_falls_back_outside_the_kernel_range():
    """beam sizes above the top-k kernel's limit and CPU tensors take the reference's own sort (beam_search.py:35-38)."""
    x = torch.randn(3, 5, 40)
    idx, val = ovqa.beam_select(x, 3, 20)
    ref_i, ref_v = R.beam_select(x, 3, 20)
    assert torch.equal(idx, ref_i) and torch.equal(val, ref_v)


def test_stateful_decode_of_a_foreign_core_keeps_the_running_keys(cpu_ops):
    """A core that is not the drop-in (the reference's memory / adaptive variants) must still see the concatenated running keys /
    values of a stateful decode (attentions.py:320-325)."""
    D, h = 32, 4
    seen = []

    class ForeignCore(torch.nn.Module):
        def forward(self, q, k, v, mask, **kw):
            seen.append(k.shape[1])
            return q, None

    from openvivqa_b200 import modules as M
    MHA = M.make_multi_head_attention(builder=lambda cfg: ForeignCore())
    m = MHA(H.attention_cfg(D, h, stateful=True)).eval()
    with torch.no_grad(), m.statefulness(2):
        for _ in range(3):
            m(torch.randn(2, 1, D), torch.randn(2, 1, D), torch.randn(2, 1, D), None)
    assert seen == [1, 2, 3]


@pytest.mark.parametrize("masked", [False, True])
def test_geometry_block_bf16_route_returns_the_geometry_gradients(cpu_ops, masked):
    """bf16 mode sends the box-geometry term to the attention kernels as an additive (B,h,N,N) bias and reduces the fc_g gradients from
    the dS the backward kernel returns; host logic of that route (bias + padding mask combination, saved tensors, gradient plumbing)."""
    D, h, B, N = 64, 4, 2, 9
    params = S.synth_params(C.att_spec(D, h, D // h, D // h, geometry_dg=D // h), 4)
    x = S.padded_features(B, N, D, 3, min_valid=4, relu_like=False)
    boxes = S.boxes(B, N, 5)
    mask = R.padding_mask(x, 0) if masked else None
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xr = x.clone().requires_grad_(True)
    ref = R.mha(pr, xr, xr, xr, mask, h=h, d_k=D // h, d_v=D // h, boxes=boxes, d_g=D // h, trig=True)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(5))
    (ref * w).sum().backward()
    m = ovqa.MultiHeadAttention(H.attention_cfg(D, h, dropout=0.0, arch="AugmentedGeometryScaledDotProductAttention", compute_dtype="bf16")).train()
    m.load_state_dict(params, strict=True)
    xi = x.clone().requires_grad_(True)
    y = m(xi, xi, xi, mask, boxes=boxes)
    (y * w).sum().backward()
    _close(y.detach(), ref.detach(), 3e-2)                               # (the emulated library rounds operands to bf16 in this mode)
    for i in range(h):
        g, gr = m.attention.fc_gs[i].weight.grad, pr[f"attention.fc_gs.{i}.weight"].grad
        assert g is not None and float((g - gr).norm() / gr.norm().clamp_min(1e-9)) < 0.1, i
        gb, gbr = m.attention.fc_gs[i].bias.grad, pr[f"attention.fc_gs.{i}.bias"].grad
        assert float((gb - gbr).abs().max()) <= 0.1 * float(gbr.abs().max()) + 1e-4


@pytest.mark.parametrize("p_drop", [0.0, 0.25])
def test_attribute_reduce_pooling_matches_oracle_autograd(cpu_ops, p_drop):
    """models/mcan.py:21-25,70-77 through the pooling head: output and every gradient (x, fc1, fc2) against autograd of the composition —
    with dropout, against the composition that applies the SAME keep-scales."""
    B, N, D = 3, 11, 32
    g = torch.Generator().manual_seed(3)
    params = S.synth_params([("fc1.weight", (D, D)), ("fc1.bias", (D,)), ("fc2.weight", (1, D)), ("fc2.bias", (1,))], 2)
    x = torch.randn((B, N, D), generator=g)
    torch.manual_seed(0)
    Fn.dropout_stream.offset = 0
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xr = x.clone().requires_grad_(True)
    xi = x.clone().requires_grad_(True)
    pi = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    y = Fn.attr_reduce_pool(xi, w1=pi["fc1.weight"], b1=pi["fc1.bias"], w2=pi["fc2.weight"], b2=pi["fc2.bias"], mode="fp32", dropout=p_drop, training=True)
    ks = cpu_ops._keep_scale(B * N * D, p_drop, Fn.dropout_stream.seed, 0).view(B, N, D)
    hid = torch.relu(torch.nn.functional.linear(xr, pr["fc1.weight"], pr["fc1.bias"])) * ks
    att = torch.softmax(torch.nn.functional.linear(hid, pr["fc2.weight"], pr["fc2.bias"]), dim=1)
    ref = (xr * att).sum(dim=1)
    w = torch.randn(ref.shape, generator=g)
    (ref * w).sum().backward()
    (y * w).sum().backward()
    _close(y.detach(), ref.detach(), 1e-5)
    _close(xi.grad, xr.grad)
    for k in params:
        _close(pi[k].grad, pr[k].grad)


def test_harness_mcan_head_goes_through_the_pooling_head(cpu_ops, golden):
    """MCAN-small end to end on the emulated library (fp32): logits against the reference-generated golden."""
    c = C.MCAN_SMALL
    gm = golden("mcan")["small"]
    m = H.MCAN(H.mcan_cfg(c["d_model"], c["d_feature"], c["layers"], c["h"], c["d_ff"], compute_dtype="fp32"), H.StubVocab(c["vocab"], c["answers"]))
    m.load_state_dict(S.synth_params(gm["spec"], 0), strict=True)
    m.eval()
    with torch.no_grad():
        logp = m(H.Batch(C.mcan_inputs(c)))
    _close(logp, gm["logp"], 1e-4)


def test_probability_dropout_block_matches_the_explicit_composition(cpu_ops):
    """HF BertSelfAttention drops attention PROBABILITIES (p = 0.1 by default): mha_block(prob_dropout=...) forward and backward against
    autograd through the explicit composition that applies the SAME keep-scales (mask stream of the call's (seed, offset))."""
    torch.manual_seed(0)
    D, h, B, L = 32, 4, 2, 9
    params = S.synth_params(C.att_spec(D, h, D // h, D // h), 3)
    x = torch.randn(B, L, D)
    mask = torch.zeros(B, 1, 1, L)
    mask[:, :, :, -2:] = R.MASK_VALUE
    pi = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xi = x.clone().requires_grad_(True)
    Fn.dropout_stream.offset = 0
    y = Fn.mha_block(xi, xi, xi, mask, wq=pi["attention.fc_q.weight"], bq=pi["attention.fc_q.bias"], wk=pi["attention.fc_k.weight"],
                     bk=pi["attention.fc_k.bias"], wv=pi["attention.fc_v.weight"], bv=pi["attention.fc_v.bias"], wo=pi["attention.fc_o.weight"],
                     bo=pi["attention.fc_o.bias"], gamma=pi["layer_norm.weight"], beta=pi["layer_norm.bias"], h=h, d_k=D // h, d_v=D // h,
                     mode="fp32", dropout=0.0, prob_dropout=0.3, training=True)
    keep = cpu_ops._keep_scale(B * h * L * L, 0.3, Fn.dropout_stream.seed, 0).view(B, h, L, L)
    pr = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    xr = x.clone().requires_grad_(True)
    lin = lambda n, t: torch.nn.functional.linear(t, pr[f"attention.{n}.weight"], pr[f"attention.{n}.bias"])
    sp = lambda t: t.view(B, L, h, D // h).permute(0, 2, 1, 3)
    att = torch.softmax(sp(lin("fc_q", xr)) @ sp(lin("fc_k", xr)).transpose(-1, -2) / math.sqrt(D // h) + mask, -1) * keep
    ctx = (att @ sp(lin("fc_v", xr))).permute(0, 2, 1, 3).reshape(B, L, D)
    ref = torch.nn.functional.layer_norm(xr + lin("fc_o", ctx), (D,), pr["layer_norm.weight"], pr["layer_norm.bias"])
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1))
    (ref * w).sum().backward()
    (y * w).sum().backward()
    _close(y.detach(), ref.detach(), 1e-5)
    _close(xi.grad, xr.grad)
    for k in params:
        _close(pi[k].grad, pr[k].grad)


def test_lstm_function_matches_nn_lstm_autograd(cpu_ops):
    """SURVEY 8f rank 4 (text_embeddings.py:234,244): the LSTM autograd function on the emulated library — which tensors are saved, the gate
    pre-activation gradients, dx / dW_ih / dW_hh (h_{t-1} operand) / both bias gradients — against autograd of nn.LSTM itself."""
    torch.manual_seed(1)
    B, L, In, Hd = 4, 6, 16, 24
    ref = torch.nn.LSTM(In, Hd, batch_first=True)
    m = torch.nn.LSTM(In, Hd, batch_first=True)
    m.load_state_dict(ref.state_dict())
    x = torch.randn(B, L, In)
    xr, xi = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr, _ = ref(xr)
    w = torch.randn(yr.shape, generator=torch.Generator().manual_seed(2))
    (yr * w).sum().backward()
    y = Fn.lstm(m, xi)
    (y * w).sum().backward()

    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-12))
    assert rel(y.detach(), yr.detach()) < 2e-2          # bf16 operands (x, W_ih, W_hh, h_{t-1}) with fp32 accumulation and state
    assert rel(xi.grad, xr.grad) < 3e-2
    for (k, p), (_, pr) in zip(m.named_parameters(), ref.named_parameters()):
        assert p.grad is not None, k
        assert rel(p.grad, pr.grad) < 3e-2, (k, rel(p.grad, pr.grad))
    assert getattr(y, "_ovqa_bf16", None) is not None      # the bf16 shadow the first encoder block consumes
    with torch.no_grad():
        assert torch.equal(Fn.lstm(m, xi), y.detach())
