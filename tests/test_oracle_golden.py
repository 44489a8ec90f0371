"""The restated oracle (oracle/restated.py) against outputs of the reference itself (tests/golden/*.pt,
made by oracle/make_golden.py from /root/reference).  CPU only.  This is what pins the oracle."""
import pytest
import torch

from oracle import cases as C
from oracle import restated as R
from oracle import synth as S

TOL = dict(rtol=2e-5, atol=2e-5)   # CPU-to-CPU fp32 (BLAS kernel choice may differ between hosts)


def _close(a, b, **kw):
    torch.testing.assert_close(a, b, **{**TOL, **kw})


@pytest.mark.parametrize("name", list(C.ATT_CASES))
def test_attention_core_and_mha(golden, name):
    B, Lq, Lk, D, h, dk, dv, _, _ = C.ATT_CASES[name]
    g = golden("attention")[name]
    x = C.att_inputs(name)
    p = S.synth_params(C.att_spec(D, h, dk, dv), 0)
    out, att = R.sdpa_core(R.sub(p, "attention"), x["queries"], x["keys"], x["values"], x["mask"], h=h, d_k=dk, d_v=dv)
    _close(out, g["core_out"])
    _close(att, g["att"])
    y = R.mha(p, x["queries"], x["keys"], x["values"], x["mask"], h=h, d_k=dk, d_v=dv)
    _close(y, g["mha_out"])


def test_all_masked_rows_are_finite_and_not_uniform_zero(golden):
    """a sample whose keys are all masked gives softmax(s - 1e5), finite, rows sum to 1."""
    att = golden("attention")["micro_30x100"]["att"]
    assert torch.isfinite(att).all()
    _close(att[0].sum(-1), torch.ones_like(att[0].sum(-1)))
    assert att[0].min() > 0       # nothing is exactly masked out in the all-masked sample


def test_mha_aoa(golden):
    B, Lq, Lk, D, h, dk, dv, _, _ = C.ATT_CASES["guided_100x30"]
    g = golden("attention")["guided_100x30+aoa"]
    x = C.att_inputs("guided_100x30")
    p = S.synth_params(C.att_spec(D, h, dk, dv, aoa=True), 0)
    y = R.mha(p, x["queries"], x["keys"], x["values"], x["mask"], h=h, d_k=dk, d_v=dv, use_aoa=True)
    _close(y, g["mha_out"])


@pytest.mark.parametrize("name,D,dff,B,L", [("pwff_512", 512, 2048, 2, 30), ("pwff_768", 768, 3072, 1, 19)])
def test_pwff(golden, name, D, dff, B, L):
    p = S.synth_params(C.pwff_spec(D, dff), 0)
    x = torch.randn((B, L, D), generator=S._gen(7, name))
    _close(R.pwff(p, x), golden("attention")[name]["out"])


@pytest.mark.parametrize("name", list(C.GEO_CASES))
def test_geometry_core(golden, name):
    B, N, D, h, trig = C.GEO_CASES[name]
    g = golden("geometry")[name]
    x = C.geo_inputs(name)
    p = S.synth_params(C.att_spec(D, h, D // h, D // h, geometry_dg=x["d_g"]), 0)
    emb = R.box_relational_embedding(x["boxes"], dim_g=x["d_g"], trignometric_embedding=trig)
    # arguments of sin/cos reach ~700: allow a few ulp of the argument
    _close(emb[:, ::5, ::3], g["emb_sample"], atol=2e-4)
    out, mn = R.geometry_core(R.sub(p, "attention"), x["x"], x["x"], x["x"], x["boxes"], None,
                              h=h, d_k=D // h, d_v=D // h, d_g=x["d_g"], trig=trig)
    _close(out, g["core_out"], atol=1e-4)
    _close(mn, g["att"], atol=1e-4)


def test_masks_and_positions(golden):
    g = golden("masks")
    feats = S.padded_features(3, 11, 8, 3, all_pad_first=True)
    toks = S.padded_tokens(3, 9, 50, 3)
    pm, tm, sm = R.padding_mask(feats, 0), R.padding_mask(toks, 0), R.sequential_mask(9)
    for a, b in ((pm, g["feat_mask"]), (tm, g["tok_mask"]), (sm, g["seq_mask"]),
                 (R.self_attention_masks(tm, sm), g["self_mask"])):
        assert a.shape == b.shape and a.dtype == b.dtype and torch.equal(a, b)
    assert pm.shape == (3, 1, 1, 11) and pm.dtype == torch.float32
    assert set(pm.unique().tolist()) <= {0.0, -100000.0}
    assert (pm[0] == -100000.0).all()                  # sample 0 is all padding
    _close(R.sinusoid_pos(torch.zeros(2, 13, 64), 64), g["pos"])


def _enc_inputs():
    D = 128
    vis = S.padded_features(2, 36, D, 31, min_valid=10, relu_like=False)
    txt = S.padded_features(2, 14, D, 32, min_valid=4, relu_like=False)
    return vis, R.padding_mask(vis, 0), txt, R.padding_mask(txt, 0)


def test_encoders(golden):
    g = golden("encoders")
    cfg = dict(d_model=128, h=4, d_k=32, d_v=32, layers=2)
    vis, vm, txt, tm = _enc_inputs()
    p = S.synth_params(g["Encoder"]["spec"], 0)
    _close(R.encoder(p, txt, tm, cfg), g["Encoder"]["out"])
    p = S.synth_params(g["GuidedAttentionEncoder"]["spec"], 0)
    _close(R.guided_encoder(p, vis, vm, txt, tm, cfg), g["GuidedAttentionEncoder"]["out"])
    p = S.synth_params(g["CoAttentionEncoder"]["spec"], 0)
    v, l = R.co_attention_encoder(p, vis, vm, txt, tm, cfg)
    _close(v, g["CoAttentionEncoder"]["out_v"]); _close(l, g["CoAttentionEncoder"]["out_l"])
    p = S.synth_params(g["CrossModalityEncoder"]["spec"], 0)
    v, l = R.cross_modality_encoder(p, vis, vm, txt, tm, cfg)
    _close(v, g["CrossModalityEncoder"]["out_v"]); _close(l, g["CrossModalityEncoder"]["out_l"])


@pytest.mark.parametrize("name", ["small", "full"])
def test_mcan(golden, name):
    c = C.MCAN_SMALL if name == "small" else C.MCAN_FULL
    g = golden("mcan")[name]
    p = S.synth_params(g["spec"], 0)
    x = C.mcan_inputs(c)
    logp = R.mcan_forward(p, x["region_features"], x["question_tokens"], d_model=c["d_model"], h=c["h"],
                          layers=c["layers"])
    _close(logp, g["logp"], atol=5e-5)
    assert (logp.argmax(-1) == g["logp"].argmax(-1)).all()
    # error floor: the reference's own fp32 vs fp64
    assert (g["logp"] - g["logp_fp64"]).abs().max() < 1e-4


def test_m4c_step_parts_and_greedy_decode(golden):
    c = C.M4C_SMALL
    g = golden("m4c")
    p = S.synth_params(g["spec"], 0)
    x = C.m4c_inputs(c)
    kw = dict(mmt_layers=c["mmt_layers"], text_layers=c["text_layers"], heads=c["heads"], ptr_qk=c["hidden"])
    enc = R.m4c_encode_inputs(p, x)
    _close(enc[1], g["step"]["obj_mmt_in"]); _close(enc[3], g["step"]["ocr_mmt_in"])
    prev = g["step"]["prev_inds"]
    _close(R.prev_pred_embeddings(R.sub(p, "mmt.prev_pred_embeddings"), p["classifier.weight"], enc[3], prev),
           g["parts"]["dec_emb"])
    _close(R.ocr_ptr_net(R.sub(p, "ocr_ptr_net"), g["step"]["mmt_dec_output"], g["step"]["mmt_ocr_output"],
                         enc[4], query_key_size=c["hidden"]), g["parts"]["ptr_scores"], atol=1e-2, rtol=1e-6)
    scores = R.m4c_step(p, x, enc, prev, **kw)
    _close(scores, g["step"]["scores"], atol=1e-2, rtol=1e-5)     # masked entries are ~ -1e5 (fp32 ulp 8e-3)
    valid = g["step"]["scores"] > -1e4
    _close(scores[valid], g["step"]["scores"][valid], atol=5e-5)
    scores, prev_inds, n = R.m4c_greedy_decode(p, x, max_iter=c["T"], **kw)
    assert torch.equal(prev_inds, g["decode"]["prev_inds"])
    _close(scores, g["decode"]["scores"], atol=1e-2, rtol=1e-5)


def _dec_case():
    c = C.DEC_CASE
    cfg = dict(h=c["h"], d_k=c["D"] // c["h"], d_v=c["D"] // c["h"])
    p = {k: v for k, v in S.synth_params(C.dec_spec(c["D"], c["h"], c["d_ff"]), 0).items() if "running_" not in k}
    return c, cfg, p, C.dec_inputs()


def test_decoder_layer_stateless_and_stepwise(golden):
    """decoders.py:13-26 over a whole sequence, and token by token with running keys / values + a beam-style state re-gather."""
    c, cfg, p, x = _dec_case()
    g = golden("decoder")
    assert g["spec"] == C.dec_spec(c["D"], c["h"], c["d_ff"])
    _close(R.decoder_layer(p, x["x"], x["enc"], x["enc"], x["full_mask"], x["enc_mask"], cfg), g["full"])
    per_step, perm = C.dec_stepwise_inputs(x)
    # the encoder side changes rows at the re-gather too, so drive the oracle in two segments that share its running state
    outs = []
    running_k = running_m = None
    for t, (xt, mt, enc, enc_mask) in enumerate(per_step):
        running_k = xt if running_k is None else torch.cat([running_k, xt], 1)
        running_m = mt if running_m is None else torch.cat([running_m, mt], -1)
        sa = R.mha(R.sub(p, "self_attn"), xt, running_k, running_k, running_m, **cfg)
        ea = R.mha(R.sub(p, "enc_attn"), sa, enc, enc, enc_mask, **cfg)
        outs.append(R.pwff(R.sub(p, "pwff"), ea))
        if t == c["regather_at"]:
            running_k, running_m = running_k[perm], running_m[perm]
    _close(torch.cat(outs, 1), g["steps"])
    # and the packaged stepwise oracle (no encoder-side permutation: compare the steps before the re-gather)
    k = c["regather_at"] + 1
    st = R.decoder_layer_stateful(p, [x["x"][:, t:t + 1] for t in range(k)], x["enc"], x["enc"], x["step_masks"][:k], x["enc_mask"], cfg)
    _close(torch.cat(st, 1), g["steps"][:, :k])


@pytest.mark.parametrize("name", list(C.BEAM_CASES))
def test_beam_select(golden, name):
    b_s, cur, V, beam = C.BEAM_CASES[name]
    idx, logprob = R.beam_select(C.beam_inputs(name), b_s, beam)
    g = golden("beam")[name]
    assert torch.equal(idx, g["idx"]) and torch.equal(logprob, g["logprob"])


def test_m4c_full_config_greedy_decode_matches_the_reference(golden):
    """BASELINE config 3 at its own shape (cases.M4C_FULL): the restated decode loop reproduces the REFERENCE's token sequences and scores."""
    c, g = C.M4C_FULL, golden("m4c_full")
    p = S.synth_params(g["spec"], 0)
    x = C.m4c_inputs(c)
    with torch.no_grad():
        scores, prev_inds, n = R.m4c_greedy_decode(p, x, max_iter=c["T"], mmt_layers=c["mmt_layers"], text_layers=c["text_layers"],
                                                   heads=c["heads"], ptr_qk=c["hidden"])
    assert torch.equal(prev_inds, g["prev_inds"])
    valid = g["scores_sample"] > -1e4
    torch.testing.assert_close(scores[:, :, ::7][valid], g["scores_sample"][valid], rtol=1e-4, atol=1e-4)
