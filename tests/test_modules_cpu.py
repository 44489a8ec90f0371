"""CPU: host-side logic of the drop-ins — parameter names/shapes/initialisers, registry contract, install()."""
import pytest
import torch

import openvivqa_b200 as ovqa
from openvivqa_b200 import functional as Fn, harness as H
from oracle import cases as C
from oracle import synth as S
from oracle import ref_loader as RL


def _spec(m):
    return [(k, tuple(v.shape)) for k, v in m.state_dict().items() if v.dtype.is_floating_point]


def test_mha_and_pwff_state_dict_keys_match_the_reference_spec():
    assert _spec(ovqa.MultiHeadAttention(H.attention_cfg())) == C.att_spec(512, 8, 64, 64)
    assert _spec(ovqa.MultiHeadAttention(H.attention_cfg(use_aoa=True))) == C.att_spec(512, 8, 64, 64, aoa=True)
    assert _spec(ovqa.MultiHeadAttention(H.attention_cfg(head=1, d_key=64, d_value=64))) == C.att_spec(512, 1, 64, 64)
    geo = ovqa.MultiHeadAttention(H.attention_cfg(arch="AugmentedGeometryScaledDotProductAttention"))
    assert _spec(geo) == C.att_spec(512, 8, 64, 64, geometry_dg=64)
    assert _spec(ovqa.PositionWiseFeedForward(H.attention_cfg())) == C.pwff_spec(512, 2048)


def test_model_state_dicts_match_reference_generated_specs(golden):
    g = golden("mcan")
    c = C.MCAN_SMALL
    m = H.MCAN(H.mcan_cfg(c["d_model"], c["d_feature"], c["layers"], c["h"], c["d_ff"]), H.StubVocab(c["vocab"], c["answers"]))
    assert _spec(m) == [tuple(x) for x in g["small"]["spec"]]
    e = golden("encoders")
    x = H.xmt_cfg("cross", 128, 256, 2, 4, 256)
    assert _spec(H.CrossModalityEncoder(x.ENCODER)) == [tuple(t) for t in e["CrossModalityEncoder"]["spec"]]
    x = H.xmt_cfg("co", 128, 256, 2, 4, 256)
    assert _spec(H.CoAttentionEncoder(x.ENCODER)) == [tuple(t) for t in e["CoAttentionEncoder"]["spec"]]


def test_initialisers_follow_the_reference():
    torch.manual_seed(0)
    m = ovqa.ScaledDotProductAttention(H.attention_cfg())
    for fc in (m.fc_q, m.fc_k, m.fc_v, m.fc_o):
        assert torch.count_nonzero(fc.bias) == 0                       # attentions.py:41-44
        bound = (6.0 / (fc.weight.shape[0] + fc.weight.shape[1])) ** 0.5   # xavier_uniform_
        assert fc.weight.abs().max() <= bound + 1e-6 and fc.weight.abs().max() > 0.9 * bound
    mha = ovqa.MultiHeadAttention(H.attention_cfg())
    assert mha.layer_norm.eps == 1e-5 and mha.dropout.p == 0.1


def test_registry_contract():
    assert "ScaledDotProductAttention" in ovqa.META_ATTENTION
    core = ovqa.build_attention(H.attention_cfg())
    assert isinstance(core, ovqa.ScaledDotProductAttention)
    with pytest.raises(KeyError):
        ovqa.META_ATTENTION.get("NoSuchAttention")
    with pytest.raises(AssertionError):                                 # duplicate names assert (registry.py:43-47)
        ovqa.META_ATTENTION.register(ovqa.ScaledDotProductAttention)


def test_stateful_module_semantics():
    m = ovqa.MultiHeadAttention(H.attention_cfg(stateful=True))
    assert m.running_keys.shape == (0, 512)
    with m.statefulness(3):
        assert m._is_stateful and m.running_keys.shape == (3, 0, 512)
    assert not m._is_stateful and m.running_keys.shape == (0, 512)


def test_decode_cache_states_stay_out_of_the_state_dict():
    """The projected K/V cache of the inference decode path is a re-gatherable state (beam search) but not a checkpoint entry."""
    m = ovqa.MultiHeadAttention(H.attention_cfg(stateful=True))
    assert set(k for k in m.state_dict() if "running" in k) == {"running_keys", "running_values"}
    assert len(list(m.states())) == 3
    with m.statefulness(2):
        assert m.running_kv_proj.shape == (2, 1, 1024)
        m.apply_to_states(lambda s: s[torch.tensor([1, 0])])
    assert m.running_kv_proj.shape == (1, 1024) and m._enc_kv is None and m._kv_len == 0
    layer = H.DecoderLayer(H.NS(SELF_ATTENTION=H.attention_cfg(stateful=True), ENC_ATTENTION=H.attention_cfg()))
    with layer.statefulness(2):
        assert layer.self_attn._is_stateful and layer.enc_attn._is_stateful and not layer.enc_attn.can_be_stateful


def test_one_call_inference_wrappers_refuse_cpu_tensors():
    """No CPU fallback and no bypass: the lean decode / feed-forward wrappers hand raw pointers to the library, so anything that is
    not a CUDA tensor must be rejected before the call (north_star: CPU tensors -> error)."""
    from openvivqa_b200 import ops
    assert not hasattr(ops, "_HOSTPROF")
    x = torch.zeros((2, 1, 64))
    cache = torch.zeros((2, 4, 128))
    with pytest.raises(ValueError):
        ops.mha_decode_step(x, None, None, True, (0,) * 6, dt=0, eps=1e-5, h=4, d_k=16, d_v=16, kv_cache=cache, length=0, scale=0.25)
    with pytest.raises(ValueError):
        ops.pwff_infer(x, None, (0,) * 6, dt=0, d_ff=256, eps=1e-5)


def test_encoder_projection_bookkeeping_across_regathers(monkeypatch):
    """Host logic of the decode session (no kernels: the three library-backed calls are stubbed): the encoder projection is reused when
    the keys are the same objects or provably what the last ``apply_to_states`` sweep made of them, and rebuilt otherwise."""
    from openvivqa_b200 import functional as Fn
    calls = []
    monkeypatch.setattr(Fn, "project_kv_fused", lambda keys, values, **k: (calls.append(1), torch.zeros(keys.shape[0], keys.shape[1], 256))[1])
    monkeypatch.setattr(Fn, "mha_decode_block", lambda plan, q, *a, **k: torch.zeros_like(q))
    monkeypatch.setattr(Fn, "mha_decode_plan", lambda **k: None)
    mha = ovqa.MultiHeadAttention(H.attention_cfg(128, 4)).eval()

    class Model(ovqa.StatefulModule):                       # owner of ``encoder_features`` (base_transformer.py:17)
        def __init__(self):
            super().__init__()
            self.m = mha
            self.register_state("encoder_features", None)

    model, perm = Model(), torch.tensor([1, 0, 2, 5, 4, 3])
    with torch.no_grad(), model.statefulness(2):
        model.encoder_features = torch.randn(2, 9, 128)
        mha(torch.randn(2, 1, 128), model.encoder_features, model.encoder_features, None)
        assert len(calls) == 1 and mha.running_enc_kv_proj.shape == (2, 9, 256)
        mha(torch.randn(2, 1, 128), model.encoder_features, model.encoder_features, None)
        assert len(calls) == 1                                                        # same objects: reused
        model.apply_to_states(lambda s: s.repeat_interleave(3, 0))                    # fan out into beams
        q = torch.randn(6, 1, 128)
        for _ in range(3):
            mha(q, model.encoder_features, model.encoder_features, None)
            model.apply_to_states(lambda s: s[perm])
        assert len(calls) == 1 and mha.running_enc_kv_proj.shape == (6, 9, 256)       # re-gathered with the states, never re-projected
        other = torch.randn_like(model.encoder_features)
        mha(q, other, other, None)
        assert len(calls) == 2                                                        # foreign tensors of the same shape: projected again
        mha(q, model.encoder_features, model.encoder_features, None)
        n = len(calls)
        model.apply_to_states(lambda s: s[perm])
        model.apply_to_states(lambda s: s[perm])                                      # two sweeps between calls: provenance chain broken
        mha(q, model.encoder_features, model.encoder_features, None)
        assert len(calls) == n + 1
        model.encoder_features.mul_(2.0)                                              # edited in place: version counter moved
        mha(q, model.encoder_features, model.encoder_features, None)
        assert len(calls) == n + 2
        assert mha._decode_fast_path(mha._session, torch.randn(6, 40, 128), other, other, None) is None   # long query block: tiled cores
    assert mha.running_enc_kv_proj.shape == (1, 256) and mha._enc_kv is None and mha._regather_memo is None


def test_compute_dtype_selection():
    assert ovqa.get_compute_dtype() in ("fp32", "bf16")
    with pytest.raises(ValueError):
        ovqa.set_compute_dtype("fp8")
    m = ovqa.PositionWiseFeedForward(H.attention_cfg(compute_dtype="bf16"))
    assert m.compute_dtype == "bf16"


@pytest.mark.skipif(not RL.available(), reason="/root/reference not mounted")
def test_install_into_the_reference_checkout_and_reference_state_dict_loads():
    RL.load()
    import models.modules.attentions as A
    import models.modules.encoders as E
    from builders.attention_builder import META_ATTENTION

    ref_mha = A.MultiHeadAttention(RL.attention_cfg())
    import models.modules.beam_search as BS
    ref_select = BS.BeamSearch.select
    try:
        ovqa.install()
        assert BS.BeamSearch.select is not ref_select and BS.BeamSearch.select.__module__.startswith("openvivqa_b200")
        assert META_ATTENTION.get("ScaledDotProductAttention") is ovqa.ScaledDotProductAttention
        assert E.MultiHeadAttention.__module__.startswith("openvivqa_b200")
        enc = E.Encoder(RL.NS(D_MODEL=512, LAYERS=1, SELF_ATTENTION=RL.attention_cfg()))   # the reference's own Encoder
        mha = enc.layers[0].mhatt
        assert type(mha).__module__.startswith("openvivqa_b200")
        from models.modules.containers import Module
        assert isinstance(mha, Module)                      # stateful recursion of the reference still sees it
        assert isinstance(enc.layers[0].pwff, ovqa.PositionWiseFeedForward)
        mha.load_state_dict(ref_mha.state_dict(), strict=True)      # reference checkpoints load unchanged
        # memory / adaptive cores stay the reference's
        assert META_ATTENTION.get("AdaptiveScaledDotProductAttention").__module__ == "models.modules.attentions"
    finally:
        ovqa.uninstall()
    assert E.MultiHeadAttention is A.MultiHeadAttention
    assert BS.BeamSearch.select is ref_select
    assert META_ATTENTION.get("ScaledDotProductAttention").__module__ == "models.modules.attentions"


def test_weight_cache_refreshes_after_fused_optimizer_step():
    """``Adam(fused=True)`` updates parameters WITHOUT bumping ``Tensor._version``; the staged (concatenated) operand
    copies must still be rebuilt after every optimiser step (process-wide optimiser post-hook)."""
    from openvivqa_b200 import functional as Fn

    a, b = torch.nn.Parameter(torch.randn(4, 8)), torch.nn.Parameter(torch.randn(4, 8))
    w0 = Fn.weights.get([a, b], torch.float32)
    assert Fn.weights.get([a, b], torch.float32) is w0                  # cached
    opt = torch.optim.Adam([a, b], lr=0.1, fused=True)
    a.grad, b.grad = torch.ones_like(a), torch.ones_like(b)
    opt.step()
    w1 = Fn.weights.get([a, b], torch.float32)
    assert w1 is not w0 and torch.equal(w1, torch.cat([a.detach(), b.detach()]))
    with torch.no_grad():
        a.mul_(2.0)                                                      # version-bump path
    assert torch.equal(Fn.weights.get([a, b], torch.float32), torch.cat([a.detach(), b.detach()]))


def test_m4c_mirror_state_dict_matches_reference_generated_spec(golden):
    c = C.M4C_SMALL
    m = H.MMF_M4C(H.m4c_cfg(c["hidden"], c["heads"], c["mmt_layers"], c["text_layers"], c["d_obj"]),
                  H.StubVocab(c["vocab"], max_answer_length=c["T"]))
    assert _spec(m) == [tuple(t) for t in golden("m4c")["spec"]]


@pytest.mark.skipif(not RL.available(), reason="/root/reference not mounted")
def test_reference_beam_search_runs_on_the_installed_drop_ins(monkeypatch, golden):
    """The REAL drop-in scenario, bookkeeping level: the reference's own BaseTransformer / Decoder / BeamSearch classes with
    ``install()`` applied — its containers.Module state recursion, its re-gather closure (wrapped), its ``select`` (rebound) — must
    reproduce the golden token sequences.  No GPU here: the four library-backed calls are replaced by plain torch restatements, so
    what is tested is every piece of host logic between them (sessions, caches with spare capacity, placeholder rows, provenance)."""
    import math
    import torch.nn.functional as F
    from openvivqa_b200 import functional as Fn, ops

    def mha_decode_block(plan, queries, kv_src, self_att, cache, length, mask, *, h, d_k, d_v, eps=1e-5, scale=None):
        w_qkv, b_qkv, w_o, b_o, gamma, beta = plan.keep
        hk = h * d_k
        B, Lq, D = queries.shape
        q = F.linear(queries, w_qkv[:hk], b_qkv[:hk])
        src = queries if self_att else kv_src
        if src is not None:
            cache[:, length:length + src.shape[1]] = F.linear(src, w_qkv[hk:], b_qkv[hk:])
            length += src.shape[1]
        k, v = cache[:, :length, :hk], cache[:, :length, hk:]
        att = q.view(B, Lq, h, d_k).permute(0, 2, 1, 3) @ k.reshape(B, length, h, d_k).permute(0, 2, 3, 1) / math.sqrt(d_k)
        att = torch.softmax(att if mask is None else att + mask, -1)
        ctx = (att @ v.reshape(B, length, h, d_v).permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, Lq, h * d_v)
        return F.layer_norm(queries + F.linear(ctx, w_o, b_o), (D,), gamma, beta, eps)

    def pwff_infer_block(plan, x, *, d_ff, eps=1e-5):
        w1, b1, w2, b2, g, b = plan.keep
        return F.layer_norm(x + F.linear(F.gelu(F.linear(x, w1, b1)), w2, b2), (x.shape[-1],), g, b, eps)

    def topk_rows(x, k):
        v, i = torch.sort(x, dim=-1, descending=True, stable=True)
        return v[:, :k].contiguous(), i[:, :k].contiguous()

    projections = []
    monkeypatch.setattr(Fn, "mha_decode_block", mha_decode_block)
    monkeypatch.setattr(Fn, "pwff_infer_block", pwff_infer_block)
    monkeypatch.setattr(Fn, "project_kv_fused", lambda keys, values, *, wk, bk, wv, bv, mode="fp32":
                        (projections.append(1), torch.cat([F.linear(keys, wk, bk), F.linear(values, wv, bv)], -1))[1])
    monkeypatch.setattr(ops, "topk_rows", topk_rows)

    RL.load()
    import models.modules.decoders as Dm
    from models.base_transformer import BaseTransformer
    c, g = C.GEN_CASE, golden("generative")
    previous_dtype = ovqa.get_compute_dtype()
    try:
        ovqa.install(compute_dtype="fp32")
        assert Dm.MultiHeadAttention.__module__.startswith("openvivqa_b200")
        att = lambda st: RL.attention_cfg(c["D"], c["h"], c["D"] // c["h"], c["D"] // c["h"], d_ff=c["d_ff"], stateful=st)
        cfg = RL.NS(D_MODEL=c["D"], LAYERS=c["layers"], ATTENTION=RL.NS(SELF_ATTENTION=att(True), ENC_ATTENTION=att(False)),
                    TEXT_EMBEDDING=RL.NS(ARCHITECTURE="UsualEmbedding", D_MODEL=c["D"], WORD_EMBEDDING=None, D_EMBEDDING=c["D"], DROPOUT=0.1))

        class Captioner(BaseTransformer):                       # as oracle/make_golden.py::gen_generative
            def __init__(self, config, vocab):
                super().__init__(config, vocab)
                self.device = "cpu"
                self.decoder = Dm.Decoder(config, vocab)        # the reference's Decoder, now built from the drop-ins

            def encoder_forward(self, feats):
                return feats["enc"], feats["enc_mask"]

        m = Captioner(cfg, RL.StubVocab(c["vocab"], max_answer_length=c["T"])).eval()
        assert S.spec_of(m.state_dict()) == g["spec"]           # reference checkpoint layout unchanged by the extra (non-persistent) states
        m.load_state_dict(S.synth_params(g["spec"], 0), strict=False)
        m.eos_idx = c["eos"]
        x = C.gen_inputs()
        with torch.no_grad():
            for beam in (1, c["beam"]):
                projections.clear()
                tokens, logp = m.beam_search(x, batch_size=c["B"], beam_size=beam, out_size=1)
                assert torch.equal(tokens, g[f"beam{beam}"]["tokens"])
                torch.testing.assert_close(logp, g[f"beam{beam}"]["logp"], rtol=1e-4, atol=1e-5)
                assert len(projections) == c["layers"]          # encoder keys projected once per layer for the whole search
    finally:
        ovqa.uninstall()
        ovqa.set_compute_dtype(previous_dtype)


@pytest.mark.skipif(not RL.available(), reason="/root/reference not mounted")
def test_install_rebinds_the_reference_lstm_text_embedding(cpu_ops, monkeypatch):
    """``install()`` rebinds ``LSTMTextEmbedding.forward`` (text_embeddings.py:236-246) on the reference's own class; the
    parameters stay the ``nn.LSTM``'s.  With the library emulated on CPU: fp32 mode keeps the stock LSTM (bit-identical to the reference's
    forward), bf16 mode goes through ``Fn.lstm`` and stays within the bf16 path's tolerance; ``uninstall()`` restores the method."""
    RL.load()
    import models.modules.text_embeddings as T
    stock = T.LSTMTextEmbedding.forward
    cfg = RL.NS(D_EMBEDDING=24, D_MODEL=32, DROPOUT=0.0, WORD_EMBEDDING=None)
    torch.manual_seed(0)
    emb = T.LSTMTextEmbedding(cfg, H.StubVocab(50, 10)).eval()
    tokens = torch.randint(1, 50, (3, 7))
    tokens[0, 5:] = 0
    ref, (pad_ref, seq_ref) = emb(tokens)
    try:
        ovqa.install(compute_dtype="fp32")
        assert T.LSTMTextEmbedding.forward is not stock
        out, (pad, seq) = emb(tokens)
        assert torch.equal(out, ref) and torch.equal(pad, pad_ref) and torch.equal(seq, seq_ref)
        # bf16 mode: the library path (here: its CPU restatement; `is_cuda` is the only gate that a CPU tensor cannot pass)
        monkeypatch.setattr(Fn, "lstm_supported", lambda module, x, mode: mode == "bf16")
        ovqa.set_compute_dtype("bf16")
        out16, _ = emb(tokens)
        assert float((out16 - ref).abs().max()) < 2e-2
        assert set(emb.state_dict()) == {"embedding.weight", "proj.weight", "proj.bias", "lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0"}
    finally:
        ovqa.set_compute_dtype("fp32")
        ovqa.uninstall()
    assert T.LSTMTextEmbedding.forward is stock
