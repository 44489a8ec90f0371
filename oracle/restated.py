"""Restated CPU oracle for OpenViVQA's cross-modal attention hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): a functional, fp32,
pure-``torch``-on-CPU restatement of the reference arithmetic, written as plain
functions over a flat ``{name: tensor}`` parameter dict (the reference module's
``state_dict()``), so that it travels to the GPU box where ``/root/reference``
does not exist.  It is pinned against outputs of the reference itself (fixtures
under ``tests/golden/`` made by ``oracle/make_golden.py``).

Every function cites the reference ``file:line`` (relative to
``/root/reference``) whose op order it follows; the order matters, because the
parity bar is 1e-4 relative in fp32 and e.g. the additive ``-1e5`` mask has to be
added *before* the row-max subtraction.

Third-party arithmetic: the M4C multimodal transformer layers are HuggingFace
``transformers`` BERT layers (unpinned by the reference; 5.5.0 here,
``models/bert/modeling_bert.py`` ``eager_attention_forward``/``BertSelfOutput``/
``BertIntermediate``/``BertOutput``) — restated in :func:`bert_layer`.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]
MASK_VALUE = -10e4  # models/utils.py:56,64,71  (= -100000.0, additive, finite)


# --------------------------------------------------------------------------- helpers
# Rounding-floor probe (tests only): with BF16_OPERANDS the oracle rounds the two operands of every dense contraction — the Linear
# layers and the two attention products — to bf16 and keeps everything else (accumulation, softmax, LayerNorm, residuals) in fp32.  That
# is the least error ANY bf16-operand implementation of the path can have against the fp32 reference; the GPU tests report the kernels'
# error next to it (BASELINE.md §4 measured the same floor on the reference itself: 1.9e-2 on MCAN log-probabilities).
BF16_OPERANDS = False


def _r(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).float() if BF16_OPERANDS else x


def _lin(p: Params, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(_r(x), _r(p[name + ".weight"]), p.get(name + ".bias"))


def _ln(p: Params, name: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], eps)


def _drop(x: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    return F.dropout(x, p, training) if (training and p > 0.0) else x


def sub(p: Params, prefix: str) -> Params:
    """View of ``p`` restricted to keys under ``prefix.`` (prefix stripped)."""
    n = len(prefix) + 1
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix + ".")}


# --------------------------------------------------------------------------- masks  (models/utils.py:44-73)
def padding_mask(sequences: torch.Tensor, padding_idx: int) -> torch.Tensor:
    """models/utils.py:44-57 — rows whose feature sum equals ``padding_idx*dim`` are padding.

    Returns an *additive* mask ``(B,1,1,L)`` with value ``-10e4`` on padded positions.  Note the
    reference builds it as ``long * -10e4`` → float32.
    """
    s = sequences.unsqueeze(-1) if sequences.dim() == 2 else sequences
    is_pad = torch.sum(s, dim=-1) == (padding_idx * s.shape[-1])
    return (is_pad.long() * MASK_VALUE).unsqueeze(1).unsqueeze(1)


def sequential_mask(seq_len: int) -> torch.Tensor:
    """models/utils.py:59-66 — strict upper triangle * -10e4, shape (1,1,L,L)."""
    return (torch.triu(torch.ones((seq_len, seq_len)), diagonal=1) * MASK_VALUE).unsqueeze(0).unsqueeze(0)


def self_attention_masks(pad: torch.Tensor, seq: torch.Tensor) -> torch.Tensor:
    """models/utils.py:68-73 — logical_or of the two, re-scaled to -10e4; (B,1,L,L)."""
    return torch.logical_or(pad != 0, seq != 0).long() * MASK_VALUE


# --------------------------------------------------------------------------- positional embedding
def sinusoid_pos(x: torch.Tensor, num_pos_feats: int, temperature: float = 10000.0) -> torch.Tensor:
    """models/modules/pos_embeddings.py:58-72 (normalize=False, mask=None → positions 1..L)."""
    B, L = x.shape[:2]
    embed = torch.ones((B, L), dtype=torch.bool).cumsum(1, dtype=torch.float32)
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    pos = embed[:, :, None] / dim_t
    return torch.stack((pos[:, :, 0::2].sin(), pos[:, :, 1::2].cos()), dim=-1).flatten(-2)


# --------------------------------------------------------------------------- attention cores
def sdpa_core(p: Params, queries, keys, values, attention_mask=None, *, h: int, d_k: int, d_v: int):
    """ScaledDotProductAttention.forward — models/modules/attentions.py:46-60.

    Returns ``(out (B,Lq,D), att (B,h,Lq,Lk))``.  ``h*d_k`` need not equal ``D`` (LoRRA, h=1).
    """
    B, nq = queries.shape[:2]
    nk = keys.shape[1]
    q = _lin(p, "fc_q", queries).view(B, nq, h, d_k).permute(0, 2, 1, 3)
    k = _lin(p, "fc_k", keys).view(B, nk, h, d_k).permute(0, 2, 3, 1)
    v = _lin(p, "fc_v", values).view(B, nk, h, d_v).permute(0, 2, 1, 3)
    att = torch.matmul(_r(q), _r(k)) / np.sqrt(d_k)   # :53
    if attention_mask is not None:
        att = att + attention_mask                     # :54-55 (in-place in the reference)
    att = torch.softmax(att, dim=-1)                   # :56  (no dropout on probabilities)
    out = torch.matmul(_r(att), _r(v)).permute(0, 2, 1, 3).contiguous().view(B, nq, h * d_v)
    return _lin(p, "fc_o", out), att


def box_relational_embedding(boxes: torch.Tensor, dim_g: int = 64, wave_len: float = 1000.0,
                             trignometric_embedding: bool = True) -> torch.Tensor:
    """models/utils.py:102-162 — (B,N,4) boxes → (B,N,N,dim_g) pairwise geometry embedding."""
    B = boxes.size(0)
    x_min, y_min, x_max, y_max = torch.chunk(boxes, 4, dim=-1)
    cx = (x_min + x_max) * 0.5
    cy = (y_min + y_max) * 0.5
    w = (x_max - x_min) + 1.0
    hh = (y_max - y_min) + 1.0
    dx = torch.log(torch.clamp(torch.abs((cx - cx.view(B, 1, -1)) / w), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy - cy.view(B, 1, -1)) / hh), min=1e-3))
    dw = torch.log(w / w.view(B, 1, -1))
    dh = torch.log(hh / hh.view(B, 1, -1))
    pos = torch.stack((dx, dy, dw, dh), dim=-1)                        # (B,N,N,4)
    if not trignometric_embedding:
        return pos
    n_freq = dim_g // 8
    feat_range = torch.arange(dim_g / 8)
    dim_mat = 1.0 / torch.pow(torch.tensor(float(wave_len)), feat_range / (dim_g / 8))
    mul = (100.0 * pos.unsqueeze(-1)) * dim_mat.view(1, 1, 1, 1, -1)  # (B,N,N,4,n_freq)
    mul = mul.reshape(B, pos.shape[1], pos.shape[2], 4 * n_freq)
    return torch.cat((torch.sin(mul), torch.cos(mul)), dim=-1)


def geometry_log_bias(p: Params, boxes: torch.Tensor, *, h: int, d_g: int, trig: bool = True) -> torch.Tensor:
    """attentions.py:113-119,131-132 — ``log(clamp(relu(fc_g_h(emb)), 1e-6))``, shape (B,h,N,N)."""
    emb = box_relational_embedding(boxes, dim_g=d_g, trignometric_embedding=trig)
    B, N = emb.shape[:2]
    flat = emb.view(-1, d_g)
    per_head = [F.linear(flat, p[f"fc_gs.{i}.weight"], p[f"fc_gs.{i}.bias"]).view(B, 1, N, N) for i in range(h)]
    g = F.relu(torch.cat(per_head, dim=1))
    return torch.log(torch.clamp(g, min=1e-6))


def geometry_core(p: Params, queries, keys, values, boxes, attention_mask=None, *, h: int, d_k: int, d_v: int,
                  d_g: int, trig: bool = True):
    """AugmentedGeometryScaledDotProductAttention.forward — attentions.py:111-137.

    DOCUMENTED DEVIATION (SURVEY.md finding 5 / Appendix C): the reference line :128-129 adds the mask to
    an undefined name ``att`` and crashes; the restatement adds it to ``a`` (the obvious fix).  With
    ``attention_mask=None`` this is exactly the reference and is pinned by a golden vector.
    """
    B, nq = queries.shape[:2]
    nk = keys.shape[1]
    logg = geometry_log_bias(p, boxes, h=h, d_g=d_g, trig=trig)
    q = _lin(p, "fc_q", queries).view(B, nq, h, d_k).permute(0, 2, 1, 3)
    k = _lin(p, "fc_k", keys).view(B, nk, h, d_k).permute(0, 2, 3, 1)
    v = _lin(p, "fc_v", values).view(B, nk, h, d_v).permute(0, 2, 1, 3)
    a = torch.matmul(q, k) / np.sqrt(d_k)
    if attention_mask is not None:
        a = a + attention_mask
    mn = torch.softmax(logg + a, dim=-1)
    out = torch.matmul(mn, v).permute(0, 2, 1, 3).contiguous().view(B, nq, h * d_v)
    return _lin(p, "fc_o", out), mn


# --------------------------------------------------------------------------- MHA / PWFF blocks
def mha(p: Params, queries, keys, values, attention_mask, *, h: int, d_k: int, d_v: int,
        dropout: float = 0.0, training: bool = False, use_aoa: bool = False, boxes=None,
        d_g: Optional[int] = None, trig: bool = True, eps: float = 1e-5):
    """MultiHeadAttention.forward — attentions.py:319-339 (stateless path)."""
    core = sub(p, "attention")
    if boxes is not None:
        out, _ = geometry_core(core, queries, keys, values, boxes, attention_mask, h=h, d_k=d_k, d_v=d_v,
                               d_g=d_g, trig=trig)
    else:
        out, _ = sdpa_core(core, queries, keys, values, attention_mask, h=h, d_k=d_k, d_v=d_v)
    out = _drop(out, dropout, training)                                  # :330
    out = _ln(p, "layer_norm", queries + out, eps)                       # :331
    if use_aoa:                                                          # :333-337
        aoa_in = torch.cat([queries, out], dim=-1)
        out = _lin(p, "informative_attention", aoa_in) * torch.sigmoid(_lin(p, "gated_attention", aoa_in))
    return out


def pwff(p: Params, x, *, dropout: float = 0.0, training: bool = False, eps: float = 1e-5):
    """PositionWiseFeedForward.forward — positionwise_feed_forward.py:23-27 (exact erf GELU)."""
    out = _drop(F.gelu(_lin(p, "fc1", x)), dropout, training)
    out = _drop(_lin(p, "fc2", out), dropout, training)
    return _ln(p, "layer_norm", x + out, eps)


# --------------------------------------------------------------------------- decoder layer (models/modules/decoders.py)
def decoder_layer(p: Params, queries, keys, values, self_mask, enc_mask, cfg):
    """DecoderLayer.forward — decoders.py:21-26: self-attention -> encoder attention -> feed-forward."""
    self_att = mha(sub(p, "self_attn"), queries, queries, queries, self_mask, **_acfg(cfg))
    enc_att = mha(sub(p, "enc_attn"), self_att, keys, values, enc_mask, **_acfg(cfg))
    return pwff(sub(p, "pwff"), enc_att, dropout=cfg.get("dropout", 0.0), training=cfg.get("training", False))


def decoder_layer_stateful(p: Params, steps, keys, values, step_masks, enc_mask, cfg, regather=None):
    """The same layer driven one token at a time inside ``statefulness()`` (beam search, base_transformer.py:50-52): the
    self-attention block appends each step's keys / values to ``running_keys`` / ``running_values`` and attends over all of them
    (attentions.py:321-325); the caller accumulates the self-attention mask the same way (decoders.py:53-55).  ``steps``: list of
    (B,1,D) inputs; ``step_masks``: list of (B,1,1,1) additive masks; ``regather``: {t: batch index tensor} applied to the states
    after step t (what beam_search.py:19-33 does through ``apply_to_states``).  Returns the list of (B,1,D) outputs."""
    running_k = running_v = running_mask = None
    outs = []
    for t, (x, m) in enumerate(zip(steps, step_masks)):
        running_k = x if running_k is None else torch.cat([running_k, x], 1)                  # attentions.py:322-325
        running_v = x if running_v is None else torch.cat([running_v, x], 1)
        running_mask = m if running_mask is None else torch.cat([running_mask, m], -1)       # decoders.py:54
        self_att = mha(sub(p, "self_attn"), x, running_k, running_v, running_mask, **_acfg(cfg))
        enc_att = mha(sub(p, "enc_attn"), self_att, keys, values, enc_mask, **_acfg(cfg))
        outs.append(pwff(sub(p, "pwff"), enc_att))
        if regather is not None and t in regather:
            idx = regather[t]
            running_k, running_v, running_mask = running_k[idx], running_v[idx], running_mask[idx]
    return outs


def beam_select(candidate_logprob: torch.Tensor, b_s: int, beam_size: int):
    """BeamSearch.select — beam_search.py:35-38: all candidates of a sample sorted by descending log-probability, the first
    ``beam_size`` kept.  Returns (flat candidate indices, their log-probabilities), both (b_s, beam_size)."""
    logprob, idx = torch.sort(candidate_logprob.view(b_s, -1), -1, descending=True)
    return idx[:, :beam_size], logprob[:, :beam_size]


# --------------------------------------------------------------------------- encoders (models/modules/encoders.py)
def _acfg(cfg) -> dict:
    return dict(h=cfg["h"], d_k=cfg["d_k"], d_v=cfg["d_v"], dropout=cfg.get("dropout", 0.0),
                training=cfg.get("training", False))


def encoder_layer(p: Params, q, k, v, mask, cfg, **kw):
    """EncoderLayer.forward — encoders.py:15-19."""
    att = mha(sub(p, "mhatt"), q, k, v, mask, **_acfg(cfg), **kw)
    return pwff(sub(p, "pwff"), att, dropout=cfg.get("dropout", 0.0), training=cfg.get("training", False))


def encoder(p: Params, features, pad_mask, cfg):
    """Encoder.forward — encoders.py:112-117."""
    out = _ln(p, "layer_norm", features) + sinusoid_pos(features, cfg["d_model"])
    for i in range(cfg["layers"]):
        out = encoder_layer(sub(p, f"layers.{i}"), out, out, out, pad_mask, cfg)
    return out


def geometric_encoder(p: Params, features, boxes, pad_mask, cfg):
    """GeometricEncoder.forward — encoders.py:130-135 (the class itself cannot be constructed upstream,
    Appendix C; this is its forward loop with ``boxes`` forwarded to the geometry core)."""
    out = _ln(p, "layer_norm", features) + sinusoid_pos(features, cfg["d_model"])
    for i in range(cfg["layers"]):
        out = encoder_layer(sub(p, f"layers.{i}"), out, out, out, pad_mask, cfg, boxes=boxes,
                            d_g=cfg["d_g"], trig=cfg.get("trig", True))
    return out


def guided_encoder(p: Params, vision, vision_mask, language, language_mask, cfg):
    """GuidedAttentionEncoder.forward + GuidedEncoderLayer.forward — encoders.py:152-164, 81-99."""
    out = _ln(p, "layer_norm", vision) + sinusoid_pos(vision, cfg["d_model"])
    a = _acfg(cfg)
    for i in range(cfg["layers"]):
        lp = sub(p, f"guided_attn_layers.{i}")
        self_att = mha(sub(lp, "self_mhatt"), out, out, out, vision_mask, **a)
        guided = mha(sub(lp, "guided_mhatt"), self_att, language, language, language_mask, **a)
        out = pwff(sub(lp, "pwff"), guided, dropout=a["dropout"], training=a["training"])
    return out


def co_attention_encoder(p: Params, vision, vision_mask, language, language_mask, cfg):
    """CoAttentionEncoder.forward — encoders.py:188-224 (ViLBERT-style, 4 EncoderLayers per level)."""
    pos = lambda t: sinusoid_pos(t, cfg["d_model"])
    v = _ln(p, "vision_layer_norm", vision) + pos(vision)
    l = _ln(p, "language_layer_norm", language) + pos(language)
    for i in range(cfg["layers"]):
        v = encoder_layer(sub(p, f"vision_language_attn_layers.{i}"), v, l, l, language_mask, cfg)
        l = encoder_layer(sub(p, f"language_vision_attn_layers.{i}"), l, v, v, vision_mask, cfg)
        v = encoder_layer(sub(p, f"vision_self_attn_layers.{i}"), v, v, v, vision_mask, cfg)
        l = encoder_layer(sub(p, f"language_self_attn_layers.{i}"), l, l, l, language_mask, cfg)
    return v, l


def cross_modality_encoder(p: Params, vision, vision_mask, language, language_mask, cfg, boxes=None):
    """CrossModalityEncoder.forward + CrossModalityEncoderLayer.forward — encoders.py:241-253, 37-72.

    The reference computes both cross-attentions and overwrites them (encoders.py:39-66, SURVEY finding 7);
    the outputs depend only on the self-attention + PWFF branch, which is what is restated.
    """
    pos = lambda t: sinusoid_pos(t, cfg["d_model"])
    v = _ln(p, "vision_layer_norm", vision) + pos(vision)
    l = _ln(p, "language_layer_norm", language) + pos(language)
    a = _acfg(cfg)
    gkw = dict(boxes=boxes, d_g=cfg.get("d_g"), trig=cfg.get("trig", True)) if boxes is not None else {}
    for i in range(cfg["layers"]):
        lp = sub(p, f"layers.{i}")
        va = mha(sub(lp, "vision_mhattn"), v, v, v, vision_mask, **a, **gkw)
        la = mha(sub(lp, "language_mhattn"), l, l, l, language_mask, **a)
        v = pwff(sub(lp, "vision_pff"), va, dropout=a["dropout"], training=a["training"])
        l = pwff(sub(lp, "language_pff"), la, dropout=a["dropout"], training=a["training"])
    return v, l


# --------------------------------------------------------------------------- MCAN (models/mcan.py)
def feature_embedding(p: Params, features, dropout=0.0, training=False):
    """FeatureEmbedding.forward — vision_embeddings.py:19-25."""
    mask = padding_mask(features, 0)
    return _drop(F.gelu(_lin(p, "proj", features)), dropout, training), mask


def lstm_forward(p: Params, x: torch.Tensor) -> torch.Tensor:
    """Single-layer batch-first ``nn.LSTM`` (text_embeddings.py:234,244), gate order i,f,g,o."""
    B, L, _ = x.shape
    H = p["weight_hh_l0"].shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    xg = F.linear(x, p["weight_ih_l0"], p["bias_ih_l0"])
    outs = []
    for t in range(L):
        g = xg[:, t] + F.linear(h, p["weight_hh_l0"], p["bias_hh_l0"])
        i, f, gg, o = g.chunk(4, dim=-1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, dim=1)


def lstm_text_embedding(p: Params, tokens, padding_idx=0, dropout=0.0, training=False):
    """LSTMTextEmbedding.forward — text_embeddings.py:236-246."""
    pad = padding_mask(tokens, padding_idx)
    seq = sequential_mask(tokens.shape[-1])
    feats = _lin(p, "proj", F.embedding(tokens, p["embedding.weight"], padding_idx=padding_idx))
    feats = _drop(feats, dropout, training)
    return lstm_forward(sub(p, "lstm"), feats), (pad, seq)


def attr_reduce_mlp(p: Params, x, dropout=0.0, training=False):
    """MLP.forward — mcan.py:21-25."""
    return _lin(p, "fc2", _drop(F.relu(_lin(p, "fc1", x)), dropout, training))


def mcan_forward(p: Params, region_features, question_tokens, *, d_model=512, h=8, layers=6,
                 dropout=0.0, training=False):
    """MCAN.forward — models/mcan.py:49-81.  Returns log-probabilities (B, n_answers)."""
    cfg = dict(d_model=d_model, h=h, d_k=d_model // h, d_v=d_model // h, layers=layers,
               dropout=dropout, training=training)
    vis, vmask = feature_embedding(sub(p, "vision_embedding"), region_features, dropout, training)
    txt, (tmask, _) = lstm_text_embedding(sub(p, "text_embedding"), question_tokens, 0, dropout, training)
    txt = encoder(sub(p, "self_encoder"), txt, tmask, cfg)                                   # SA   :57-60
    vis = guided_encoder(sub(p, "guided_encoder"), vis, vmask, txt, tmask, cfg)              # SGA  :63-68
    av = F.softmax(attr_reduce_mlp(sub(p, "vision_attr_reduce"), vis, dropout, training), dim=1)
    at = F.softmax(attr_reduce_mlp(sub(p, "text_attr_reduce"), txt, dropout, training), dim=1)
    wv = (vis * av).sum(dim=1)
    wt = (txt * at).sum(dim=1)
    out = _ln(p, "layer_norm", _lin(p, "vision_proj", wv) + _lin(p, "text_proj", wt))
    return F.log_softmax(_lin(p, "classify", out), dim=-1)


# --------------------------------------------------------------------------- classifier heads of the XMT / ViLBERT models
def xmt_forward(p: Params, region_features, question_tokens, *, kind: str, d_model=512, h=8, layers=2,
                boxes=None, d_g=None):
    """CrossModalityTransformer / VisiolinguisticTransformer forward —
    models/cross_modality_transformer.py:49-78, models/visiolinguistic_transformer.py:49-78 (same head as
    MCAN but returns raw logits).  ``kind`` is "cross" (CrossModalityEncoder) or "co" (CoAttentionEncoder)."""
    cfg = dict(d_model=d_model, h=h, d_k=d_model // h, d_v=d_model // h, layers=layers, d_g=d_g)
    vis, vmask = feature_embedding(sub(p, "vision_embedding"), region_features)
    txt, (tmask, _) = lstm_text_embedding(sub(p, "text_embedding"), question_tokens)
    if kind == "cross":
        vis, txt = cross_modality_encoder(sub(p, "encoder"), vis, vmask, txt, tmask, cfg, boxes=boxes)
    else:
        vis, txt = co_attention_encoder(sub(p, "encoder"), vis, vmask, txt, tmask, cfg)
    av = F.softmax(attr_reduce_mlp(sub(p, "vision_attr_reduce"), vis), dim=1)
    at = F.softmax(attr_reduce_mlp(sub(p, "text_attr_reduce"), txt), dim=1)
    out = _ln(p, "layer_norm", _lin(p, "vision_proj", (vis * av).sum(1)) + _lin(p, "text_proj", (txt * at).sum(1)))
    return _lin(p, "classify", out)


# --------------------------------------------------------------------------- M4C (models/mmf_m4c.py)
def ocr_ptr_net(p: Params, query_inputs, key_inputs, attention_mask, *, query_key_size: int):
    """OcrPtrNet.forward — mmf_m4c.py:379-396: (Wq·dec)(Wk·ocr)ᵀ/√dqk + mask(B,1,Locr)."""
    ext = attention_mask.squeeze(1)
    ql = _lin(p, "query", query_inputs)
    squeeze = ql.dim() == 2
    if squeeze:
        ql = ql.unsqueeze(1)
    kl = _lin(p, "key", key_inputs)
    scores = torch.matmul(ql, kl.transpose(-1, -2)) / math.sqrt(query_key_size) + ext
    return scores.squeeze(1) if squeeze else scores


def dynamic_pointer_network(p: Params, query_inputs, key_inputs, key_is_pad, *, d_model: int):
    """DynamicPointerNetwork.forward — m4c.py:27-33 (author's variant; SPEC ONLY, not runnable upstream):
    scores = QKᵀ/√d_model with masked_fill(-inf) on padded keys."""
    ql = _lin(p, "fc_q", query_inputs)
    kl = _lin(p, "fc_k", key_inputs)
    scores = torch.matmul(ql, kl.transpose(-1, -2)) / math.sqrt(d_model)
    return scores.masked_fill(key_is_pad.view(key_is_pad.shape[0], 1, -1), float("-inf"))


def batch_gather(x: torch.Tensor, inds: torch.Tensor) -> torch.Tensor:
    """_batch_gather — mmf_m4c.py:447-459: rows ``inds[b,t]`` of ``x[b]``."""
    B, L, D = x.shape
    flat = x.reshape(B * L, D)
    return F.embedding(torch.arange(B).unsqueeze(-1) * L + inds, flat)


def prev_pred_embeddings(p: Params, ans_emb, ocr_emb, prev_inds, *, eps: float = 1e-12):
    """PrevPredEmbeddings.forward — mmf_m4c.py:416-445 (eval: emb_dropout is identity)."""
    B, T = prev_inds.shape
    ans_num = ans_emb.size(0)
    ans = _ln(p, "ans_layer_norm", ans_emb, eps).unsqueeze(0).expand(B, -1, -1)
    ocr = _ln(p, "ocr_layer_norm", ocr_emb, eps)
    raw = batch_gather(torch.cat([ans, ocr], dim=1), prev_inds)
    pos = F.embedding(torch.arange(T).unsqueeze(0).expand(B, T), p["position_embeddings.weight"])
    typ = F.embedding(prev_inds.ge(ans_num).long(), p["token_type_embeddings.weight"])
    return raw + _ln(p, "emb_layer_norm", pos + typ, eps)


def bert_layer(p: Params, x, mask, *, heads: int, eps: float = 1e-12):
    """One HF BertLayer in eval mode (third-party arithmetic, see module docstring):
    softmax(QKᵀ·d^-½ + mask)V → dense → LN(·+x) → dense → GELU(erf) → dense → LN(·+·)."""
    B, L, D = x.shape
    dh = D // heads
    shp = (B, L, heads, dh)
    q = _lin(p, "attention.self.query", x).view(shp).transpose(1, 2)
    k = _lin(p, "attention.self.key", x).view(shp).transpose(1, 2)
    v = _lin(p, "attention.self.value", x).view(shp).transpose(1, 2)
    w = torch.matmul(q, k.transpose(2, 3)) * (dh ** -0.5)
    if mask is not None:
        w = w + mask
    w = F.softmax(w, dim=-1)
    ctx = torch.matmul(w, v).transpose(1, 2).contiguous().reshape(B, L, D)
    att = _ln(p, "attention.output.LayerNorm", _lin(p, "attention.output.dense", ctx) + x, eps)
    inter = F.gelu(_lin(p, "intermediate.dense", att))
    return _ln(p, "output.LayerNorm", _lin(p, "output.dense", inter) + att, eps)


def bert_encoder(p: Params, x, mask, *, layers: int, heads: int, eps: float = 1e-12):
    for i in range(layers):
        x = bert_layer(sub(p, f"layer.{i}"), x, mask, heads=heads, eps=eps)
    return x


def text_bert(p: Params, txt_inds, txt_mask, *, layers: int, heads: int, eps: float = 1e-12):
    """TextBert.forward — mmf_m4c.py:266-281 (HF BertEmbeddings: word + type(0) + position → LN)."""
    B, L = txt_inds.shape
    e = F.embedding(txt_inds, p["embeddings.word_embeddings.weight"])
    e = e + F.embedding(torch.zeros((B, L), dtype=torch.long), p["embeddings.token_type_embeddings.weight"])
    e = e + F.embedding(torch.arange(L).unsqueeze(0), p["embeddings.position_embeddings.weight"])
    e = _ln(p, "embeddings.LayerNorm", e, eps)
    return bert_encoder(sub(p, "encoder"), e, txt_mask, layers=layers, heads=heads, eps=eps)


def mmt_mask(txt_mask, obj_mask, ocr_mask, dec_len: int) -> torch.Tensor:
    """The (B,1,L,L) prefix-LM additive mask of MMT.forward — mmf_m4c.py:310-343.

    Decoder *columns* are 0 (attend-able) for encoder rows (``dec_mask = zeros``), and the dec×dec block is
    the causal mask — reproduced exactly (SURVEY §3.3 note ii, Appendix C)."""
    B = txt_mask.shape[0]
    dec_mask = torch.zeros(B, dec_len, dtype=torch.float32).unsqueeze(1).unsqueeze(2)
    am = torch.cat([txt_mask, obj_mask, ocr_mask, dec_mask], dim=-1)
    L = am.size(-1)
    ext = am.repeat(1, 1, L, 1)
    ext[:, :, -dec_len:, -dec_len:] = sequential_mask(dec_len)
    return ext


def mmt_forward(p: Params, txt_emb, txt_mask, obj_emb, obj_mask, ocr_emb, ocr_mask, fixed_ans_emb, prev_inds, *,
                layers: int, heads: int, eps: float = 1e-12):
    """MMT.forward — mmf_m4c.py:290-364.  Returns (mmt_ocr_output, mmt_dec_output)."""
    dec_emb = prev_pred_embeddings(sub(p, "prev_pred_embeddings"), fixed_ans_emb, ocr_emb, prev_inds, eps=eps)
    T = dec_emb.size(1)
    x = torch.cat([txt_emb, obj_emb, ocr_emb, dec_emb], dim=1)
    ext = mmt_mask(txt_mask, obj_mask, ocr_mask, T)
    y = bert_encoder(sub(p, "encoder"), x, ext, layers=layers, heads=heads, eps=eps)
    ocr_begin = txt_mask.size(-1) + obj_mask.size(-1)
    return y[:, ocr_begin:ocr_begin + ocr_mask.size(-1)], y[:, -T:]


def m4c_encode_inputs(p: Params, batch: dict):
    """MMF_M4C._forward_{txt,obj,ocr}_encoding — mmf_m4c.py:128-200 (eval)."""
    txt_mask = padding_mask(batch["question_tokens"], 0)
    obj = _ln(p, "obj_feat_layer_norm", _lin(p, "linear_obj_feat_to_mmt_in", batch["region_features"])) + \
        _ln(p, "obj_bbox_layer_norm", _lin(p, "linear_obj_bbox_to_mmt_in", batch["region_boxes"]))
    obj_mask = padding_mask(batch["region_features"], 0)
    ocr_feat = torch.cat([F.normalize(batch["ocr_fasttext_features"], dim=-1),
                          F.normalize(batch["ocr_rec_features"], dim=-1),
                          F.normalize(batch["ocr_det_features"], dim=-1)], dim=-1)
    ocr = _ln(p, "ocr_feat_layer_norm", _lin(p, "linear_ocr_feat_to_mmt_in", ocr_feat)) + \
        _ln(p, "ocr_bbox_layer_norm", _lin(p, "linear_ocr_bbox_to_mmt_in", batch["ocr_boxes"]))
    ocr_mask = padding_mask(ocr_feat, 0)
    return txt_mask, obj, obj_mask, ocr, ocr_mask


def m4c_step(p: Params, batch: dict, enc, prev_inds, *, mmt_layers: int, text_layers: int, heads: int,
             ptr_qk: int):
    """One _forward_mmt + _forward_output — mmf_m4c.py:202-229.  Returns scores (B,T,V+Locr)."""
    txt_mask, obj, obj_mask, ocr, ocr_mask = enc
    txt = text_bert(sub(p, "text_bert"), batch["question_tokens"], txt_mask, layers=text_layers, heads=heads)
    if "text_bert_out_linear.weight" in p:
        txt = _lin(p, "text_bert_out_linear", txt)
    ocr_out, dec_out = mmt_forward(sub(p, "mmt"), txt, txt_mask, obj, obj_mask, ocr, ocr_mask,
                                   p["classifier.weight"], prev_inds, layers=mmt_layers, heads=heads)
    fixed = _lin(p, "classifier", dec_out)
    dyn = ocr_ptr_net(sub(p, "ocr_ptr_net"), dec_out, ocr_out, ocr_mask, query_key_size=ptr_qk)
    return torch.cat([fixed, dyn], dim=-1)


def m4c_greedy_decode(p: Params, batch: dict, *, max_iter: int = 12, mmt_layers: int = 4, text_layers: int = 3,
                      heads: int = 8, ptr_qk: int = 768, bos_idx: int = 1, eos_idx: int = 2,
                      early_exit: bool = True):
    """MMF_M4C._forward_mmt_and_output (eval branch) — mmf_m4c.py:236-256.

    Returns ``(scores of the last executed iteration, prev_inds, n_iterations)``."""
    B = batch["question_tokens"].shape[0]
    enc = m4c_encode_inputs(p, batch)
    prev_inds = torch.zeros((B, max_iter), dtype=torch.long)
    prev_inds[:, 0] = bos_idx
    last_ids = torch.zeros((B,))
    scores, n = None, 0
    for ith in range(max_iter):
        scores = m4c_step(p, batch, enc, prev_inds, mmt_layers=mmt_layers, text_layers=text_layers,
                          heads=heads, ptr_qk=ptr_qk)
        n += 1
        argmax_inds = scores.argmax(dim=-1)
        prev_inds[:, 1:] = argmax_inds[:, :-1]
        last_ids = torch.where(last_ids == eos_idx, last_ids, argmax_inds[:, ith].to(last_ids.dtype))
        if early_exit and last_ids.mean() == eos_idx:      # :255 (float mean == int: all rows emitted eos)
            break
    return scores, prev_inds, n
