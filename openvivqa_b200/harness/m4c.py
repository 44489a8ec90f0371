"""Host-side mirror of ``MMF_M4C`` (models/mmf_m4c.py:24-459) wired to the B200 kernels — BASELINE config 3.

Same ``state_dict()`` names as the reference module (HF ``BertEncoder`` naming for ``text_bert`` / ``mmt.encoder``), so
reference checkpoints and the oracle's parameter dicts load with ``strict=True``.  The classes below are parameter containers;
ALL arithmetic — the HF-BERT layers through the fused blocks, ``PrevPredEmbeddings``, the prefix-LM mask, the pointer scores and the
restructured greedy loop — lives in :mod:`openvivqa_b200.m4c`, the same functions ``install()`` binds over the reference's own classes.
This mirror exists because the GPU box has no reference checkout.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as Fn
from .. import m4c as M4
from ..modules import OcrPtrNet, get_compute_dtype
from .models import generate_padding_mask


class _BertSelf(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)


class _BertDenseLN(nn.Module):
    def __init__(self, d_in, d_out, eps):
        super().__init__()
        self.dense = nn.Linear(d_in, d_out)
        self.LayerNorm = nn.LayerNorm(d_out, eps=eps)


class _BertAttention(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.self = _BertSelf(d)
        self.output = _BertDenseLN(d, d, eps)


class _BertIntermediate(nn.Module):
    def __init__(self, d, d_ff):
        super().__init__()
        self.dense = nn.Linear(d, d_ff)


class BertLayer(nn.Module):
    """HF ``BertLayer`` parameters (no dropout modules: the mirror is the inference model of BASELINE config 3)."""

    def __init__(self, d, heads, d_ff=3072, eps=1e-12):
        super().__init__()
        self.attention = _BertAttention(d, eps)
        self.intermediate = _BertIntermediate(d, d_ff)
        self.output = _BertDenseLN(d_ff, d, eps)
        self.heads = heads

    def forward(self, x, mask, mode, q_rows=None):
        return M4.bert_layer(self, x, mask, mode, heads=self.heads, q_rows=q_rows, training=self.training)


class BertEncoder(nn.Module):
    def __init__(self, d, heads, layers, d_ff=3072, eps=1e-12):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(d, heads, d_ff, eps) for _ in range(layers)])
        self.heads = heads

    def forward(self, x, mask, mode, last_q_rows=None):
        return M4.bert_encoder(self, x, mask, mode, heads=self.heads, last_q_rows=last_q_rows, training=self.training)


class _BertEmbeddings(nn.Module):
    def __init__(self, d, vocab_size=30522, max_pos=512, eps=1e-12):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab_size, d, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, d)
        self.token_type_embeddings = nn.Embedding(2, d)
        self.LayerNorm = nn.LayerNorm(d, eps=eps)

    def forward(self, ids):
        return M4.bert_embeddings(self, ids)


class TextBert(nn.Module):
    """mmf_m4c.py:266-281."""

    def __init__(self, d, heads, layers, vocab_size=30522):
        super().__init__()
        self.embeddings = _BertEmbeddings(d, vocab_size)
        self.encoder = BertEncoder(d, heads, layers)
        self.heads = heads

    def forward(self, txt_inds, txt_mask, mode):
        return M4.text_bert(self, txt_inds, txt_mask, mode, heads=self.heads)


class PrevPredEmbeddings(nn.Module):
    """mmf_m4c.py:399-445 — parameters; the arithmetic is ``m4c.prev_pred_embeddings`` (``ovqa_prev_pred_embed``)."""

    def __init__(self, d, eps=1e-12, max_dec_length=100, max_type_num=5):
        super().__init__()
        self.position_embeddings = nn.Embedding(max_dec_length, d)
        self.token_type_embeddings = nn.Embedding(max_type_num, d)
        self.ans_layer_norm = nn.LayerNorm(d, eps=eps)
        self.ocr_layer_norm = nn.LayerNorm(d, eps=eps)
        self.emb_layer_norm = nn.LayerNorm(d, eps=eps)

    def forward(self, ans_emb, ocr_emb, prev_inds):
        return M4.prev_pred_embeddings(self, ans_emb, ocr_emb, prev_inds)


class MMT(nn.Module):
    """mmf_m4c.py:283-364."""

    def __init__(self, d, heads, layers):
        super().__init__()
        self.prev_pred_embeddings = PrevPredEmbeddings(d)
        self.encoder = BertEncoder(d, heads, layers)
        self.heads = heads

    def forward(self, txt_emb, txt_mask, obj_emb, obj_mask, ocr_emb, ocr_mask, fixed_ans_emb, prev_inds, mode="fp32"):
        return M4.mmt_forward(self, txt_emb, txt_mask, obj_emb, obj_mask, ocr_emb, ocr_mask, fixed_ans_emb, prev_inds, heads=self.heads,
                              mode=mode, training=self.training)


class MMF_M4C(nn.Module):
    """models/mmf_m4c.py:24-256, inference (greedy decode) path."""

    def __init__(self, config, vocab):
        super().__init__()
        d = config.MMT.HIDDEN_SIZE
        heads = config.MMT.NUM_ATTENTION_HEADS
        self.d, self.heads = d, heads
        self.max_iter = vocab.max_answer_length
        self.bos_idx, self.eos_idx, self.padding_idx = vocab.bos_idx, vocab.eos_idx, vocab.padding_idx
        self.text_bert = TextBert(config.TEXT_BERT.HIDDEN_SIZE, heads, config.TEXT_BERT.NUM_HIDDEN_LAYERS)
        if d != 768:                                                        # mmf_m4c.py:62-72 (compares with the constant 768)
            self.text_bert_out_linear = nn.Linear(config.TEXT_BERT.HIDDEN_SIZE, d)
        else:
            self.text_bert_out_linear = nn.Identity()
        self.linear_obj_feat_to_mmt_in = nn.Linear(config.OBJECT_EMBEDDING.D_FEATURE, d)
        self.linear_obj_bbox_to_mmt_in = nn.Linear(4, d)
        self.obj_feat_layer_norm, self.obj_bbox_layer_norm = nn.LayerNorm(d), nn.LayerNorm(d)
        self.linear_ocr_feat_to_mmt_in = nn.Linear(config.OCR_EMBEDDING.D_FEATURE, d)
        self.linear_ocr_bbox_to_mmt_in = nn.Linear(4, d)
        self.ocr_feat_layer_norm, self.ocr_bbox_layer_norm, self.ocr_text_layer_norm = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)
        self.mmt = MMT(d, heads, config.MMT.NUM_HIDDEN_LAYERS)
        self.ocr_ptr_net = OcrPtrNet(config.OCR_PTR_NET.HIDDEN_SIZE, config.OCR_PTR_NET.QUERY_KEY_SIZE)
        self.classifier = nn.Linear(d, len(vocab))
        self.compute_dtype = getattr(config, "COMPUTE_DTYPE", None)
        self.ocr_ptr_net.compute_dtype = self.compute_dtype          # the pointer network follows the model's compute dtype
        self.prune_last_layer = True

    # ------------------------------------------------------------------ step-invariant part (once per batch)
    def _mode(self):
        return self.compute_dtype or get_compute_dtype()

    def _lin_ln(self, lin, ln, x, mode, post_add=None):
        # bf16 operands need a 16-byte row pitch (TMA): the 4-wide box projections stay on the fp32 path, wider inputs whose
        # width is not a multiple of 8 (OCR features: 256 + 256 + 300 = 812) are zero-padded to one (once per batch)
        K = x.shape[-1]
        if mode != "fp32" and K % 8 != 0 and K >= 64:
            pad = 8 - K % 8
            y = Fn.linear(F.pad(x, (0, pad)), F.pad(lin.weight, (0, pad)), lin.bias, mode)
            return Fn.layer_norm(y, ln.weight, ln.bias, ln.eps, post_add=post_add)
        m = mode if K % 8 == 0 else "fp32"
        y = Fn.linear(x, lin.weight, lin.bias, m)
        return Fn.layer_norm(y, ln.weight, ln.bias, ln.eps, post_add=post_add)

    def encode_inputs(self, items, mode):
        """_forward_{txt,obj,ocr}_encoding (mmf_m4c.py:128-200): masks and the object / OCR inputs of MMT."""
        txt_mask = generate_padding_mask(items.question_tokens, self.padding_idx)
        obj_box = self._lin_ln(self.linear_obj_bbox_to_mmt_in, self.obj_bbox_layer_norm, items.region_boxes, mode)
        obj = self._lin_ln(self.linear_obj_feat_to_mmt_in, self.obj_feat_layer_norm, items.region_features, mode, post_add=None) + obj_box
        obj_mask = generate_padding_mask(items.region_features, 0)
        ocr_feat = torch.cat([F.normalize(items.ocr_fasttext_features, dim=-1), F.normalize(items.ocr_rec_features, dim=-1),
                              F.normalize(items.ocr_det_features, dim=-1)], dim=-1)
        ocr_box = self._lin_ln(self.linear_ocr_bbox_to_mmt_in, self.ocr_bbox_layer_norm, items.ocr_boxes, mode)
        ocr = self._lin_ln(self.linear_ocr_feat_to_mmt_in, self.ocr_feat_layer_norm, ocr_feat, mode) + ocr_box
        ocr_mask = generate_padding_mask(ocr_feat, 0)
        return txt_mask, obj, obj_mask, ocr, ocr_mask

    mmt_mask = staticmethod(M4.mmt_mask)

    # ------------------------------------------------------------------ decode
    def prepare(self, items):
        """Allocate the persistent decode buffers for a batch and run the step-invariant part (incl. the hoisted TextBert pass)."""
        mode = self._mode()
        txt_mask, obj, obj_mask, ocr, ocr_mask = self.encode_inputs(items, mode)
        txt = M4.encode_text(self, items.question_tokens, txt_mask, mode)
        return M4.prepare_decode(self, txt, txt_mask, obj, obj_mask, ocr, ocr_mask, mode)

    def step(self, st, ith):
        """One _forward_mmt + _forward_output + greedy update (mmf_m4c.py:202-229,247-253)."""
        M4.decode_step(self, st, ith, self.prune_last_layer)

    @torch.no_grad()
    def forward(self, items, early_exit=True):
        """Greedy decode.  Returns ``{"scores": (B,T,V+Locr) of the last executed iteration, "prev_inds": (B,T), "iters": n}``."""
        st = self.prepare(items)
        n = 0
        for ith in range(self.max_iter):
            self.step(st, ith)
            n += 1
            if early_exit and st.all_eos.item():                           # mmf_m4c.py:255 (one host sync per step, as upstream)
                break
        return {"scores": st.scores[..., :st.W], "prev_inds": st.prev_inds, "iters": n}
