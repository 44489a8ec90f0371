"""Host-side mirror of ``MMF_M4C`` (models/mmf_m4c.py:24-459) wired to the B200 kernels — BASELINE config 3.

Same ``state_dict()`` names as the reference module (HF ``BertEncoder`` naming for ``text_bert`` / ``mmt.encoder``), so
reference checkpoints and the oracle's parameter dicts load with ``strict=True``.

B200-first restructuring of the greedy decode loop (mmf_m4c.py:236-256) — results identical to the reference's:

* the BERT layers (third-party arithmetic) run through the SAME fused blocks as the repo's own
  attention (``functional.mha_block`` / ``pwff_block`` with eps 1e-12 and d_ff 3072): one fused QKV GEMM, the
  tensor-core attention core with the (B,1,L,L) prefix-LM mask, output projection + residual + LayerNorm;
* everything that does not depend on ``prev_inds`` is computed ONCE per batch instead of once per step: the
  object / OCR input encodings, the masks, and the TextBert pass + ``text_bert_out_linear`` (the reference re-runs
  TextBert in every one of the 12 iterations, mmf_m4c.py:203-206; in eval mode it is a pure function of the tokens);
* the joint ``(B, L, D)`` encoder input is one persistent buffer: text / object / OCR rows are written once, the
  decoder rows are written in place each step by ``ovqa_prev_pred_embed`` (no ``(B, V+Locr, D)`` table, no cat);
* scores live in one ``(B, T, V+Locr)`` buffer: the classifier GEMM writes columns ``[0, V)``, the pointer kernel
  writes ``[V, V+Locr)`` (no ``torch.cat``, mmf_m4c.py:227);
* argmax / shift / EOS latch are one kernel pair; the host reads the all-EOS flag only when ``early_exit`` is on.

MMT must be re-run every step: encoder rows attend to the (unmasked) decoder columns (mmf_m4c.py:310-314), so
there is no prefix to cache.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .. import functional as Fn
from .. import ops
from ..modules import OcrPtrNet, get_compute_dtype
from .models import MASK_VALUE, generate_padding_mask, generate_sequential_mask


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
    """HF ``BertLayer`` (eval): ``LN(dense(softmax(QK^T d^-1/2 + mask) V) + x)`` then ``LN(dense(gelu(dense ·)) + ·)``,
    eps 1e-12 — computed by the fused MHA / PWFF blocks of the hot path."""

    def __init__(self, d, heads, d_ff=3072, eps=1e-12):
        super().__init__()
        self.attention = _BertAttention(d, eps)
        self.intermediate = _BertIntermediate(d, d_ff)
        self.output = _BertDenseLN(d_ff, d, eps)
        self.heads, self.d_head, self.eps = heads, d // heads, eps

    def forward(self, x, mask, mode, q_rows=None):
        """``q_rows`` (slice): compute the layer only for those query rows (keys/values: all rows) — used for the LAST
        layer of the decode step, whose other rows nobody reads."""
        if self.training:
            raise RuntimeError("harness BertLayer is inference-only (BASELINE config 3 is M4C *decode*)")
        a, s = self.attention, self.attention.self
        xq, mq = x, mask
        if q_rows is not None:
            xq = x[:, q_rows].contiguous()
            mq = mask[:, :, q_rows] if (mask is not None and mask.shape[2] > 1) else mask
            hint = getattr(mask, "_ovqa_prefix_lm", None)
            if hint is not None and mq is not mask and q_rows.stop in (None, mask.shape[2]) and mq.shape[2] >= hint[1]:
                mq._ovqa_prefix_lm = hint                     # the row slice still ends with the causal tail
        att = Fn.mha_block(xq, x, x, mq, wq=s.query.weight, bq=s.query.bias, wk=s.key.weight, bk=s.key.bias,
                           wv=s.value.weight, bv=s.value.bias, wo=a.output.dense.weight, bo=a.output.dense.bias,
                           gamma=a.output.LayerNorm.weight, beta=a.output.LayerNorm.bias, h=self.heads, d_k=self.d_head,
                           d_v=self.d_head, mode=mode, eps=self.eps)
        return Fn.pwff_block(att, w1=self.intermediate.dense.weight, b1=self.intermediate.dense.bias,
                             w2=self.output.dense.weight, b2=self.output.dense.bias, gamma=self.output.LayerNorm.weight,
                             beta=self.output.LayerNorm.bias, mode=mode, eps=self.eps)


class BertEncoder(nn.Module):
    def __init__(self, d, heads, layers, d_ff=3072, eps=1e-12):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(d, heads, d_ff, eps) for _ in range(layers)])

    def forward(self, x, mask, mode, last_q_rows=None):
        n = len(self.layer)
        for i, layer in enumerate(self.layer):
            x = layer(x, mask, mode, q_rows=last_q_rows if i == n - 1 else None)
        return x


class _BertEmbeddings(nn.Module):
    def __init__(self, d, vocab_size=30522, max_pos=512, eps=1e-12):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab_size, d, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, d)
        self.token_type_embeddings = nn.Embedding(2, d)
        self.LayerNorm = nn.LayerNorm(d, eps=eps)

    def forward(self, ids):
        L = ids.shape[1]
        pos_type = self.position_embeddings.weight[:L] + self.token_type_embeddings.weight[0]
        e = self.word_embeddings(ids)                                   # gather: library op off the hot path
        return Fn.layer_norm(e + pos_type, self.LayerNorm.weight, self.LayerNorm.bias, self.LayerNorm.eps)


class TextBert(nn.Module):
    """mmf_m4c.py:266-281."""

    def __init__(self, d, heads, layers, vocab_size=30522):
        super().__init__()
        self.embeddings = _BertEmbeddings(d, vocab_size)
        self.encoder = BertEncoder(d, heads, layers)

    def forward(self, txt_inds, txt_mask, mode):
        return self.encoder(self.embeddings(txt_inds), txt_mask, mode)


class PrevPredEmbeddings(nn.Module):
    """mmf_m4c.py:399-445 — parameters only; the arithmetic is ``ovqa_prev_pred_embed``."""

    def __init__(self, d, eps=1e-12, max_dec_length=100, max_type_num=5):
        super().__init__()
        self.position_embeddings = nn.Embedding(max_dec_length, d)
        self.token_type_embeddings = nn.Embedding(max_type_num, d)
        self.ans_layer_norm = nn.LayerNorm(d, eps=eps)
        self.ocr_layer_norm = nn.LayerNorm(d, eps=eps)
        self.emb_layer_norm = nn.LayerNorm(d, eps=eps)
        self.eps = eps

    def write(self, ans_emb, ocr_emb, prev_inds, out):
        ln = lambda m: (m.weight, m.bias)
        return ops.prev_pred_embed(ans_emb, ocr_emb, prev_inds, self.position_embeddings.weight, self.token_type_embeddings.weight,
                                   ln(self.ans_layer_norm), ln(self.ocr_layer_norm), ln(self.emb_layer_norm), self.eps, out=out)


class MMT(nn.Module):
    """mmf_m4c.py:283-364 (parameters; the forward is driven by :class:`MMF_M4C`)."""

    def __init__(self, d, heads, layers):
        super().__init__()
        self.prev_pred_embeddings = PrevPredEmbeddings(d)
        self.encoder = BertEncoder(d, heads, layers)


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
        self.ocr_ptr_net.compute_dtype = self.compute_dtype          # the pointer network follows the model's compute dtype (it fell back
                                                                     # to the global default: two fp32 CUDA-core GEMMs per decode step, 20 % of the decode)
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
        """_forward_{txt,obj,ocr}_encoding (mmf_m4c.py:128-200) + the TextBert pass hoisted out of the loop."""
        txt_mask = generate_padding_mask(items.question_tokens, self.padding_idx)
        obj_box = self._lin_ln(self.linear_obj_bbox_to_mmt_in, self.obj_bbox_layer_norm, items.region_boxes, mode)
        obj = self._lin_ln(self.linear_obj_feat_to_mmt_in, self.obj_feat_layer_norm, items.region_features, mode, post_add=None) + obj_box
        obj_mask = generate_padding_mask(items.region_features, 0)
        ocr_feat = torch.cat([F.normalize(items.ocr_fasttext_features, dim=-1), F.normalize(items.ocr_rec_features, dim=-1),
                              F.normalize(items.ocr_det_features, dim=-1)], dim=-1)
        ocr_box = self._lin_ln(self.linear_ocr_bbox_to_mmt_in, self.ocr_bbox_layer_norm, items.ocr_boxes, mode)
        ocr = self._lin_ln(self.linear_ocr_feat_to_mmt_in, self.ocr_feat_layer_norm, ocr_feat, mode) + ocr_box
        ocr_mask = generate_padding_mask(ocr_feat, 0)
        txt = self.text_bert(items.question_tokens, txt_mask, mode)
        if not isinstance(self.text_bert_out_linear, nn.Identity):
            txt = Fn.linear(txt, self.text_bert_out_linear.weight, self.text_bert_out_linear.bias, mode)
        return txt, txt_mask, obj, obj_mask, ocr, ocr_mask

    @staticmethod
    def mmt_mask(txt_mask, obj_mask, ocr_mask, T):
        """(B,1,L,L) prefix-LM mask of MMT.forward (mmf_m4c.py:310-343), decoder columns open for encoder rows."""
        B, dev = txt_mask.shape[0], txt_mask.device
        am = torch.cat([txt_mask, obj_mask, ocr_mask, torch.zeros((B, 1, 1, T), device=dev)], dim=-1).float()
        L = am.shape[-1]
        ext = am.repeat(1, 1, L, 1)
        ext[:, :, -T:, -T:] = generate_sequential_mask(T, dev)
        ext = ext.contiguous()
        # structure hint for the attention core: the dense tensor is a per-key row plus a causal T x T tail, which the
        # tensor-memory kernel evaluates without reading (B, L, L) floats per launch (ops.sdpa_fwd; dense mask otherwise)
        ext._ovqa_prefix_lm = (am.contiguous(), T, float(MASK_VALUE))
        return ext

    # ------------------------------------------------------------------ decode
    class State:
        pass

    def prepare(self, items):
        """Allocate the persistent decode buffers for a batch and run the step-invariant part."""
        mode = self._mode()
        txt, txt_mask, obj, obj_mask, ocr, ocr_mask = self.encode_inputs(items, mode)
        B, dev = txt.shape[0], txt.device
        T, D = self.max_iter, self.d
        Lt, No, Nc = txt.shape[1], obj.shape[1], ocr.shape[1]
        st = self.State()
        st.mode, st.B, st.T, st.Lt, st.No, st.Nc = mode, B, T, Lt, No, Nc
        st.L = Lt + No + Nc + T
        st.x = torch.empty((B, st.L, D), device=dev, dtype=torch.float32)
        st.x[:, :Lt], st.x[:, Lt:Lt + No], st.x[:, Lt + No:Lt + No + Nc] = txt, obj, ocr
        st.ocr_in = ocr.contiguous()
        st.ocr_mask = ocr_mask
        st.mask = self.mmt_mask(txt_mask, obj_mask, ocr_mask, T)
        st.V = self.classifier.weight.shape[0]
        st.W = st.V + Nc
        st.Wp = (st.W + 3) // 4 * 4                                        # row pitch of the score buffer: 16-byte rows
        st.scores = torch.full((B, T, st.Wp), float("-inf"), device=dev, dtype=torch.float32)
        st.prev_inds = torch.zeros((B, T), dtype=torch.long, device=dev)
        st.prev_inds[:, 0] = self.bos_idx
        st.last_ids = torch.zeros((B,), dtype=torch.long, device=dev)
        st.argmax = torch.empty((B, T), dtype=torch.long, device=dev)
        st.all_eos = torch.zeros((1,), dtype=torch.int32, device=dev)
        st.rows = slice(Lt + No, st.L)                                     # OCR + decoder rows: all the last layer must produce
        return st

    def step(self, st, ith):
        """One _forward_mmt + _forward_output + greedy update (mmf_m4c.py:202-229,247-253)."""
        mode = st.mode
        self.mmt.prev_pred_embeddings.write(self.classifier.weight, st.ocr_in, st.prev_inds, st.x[:, -st.T:])
        st.x._ovqa_bf16 = None                                             # the buffer changed in place: drop any shadow
        if self.prune_last_layer:
            y = self.mmt.encoder(st.x, st.mask, mode, last_q_rows=st.rows)  # (B, Nc+T, D)
            ocr_out, dec_out = y[:, :st.Nc], y[:, st.Nc:]
        else:
            y = self.mmt.encoder(st.x, st.mask, mode)
            ocr_out, dec_out = y[:, st.Lt + st.No:st.Lt + st.No + st.Nc], y[:, -st.T:]
        s2 = st.scores.view(st.B * st.T, st.Wp)
        dec2 = Fn._operand(dec_out.contiguous(), mode)
        ops.gemm(dec2, Fn.weights.get([self.classifier.weight], Fn._cdt(mode)), bias=Fn.weights.get([self.classifier.bias], torch.float32),
                 out=s2[:, :st.V])
        self.ocr_ptr_net(dec_out.contiguous(), ocr_out.contiguous(), st.ocr_mask, scores_out=st.scores, col_offset=st.V)
        ops.greedy_step(st.scores, st.prev_inds, st.last_ids, st.argmax, st.all_eos, ith, self.eos_idx)

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
