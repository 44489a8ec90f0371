"""Attribute-style config trees with the reference YAML key names (configs/mcan.yaml:49-112 etc.); stand-in for
yacs.CfgNode, which the image does not have.  Scalars default to BASELINE.json's numbers."""
from __future__ import annotations

import types


class NS(types.SimpleNamespace):
    def __contains__(self, k):
        return hasattr(self, k)


class StubVocab:
    """The attributes the hot-path models read from a vocab (data_utils/vocabs/vocab.py:34,55-58)."""

    padding_idx, bos_idx, eos_idx, unk_idx = 0, 1, 2, 3

    def __init__(self, size=4096, total_answers=353, max_answer_length=12):
        self._size, self.total_answers, self.max_answer_length = size, total_answers, max_answer_length

    def __len__(self):
        return self._size


def attention_cfg(d_model=512, head=8, d_key=None, d_value=None, d_ff=2048, dropout=0.1,
                  arch="ScaledDotProductAttention", use_aoa=False, stateful=False, trig=True, compute_dtype=None):
    d_key = d_model // head if d_key is None else d_key
    d_value = d_model // head if d_value is None else d_value
    c = NS(ARCHITECTURE=arch, HEAD=head, D_MODEL=d_model, D_KEY=d_key, D_VALUE=d_value, D_FF=d_ff, D_FEATURE=2048,
           USE_AOA=use_aoa, CAN_BE_STATEFUL=stateful, DROPOUT=dropout, TRIGNOMETRIC_EMBEDDING=trig)
    if compute_dtype is not None:
        c.COMPUTE_DTYPE = compute_dtype
    return c


def _embeddings(d_model, d_feature, dropout, compute_dtype=None):
    vis = NS(ARCHITECTURE="FeatureEmbedding", D_FEATURE=d_feature, D_MODEL=d_model, DROPOUT=dropout)
    if compute_dtype is not None:
        vis.COMPUTE_DTYPE = compute_dtype
    return dict(
        VISION_EMBEDDING=vis,
        TEXT_EMBEDDING=NS(ARCHITECTURE="LSTMTextEmbedding", D_MODEL=d_model, D_EMBEDDING=300, DROPOUT=dropout,
                          WORD_EMBEDDING=None, WORD_EMBEDDING_CACHE=None),
        VISION_ATTR_REDUCE=NS(D_MODEL=d_model, DROPOUT=dropout), TEXT_ATTR_REDUCE=NS(D_MODEL=d_model, DROPOUT=dropout))


def mcan_cfg(d_model=512, d_feature=2048, layers=6, head=8, d_ff=2048, dropout=0.1, compute_dtype=None):
    att = lambda: attention_cfg(d_model, head, d_ff=d_ff, dropout=dropout, compute_dtype=compute_dtype)
    return NS(ARCHITECTURE="MCAN", NAME="mcan", DEVICE="cuda", D_MODEL=d_model,
              SELF_ENCODER=NS(ARCHITECTURE="Encoder", D_MODEL=d_model, LAYERS=layers, SELF_ATTENTION=att()),
              GUIDED_ENCODER=NS(ARCHITECTURE="GuidedAttentionEncoder", D_MODEL=d_model, LAYERS=layers,
                                SELF_ATTENTION=att(), GUIDED_ATTENTION=att()),
              **_embeddings(d_model, d_feature, dropout, compute_dtype))


def xmt_cfg(kind="cross", d_model=512, d_feature=2048, layers=2, head=8, d_ff=2048, dropout=0.1, geometry=False,
            compute_dtype=None):
    """configs/cross_modality_transformer*.yaml / visiolinguistic_transformer*.yaml structure; ``geometry`` switches
    the vision self-attention core to AugmentedGeometryScaledDotProductAttention (BASELINE config 4)."""
    att = lambda arch="ScaledDotProductAttention": attention_cfg(d_model, head, d_ff=d_ff, dropout=dropout, arch=arch,
                                                                 compute_dtype=compute_dtype)
    enc = NS(ARCHITECTURE="CrossModalityEncoder" if kind == "cross" else "CoAttentionEncoder", D_MODEL=d_model, LAYERS=layers,
             VISION_LANGUAGE_ATTENTION=att(), LANGUAGE_VISION_ATTENTION=att(),
             VISION_SELF_ATTENTION=att("AugmentedGeometryScaledDotProductAttention" if geometry else "ScaledDotProductAttention"),
             LANGUAGE_SELF_ATTENTION=att())
    return NS(ARCHITECTURE="CrossModalityTransformer", NAME="xmt", DEVICE="cuda", D_MODEL=d_model, ENCODER=enc,
              USE_BOXES=geometry, **_embeddings(d_model, d_feature, dropout, compute_dtype))


def m4c_cfg(hidden=768, heads=8, mmt_layers=4, text_layers=3, d_obj=1024, d_ocr=812, ptr_qk=None, compute_dtype=None):
    """configs/mmf_m4c.yaml:64-100 structure with BASELINE config 3's numbers (768 hidden, 8 heads -> d_head 96,
    4 MMT layers, 3 TextBert layers, random init: ``LOAD_PRETRAINED False``)."""
    c = NS(ARCHITECTURE="MMF_M4C", NAME="mmf_m4c", DEVICE="cuda",
           MMT=NS(HIDDEN_SIZE=hidden, NUM_HIDDEN_LAYERS=mmt_layers, NUM_ATTENTION_HEADS=heads),
           TEXT_BERT=NS(HIDDEN_SIZE=hidden, NUM_HIDDEN_LAYERS=text_layers, LOAD_PRETRAINED=False, PRETRAINED_NAME=None),
           OBJECT_EMBEDDING=NS(D_FEATURE=d_obj, DROPOUT=0.1), OCR_EMBEDDING=NS(D_FEATURE=d_ocr, DROPOUT=0.1),
           OCR_PTR_NET=NS(HIDDEN_SIZE=hidden, QUERY_KEY_SIZE=hidden if ptr_qk is None else ptr_qk))
    if compute_dtype is not None:
        c.COMPUTE_DTYPE = compute_dtype
    return c
