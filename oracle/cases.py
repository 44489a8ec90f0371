"""Parity cases: one table that drives the golden generator, the oracle tests and the GPU parity tests.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  A case is ``(kind, cfg)``; inputs and weights are
regenerated deterministically from seeds (``oracle/synth.py``), only reference OUTPUTS are stored in
``tests/golden/``.  Shape classes follow SURVEY.md §8a: MCAN (Lq,Lk) ∈ {(30,30),(100,100),(100,30)},
microbench (30,100), LoRRA h=1 with h·d_k ≠ D, masks (B,1,1,Lk) / (B,1,L,L) / (1,1,L,L) / None,
a sample whose keys are ALL masked, ragged lengths not multiples of 8/16.
"""
from __future__ import annotations

import torch

from . import restated as R
from . import synth as S

# ----------------------------------------------------------------------------- attention-level cases
# name: (B, Lq, Lk, D, h, dk, dv, mask_kind, self_attention)
ATT_CASES = {
    "sa_text_30":       (2, 30, 30, 512, 8, 64, 64, "pad", True),
    "sa_vision_100":    (2, 100, 100, 512, 8, 64, 64, "pad", True),
    "guided_100x30":    (2, 100, 30, 512, 8, 64, 64, "pad", False),
    "micro_30x100":     (3, 30, 100, 512, 8, 64, 64, "pad_allmasked", False),
    "lorra_h1":         (2, 20, 37, 512, 1, 64, 64, "pad", False),
    "dec_causal_13":    (2, 13, 13, 512, 8, 64, 64, "pad+causal", True),
    "seq_only_17":      (2, 17, 17, 256, 4, 64, 64, "causal", True),
    "nomask_ragged":    (1, 7, 45, 256, 4, 64, 64, "none", False),
    "m4c_d96":          (2, 30, 182, 768, 8, 96, 96, "pad", False),
    "long_600":         (1, 30, 600, 512, 8, 64, 64, "pad", False),
}


def att_spec(D, h, dk, dv, aoa=False, geometry_dg=None):
    spec = []
    if aoa:
        spec += [("informative_attention.weight", (D, 2 * D)), ("informative_attention.bias", (D,)),
                 ("gated_attention.weight", (D, 2 * D)), ("gated_attention.bias", (D,))]
    spec += [("attention.fc_q.weight", (h * dk, D)), ("attention.fc_q.bias", (h * dk,)),
             ("attention.fc_k.weight", (h * dk, D)), ("attention.fc_k.bias", (h * dk,)),
             ("attention.fc_v.weight", (h * dv, D)), ("attention.fc_v.bias", (h * dv,)),
             ("attention.fc_o.weight", (D, h * dv)), ("attention.fc_o.bias", (D,))]
    if geometry_dg is not None:
        for i in range(h):
            spec += [(f"attention.fc_gs.{i}.weight", (1, geometry_dg)), (f"attention.fc_gs.{i}.bias", (1,))]
    spec += [("layer_norm.weight", (D,)), ("layer_norm.bias", (D,))]
    return spec


def pwff_spec(D, d_ff):
    return [("fc1.weight", (d_ff, D)), ("fc1.bias", (d_ff,)), ("fc2.weight", (D, d_ff)), ("fc2.bias", (D,)),
            ("layer_norm.weight", (D,)), ("layer_norm.bias", (D,))]


def att_inputs(name: str, seed: int = 11):
    B, Lq, Lk, D, h, dk, dv, mkind, self_att = ATT_CASES[name]
    g = S._gen(seed, name)
    keys = S.padded_features(B, Lk, D, seed + 1, min_valid=max(1, Lk // 3), relu_like=False,
                             all_pad_first=(mkind == "pad_allmasked"))
    if self_att:
        queries = keys
    else:
        queries = torch.randn((B, Lq, D), generator=g)
    mask = None
    if mkind in ("pad", "pad_allmasked"):
        mask = R.padding_mask(keys, 0)
    elif mkind == "causal":
        mask = R.sequential_mask(Lk)
    elif mkind == "pad+causal":
        mask = R.self_attention_masks(R.padding_mask(keys, 0), R.sequential_mask(Lk))
    return dict(queries=queries, keys=keys, values=keys, mask=mask, h=h, d_k=dk, d_v=dv, D=D)


# ----------------------------------------------------------------------------- geometry cases  (B, N, D, h, trig)
GEO_CASES = {
    "geo_36": (2, 36, 512, 8, True),
    "geo_100": (1, 100, 512, 8, True),
    "geo_raw4": (2, 19, 256, 4, False),
}


def geo_inputs(name: str, seed: int = 23):
    B, N, D, h, trig = GEO_CASES[name]
    g = S._gen(seed, name)
    x = torch.randn((B, N, D), generator=g)
    return dict(x=x, boxes=S.boxes(B, N, seed), h=h, d_k=D // h, d_v=D // h, D=D,
                d_g=(D // h if trig else 4), trig=trig)


# ----------------------------------------------------------------------------- model-level cases
MCAN_SMALL = dict(B=3, Lt=30, Nv=100, d_feature=256, d_model=128, h=4, layers=2, d_ff=256, vocab=200, answers=53)
MCAN_FULL = dict(B=2, Lt=30, Nv=100, d_feature=2048, d_model=512, h=8, layers=6, d_ff=2048, vocab=4096, answers=353)

M4C_SMALL = dict(B=3, Lt=7, Nobj=9, Nocr=5, d_obj=32, hidden=96, heads=4, mmt_layers=2, text_layers=2,
                 vocab=60, T=6)


# BASELINE config 3 at its own shape (what bench.py's m4c_decode workload runs): 20 + 100 + 50 + 12 = 182 positions, hidden 768, 8 heads of
# 96, 4 MMT + 3 TextBert layers, vocabulary 3000 -> 3050 score columns, 12 greedy steps.  Batch 4 keeps the CPU reference run short.
M4C_FULL = dict(B=4, Lt=20, Nobj=100, Nocr=50, d_obj=1024, hidden=768, heads=8, mmt_layers=4, text_layers=3, vocab=3000, T=12)


def mcan_inputs(c, seed: int = 1):
    return dict(region_features=S.padded_features(c["B"], c["Nv"], c["d_feature"], seed, min_valid=36),
                question_tokens=S.padded_tokens(c["B"], c["Lt"], c["vocab"], seed, min_len=5))


def m4c_inputs(c, seed: int = 5):
    B = c["B"]
    ocr_n = dict(min_valid=0)
    f = S.padded_features
    # one shared validity pattern for the three OCR feature streams (zero rows == padding, mmf_m4c.py:196-199)
    det = f(B, c["Nocr"], 256, seed + 3, relu_like=False, all_pad_first=True, **ocr_n)
    valid = (det.abs().sum(-1, keepdim=True) > 0).float()
    g = S._gen(seed, "m4c")
    return dict(
        question_tokens=S.padded_tokens(B, c["Lt"], c["vocab"], seed, min_len=2),
        region_features=f(B, c["Nobj"], c["d_obj"], seed + 1, min_valid=3),
        region_boxes=S.boxes(B, c["Nobj"], seed + 2, scale=1.0),
        ocr_det_features=det,
        ocr_rec_features=torch.randn((B, c["Nocr"], 256), generator=g) * valid,
        ocr_fasttext_features=torch.randn((B, c["Nocr"], 300), generator=g) * valid,
        ocr_boxes=S.boxes(B, c["Nocr"], seed + 4, scale=1.0),
    )


# ----------------------------------------------------------------------------- decoder layer, stateless and stepwise (decoders.py:13-26)
DEC_CASE = dict(B=3, T=7, N=11, D=128, h=4, d_ff=256, regather_at=3, perm=(2, 0, 1))


def dec_spec(D: int, h: int, d_ff: int):
    """state_dict layout of the reference DecoderLayer (running_keys / running_values buffers of the stateful self-attention incl.)."""
    spec = []
    for blk, stateful in (("self_attn", True), ("enc_attn", False)):
        if stateful:                                                     # a module's own buffers precede its children's entries
            spec += [(f"{blk}.running_keys", (0, D)), (f"{blk}.running_values", (0, D))]
        spec += [(f"{blk}.{k}", sh) for k, sh in att_spec(D, h, D // h, D // h)]
    return spec + [(f"pwff.{k}", sh) for k, sh in pwff_spec(D, d_ff)]


def dec_inputs(seed: int = 41):
    c = DEC_CASE
    g = S._gen(seed, "decoder")
    x = torch.randn((c["B"], c["T"], c["D"]), generator=g)
    tok_pad = torch.zeros((c["B"], c["T"]), dtype=torch.bool)
    tok_pad[1, 4:] = True                                               # sample 1 ends after 4 tokens: later positions are padding
    enc = S.padded_features(c["B"], c["N"], c["D"], seed + 1, min_valid=4, relu_like=False)
    enc_mask = R.padding_mask(enc, 0)
    pad_mask = (tok_pad.long() * -10e4).float().view(c["B"], 1, 1, c["T"])                 # models/utils.py:44-60 form
    full_mask = R.self_attention_masks(pad_mask, R.sequential_mask(c["T"]))
    step_masks = [pad_mask[..., t:t + 1] for t in range(c["T"])]
    return dict(x=x, enc=enc, enc_mask=enc_mask, full_mask=full_mask, step_masks=step_masks)


def dec_stepwise_inputs(x):
    """Per-step (x_t, mask_t, enc, enc_mask) of the golden stepwise run: after step ``regather_at`` the batch rows change places
    (make_golden.gen_decoder), as they do when beam search re-gathers the surviving beams."""
    c = DEC_CASE
    perm, cur, per_step = torch.tensor(c["perm"]), torch.arange(c["B"]), []
    for t in range(c["T"]):
        per_step.append((x["x"][cur, t:t + 1], x["step_masks"][t][cur], x["enc"][cur], x["enc_mask"][cur]))
        if t == c["regather_at"]:
            cur = cur[perm]
    return per_step, perm


# ----------------------------------------------------------------------------- beam selection (beam_search.py:35-38)
BEAM_CASES = {"first_step": (3, 1, 37, 4), "later_step": (3, 4, 37, 4), "wide": (2, 5, 3000, 5)}     # (b_s, current beams, vocab, beam_size)


def beam_inputs(name: str, seed: int = 53):
    b_s, cur, V, beam = BEAM_CASES[name]
    x = torch.log_softmax(torch.randn((b_s, cur, V), generator=S._gen(seed, name)) * 3.0, -1)
    x = x + torch.randn((b_s, cur, 1), generator=S._gen(seed + 1, name))                  # running sequence log-probabilities
    if cur > 1:
        x[0, 1, 1:] = -999.0                                                              # a finished beam: beam_search.py:54-56
    return x


# ----------------------------------------------------------------------------- beam search over a stateful decoder (base_transformer.py:30-54)
GEN_CASE = dict(B=3, N=7, D=64, h=4, d_ff=128, layers=2, vocab=23, T=6, beam=3, eos=10)    # eos: a token these weights emit early


def gen_inputs(seed: int = 61):
    c = GEN_CASE
    enc = S.padded_features(c["B"], c["N"], c["D"], seed, min_valid=3, relu_like=False)
    return dict(enc=enc, enc_mask=R.padding_mask(enc, 0))
