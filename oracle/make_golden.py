"""Generate ``tests/golden/*.pt`` by running the UNMODIFIED reference on CPU (fp32).

Run in the authoring container only:  ``python -m oracle.make_golden``.
Imports ``/root/reference`` through ``oracle/ref_loader.py``; the fixtures store the reference OUTPUTS plus
the ``(name, shape)`` parameter spec — weights/inputs are regenerated from seeds (``oracle/synth.py``,
``oracle/cases.py``).  TEST INFRASTRUCTURE (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import os
import sys

import torch

from . import cases as C
from . import ref_loader as RL
from . import synth as S

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 0


def _load(module, seed=SEED):
    spec = S.spec_of(module.state_dict())
    params = S.synth_params(spec, seed)
    missing, unexpected = module.load_state_dict(params, strict=False)
    assert not unexpected, unexpected
    assert all(not module.state_dict()[m].dtype.is_floating_point for m in missing), missing
    return spec


def gen_attention():
    import models.modules.attentions as A
    from models.modules.positionwise_feed_forward import PositionWiseFeedForward

    out = {}
    for name, (B, Lq, Lk, D, h, dk, dv, mk, sa) in C.ATT_CASES.items():
        for aoa in ([False, True] if name == "guided_100x30" else [False]):
            cfg = RL.attention_cfg(D, h, dk, dv, use_aoa=aoa, dropout=0.1)
            m = A.MultiHeadAttention(cfg).eval()
            spec = _load(m)
            assert spec == C.att_spec(D, h, dk, dv, aoa=aoa), (name, spec)
            x = C.att_inputs(name)
            with torch.no_grad():
                mask = None if x["mask"] is None else x["mask"].clone()
                core_out, att = m.attention(x["queries"], x["keys"], x["values"], mask)
                mask = None if x["mask"] is None else x["mask"].clone()
                mha_out = m(x["queries"], x["keys"], x["values"], mask)
            out[name + ("+aoa" if aoa else "")] = dict(core_out=core_out, att=att, mha_out=mha_out)
    # PWFF
    for name, (D, dff, B, L) in {"pwff_512": (512, 2048, 2, 30), "pwff_768": (768, 3072, 1, 19)}.items():
        cfg = RL.attention_cfg(D, 8, D // 8, D // 8, d_ff=dff)
        m = PositionWiseFeedForward(cfg).eval()
        spec = _load(m)
        assert spec == C.pwff_spec(D, dff)
        x = torch.randn((B, L, D), generator=S._gen(7, name))
        with torch.no_grad():
            out[name] = dict(out=m(x))
    torch.save(out, os.path.join(OUT, "attention.pt"))
    print("attention.pt", len(out))


def gen_geometry():
    import models.modules.attentions as A
    from models.utils import box_relational_embedding

    out = {}
    for name, (B, N, D, h, trig) in C.GEO_CASES.items():
        cfg = RL.attention_cfg(D, h, D // h, D // h, arch="AugmentedGeometryScaledDotProductAttention", trig=trig)
        core = A.AugmentedGeometryScaledDotProductAttention(cfg).eval()
        x = C.geo_inputs(name)
        spec = [(k[len("attention."):], s) for k, s in C.att_spec(D, h, D // h, D // h, geometry_dg=x["d_g"])
                if k.startswith("attention.")]
        params = S.synth_params([("attention." + k, s) for k, s in spec], SEED)
        core.load_state_dict({k[len("attention."):]: v for k, v in params.items()}, strict=True)
        with torch.no_grad():
            o, mn = core(x["x"], x["x"], x["x"], x["boxes"])      # the only call the reference can execute
            emb = box_relational_embedding(x["boxes"], dim_g=x["d_g"], trignometric_embedding=trig)
        # the (B,N,N,d_g) embedding is large: keep a strided sample for the check
        out[name] = dict(core_out=o, att=mn, emb_sample=emb[:, ::5, ::3].contiguous())
    torch.save(out, os.path.join(OUT, "geometry.pt"))
    print("geometry.pt", len(out))


def gen_masks():
    from models.utils import generate_padding_mask, generate_sequential_mask, generate_self_attention_masks
    from models.modules.pos_embeddings import SinusoidPositionalEmbedding

    feats = S.padded_features(3, 11, 8, 3, all_pad_first=True)
    toks = S.padded_tokens(3, 9, 50, 3)
    pm, tm, sm = generate_padding_mask(feats, 0), generate_padding_mask(toks, 0), generate_sequential_mask(9)
    out = dict(feat_mask=pm, tok_mask=tm, seq_mask=sm, self_mask=generate_self_attention_masks(tm, sm),
               pos=SinusoidPositionalEmbedding(64)(torch.zeros(2, 13, 64)))
    torch.save(out, os.path.join(OUT, "masks.pt"))
    print("masks.pt")


def gen_encoders():
    import models.modules.encoders as E

    D, h, L = 128, 4, 2
    att = lambda: RL.attention_cfg(D, h, D // h, D // h, d_ff=256)
    vis = S.padded_features(2, 36, D, 31, min_valid=10, relu_like=False)
    txt = S.padded_features(2, 14, D, 32, min_valid=4, relu_like=False)
    from models.utils import generate_padding_mask
    vm, tm = generate_padding_mask(vis, 0), generate_padding_mask(txt, 0)
    out = {}
    with torch.no_grad():
        m = E.Encoder(RL.NS(D_MODEL=D, LAYERS=L, SELF_ATTENTION=att())).eval()
        out["Encoder"] = dict(spec=_load(m), out=m(txt, tm))
        m = E.GuidedAttentionEncoder(RL.NS(D_MODEL=D, LAYERS=L, SELF_ATTENTION=att(), GUIDED_ATTENTION=att())).eval()
        out["GuidedAttentionEncoder"] = dict(spec=_load(m), out=m(vis, vm, txt, tm))
        four = dict(VISION_LANGUAGE_ATTENTION=att(), LANGUAGE_VISION_ATTENTION=att(),
                    VISION_SELF_ATTENTION=att(), LANGUAGE_SELF_ATTENTION=att())
        m = E.CoAttentionEncoder(RL.NS(D_MODEL=D, LAYERS=L, **four)).eval()
        v, l = m(vis, vm, txt, tm)
        out["CoAttentionEncoder"] = dict(spec=_load(m), out_v=None, out_l=None)
        v, l = m(vis, vm, txt, tm)
        out["CoAttentionEncoder"].update(out_v=v, out_l=l)
        m = E.CrossModalityEncoder(RL.NS(D_MODEL=D, LAYERS=L, **four)).eval()
        spec = _load(m)
        v, l = m(vis, vm, txt, tm)
        out["CrossModalityEncoder"] = dict(spec=spec, out_v=v, out_l=l)
    torch.save(out, os.path.join(OUT, "encoders.pt"))
    print("encoders.pt")


def gen_mcan():
    import models.mcan as M
    from utils.instance import Instance

    out = {}
    for name, c in (("small", C.MCAN_SMALL), ("full", C.MCAN_FULL)):
        cfg = RL.mcan_cfg(c["d_model"], c["d_feature"], c["layers"], c["h"], c["d_ff"])
        m = M.MCAN(cfg, RL.StubVocab(c["vocab"], c["answers"])).eval()
        spec = _load(m)
        x = C.mcan_inputs(c)
        with torch.no_grad():
            logp = m(Instance(**x))
            logp64 = m.double()(Instance(region_features=x["region_features"].double(),
                                         question_tokens=x["question_tokens"])).float()
        out[name] = dict(spec=spec, logp=logp, logp_fp64=logp64)
        print("mcan", name, "fp32-vs-fp64 max abs", (logp - logp64).abs().max().item())
    torch.save(out, os.path.join(OUT, "mcan.pt"))
    print("mcan.pt")


def gen_m4c():
    RL.load_m4c()
    import models.mmf_m4c as M4

    c = C.M4C_SMALL
    cfg = RL.NS(
        DEVICE="cpu",
        MMT=RL.NS(HIDDEN_SIZE=c["hidden"], NUM_HIDDEN_LAYERS=c["mmt_layers"], NUM_ATTENTION_HEADS=c["heads"]),
        TEXT_BERT=RL.NS(HIDDEN_SIZE=c["hidden"], NUM_HIDDEN_LAYERS=c["text_layers"], LOAD_PRETRAINED=False,
                        PRETRAINED_NAME=None),
        OBJECT_EMBEDDING=RL.NS(D_FEATURE=c["d_obj"], DROPOUT=0.1),
        OCR_EMBEDDING=RL.NS(D_FEATURE=812, DROPOUT=0.1),
        OCR_PTR_NET=RL.NS(HIDDEN_SIZE=c["hidden"], QUERY_KEY_SIZE=c["hidden"]),
    )
    vocab = RL.StubVocab(c["vocab"], max_answer_length=c["T"])
    m = M4.MMF_M4C(cfg, vocab).eval()
    spec = _load(m)
    x = C.m4c_inputs(c)

    class Items(dict):
        __getattr__ = dict.__getitem__

    items = Items(batch_size=c["B"], **x)
    out = {}
    with torch.no_grad():
        # full greedy decode as the reference runs it (with its EOS early exit)
        fwd = {}
        m._forward_txt_encoding(items, fwd)
        m._forward_obj_encoding(items, fwd)
        m._forward_ocr_encoding(items, fwd)
        m._forward_mmt_and_output(items, fwd)
        out["decode"] = dict(scores=fwd["scores"], prev_inds=fwd["prev_inds"])
        # single teacher-forced step with a fixed prev_inds touching vocab AND ocr rows
        prev = torch.tensor([[1, 5, 61, 7, 60, 2], [1, 64, 3, 3, 9, 10], [1, 2, 0, 0, 0, 0]])
        fwd["prev_inds"] = prev.clone()
        m._forward_mmt(items, fwd)
        m._forward_output(items, fwd)
        out["step"] = dict(prev_inds=prev, scores=fwd["scores"], mmt_dec_output=fwd["mmt_dec_output"],
                           mmt_ocr_output=fwd["mmt_ocr_output"], obj_mmt_in=fwd["obj_mmt_in"],
                           ocr_mmt_in=fwd["ocr_mmt_in"], txt_emb=fwd["txt_emb"])
        dec_emb = m.mmt.prev_pred_embeddings(m.classifier.weight, fwd["ocr_mmt_in"], prev)
        ptr = m.ocr_ptr_net(fwd["mmt_dec_output"], fwd["mmt_ocr_output"], fwd["ocr_mask"])
        out["parts"] = dict(dec_emb=dec_emb, ptr_scores=ptr)
    out["spec"] = spec
    torch.save(out, os.path.join(OUT, "m4c.pt"))
    print("m4c.pt decode prev_inds:\n", out["decode"]["prev_inds"])


def gen_m4c_full():
    """The reference MMF_M4C at BASELINE config 3's own shape (cases.M4C_FULL): greedy token sequences (early exit as upstream) and a
    sample of the scores.  78 M parameters regenerated from seeds; only outputs are stored."""
    RL.load_m4c()
    import models.mmf_m4c as M4

    c = C.M4C_FULL
    cfg = RL.NS(
        DEVICE="cpu",
        MMT=RL.NS(HIDDEN_SIZE=c["hidden"], NUM_HIDDEN_LAYERS=c["mmt_layers"], NUM_ATTENTION_HEADS=c["heads"]),
        TEXT_BERT=RL.NS(HIDDEN_SIZE=c["hidden"], NUM_HIDDEN_LAYERS=c["text_layers"], LOAD_PRETRAINED=False, PRETRAINED_NAME=None),
        OBJECT_EMBEDDING=RL.NS(D_FEATURE=c["d_obj"], DROPOUT=0.1), OCR_EMBEDDING=RL.NS(D_FEATURE=812, DROPOUT=0.1),
        OCR_PTR_NET=RL.NS(HIDDEN_SIZE=c["hidden"], QUERY_KEY_SIZE=c["hidden"]),
    )
    m = M4.MMF_M4C(cfg, RL.StubVocab(c["vocab"], max_answer_length=c["T"])).eval()
    spec = _load(m)

    class Items(dict):
        __getattr__ = dict.__getitem__

    items = Items(batch_size=c["B"], **C.m4c_inputs(c))
    with torch.no_grad():
        fwd = {}
        m._forward_txt_encoding(items, fwd)
        m._forward_obj_encoding(items, fwd)
        m._forward_ocr_encoding(items, fwd)
        # first iteration alone (prev_inds = [bos, 0, ...]): the scores a bf16 run is compared against before any argmax feedback
        fwd["prev_inds"] = torch.zeros((c["B"], c["T"]), dtype=torch.long)
        fwd["prev_inds"][:, 0] = 1
        m._forward_mmt(items, fwd)
        m._forward_output(items, fwd)
        first = fwd["scores"].clone()
        m._forward_mmt_and_output(items, fwd)
    top2 = fwd["scores"].topk(2, dim=-1).values
    out = dict(spec=spec, prev_inds=fwd["prev_inds"], first_step_scores=first[:, 0].contiguous(), scores_sample=fwd["scores"][:, :, ::7].contiguous(),
               top2_margin=(top2[..., 0] - top2[..., 1]).contiguous())
    torch.save(out, os.path.join(OUT, "m4c_full.pt"))
    print("m4c_full.pt prev_inds:\n", out["prev_inds"], "\nmin top-2 margin", float(out["top2_margin"].min()))


def gen_decoder():
    """DecoderLayer (decoders.py:13-26) run stateless over the whole sequence and one token at a time inside ``statefulness()``
    with a beam-style re-gather of the states half-way (containers.py:26-31, beam_search.py:19-33)."""
    import models.modules.decoders as Dm
    from models.utils import generate_sequential_mask, generate_self_attention_masks

    c = C.DEC_CASE
    att = lambda st: RL.attention_cfg(c["D"], c["h"], c["D"] // c["h"], c["D"] // c["h"], d_ff=c["d_ff"], stateful=st)
    layer = Dm.DecoderLayer(RL.NS(SELF_ATTENTION=att(True), ENC_ATTENTION=att(False))).eval()
    spec = _load(layer)
    assert spec == C.dec_spec(c["D"], c["h"], c["d_ff"]), spec
    x = C.dec_inputs()
    out = dict(spec=spec)
    with torch.no_grad():
        pad = x["full_mask"].new_zeros(()) + torch.stack([m for m in x["step_masks"]], -1).view(c["B"], 1, 1, c["T"])
        full_mask = generate_self_attention_masks(pad, generate_sequential_mask(c["T"]))
        assert torch.equal(full_mask.float(), x["full_mask"].float())
        out["full"] = layer(x["x"], x["enc"], x["enc"], full_mask, x["enc_mask"])
        steps = []
        perm = torch.tensor(c["perm"])
        xs, enc, enc_mask, masks = x["x"], x["enc"], x["enc_mask"], x["step_masks"]
        with layer.statefulness(c["B"]):
            running_mask = torch.zeros((c["B"], 1, 1, 0))
            for t in range(c["T"]):
                running_mask = torch.cat([running_mask, masks[t]], -1)                       # decoders.py:53-55
                steps.append(layer(xs[:, t:t + 1], enc, enc, running_mask, enc_mask))
                if t == c["regather_at"]:                                                    # the batch rows change places, states follow
                    layer.apply_to_states(lambda s_: s_[perm])
                    running_mask = running_mask[perm]
                    xs, enc, enc_mask = xs[perm], enc[perm], enc_mask[perm]
                    masks = [m[perm] for m in masks]
            assert layer.self_attn.running_keys.shape == (c["B"], c["T"], c["D"])
        out["steps"] = torch.cat(steps, 1)
    torch.save(out, os.path.join(OUT, "decoder.pt"))
    print("decoder.pt: stepwise-vs-full max abs before the re-gather",
          (out["steps"][:, :c["regather_at"] + 1] - out["full"][:, :c["regather_at"] + 1]).abs().max().item())


def gen_beam():
    """BeamSearch.select (beam_search.py:35-38) on seeded candidate tables, incl. one with a finished beam's -999 entries."""
    from models.modules.beam_search import BeamSearch

    out = {}
    for name, (b_s, cur, V, beam) in C.BEAM_CASES.items():
        bs = BeamSearch(model=None, b_s=b_s, max_len=4, eos_idx=2, beam_size=beam, device="cpu")
        idx, logprob = bs.select(C.beam_inputs(name))
        out[name] = dict(idx=idx, logprob=logprob)
    torch.save(out, os.path.join(OUT, "beam.pt"))
    print("beam.pt", {k: tuple(v["idx"].shape) for k, v in out.items()})


def gen_generative():
    """The reference's own beam search (BaseTransformer.beam_search / step, base_transformer.py:30-54; BeamSearch, beam_search.py;
    Decoder, decoders.py:28-76) over a small stateful decoder: token sequences and their log-probabilities."""
    import models.modules.decoders as Dm
    from models.base_transformer import BaseTransformer

    c = C.GEN_CASE
    att = lambda st: RL.attention_cfg(c["D"], c["h"], c["D"] // c["h"], c["D"] // c["h"], d_ff=c["d_ff"], stateful=st)
    cfg = RL.NS(D_MODEL=c["D"], LAYERS=c["layers"], ATTENTION=RL.NS(SELF_ATTENTION=att(True), ENC_ATTENTION=att(False)),
                TEXT_EMBEDDING=RL.NS(ARCHITECTURE="UsualEmbedding", D_MODEL=c["D"], WORD_EMBEDDING=None, D_EMBEDDING=c["D"], DROPOUT=0.1))

    class Captioner(BaseTransformer):
        def __init__(self, config, vocab):
            super().__init__(config, vocab)
            self.device = "cpu"
            self.decoder = Dm.Decoder(config, vocab)

        def encoder_forward(self, feats):
            return feats["enc"], feats["enc_mask"]

    m = Captioner(cfg, RL.StubVocab(c["vocab"], max_answer_length=c["T"])).eval()
    m.eos_idx = c["eos"]                       # (so that finished beams — beam_search.py:50-56 — occur with these seeded weights)
    spec = _load(m)
    x = C.gen_inputs()
    out = dict(spec=spec)
    with torch.no_grad():
        for beam in (1, c["beam"]):
            tokens, logp = m.beam_search(x, batch_size=c["B"], beam_size=beam, out_size=1)
            out[f"beam{beam}"] = dict(tokens=tokens, logp=logp)
        # the decoder as a whole-sequence scorer (training-time call, decoders.py:47-76) on the beam-3 result
        toks = torch.cat([torch.full((c["B"], 1), 1, dtype=torch.long), out[f"beam{c['beam']}"]["tokens"][:, :-1]], 1)
        out["teacher_forced"] = m.decoder(toks, x["enc"], x["enc_mask"])
    torch.save(out, os.path.join(OUT, "generative.pt"))
    print("generative.pt", out["beam1"]["tokens"].tolist(), out[f"beam{c['beam']}"]["tokens"].tolist())


def main():
    os.makedirs(OUT, exist_ok=True)
    RL.load()
    torch.manual_seed(0)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    which = sys.argv[1:] or ["attention", "geometry", "masks", "encoders", "mcan", "m4c", "m4c_full", "decoder", "beam", "generative"]
    for w in which:
        globals()["gen_" + w]()


if __name__ == "__main__":
    main()
