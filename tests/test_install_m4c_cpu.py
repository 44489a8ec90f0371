"""CPU: the reference's OWN ``MMF_M4C`` (models/mmf_m4c.py) after ``install()`` — its constructor, HF initialisation, ``state_dict`` and
control flow, with ``MMT.forward`` / ``TextBert.forward`` / ``PrevPredEmbeddings.forward`` / ``OcrPtrNet.forward`` / ``_batch_gather`` and
the greedy loop rebound to the B200 path — must reproduce the golden vectors the UNPATCHED reference produced (tests/golden/m4c.pt).
The library calls are replaced by their plain-torch restatements (tests/cpu_ops.py): what is checked is everything between them."""
import pytest
import torch

import openvivqa_b200 as ovqa
from oracle import cases as C, ref_loader as RL, synth as S

pytestmark = pytest.mark.skipif(not RL.available(), reason="/root/reference not mounted")


class Items(dict):
    __getattr__ = dict.__getitem__


def _reference_model(c):
    RL.load_m4c()
    import models.mmf_m4c as M4
    cfg = RL.NS(DEVICE="cpu", MMT=RL.NS(HIDDEN_SIZE=c["hidden"], NUM_HIDDEN_LAYERS=c["mmt_layers"], NUM_ATTENTION_HEADS=c["heads"]),
                TEXT_BERT=RL.NS(HIDDEN_SIZE=c["hidden"], NUM_HIDDEN_LAYERS=c["text_layers"], LOAD_PRETRAINED=False, PRETRAINED_NAME=None),
                OBJECT_EMBEDDING=RL.NS(D_FEATURE=c["d_obj"], DROPOUT=0.1), OCR_EMBEDDING=RL.NS(D_FEATURE=812, DROPOUT=0.1),
                OCR_PTR_NET=RL.NS(HIDDEN_SIZE=c["hidden"], QUERY_KEY_SIZE=c["hidden"]))
    return M4, M4.MMF_M4C(cfg, RL.StubVocab(c["vocab"], max_answer_length=c["T"]))


@pytest.fixture
def installed(cpu_ops):
    prev = ovqa.get_compute_dtype()
    RL.load_m4c()
    import models.mmf_m4c  # noqa: F401  (install() patches what is imported)
    import models.mmf_sal  # noqa: F401
    ovqa.install(compute_dtype="fp32")
    yield
    ovqa.uninstall()
    ovqa.set_compute_dtype(prev)


def test_reference_m4c_greedy_decode_after_install_reproduces_the_golden(installed, golden):
    c, g = C.M4C_SMALL, golden("m4c")
    M4, m = _reference_model(c)
    assert M4.MMT.forward.__module__.startswith("openvivqa_b200") and M4._batch_gather.__module__.startswith("openvivqa_b200")
    assert S.spec_of(m.state_dict()) == g["spec"]                        # the reference's checkpoint layout is untouched
    m.load_state_dict(S.synth_params(g["spec"], 0), strict=False)
    m.eval()
    items = Items(batch_size=c["B"], **C.m4c_inputs(c))
    with torch.no_grad():
        fwd = {}
        m._forward_txt_encoding(items, fwd)
        m._forward_obj_encoding(items, fwd)
        m._forward_ocr_encoding(items, fwd)
        m._forward_mmt_and_output(items, fwd)                             # the rebound greedy loop
        assert torch.equal(fwd["prev_inds"], g["decode"]["prev_inds"])   # identical greedy token sequences
        valid = g["decode"]["scores"] > -1e4
        torch.testing.assert_close(fwd["scores"][valid], g["decode"]["scores"][valid], rtol=1e-4, atol=1e-4)
        assert (fwd["scores"][~valid] < -9e4).all()
        out = m(items)                                                    # the public entry point returns the same scores
        torch.testing.assert_close(out["scores"][valid], g["decode"]["scores"][valid], rtol=1e-4, atol=1e-4)
        # one teacher-forced step through the reference's own _forward_mmt / _forward_output (rebound MMT, TextBert, OcrPtrNet)
        st = g["step"]
        fwd["prev_inds"] = st["prev_inds"].clone()
        m._forward_mmt(items, fwd)
        m._forward_output(items, fwd)
        torch.testing.assert_close(fwd["mmt_dec_output"], st["mmt_dec_output"], rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(fwd["mmt_ocr_output"], st["mmt_ocr_output"], rtol=1e-4, atol=1e-4)
        v2 = st["scores"] > -1e4
        torch.testing.assert_close(fwd["scores"][v2], st["scores"][v2], rtol=1e-4, atol=1e-4)
        dec_emb = m.mmt.prev_pred_embeddings(m.classifier.weight, fwd["ocr_mmt_in"], st["prev_inds"])
        torch.testing.assert_close(dec_emb, g["parts"]["dec_emb"], rtol=1e-5, atol=1e-5)
        ptr = m.ocr_ptr_net(fwd["mmt_dec_output"], fwd["mmt_ocr_output"], fwd["ocr_mask"])
        pv = g["parts"]["ptr_scores"] > -1e4
        torch.testing.assert_close(ptr[pv], g["parts"]["ptr_scores"][pv], rtol=1e-4, atol=1e-4)
        x = torch.randn(2, 7, 5)
        inds = torch.tensor([[0, 6, 3], [2, 2, 1]])
        assert torch.equal(M4._batch_gather(x, inds), torch.stack([x[0, inds[0]], x[1, inds[1]]]))


def test_reference_m4c_training_step_after_install_has_pointer_gradients(installed):
    """Teacher-forced training through the rebound classes: with dropout on the attention probabilities (the HF default) the layers stay
    the reference's own; the pointer network — now differentiable — and everything upstream of it receive gradients."""
    c = C.M4C_SMALL
    M4, m = _reference_model(c)
    m.train()
    items = Items(batch_size=c["B"], answer_tokens=torch.randint(0, c["vocab"] + 5, (c["B"], c["T"])), **C.m4c_inputs(c))
    scores = m(items)["scores"]
    assert scores.shape == (c["B"], c["T"], c["vocab"] + 5)
    valid = scores > -1e4
    scores[valid].square().mean().backward()
    for name in ("ocr_ptr_net.query.weight", "ocr_ptr_net.key.weight", "mmt.encoder.layer.0.attention.self.query.weight",
                 "mmt.prev_pred_embeddings.ans_layer_norm.weight", "classifier.weight"):
        grad = dict(m.named_parameters())[name].grad
        assert grad is not None and float(grad.abs().max()) > 0.0, name


def test_reference_m4c_training_without_probability_dropout_runs_the_fused_layers(installed):
    """attention_probs_dropout_prob = 0: the fused blocks cover the training arithmetic; gradients must match the unpatched model."""
    c = C.M4C_SMALL
    M4, m = _reference_model(c)
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train()
    items = Items(batch_size=c["B"], answer_tokens=torch.randint(0, c["vocab"] + 5, (c["B"], c["T"]), generator=torch.Generator().manual_seed(0)),
                  **C.m4c_inputs(c))
    s_new = m(items)["scores"]
    valid = s_new > -1e4
    s_new[valid].square().mean().backward()
    g_new = {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}
    ovqa.uninstall()
    try:
        m.zero_grad()
        s_ref = m(items)["scores"]
        s_ref[valid].square().mean().backward()
    finally:
        ovqa.install(compute_dtype="fp32")
    torch.testing.assert_close(s_new[valid], s_ref[valid], rtol=1e-4, atol=1e-4)
    for k, v in m.named_parameters():
        if v.grad is None:
            continue
        scale = max(1e-6, float(v.grad.abs().max()))
        assert k in g_new, k
        assert float((g_new[k] - v.grad).abs().max()) <= 2e-3 * scale + 1e-7, k


def test_install_covers_the_sal_copies_and_uninstall_restores(cpu_ops):
    RL.load_m4c()
    import models.mmf_m4c as M4
    import models.mmf_sal as SAL
    orig = (M4.MMT.forward, M4.OcrPtrNet.forward, SAL.MMT.forward, SAL._batch_gather, M4.MMF_M4C._forward_mmt_and_output)
    ovqa.install()
    try:
        for fn in (M4.MMT.forward, M4.TextBert.forward, M4.PrevPredEmbeddings.forward, M4.OcrPtrNet.forward, M4._batch_gather,
                   M4.MMF_M4C._forward_mmt_and_output, SAL.MMT.forward, SAL.PrevPredEmbeddings.forward, SAL.OcrPtrNet.forward,
                   SAL._batch_gather, SAL.MMF_SAL._forward_mmt_and_output):
            assert fn.__module__.startswith("openvivqa_b200"), fn
    finally:
        ovqa.uninstall()
    assert orig == (M4.MMT.forward, M4.OcrPtrNet.forward, SAL.MMT.forward, SAL._batch_gather, M4.MMF_M4C._forward_mmt_and_output)


def test_reference_mcan_after_install_matches_the_stock_model(cpu_ops):
    """The reference's own MCAN (models/mcan.py) built after install(): drop-in attention / feed-forward / vision embedding and the rebound
    forward whose classifier tail runs the library's pooling head — same log-probabilities as the stock model, reference state_dict loads."""
    from openvivqa_b200 import harness as H
    RL.load()
    import models.mcan as MC
    from builders.model_builder import META_ARCHITECTURE
    c = C.MCAN_SMALL
    vocab = RL.StubVocab(c["vocab"], c["answers"])
    cfg = RL.mcan_cfg(c["d_model"], c["d_feature"], c["layers"], c["h"], c["d_ff"], device="cpu")
    torch.manual_seed(0)
    stock = META_ARCHITECTURE.get("MCAN")(cfg, vocab).eval()
    x = H.Batch(C.mcan_inputs(c))
    with torch.no_grad():
        ref = stock(x)
    prev = ovqa.get_compute_dtype()
    try:
        ovqa.install(compute_dtype="fp32")
        assert MC.MCAN.forward.__module__.startswith("openvivqa_b200")
        ours = META_ARCHITECTURE.get("MCAN")(cfg, vocab)
        ours.load_state_dict(stock.state_dict(), strict=True)
        ours.eval()
        with torch.no_grad():
            got = ours(x)
    finally:
        ovqa.uninstall()
        ovqa.set_compute_dtype(prev)
    assert not MC.MCAN.forward.__module__.startswith("openvivqa_b200")
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)
