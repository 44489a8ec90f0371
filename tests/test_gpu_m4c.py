"""GPU: the MMF_M4C greedy decode through the B200 blocks against the reference-generated golden decode and the
CPU oracle (mmf_m4c.py:231-256).  fp32 path: identical greedy token sequences, scores <= 1e-4 relative."""
import pytest
import torch

import openvivqa_b200 as ovqa
from openvivqa_b200 import harness as H
from oracle import cases as C
from oracle import restated as R
from oracle import synth as S

pytestmark = pytest.mark.gpu


def _build(c, params, mode, dev):
    m = H.MMF_M4C(H.m4c_cfg(c["hidden"], c["heads"], c["mmt_layers"], c["text_layers"], c["d_obj"], compute_dtype=mode),
                  H.StubVocab(c["vocab"], max_answer_length=c["T"]))
    m.load_state_dict(params, strict=True)
    return m.to(dev).eval()


def _valid_rel(got, ref):
    valid = ref > -1e4
    assert (got[~valid] < -9e4).all()
    return ((got[valid] - ref[valid]).abs().max() / ref[valid].abs().max()).item()


@pytest.mark.parametrize("prune", [True, False])
def test_m4c_decode_matches_reference_golden_fp32(golden, prune):
    dev = torch.device("cuda:0")
    H.strict_fp32_library_math()
    c, g = C.M4C_SMALL, golden("m4c")
    params = S.synth_params(g["spec"], 0)
    m = _build(c, params, "fp32", dev)
    m.prune_last_layer = prune
    out = m(H.Batch(C.m4c_inputs(c)).to(dev))
    assert torch.equal(out["prev_inds"].cpu(), g["decode"]["prev_inds"])              # identical greedy token sequences
    assert _valid_rel(out["scores"].cpu(), g["decode"]["scores"]) <= 1e-4


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_m4c_decode_vs_oracle_wider(mode):
    """d_head 32 (tensor-core attention kernel on the bf16 path), ragged OCR counts incl. an image with none."""
    dev = torch.device("cuda:0")
    H.strict_fp32_library_math()
    c = dict(B=5, Lt=9, Nobj=12, Nocr=10, d_obj=64, hidden=128, heads=4, mmt_layers=2, text_layers=1, vocab=90, T=5)
    spec_model = H.MMF_M4C(H.m4c_cfg(c["hidden"], c["heads"], c["mmt_layers"], c["text_layers"], c["d_obj"]),
                           H.StubVocab(c["vocab"], max_answer_length=c["T"]))
    params = S.synth_params([(k, tuple(v.shape)) for k, v in spec_model.state_dict().items()], 3)
    x = C.m4c_inputs(c, seed=9)
    ref_scores, ref_prev, n = R.m4c_greedy_decode(params, x, max_iter=c["T"], mmt_layers=c["mmt_layers"], text_layers=c["text_layers"],
                                                  heads=c["heads"], ptr_qk=c["hidden"], early_exit=False)
    m = _build(c, params, mode, dev)
    out = m(H.Batch(x).to(dev), early_exit=False)
    assert out["iters"] == c["T"]
    got = out["scores"].cpu()
    if mode == "fp32":
        assert torch.equal(out["prev_inds"].cpu(), ref_prev)
        assert _valid_rel(got, ref_scores) <= 1e-4
    else:
        # bf16: compare the FIRST step's scores (later steps depend on argmax ties the bf16 path may break differently)
        valid = ref_scores > -1e4
        same = torch.equal(out["prev_inds"].cpu(), ref_prev)
        if same:
            assert (got[valid] - ref_scores[valid]).abs().max().item() <= 2e-2 * max(1.0, ref_scores[valid].abs().max().item() / 2)
        m2 = _build(c, params, "bf16", dev)
        st = m2.prepare(H.Batch(x).to(dev))
        m2.step(st, 0)
        s1 = R.m4c_step(params, x, R.m4c_encode_inputs(params, x), torch.tensor([[1] + [0] * (c["T"] - 1)] * c["B"]),
                        mmt_layers=c["mmt_layers"], text_layers=c["text_layers"], heads=c["heads"], ptr_qk=c["hidden"])
        g1 = st.scores[..., :st.W].cpu()
        v1 = s1 > -1e4
        assert (g1[v1] - s1[v1]).abs().max().item() <= 2e-2 * max(1.0, s1[v1].abs().max().item() / 2)


def test_m4c_early_exit_matches_oracle():
    dev = torch.device("cuda:0")
    H.strict_fp32_library_math()
    c, = (C.M4C_SMALL,)
    spec_model = H.MMF_M4C(H.m4c_cfg(c["hidden"], c["heads"], c["mmt_layers"], c["text_layers"], c["d_obj"]),
                           H.StubVocab(c["vocab"], max_answer_length=c["T"]))
    params = S.synth_params([(k, tuple(v.shape)) for k, v in spec_model.state_dict().items()], 0)
    params["classifier.bias"] = params["classifier.bias"].clone()
    params["classifier.bias"][2] += 50.0                                               # force EOS at once -> loop exits after 1 step
    x = C.m4c_inputs(c)
    _, ref_prev, n_ref = R.m4c_greedy_decode(params, x, max_iter=c["T"], mmt_layers=c["mmt_layers"], text_layers=c["text_layers"],
                                             heads=c["heads"], ptr_qk=c["hidden"])
    out = _build(c, params, "fp32", dev)(H.Batch(x).to(dev))
    assert out["iters"] == n_ref == 1
    assert torch.equal(out["prev_inds"].cpu(), ref_prev)
