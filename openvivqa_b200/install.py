"""Install the drop-ins into an importable OpenViVQA checkout.

``builders.registry.Registry._do_register`` asserts on duplicate names (registry.py:43-47), so the registry's
``_obj_map`` is overwritten directly; ``MultiHeadAttention`` / ``PositionWiseFeedForward`` are plain classes that
``encoders.py:4-5``, ``decoders.py:5,7`` (and a few models) import by name, so those module attributes are rebound.
Call after ``import builders`` (or after importing ``models.modules.*``) and before ``build_model``.
"""
from __future__ import annotations

import importlib
import sys

from . import m4c as M4
from . import modules as M

_saved = {}

# reference modules that define the M4C classes (mmf_m4c.py:259-459) or verbatim-equivalent copies of them (mmf_sal.py:207-375), with
# the model class whose greedy loop (mmf_m4c.py:231-256 / mmf_sal.py:179-205) is rebound and whether that model owns a TextBert.
# Classes other modules import FROM these (mmf_regional_m4c, mmf_iterative_*, mmf_language_adaptive_m4c: `from .mmf_m4c import OcrPtrNet,
# MMT, PrevPredEmbeddings`) are the same class objects, so patching the methods in place reaches them too.  The variants that use
# binary masks (mmf_improved_decoding_m4c.py, experimental_mmf_m4c.py) are different arithmetic and stay untouched.
_M4C_MODULES = {"models.mmf_m4c": ("MMF_M4C", True), "models.mmf_sal": ("MMF_SAL", False)}

_REBIND = {
    "MultiHeadAttention": ("models.modules.attentions", "models.modules.encoders", "models.modules.decoders",
                           "models.modules.SCP"),
    "PositionWiseFeedForward": ("models.modules.positionwise_feed_forward", "models.modules.encoders",
                                "models.modules.decoders", "models.iterative_mcan"),
}


def _maybe(modname):
    if modname in sys.modules:
        return sys.modules[modname]
    return None


def install(compute_dtype: str | None = None, verbose: bool = False):
    """Patch the already-imported reference modules.  Returns the dict of classes that were installed."""
    if compute_dtype is not None:
        M.set_compute_dtype(compute_dtype)
    att_builder = _maybe("builders.attention_builder")
    if att_builder is None:
        att_builder = importlib.import_module("builders.attention_builder")
    reg = att_builder.META_ATTENTION
    for name in ("ScaledDotProductAttention", "AugmentedGeometryScaledDotProductAttention"):
        _saved.setdefault(("registry", name), reg._obj_map.get(name))
        reg._obj_map[name] = getattr(M, name)
    # derive MHA from the reference's stateful base so containers.py's isinstance recursion still sees it
    containers = _maybe("models.modules.containers")
    base = containers.Module if containers is not None else M.StatefulModule
    mha = M.make_multi_head_attention(base=base, builder=att_builder.build_attention)
    classes = {"MultiHeadAttention": mha, "PositionWiseFeedForward": M.PositionWiseFeedForward,
               "ScaledDotProductAttention": M.ScaledDotProductAttention,
               "AugmentedGeometryScaledDotProductAttention": M.AugmentedGeometryScaledDotProductAttention}
    # vision front-end: registered in META_VISION_EMBEDDING (vision_embeddings.py:10)
    vis_builder = _maybe("builders.vision_embedding_builder")
    if vis_builder is not None:
        vreg = vis_builder.META_VISION_EMBEDDING
        _saved.setdefault(("vision_registry", "FeatureEmbedding"), vreg._obj_map.get("FeatureEmbedding"))
        vreg._obj_map["FeatureEmbedding"] = M.FeatureEmbedding
        classes["FeatureEmbedding"] = M.FeatureEmbedding
    # beam selection: BeamSearch.select's full sort -> row-wise top-k (beam_search.py:35-38)
    beam_mod = _maybe("models.modules.beam_search")
    if beam_mod is not None and hasattr(beam_mod, "BeamSearch"):
        _saved.setdefault(("beam_select", "select"), beam_mod.BeamSearch.select)
        beam_mod.BeamSearch.select = M._beam_search_select
        classes["BeamSearch.select"] = M._beam_search_select
        # the re-gather closure records its input -> output tensors, so encoder attention can keep its projected keys across beams
        _saved.setdefault(("beam_expand", "_expand_state"), beam_mod.BeamSearch._expand_state)
        beam_mod.BeamSearch._expand_state = M._beam_search_expand_state(_saved[("beam_expand", "_expand_state")])
    # M4C family: methods of the reference's own classes are rebound in place (state_dict, constructors and HF initialisation stay theirs)
    for modname, (model_name, with_text_bert) in _M4C_MODULES.items():
        mod = _maybe(modname)
        if mod is None:
            continue
        bindings = M4.make_reference_bindings(mod, with_text_bert)
        if hasattr(mod, model_name):
            bindings[(model_name, "_forward_mmt_and_output")] = M4.make_model_binding(getattr(mod, model_name), with_text_bert)
        for (owner, attr), fn in bindings.items():
            target = mod if owner is None else getattr(mod, owner)
            _saved.setdefault(("m4c", modname, owner, attr), getattr(target, attr))
            setattr(target, attr, fn)
            classes[f"{modname}.{owner + '.' if owner else ''}{attr}"] = fn
            if verbose:
                print(f"openvivqa_b200: {modname}.{owner + '.' if owner else ''}{attr} -> B200 path")
    # classifier tail of MCAN (models/mcan.py:70-80): the forward is rebound so that the attribute-reduce pooling runs in the library
    mcan_mod = _maybe("models.mcan")
    if mcan_mod is not None and hasattr(mcan_mod, "MCAN"):
        _saved.setdefault(("m4c", "models.mcan", "MCAN", "forward"), mcan_mod.MCAN.forward)
        mcan_mod.MCAN.forward = M.mcan_forward
        classes["models.mcan.MCAN.forward"] = M.mcan_forward
    # LSTM text embedding (text_embeddings.py:222-246): the recurrence runs in the library's cluster-resident kernels
    txt_mod = _maybe("models.modules.text_embeddings")
    if txt_mod is not None and hasattr(txt_mod, "LSTMTextEmbedding"):
        _saved.setdefault(("m4c", "models.modules.text_embeddings", "LSTMTextEmbedding", "forward"), txt_mod.LSTMTextEmbedding.forward)
        txt_mod.LSTMTextEmbedding.forward = M.make_lstm_text_embedding_forward(txt_mod)
        classes["models.modules.text_embeddings.LSTMTextEmbedding.forward"] = txt_mod.LSTMTextEmbedding.forward
    for name, mods in _REBIND.items():
        for modname in mods:
            mod = _maybe(modname)
            if mod is not None and hasattr(mod, name):
                _saved.setdefault((modname, name), getattr(mod, name))
                setattr(mod, name, classes[name])
                if verbose:
                    print(f"openvivqa_b200: {modname}.{name} -> B200 drop-in")
    return classes


def uninstall():
    """Undo :func:`install` (used by the tests that compare with/without the drop-ins in one process)."""
    for key, obj in list(_saved.items()):
        if key[0] == "m4c":
            _, modname, owner, attr = key
            mod = sys.modules[modname]
            setattr(mod if owner is None else getattr(mod, owner), attr, obj)
            continue
        where, name = key
        if where == "vision_registry":
            vreg = sys.modules["builders.vision_embedding_builder"].META_VISION_EMBEDDING
            if obj is None:
                vreg._obj_map.pop(name, None)
            else:
                vreg._obj_map[name] = obj
        elif where == "beam_select":
            sys.modules["models.modules.beam_search"].BeamSearch.select = obj
        elif where == "beam_expand":
            sys.modules["models.modules.beam_search"].BeamSearch._expand_state = obj
        elif where == "registry":
            reg = sys.modules["builders.attention_builder"].META_ATTENTION
            if obj is None:
                reg._obj_map.pop(name, None)
            else:
                reg._obj_map[name] = obj
        else:
            setattr(sys.modules[where], name, obj)
    _saved.clear()
