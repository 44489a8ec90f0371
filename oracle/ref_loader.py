"""Import the (read-only, untrusted) OpenViVQA reference in THIS container.

Only used to generate golden vectors (``oracle/make_golden.py``) and to
cross-check the restated oracle when ``/root/reference`` is mounted.  Nothing
here runs on the GPU box (``/root/reference`` does not exist there).

Why a loader is needed (SURVEY.md finding 8 / Appendix A): ``import builders``
pulls in ``evaluation/meteor/meteor.py`` which imports a module that is not in
the repository, ``termcolor``/``yacs`` are not installed, and
``transformers.ViTFeatureExtractor`` was removed in transformers 5.  We bypass
the package ``__init__`` files with namespace stubs and stub the missing names.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("OVQA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "models", "modules"))


_loaded = False


def load():
    """Make ``models.*``, ``builders.*``, ``utils.*`` importable from the reference."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    sys.dont_write_bytecode = True  # the mount is read-only
    for name in ("builders", "models", "utils", "tasks"):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, name)]
        sys.modules[name] = m
    mm = types.ModuleType("models.modules")
    mm.__path__ = [os.path.join(REF, "models", "modules")]
    sys.modules["models.modules"] = mm
    if REF not in sys.path:
        sys.path.insert(0, REF)
    tc = types.ModuleType("termcolor")
    tc.colored = lambda s, *a, **k: s
    sys.modules.setdefault("termcolor", tc)
    _stub_vit()
    _loaded = True
    # populate the registries in dependency order (each import fires @META_*.register())
    import models.modules.attentions  # noqa: F401
    import models.modules.encoders  # noqa: F401
    import models.modules.text_embeddings  # noqa: F401  (re-imports transformers lazily)
    _stub_vit()
    import models.modules.vision_embeddings  # noqa: F401


def _stub_vit():
    # ViTFeatureExtractor was removed in transformers 5; vision_embeddings.py:3 imports it by name.
    # transformers swaps its own sys.modules entry for a lazy module, so patch whatever is live.
    import transformers  # noqa: F401

    for mod in {id(m): m for m in (transformers, sys.modules["transformers"])}.values():
        mod.ViTFeatureExtractor = type("ViTFeatureExtractor", (), {})


def load_m4c():
    """Extra shims for ``models/mmf_m4c.py`` on transformers>=5."""
    load()
    from transformers.modeling_utils import PreTrainedModel

    if not getattr(PreTrainedModel, "_ovqa_patched", False):
        orig = PreTrainedModel.init_weights

        def init_weights(self):
            if not hasattr(self, "all_tied_weights_keys") and not getattr(self, "_ovqa_in_post", False):
                self._ovqa_in_post = True
                try:
                    self.post_init()
                finally:
                    self._ovqa_in_post = False
                return
            return orig(self)

        PreTrainedModel.init_weights = init_weights
        PreTrainedModel._ovqa_patched = True


class NS(types.SimpleNamespace):
    """Attribute-style config node (stand-in for yacs.CfgNode; yacs is not installed)."""

    def __contains__(self, k):
        return hasattr(self, k)


def attention_cfg(d_model=512, head=8, d_key=64, d_value=64, d_ff=2048, dropout=0.1,
                  arch="ScaledDotProductAttention", use_aoa=False, stateful=False, trig=True):
    return NS(ARCHITECTURE=arch, HEAD=head, D_MODEL=d_model, D_KEY=d_key, D_VALUE=d_value,
              D_FF=d_ff, D_FEATURE=2048, USE_AOA=use_aoa, CAN_BE_STATEFUL=stateful,
              DROPOUT=dropout, TRIGNOMETRIC_EMBEDDING=trig)


def mcan_cfg(d_model=512, d_feature=2048, layers=6, head=8, d_ff=2048, dropout=0.1, device="cpu"):
    att = lambda: attention_cfg(d_model, head, d_model // head, d_model // head, d_ff, dropout)
    return NS(
        ARCHITECTURE="MCAN", NAME="mcan", DEVICE=device, D_MODEL=d_model,
        VISION_EMBEDDING=NS(ARCHITECTURE="FeatureEmbedding", D_FEATURE=d_feature, D_MODEL=d_model, DROPOUT=dropout),
        TEXT_EMBEDDING=NS(ARCHITECTURE="LSTMTextEmbedding", D_MODEL=d_model, D_EMBEDDING=300, DROPOUT=dropout,
                          WORD_EMBEDDING=None, WORD_EMBEDDING_CACHE=None),
        SELF_ENCODER=NS(ARCHITECTURE="Encoder", D_MODEL=d_model, LAYERS=layers, SELF_ATTENTION=att()),
        GUIDED_ENCODER=NS(ARCHITECTURE="GuidedAttentionEncoder", D_MODEL=d_model, LAYERS=layers,
                          SELF_ATTENTION=att(), GUIDED_ATTENTION=att()),
        VISION_ATTR_REDUCE=NS(D_MODEL=d_model, DROPOUT=dropout),
        TEXT_ATTR_REDUCE=NS(D_MODEL=d_model, DROPOUT=dropout),
    )


class StubVocab:
    """The 6 attributes the hot-path models read from a vocab (data_utils/vocabs/vocab.py:34,55-58)."""

    padding_idx, bos_idx, eos_idx, unk_idx = 0, 1, 2, 3

    def __init__(self, size=4096, total_answers=353, max_answer_length=12):
        self._size = size
        self.total_answers = total_answers
        self.max_answer_length = max_answer_length

    def __len__(self):
        return self._size
