"""openvivqa_b200 — B200-native (sm_100a) drop-in for OpenViVQA's cross-modal attention hot path.

Public surface:
  * modules:   ScaledDotProductAttention, AugmentedGeometryScaledDotProductAttention, MultiHeadAttention,
               PositionWiseFeedForward, OcrPtrNet, DynamicPointerNetwork, FeatureEmbedding (vision front-end), beam_select, META_ATTENTION,
               build_attention
  * install(): patch a reference checkout's registry / imported symbols so every model and YAML picks them up
  * ops / functional: the C-ABI wrappers and autograd blocks
The CUDA library (openvivqa_b200/libovqa_b200.so) is required; there is no CPU or PyTorch fallback.
"""
from .modules import (AugmentedGeometryScaledDotProductAttention, DynamicPointerNetwork, FeatureEmbedding, META_ATTENTION,  # noqa: F401
                      MultiHeadAttention, OcrPtrNet, PositionWiseFeedForward, ScaledDotProductAttention,
                      StatefulModule, beam_select, build_attention, get_compute_dtype, set_compute_dtype)
from .install import install, uninstall  # noqa: F401

__version__ = "0.1.0"
