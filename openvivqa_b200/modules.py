"""Drop-in modules: same class names, ``__init__(config)`` keys, parameter names/shapes/initialisers and
``forward`` signatures as the reference, computing through libovqa_b200.

Reference classes mirrored (``/root/reference``):
  * ``ScaledDotProductAttention``                      models/modules/attentions.py:10-60
  * ``AugmentedGeometryScaledDotProductAttention``     models/modules/attentions.py:62-137
  * ``MultiHeadAttention``                             models/modules/attentions.py:293-339
  * ``PositionWiseFeedForward``                        models/modules/positionwise_feed_forward.py:5-27
  * ``OcrPtrNet`` / ``DynamicPointerNetwork``          models/mmf_m4c.py:367-396 / models/m4c.py:19-33

Compute dtype: ``"fp32"`` (parity path, default) or ``"bf16"`` (tcgen05 path); chosen by, in order, an optional
config key ``COMPUTE_DTYPE``, :func:`set_compute_dtype`, env ``OVQA_COMPUTE_DTYPE``.  No new required config keys.
"""
from __future__ import annotations

import copy
import math
import os
from contextlib import contextmanager
from typing import Optional

import torch
from torch import nn

from . import functional as Fn
from . import ops

_default_mode = os.environ.get("OVQA_COMPUTE_DTYPE", "fp32")


def set_compute_dtype(mode: str):
    """Process-wide default compute dtype for modules built without a ``COMPUTE_DTYPE`` key: "fp32" or "bf16"."""
    global _default_mode
    if mode not in ("fp32", "bf16"):
        raise ValueError("compute dtype must be 'fp32' or 'bf16'")
    _default_mode = mode


def get_compute_dtype() -> str:
    return _default_mode


MASK_VALUE = -10e4          # models/utils.py:56
DECODE_KV_CACHE = True      # inference-time stateful decode keeps projected keys / values (False: the reference's bookkeeping, for A/B timing)


def _mode_of(module) -> str:
    return module.compute_dtype or _default_mode


DECODE_MAX_KEYS = 200 * 1024 // (4 * 4)   # sdpa_decode_kernel keeps one fp32 score row per warp in shared memory (csrc/decode.cu: 4 warps, 200 KB)
DECODE_MAX_QUERY_ROWS = 16   # the one-warp-per-(row, head) decode kernel is for a handful of query rows; longer blocks use the tiled cores


def tracked_regather(fn):
    """Wrap a state re-gather function (``BeamSearch._expand_state``'s closure, beam_search.py:19-33) so that it remembers which
    output tensor it produced from which input tensor during one ``apply_to_states`` sweep.  The sweep transforms the model's own
    states (e.g. ``encoder_features``, base_transformer.py:17) and, with the SAME function, the projected-K/V caches of the attention
    blocks below it; the record lets a block prove at its next call that the keys it is handed ARE ``fn(previous keys)`` — so its
    re-gathered projection is still theirs — by object identity, without comparing contents."""
    if hasattr(fn, "_ovqa_memo"):
        return fn
    memo = {}

    def tracked(s):
        out = fn(s)
        if isinstance(s, torch.Tensor) and isinstance(out, torch.Tensor):
            memo[id(s)] = (s, out)                     # (the input is held so that its id stays unique while the record lives)
        return out

    tracked._ovqa_memo = memo
    return tracked


class StatefulModule(nn.Module):
    """Minimal stand-in for the reference's ``containers.Module`` (models/modules/containers.py:4-70): buffers
    registered as decode "states" that ``statefulness(batch_size)`` expands per batch and resets afterwards.
    When installed into a reference checkout the drop-ins derive from the reference's own class instead, so its
    ``isinstance(m, Module)`` recursion (containers.py:22-24) keeps working."""

    def __init__(self):
        super().__init__()
        self._is_stateful = False
        self._state_names = []
        self._state_defaults = {}

    def register_state(self, name, default):
        self._state_names.append(name)
        self._state_defaults[name] = None if default is None else default.clone().detach()
        self.register_buffer(name, default)

    def _children_stateful(self):
        return [m for m in self.children() if isinstance(m, StatefulModule)]

    def states(self):
        for n in self._state_names:
            yield self._buffers[n]
        for m in self._children_stateful():
            yield from m.states()

    def apply_to_states(self, fn):
        fn = tracked_regather(fn)
        for n in self._state_names:
            self._buffers[n] = fn(self._buffers[n])
        for m in self._children_stateful():
            m.apply_to_states(fn)

    def enable_statefulness(self, batch_size):
        for m in self._children_stateful():
            m.enable_statefulness(batch_size)
        for n in self._state_names:
            d = self._state_defaults[n]
            if d is None:
                self._buffers[n] = None
            else:
                t = self._default_on(d, self._buffers[n].device).unsqueeze(0)
                self._buffers[n] = t.expand([batch_size] + list(t.shape[1:])).contiguous()
        self._is_stateful = True

    @staticmethod
    def _default_on(d, device):
        """A state's default on ``device``.  All-zero defaults (every one the path registers) are made by a fill on the device: no
        host-to-device copy, so a decode session may be opened inside a CUDA-graph capture."""
        if d.device == device:
            return d.clone().detach()
        if d.numel() == 0 or not bool(d.any()):
            return torch.zeros(d.shape, dtype=d.dtype, device=device)
        return d.clone().detach().to(device)

    def disable_statefulness(self):
        for m in self._children_stateful():
            m.disable_statefulness()
        for n in self._state_names:
            d = self._state_defaults[n]
            self._buffers[n] = None if d is None else self._default_on(d, self._buffers[n].device)
        self._is_stateful = False

    @contextmanager
    def statefulness(self, batch_size):
        self.enable_statefulness(batch_size)
        try:
            yield
        finally:
            self.disable_statefulness()


def _init_core_linears(m):
    # attentions.py:36-44: Xavier-uniform q,k,v,o in that order, zero biases
    for fc in (m.fc_q, m.fc_k, m.fc_v, m.fc_o):
        nn.init.xavier_uniform_(fc.weight)
    for fc in (m.fc_q, m.fc_k, m.fc_v, m.fc_o):
        nn.init.constant_(fc.bias, 0)


class ScaledDotProductAttention(nn.Module):
    """attentions.py:10-60.  ``forward(queries, keys, values, attention_mask=None, **kwargs) -> (out, att)``;
    ``att`` (B,h,Lq,Lk) is materialised here because direct callers consume it (mmf_lorra.py:118-146); the
    ``MultiHeadAttention`` wrapper uses the fused block instead and never materialises it."""

    def __init__(self, config):
        super().__init__()
        d_model, h, d_k, d_v = config.D_MODEL, config.HEAD, config.D_KEY, config.D_VALUE
        self.fc_q = nn.Linear(d_model, h * d_k)
        self.fc_k = nn.Linear(d_model, h * d_k)
        self.fc_v = nn.Linear(d_model, h * d_v)
        self.fc_o = nn.Linear(h * d_v, d_model)
        self.d_model, self.d_k, self.d_v, self.h = d_model, d_k, d_v, h
        self.compute_dtype = getattr(config, "COMPUTE_DTYPE", None)
        self.init_weights()

    def init_weights(self):
        _init_core_linears(self)

    def _geo(self, boxes):
        return {}

    def forward(self, queries, keys, values, attention_mask=None, **kwargs):
        mode = _mode_of(self)
        odt = torch.float32 if mode == "fp32" else torch.bfloat16
        q = Fn.linear(queries, self.fc_q.weight, self.fc_q.bias, mode, odt)
        k = Fn.linear(keys, self.fc_k.weight, self.fc_k.bias, mode, odt)
        v = Fn.linear(values, self.fc_v.weight, self.fc_v.bias, mode, odt)
        # `att` is differentiable: direct callers build features from the weights alone (mmf_lorra.py:127-146)
        ctx, att = Fn.sdpa(q, k, v, attention_mask, h=self.h, d_k=self.d_k, d_v=self.d_v, need_att=True, **self._geo(kwargs.get("boxes")))
        out = Fn.linear(ctx, self.fc_o.weight, self.fc_o.bias, mode)
        return out, att


class AugmentedGeometryScaledDotProductAttention(ScaledDotProductAttention):
    """attentions.py:62-137: the core with ``log(clamp(relu(fc_g_h(emb(boxes))), 1e-6))`` added to the scores.

    Documented deviations (SURVEY.md finding 5): ``attention_mask`` is honoured (the reference crashes on it,
    attentions.py:128-129) and ``boxes`` may also arrive as a keyword (so it works through ``MultiHeadAttention``)."""

    def __init__(self, config):
        nn.Module.__init__(self)
        d_model, h, d_k, d_v = config.D_MODEL, config.HEAD, config.D_KEY, config.D_VALUE
        self.trignometric_embedding = config.TRIGNOMETRIC_EMBEDDING
        self.d_g = d_model // h if self.trignometric_embedding else 4
        self.fc_q = nn.Linear(d_model, h * d_k)
        self.fc_k = nn.Linear(d_model, h * d_k)
        self.fc_v = nn.Linear(d_model, h * d_v)
        self.fc_o = nn.Linear(h * d_v, d_model)
        fc_g = nn.Linear(self.d_g, 1)
        self.fc_gs = nn.ModuleList([copy.deepcopy(fc_g) for _ in range(h)])   # models/utils.py:40-42 clones()
        self.d_model, self.d_k, self.d_v, self.h = d_model, d_k, d_v, h
        self.compute_dtype = getattr(config, "COMPUTE_DTYPE", None)
        self.init_weights()

    def init_weights(self):
        for fc in (self.fc_q, self.fc_k, self.fc_v, self.fc_o):
            nn.init.xavier_uniform_(fc.weight)
        for g in self.fc_gs:
            nn.init.xavier_uniform_(g.weight)
        for fc in (self.fc_q, self.fc_k, self.fc_v, self.fc_o):
            nn.init.constant_(fc.bias, 0)
        for g in self.fc_gs:
            nn.init.constant_(g.bias, 0)

    def geo_weights(self):
        w = torch.cat([g.weight for g in self.fc_gs], dim=0)        # (h, d_g)
        b = torch.cat([g.bias for g in self.fc_gs], dim=0)          # (h,)
        return w, b

    def _geo(self, boxes):
        if boxes is None:
            raise ValueError("AugmentedGeometryScaledDotProductAttention needs `boxes` (B,N,4)")
        w, b = self.geo_weights()                       # (cat of the h Linear(d_g, 1): autograd routes the gradient back to each)
        return dict(boxes=boxes, geo_w=w, geo_b=b, geo_trig=self.trignometric_embedding)

    def forward(self, queries, keys, values, boxes=None, attention_mask=None, **kwargs):
        return super().forward(queries, keys, values, attention_mask, boxes=boxes)


class AttentionRegistry:
    """Name → class map with the reference's ``build_attention`` contract (builders/attention_builder.py:3-8,
    builders/registry.py:42-48 incl. the duplicate-name assertion)."""

    def __init__(self):
        self._obj_map = {}

    def register(self, obj=None):
        def deco(c):
            assert c.__name__ not in self._obj_map, f"An object named '{c.__name__}' was already registered!"
            self._obj_map[c.__name__] = c
            return c
        return deco if obj is None else deco(obj)

    def get(self, name):
        if name not in self._obj_map:
            raise KeyError(f"No object named '{name}' found in 'META_ATTENTION' registry!")
        return self._obj_map[name]

    def __contains__(self, name):
        return name in self._obj_map


META_ATTENTION = AttentionRegistry()
META_ATTENTION.register(ScaledDotProductAttention)
META_ATTENTION.register(AugmentedGeometryScaledDotProductAttention)


def build_attention(config):
    return META_ATTENTION.get(config.ARCHITECTURE)(config)


def make_multi_head_attention(base=StatefulModule, builder=None):
    """Class factory so that, inside a reference checkout, the drop-in derives from the reference's own
    ``containers.Module`` and builds its core through the reference's registry."""
    builder = builder or build_attention

    class MultiHeadAttention(base):
        """attentions.py:293-339: (stateful K/V concat) → core → dropout → LayerNorm(queries + out) → (AoA gate)."""

        def __init__(self, config):
            super().__init__()
            d_model = config.D_MODEL
            self.use_aoa = config.USE_AOA
            if self.use_aoa:
                self.informative_attention = nn.Linear(2 * d_model, d_model)
                self.gated_attention = nn.Linear(2 * d_model, d_model)
            self.attention = builder(config)
            self.dropout = nn.Dropout(p=config.DROPOUT)
            self.layer_norm = nn.LayerNorm(d_model)
            self.can_be_stateful = config.CAN_BE_STATEFUL
            if self.can_be_stateful:
                self.register_state("running_keys", torch.zeros((0, d_model)))
                self.register_state("running_values", torch.zeros((0, d_model)))
                if isinstance(self.attention, ScaledDotProductAttention):
                    # decode-time cache of the PROJECTED keys / values, (B, capacity, h*d_k + h*d_v) with ``_kv_len`` valid positions
                    # (inference fast path below).  Registered as a state so that beam search re-gathers it with the rest
                    # (beam_search.py:19-33), but kept out of the state_dict: the reference's checkpoints do not have it.
                    core = self.attention
                    # (one placeholder row, not zero: the reference's re-gather closure cannot reshape an EMPTY state — `view(-1, …)`
                    # of zero elements is ambiguous, beam_search.py:31)
                    self.register_state("running_kv_proj", torch.zeros((1, core.h * (core.d_k + core.d_v))))
                    self._non_persistent_buffers_set.add("running_kv_proj")
            elif isinstance(self.attention, ScaledDotProductAttention):
                # encoder attention inside a decode session: the encoder keys / values are projected once; the projection is a state too,
                # so a beam re-gather of the model's ``encoder_features`` (base_transformer.py:17) re-gathers it the same way
                core = self.attention
                self.register_state("running_enc_kv_proj", torch.zeros((1, core.h * (core.d_k + core.d_v))))      # (placeholder row, as above)
                self._non_persistent_buffers_set.add("running_enc_kv_proj")
            self._kv_len = 0
            self._enc_kv = None          # (tag, keys, values) of the tensors whose projection ``running_enc_kv_proj`` holds
            self._regather_memo = None   # record of the last ``apply_to_states`` sweep (tracked_regather)
            self._session = None         # per-session decode plan: None = not decided yet, False = reference bookkeeping
            self.compute_dtype = getattr(config, "COMPUTE_DTYPE", None)

        def enable_statefulness(self, batch_size):
            super().enable_statefulness(batch_size)
            self._kv_len, self._enc_kv, self._session, self._regather_memo = 0, None, None, None

        def disable_statefulness(self):
            super().disable_statefulness()
            self._kv_len, self._enc_kv, self._session, self._regather_memo = 0, None, None, None

        def apply_to_states(self, fn):
            super().apply_to_states(fn)                  # (re-gathers running_kv_proj / running_enc_kv_proj with everything else)
            if not self.can_be_stateful:
                self._regather_memo = getattr(fn, "_ovqa_memo", None)

        def _encoder_projection(self, sess, keys, values):
            """[K | V] of the encoder keys / values, projected once per session.  Still valid when the tensors are the very objects
            seen last time (and unmodified), or when they are what the last ``apply_to_states`` sweep made of those objects — the
            sweep re-gathered the projection with the same function.  Anything else: project again."""
            tag = (keys.data_ptr(), keys._version, keys.shape, values.data_ptr(), values._version)
            enc, memo = self._enc_kv, self._regather_memo
            self._regather_memo = None
            if enc is not None and enc[0] == tag:
                return self._buffers["running_enc_kv_proj"]
            if enc is not None and memo is not None:
                rk, rv = memo.get(id(enc[1])), memo.get(id(enc[2]))
                if rk is not None and rv is not None and rk[0] is enc[1] and rv[0] is enc[2] and rk[1] is keys and rv[1] is values \
                        and self._buffers["running_enc_kv_proj"].shape[:2] == keys.shape[:2]:
                    self._enc_kv = (tag, keys, values)
                    cache = self._buffers["running_enc_kv_proj"]
                    if not cache.is_contiguous():
                        self._buffers["running_enc_kv_proj"] = cache = cache.contiguous()
                    return cache
            self._buffers["running_enc_kv_proj"] = cache = Fn.project_kv_fused(keys, values, **sess["kv"]).contiguous()
            self._enc_kv = (tag, keys, values)               # (the tensors are held so that the tag stays unique)
            return cache

        def _open_session(self, core):
            self._session = self._open_decode_session(core)
            return self._session is not False

        def _open_decode_session(self, core):
            """Decide once per ``statefulness()`` session whether the inference fast path applies, and resolve its operands."""
            ok = (DECODE_KV_CACHE and type(core) is ScaledDotProductAttention and (not self.training or self.dropout.p == 0.0)
                  and (not self.can_be_stateful or self.running_keys.shape[1] == 0))
            mode = self.compute_dtype or core.compute_dtype or get_compute_dtype()
            vec = 4 if mode == "fp32" else 8                    # elements per 16-byte vector: the decode kernel's row granularity and width limit
            ok = ok and all(d % vec == 0 and d <= 32 * vec for d in (core.d_k, core.d_v))
            if not ok:
                return False                                    # shapes outside the one-call kernel's range: the tiled blocks (reference bookkeeping)
            kv = dict(wk=core.fc_k.weight, bk=core.fc_k.bias, wv=core.fc_v.weight, bv=core.fc_v.bias, mode=mode)
            plan = Fn.mha_decode_plan(wq=core.fc_q.weight, bq=core.fc_q.bias, wo=core.fc_o.weight, bo=core.fc_o.bias,
                                      gamma=self.layer_norm.weight, beta=self.layer_norm.bias, **kv)
            return dict(plan=plan, kv=kv, mode=mode, h=core.h, d_k=core.d_k, d_v=core.d_v, eps=float(self.layer_norm.eps),
                        stateful=bool(self.can_be_stateful), cdt=torch.float32 if mode == "fp32" else torch.bfloat16)

        def _decode_fast_path(self, sess, queries, keys, values, attention_mask):
            """Inference-time incremental decoding.  The reference re-projects the WHOLE running key / value
            sequence at every step (attentions.py:321-325 then :48-50) and re-projects the unchanged encoder output in every
            cross-attention call; here only the new rows go through fc_k / fc_v and the projected rows are kept — in a buffer with spare
            capacity, so nothing is concatenated either — and the whole block is enqueued by one library call (csrc/decode.cu).  A T-step
            decode does O(T) instead of O(T^2) projection work.  Same outputs: fc_k / fc_v act row by row."""
            kv_src, self_att = None, False
            if (self._kv_len if sess["stateful"] else 0) + keys.shape[1] > DECODE_MAX_KEYS:
                if sess["stateful"] and self._kv_len:
                    raise RuntimeError(f"MultiHeadAttention: the projected-K/V decode cache holds at most {DECODE_MAX_KEYS} positions")
                if sess["stateful"]:
                    self._session = False                        # nothing cached yet: this session keeps the reference's bookkeeping
                return None
            if sess["stateful"]:
                if self._kv_len == 0 and self._buffers["running_keys"].shape[1] == 0:
                    # the raw running keys / values stay unused on this path; give the states one element per row so that the
                    # reference's re-gather closure (beam_search.py:19-33) can still reshape them
                    self._buffers["running_keys"] = self._buffers["running_values"] = queries.new_zeros((queries.shape[0], 1, 1))
                need = self._kv_len + keys.shape[1]
                cache = self._buffers["running_kv_proj"]
                if cache.shape[1] < need or cache.dtype != sess["cdt"]:              # grow geometrically: amortised O(1) copies per step
                    grown = torch.empty((queries.shape[0], max(32, 2 * need), cache.shape[2]), device=queries.device, dtype=sess["cdt"])
                    if self._kv_len:
                        grown[:, :self._kv_len] = cache[:, :self._kv_len]
                    self._buffers["running_kv_proj"] = cache = grown
                elif not cache.is_contiguous():
                    self._buffers["running_kv_proj"] = cache = cache.contiguous()
                self_att = keys is queries and values is queries
                length = self._kv_len
                if not self_att:
                    if values is keys:
                        kv_src = keys
                    else:                                                            # distinct key / value sources: project, then file
                        cache[:, length:need] = Fn.project_kv_fused(keys, values, **sess["kv"])
                        length = need
                self._kv_len = need
            else:
                if queries.shape[1] > DECODE_MAX_QUERY_ROWS:
                    return None                                                      # a long block (an encoder run inside the session): tiled cores
                cache, length = self._encoder_projection(sess, keys, values), keys.shape[1]
            return Fn.mha_decode_block(sess["plan"], queries, kv_src, self_att, cache, length, attention_mask, h=sess["h"], d_k=sess["d_k"],
                                       d_v=sess["d_v"], eps=sess["eps"])

        def forward(self, queries, keys, values, attention_mask, **kwargs):
            core = self.attention
            if not isinstance(core, ScaledDotProductAttention):
                # a core we do not accelerate (memory / adaptive variants): reference composition, their own math — including the
                # running keys / values of a stateful decode, which the reference concatenates for EVERY core (attentions.py:320-325)
                if self.can_be_stateful and self._is_stateful:
                    self.running_keys = torch.cat([self.running_keys, keys], 1)
                    keys = self.running_keys
                    self.running_values = torch.cat([self.running_values, values], 1)
                    values = self.running_values
                out, _ = core(queries, keys, values, attention_mask, **kwargs)
                out = self.layer_norm(queries + self.dropout(out))
            elif self._is_stateful and self._session is not False and not torch.is_grad_enabled() and (
                    self._session is not None or self._open_session(core)):
                out = self._decode_fast_path(self._session, queries, keys, values, attention_mask)
            else:
                out = None
            if out is None and isinstance(core, ScaledDotProductAttention):
                if self.can_be_stateful and self._is_stateful:
                    if self._kv_len != 0:
                        raise RuntimeError("MultiHeadAttention: a stateful decode may not switch between no_grad and grad mode half-way")
                    self.running_keys = torch.cat([self.running_keys, keys], 1)
                    keys = self.running_keys
                    self.running_values = torch.cat([self.running_values, values], 1)
                    values = self.running_values
                geo = isinstance(core, AugmentedGeometryScaledDotProductAttention)
                boxes = kwargs.get("boxes") if geo else None
                if geo and boxes is None:
                    raise ValueError("geometry attention needs `boxes=` (B,N,4)")
                gw, gb = core.geo_weights() if geo else (None, None)
                mode = self.compute_dtype or core.compute_dtype or get_compute_dtype()
                out = Fn.mha_block(queries, keys, values, attention_mask,
                                   wq=core.fc_q.weight, bq=core.fc_q.bias, wk=core.fc_k.weight, bk=core.fc_k.bias,
                                   wv=core.fc_v.weight, bv=core.fc_v.bias, wo=core.fc_o.weight, bo=core.fc_o.bias,
                                   gamma=self.layer_norm.weight, beta=self.layer_norm.bias,
                                   h=core.h, d_k=core.d_k, d_v=core.d_v, mode=mode, dropout=self.dropout.p,
                                   training=self.training, eps=self.layer_norm.eps, boxes=boxes, geo_w=gw, geo_b=gb,
                                   geo_trig=getattr(core, "trignometric_embedding", True))
            if self.use_aoa:                                                     # attentions.py:333-337
                mode = self.compute_dtype or get_compute_dtype()
                aoa_in = torch.cat([queries, out], dim=-1)
                i = Fn.linear(aoa_in, self.informative_attention.weight, self.informative_attention.bias, mode)
                g = torch.sigmoid(Fn.linear(aoa_in, self.gated_attention.weight, self.gated_attention.bias, mode))
                out = i * g
            return out

    return MultiHeadAttention


MultiHeadAttention = make_multi_head_attention()


class PositionWiseFeedForward(nn.Module):
    """positionwise_feed_forward.py:5-27: ``LN(x + drop(fc2(drop(gelu(fc1 x)))))`` (erf GELU, eps 1e-5)."""

    def __init__(self, config):
        super().__init__()
        d_model, d_ff, dropout = config.D_MODEL, config.D_FF, config.DROPOUT
        self.fc1 = nn.Linear(d_model, d_ff)
        self.fc2 = nn.Linear(d_ff, d_model)
        self.dropout_1 = nn.Dropout(p=dropout)
        self.dropout_2 = nn.Dropout(p=dropout)
        self.layer_norm = nn.LayerNorm(d_model)
        self.compute_dtype = getattr(config, "COMPUTE_DTYPE", None)

    def forward(self, input):
        if not torch.is_grad_enabled() and not self.training:                  # inference: the block from one library call
            plan = self.__dict__.get("_plan")
            if plan is None or not plan[0].valid() or plan[0].mode != _mode_of(self):
                plan = (Fn.pwff_plan(w1=self.fc1.weight, b1=self.fc1.bias, w2=self.fc2.weight, b2=self.fc2.bias, gamma=self.layer_norm.weight,
                                     beta=self.layer_norm.bias, mode=_mode_of(self)), self.fc1.weight.shape[0], float(self.layer_norm.eps))
                self.__dict__["_plan"] = plan
            return Fn.pwff_infer_block(plan[0], input, d_ff=plan[1], eps=plan[2])
        return Fn.pwff_block(input, w1=self.fc1.weight, b1=self.fc1.bias, w2=self.fc2.weight, b2=self.fc2.bias,
                             gamma=self.layer_norm.weight, beta=self.layer_norm.bias, mode=_mode_of(self),
                             dropout=self.dropout_1.p, training=self.training, eps=self.layer_norm.eps)


TOPK_MAX = 16       # kTopkMax of csrc/decode.cu


def beam_select(candidate_logprob: torch.Tensor, b_s: int, beam_size: int):
    """``BeamSearch.select`` (beam_search.py:35-38) without the full sort: the ``beam_size`` best of the beam x vocabulary candidates of
    every sample from the row-wise top-k kernel (csrc/decode.cu).  Returns (flat candidate indices (b_s, beam_size) int64, their
    log-probabilities), as the reference does.  Under autograd (self-critical training calls beam search with gradients enabled) the
    values are gathered from the input so they stay differentiable."""
    flat = candidate_logprob.reshape(b_s, -1)
    if not flat.is_cuda or beam_size > TOPK_MAX or b_s > 65535 or beam_size > flat.shape[1]:
        # outside the top-k kernel's range (or a CPU run of the reference): the reference's own selection, beam_search.py:36-38
        logprob, idx = torch.sort(flat, -1, descending=True)
        return idx[:, :beam_size], logprob[:, :beam_size]
    val, idx = ops.topk_rows(flat.detach().float(), beam_size)
    if torch.is_grad_enabled() and flat.requires_grad:
        val = torch.gather(flat, 1, idx)
    return idx, val.to(candidate_logprob.dtype)


def _beam_search_expand_state(orig):
    """``BeamSearch._expand_state`` (beam_search.py:19-33) returning a tracked re-gather function (see ``tracked_regather``)."""
    def _expand_state(self, selected_beam, cur_beam_size):
        return tracked_regather(orig(self, selected_beam, cur_beam_size))
    return _expand_state


def _beam_search_select(self, candidate_logprob):
    """Bound over the reference's ``BeamSearch.select`` by ``install()``."""
    return beam_select(candidate_logprob, self.b_s, self.beam_size)


class FeatureEmbedding(nn.Module):
    """vision_embeddings.py:10-25 (``META_VISION_EMBEDDING``): ``dropout(gelu(proj(features)))`` and the padding mask
    of the all-zero region rows (models/utils.py:44-57).  Same parameter names (``proj.weight`` / ``proj.bias``) and config keys
    (``D_FEATURE``, ``D_MODEL``, ``DROPOUT``).  bf16 mode: one pass over the fp32 features (bf16 operand + mask) and one tensor-core GEMM
    with bias + GELU + dropout fused; fp32 mode (the 1e-4 parity path): the library's fp32 GEMM, then the reference's own element-wise ops."""

    def __init__(self, config):
        super().__init__()
        self.proj = nn.Linear(config.D_FEATURE, config.D_MODEL)
        self.gelu = nn.GELU()
        self.dropout = nn.Dropout(config.DROPOUT)
        self.compute_dtype = getattr(config, "COMPUTE_DTYPE", None)

    def forward(self, features):
        mode = _mode_of(self)
        if mode == "bf16" and features.dim() == 3 and features.dtype in (torch.float32, torch.bfloat16) and features.shape[-1] % 8 == 0:
            return Fn.feature_embedding(features, self.proj.weight, self.proj.bias, dropout=self.dropout.p, training=self.training,
                                        mask_value=MASK_VALUE)
        masks = (torch.sum(features, dim=-1) == 0).long() * MASK_VALUE
        masks = masks.unsqueeze(1).unsqueeze(1).to(features.dtype)
        y = Fn.linear(features, self.proj.weight, self.proj.bias, "fp32")
        return self.dropout(self.gelu(y)), masks


def mcan_forward(self, input_features):
    """``MCAN.forward`` (models/mcan.py:49-81) with the classifier tail (:70-80: attribute-reduce pooling, fusion LayerNorm, classifier)
    through the library's pooling head; bound over the reference's class by ``install()``.  The embeddings and encoders it calls are the
    reference's own objects (already built from the drop-ins)."""
    vision_features, vision_padding_mask = self.vision_embedding(input_features.region_features)
    text_features, (text_padding_mask, _) = self.text_embedding(input_features.question_tokens)
    text_features = self.self_encoder(features=text_features, padding_mask=text_padding_mask)
    vision_features = self.guided_encoder(vision_features=vision_features, vision_padding_mask=vision_padding_mask,
                                          language_features=text_features, language_padding_mask=text_padding_mask)
    return Fn.pool_and_classify(self, vision_features, text_features, get_compute_dtype(), log_softmax=True)


def make_lstm_text_embedding_forward(mod):
    """``LSTMTextEmbedding.forward`` (models/modules/text_embeddings.py:236-246) with the recurrence through the library's LSTM kernels when
    the compute dtype is bf16 and the hidden size is one the kernels are built for; bound over the reference's class by
    ``install()``.  ``mod`` is the reference module (its own mask generators are used); parameters stay the ``nn.LSTM``'s."""
    gen_pad, gen_seq = mod.generate_padding_mask, mod.generate_sequential_mask

    def forward(self, tokens):
        padding_masks = gen_pad(tokens, padding_idx=self.padding_idx).to(tokens.device)
        sequential_masks = gen_seq(tokens.shape[-1]).to(tokens.device)
        features = self.dropout(self.proj(self.embedding(tokens)))
        if Fn.lstm_supported(self.lstm, features, get_compute_dtype()):
            features = Fn.lstm(self.lstm, features)
        else:
            features, _ = self.lstm(features)
        return features, (padding_masks, sequential_masks)

    return forward


class OcrPtrNet(nn.Module):
    """mmf_m4c.py:367-396: ``(Wq·dec + b)(Wk·ocr + b)ᵀ / √d_qk + mask``."""

    def __init__(self, hidden_size, query_key_size=None):
        super().__init__()
        query_key_size = hidden_size if query_key_size is None else query_key_size
        self.hidden_size, self.query_key_size = hidden_size, query_key_size
        self.query = nn.Linear(hidden_size, query_key_size)
        self.key = nn.Linear(hidden_size, query_key_size)
        self.compute_dtype = None

    def forward(self, query_inputs, key_inputs, attention_mask, scores_out=None, col_offset=0):
        """``scores_out`` (B,T,W) fp32 + ``col_offset``: write the pointer scores straight into the joint score
        tensor (replaces the ``torch.cat`` at mmf_m4c.py:227).  Inference-only fast path; returns (B,T,Locr)."""
        mode = getattr(self, "compute_dtype", None) or _default_mode          # (also bound over the reference's own class: no such attribute there)
        squeeze = query_inputs.dim() == 2
        if squeeze:
            query_inputs = query_inputs.unsqueeze(1)
        odt = torch.float32 if mode == "fp32" else torch.bfloat16
        q = Fn.linear(query_inputs, self.query.weight, self.query.bias, mode, odt)
        k = Fn.linear(key_inputs, self.key.weight, self.key.bias, mode, odt)
        B, T, L = q.shape[0], q.shape[1], k.shape[1]
        scale = 1.0 / math.sqrt(self.query_key_size)
        if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad):
            if scores_out is not None:
                raise ValueError("OcrPtrNet: `scores_out` (in-place scores) is an inference-only fast path; call it under torch.no_grad()")
            res = Fn.ptr_scores(q, k, attention_mask, scale)                # differentiable (mmf_m4c.py:379-396 trains through it)
        else:
            if scores_out is None:
                scores_out, col_offset = torch.empty((B, T, L), device=q.device, dtype=torch.float32), 0
            ops.ptr_scores(q, k, attention_mask, scores_out, col_offset, scale)
            res = scores_out[..., col_offset:col_offset + L]
        return res.squeeze(1) if squeeze else res


class DynamicPointerNetwork(nn.Module):
    """m4c.py:19-33 (author's variant): ``QKᵀ/√d_model`` with ``masked_fill(-inf)`` on a boolean key mask."""

    def __init__(self, config):
        super().__init__()
        self.d_model = config.D_MODEL
        self.fc_q = nn.Linear(config.D_MODEL, config.D_MODEL)
        self.fc_k = nn.Linear(config.D_MODEL, config.D_MODEL)
        self.compute_dtype = getattr(config, "COMPUTE_DTYPE", None)

    def forward(self, query_inputs, key_inputs, attention_mask):
        mode = _mode_of(self)
        odt = torch.float32 if mode == "fp32" else torch.bfloat16
        q = Fn.linear(query_inputs, self.fc_q.weight, self.fc_q.bias, mode, odt)
        k = Fn.linear(key_inputs, self.fc_k.weight, self.fc_k.bias, mode, odt)
        return Fn.ptr_scores(q, k, attention_mask, 1.0 / math.sqrt(self.d_model), mask_is_bool=True)
