"""Autograd blocks of the hot path, built from the library calls in :mod:`openvivqa_b200.ops`.

Two fused blocks mirror the reference modules one-to-one:

* :func:`mha_block`  — ``MultiHeadAttention.forward`` (attentions.py:319-339): QKV projection(s) → attention core →
  output projection with bias + dropout + residual fused in the GEMM epilogue → LayerNorm.
* :func:`pwff_block` — ``PositionWiseFeedForward.forward`` (positionwise_feed_forward.py:23-27): fc1 + GELU(erf) +
  dropout in one epilogue, fc2 + dropout + residual in another, → LayerNorm.

Both have hand-written backward passes (dgrad/wgrad through the same GEMM kernel reading the saved tensors in
place, LayerNorm/attention/GELU backward kernels, dropout masks regenerated from (seed, offset)).

``mode`` is "fp32" (exact FFMA GEMMs, the 1e-4 parity path) or "bf16" (tcgen05 GEMMs with bf16 operands and fp32
accumulation; residual stream, softmax, LayerNorm statistics and masks stay fp32).
"""
from __future__ import annotations

import math
import weakref
from typing import List, Optional, Sequence

import torch
from torch.optim.optimizer import register_optimizer_step_post_hook

from . import ops

# ----------------------------------------------------------------------------- dropout counter-RNG stream
class _DropoutStream:
    """(seed, offset) provider for the stateless Philox masks; deterministic after torch.manual_seed()."""

    def __init__(self):
        self.offset = 0
        self._seed_of = None
        self.seed = 0
        self._base = {}

    def base(self, device) -> torch.Tensor:
        """Device word added to every offset inside the kernels (``offset_base`` of the C ABI).  Zero unless a
        CUDA-graph user advances it between replays with :meth:`advance_on_device` — the launch parameters
        (seed, offset) are frozen inside a captured graph, this word is what makes each replay draw new masks."""
        key = (device.type, device.index if (device.index is not None or device.type != "cuda") else torch.cuda.current_device())
        t = self._base.get(key)
        if t is None:
            t = self._base[key] = torch.zeros((1,), dtype=torch.int64, device=device)
        return t

    def advance_on_device(self, device, delta: int = 1 << 32):
        """Enqueue ``base += delta`` (capturable).  ``delta`` must exceed the offsets one step consumes."""
        ops.advance_u64(self.base(device), delta)

    def next(self, numel: int):
        s = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        if s != self._seed_of:
            self._seed_of, self.seed, self.offset = s, s ^ 0x5DEECE66D, 0
        off = self.offset
        self.offset += (numel + 3) // 4 + 1
        return self.seed, off


dropout_stream = _DropoutStream()


# ----------------------------------------------------------------------------- operand staging caches
class _Entry:
    __slots__ = ("refs", "dtype", "dst", "ver", "epoch", "staged")


class _WeightCache:
    """bf16 (and concatenated) copies of parameters.  Keeps fc_q/fc_k/fc_v as separately named parameters
    (state_dict compatibility) while the kernels see one fused [3*h*d, D] operand.

    A copy is refreshed when (a) a parameter's version counter or storage moves (``copy_``, ``load_state_dict``, foreach
    optimisers) or (b) ANY ``torch.optim`` optimiser has stepped since it was made: fused optimisers
    (``Adam(fused=True)``) update parameters without bumping ``_version``, so a process-wide optimiser-step
    post-hook advances ``epoch`` instead.  Code that writes parameters behind both mechanisms (raw
    ``torch._foreach_*`` on ``.data``) must call :func:`invalidate_weight_cache`.

    Case (b) — once per training step — re-stages EVERY registered copy of the device with one ``ovqa_stage_many``
    launch (a descriptor table of (src, dst, n, dtype) kept on the device) instead of a cat + cast per operand; the
    copies keep their addresses, so a captured CUDA graph replays that one node."""

    def __init__(self):
        self._store = {}
        self.epoch = 0
        self._tables = {}            # device -> {"dirty": bool, "host": pinned int64 (n,4), "dev": int64 (n,4), "n": int}

    def invalidate(self):
        self.epoch += 1

    @staticmethod
    def _versions(params):
        return tuple((p._version, p.data_ptr()) for p in params)

    def get(self, params: Sequence[torch.Tensor], dtype: torch.dtype) -> torch.Tensor:
        key = tuple(id(p) for p in params) + (dtype,)
        e = self._store.get(key)
        if e is not None and all(r() is p for r, p in zip(e.refs, params)):
            if e.ver == self._versions(params):
                if e.epoch == self.epoch:
                    return e.dst
                if e.staged:
                    self._restage_all(e.dst.device)
                    return e.dst
        return self._refresh(key, params, dtype)

    def _refresh(self, key, params, dtype):
        """(Re)build one copy the simple way: first use, or a parameter was rewritten / re-allocated."""
        with torch.no_grad():
            src = params[0].detach() if len(params) == 1 else torch.cat([p.detach() for p in params], dim=0)
            src = src.contiguous()
            val = ops.to_bf16(src) if dtype == torch.bfloat16 else src.float()
        e = self._store.get(key)
        direct = len(params) == 1 and val.data_ptr() == params[0].data_ptr()       # fp32 view of the parameter itself: nothing staged
        if e is not None and not direct and e.staged and e.dst.shape == val.shape and e.dst.dtype == val.dtype and e.dst.device == val.device:
            e.dst.copy_(val)                                                         # keep the staged address (descriptor table, graphs)
        else:
            e = _Entry()
            e.dst = val
            e.staged = (not direct) and val.is_cuda and all(p.is_contiguous() and p.dtype == torch.float32 for p in params)
            self._store[key] = e
        e.refs, e.dtype, e.ver, e.epoch = [weakref.ref(p) for p in params], dtype, self._versions(params), self.epoch
        if e.staged:
            self._tables.setdefault(val.device, {})["dirty"] = True                  # new entry or new source addresses
        if len(self._store) > 4096:
            self._prune()
        return e.dst

    def _prune(self):
        dead = [k for k, v in self._store.items() if any(r() is None for r in v.refs)]
        for k in dead:
            e = self._store.pop(k)
            if e.staged:
                self._tables.setdefault(e.dst.device, {})["dirty"] = True

    def _restage_all(self, device):
        """One launch refreshes every staged copy on ``device`` from the current parameter values."""
        t = self._tables.setdefault(device, {"dirty": True})
        if any(r() is None for e in self._store.values() if e.staged and e.dst.device == device for r in e.refs):
            self._prune()
        entries = [e for e in self._store.values() if e.staged and e.dst.device == device]
        if t.get("dirty", True):
            rows = []
            for e in entries:
                off, esz = 0, e.dst.element_size()
                for r in e.refs:
                    p = r()
                    rows.append((p.data_ptr(), e.dst.data_ptr() + off * esz, p.numel(), 1 if e.dst.dtype == torch.bfloat16 else 0))
                    off += p.numel()
            host = torch.tensor(rows, dtype=torch.int64).reshape(-1, 4).pin_memory()
            t.update(host=host, dev=host.to(device, non_blocking=True), n=len(rows), dirty=False)   # host kept alive (async copy, graphs)
        if t["n"]:
            ops.stage_many(t["dev"], t["n"])
        for e in entries:
            e.epoch, e.ver = self.epoch, self._versions([r() for r in e.refs])


weights = _WeightCache()
register_optimizer_step_post_hook(lambda *_a, **_k: weights.invalidate())


def invalidate_weight_cache():
    """Force the staged bf16 / fused operand copies to be rebuilt at their next use."""
    weights.invalidate()


def _operand(x: torch.Tensor, mode: str) -> torch.Tensor:
    """2-D GEMM operand of x (…,D) in the compute dtype; reuses the bf16 shadow a previous block attached."""
    D = x.shape[-1]
    if mode == "fp32":
        return x.reshape(-1, D).float().contiguous() if x.dtype != torch.float32 or not x.is_contiguous() else x.reshape(-1, D)
    sh = getattr(x, "_ovqa_bf16", None)
    if sh is not None and sh[0] == x._version and sh[1].shape == x.shape:
        return sh[1].reshape(-1, D)
    if x.dtype == torch.bfloat16:
        return x.reshape(-1, D).contiguous()
    return ops.to_bf16(x.reshape(-1, D).float())


def _attach_shadow(y: torch.Tensor, y16: Optional[torch.Tensor]):
    if y16 is not None:
        y._ovqa_bf16 = (y._version, y16)


def _cdt(mode: str) -> torch.dtype:
    return torch.float32 if mode == "fp32" else torch.bfloat16


def _wgrad(dy2: torch.Tensor, x2: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dW[n_out, n_in] = dy2^T x2 (fp32), both operands read where they lie ("T","T").  C has few tiles and the
    reduction over B*L rows is long: ``split_k=0`` lets the library split it to fill the SMs.  ``out``: zero-filled fp32 [n_out, n_in]."""
    return ops.gemm(dy2, x2, a_op="T", b_op="T", out=out, out_dtype=torch.float32, split_k=0)


def _zero_slab(device, *shapes):
    """fp32 zero-filled tensors of the given shapes carved from ONE allocation (one fill kernel instead of one per
    accumulator: every split-K weight gradient and every column-sum accumulator of a block's backward pass starts from zero)."""
    sizes = [(math.prod(sh) + 3) // 4 * 4 for sh in shapes]                 # 16-byte aligned starts (vector atomics)
    slab = torch.zeros((sum(sizes),), device=device, dtype=torch.float32)
    out, o = [], 0
    for sh, n in zip(shapes, sizes):
        out.append(slab[o:o + math.prod(sh)].view(sh))
        o += n
    return out


def _masked_operand(ds: torch.Tensor, ds16: Optional[torch.Tensor], mode: str, p: float, seed: int, off: int, rng_base=None) -> torch.Tensor:
    if mode == "bf16":
        return ds16.reshape(-1, ds.shape[-1])
    d2 = ds.reshape(-1, ds.shape[-1])
    return d2 if p == 0.0 else ops.dropout_apply(d2, p, seed, off, offset_base=rng_base)


def _proj_residual_ln(a2, w_c, bias, residual, gamma, beta, *, eps, p, seed, off, rng_base, mode, need_grad):
    """``LayerNorm(dropout(a2 w_c^T + bias) + residual)``: the tail of both blocks (attentions.py:58,331; positionwise_feed_forward.py:25-26).
    The GEMM with the fused bias / dropout / residual epilogue followed by the LayerNorm kernel; or, opt-in (``OVQA_GEMM_LN=1``, see
    ``ops.gemm_ln_enabled``: parity-green but currently slower), ONE kernel when the width is the fused kernel's (bf16 operands, 512 columns).  Returns (s, y, y16, mean, rstd); s / mean / rstd only when ``need_grad``."""
    if mode == "bf16" and ops.gemm_ln_enabled() and ops.gemm_ln_supported(a2, w_c):
        return ops.gemm_ln(a2, w_c, bias, residual, gamma, beta, eps=eps, dropout_p=p, seed=seed, offset=off, offset_base=rng_base,
                           save=need_grad, want_bf16=True)
    s = ops.gemm(a2, w_c, bias=bias, out_dtype=torch.float32, dropout_p=p, seed=seed, offset=off, offset_base=rng_base, residual=residual)
    y, y16, _, mean, rstd = ops.layernorm_fwd(s, gamma, beta, eps=eps, want_bf16=(mode == "bf16"), save_stats=need_grad)
    return s, y, y16, mean, rstd


# ----------------------------------------------------------------------------- generic Linear (AoA gate, pointer net, ...)
class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, mode, out_dtype):
        x2 = _operand(x, mode)
        w_c = weights.get([w], _cdt(mode))
        y = ops.gemm(x2, w_c, bias=weights.get([b], torch.float32) if b is not None else None, out_dtype=out_dtype)
        ctx.save_for_backward(x2, w_c)
        ctx.mode, ctx.has_b, ctx.xshape = mode, b is not None, x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w_c = ctx.saved_tensors
        mode = ctx.mode
        dy2 = dy.reshape(-1, dy.shape[-1])
        dyc = dy2.float().contiguous() if mode == "fp32" else ops.to_bf16(dy2.float() if dy2.dtype != torch.bfloat16 else dy2)
        dx = ops.gemm(dyc, w_c, b_op="T", out_dtype=torch.float32).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = _wgrad(dyc, x2) if ctx.needs_input_grad[1] else None
        db = ops.colsum(dyc) if ctx.has_b and ctx.needs_input_grad[2] else None
        return dx, dw, db, None, None


def linear(x, w, b=None, mode="fp32", out_dtype=torch.float32):
    """``F.linear`` through ovqa_gemm; compute dtype per ``mode``; output fp32 unless ``out_dtype`` says otherwise."""
    return _LinearFn.apply(x, w, b, mode, out_dtype)


# ----------------------------------------------------------------------------- fused MHA block
class _MHAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, queries, keys, values, mask, boxes, wq, bq, wk, bk, wv, bv, wo, bo, gamma, beta, geo_w, geo_b, cfg):
        ctx.set_materialize_grads(False)      # no zero-filled "gradient" tensor for the bf16 shadow output
        mode, h, dk, dv = cfg["mode"], cfg["h"], cfg["d_k"], cfg["d_v"]
        p = cfg["dropout"] if cfg["training"] else 0.0
        cdt = _cdt(mode)
        B, Lq, D = queries.shape
        Lk = keys.shape[1]
        self_att = (keys is queries) and (values is queries)
        kv_same = values is keys
        xq = _operand(queries, mode)
        xk = xq if keys is queries else _operand(keys, mode)
        xv = xk if kv_same else _operand(values, mode)
        hk, hv = h * dk, h * dv
        fb = lambda *bs: weights.get(list(bs), torch.float32)
        if self_att:
            qkv = ops.gemm(xq, weights.get([wq, wk, wv], cdt), bias=fb(bq, bk, bv), out_dtype=cdt).view(B, Lq, 2 * hk + hv)
            q, k, v = qkv[..., :hk], qkv[..., hk:2 * hk], qkv[..., 2 * hk:]
            proj = (qkv,)
        elif kv_same:
            qb = ops.gemm(xq, weights.get([wq], cdt), bias=fb(bq), out_dtype=cdt).view(B, Lq, hk)
            kvb = ops.gemm(xk, weights.get([wk, wv], cdt), bias=fb(bk, bv), out_dtype=cdt).view(B, Lk, hk + hv)
            q, k, v = qb, kvb[..., :hk], kvb[..., hk:]
            proj = (qb, kvb)
        else:
            q = ops.gemm(xq, weights.get([wq], cdt), bias=fb(bq), out_dtype=cdt).view(B, Lq, hk)
            k = ops.gemm(xk, weights.get([wk], cdt), bias=fb(bk), out_dtype=cdt).view(B, Lk, hk)
            v = ops.gemm(xv, weights.get([wv], cdt), bias=fb(bv), out_dtype=cdt).view(B, Lk, hv)
            proj = (q, k, v)
        geo = ops.GeoParams(boxes, geo_w.detach(), geo_b.detach(), cfg["geo_trig"]) if boxes is not None else None
        need_grad = cfg["grad"] and any(ctx.needs_input_grad)     # (grad mode is always off INSIDE forward: flag taken outside)
        geo_as_bias = geo is not None and mode == "bf16"
        if geo_as_bias:
            # tensor-core route of the geometry core: the box bias (pair embedding evaluated once for all heads) + the padding mask become
            # one additive (B,h,N,N) tensor for the bf16 attention kernels; backward gets dS back and reduces the fc_g gradients from it
            per_key = mask is None or (mask.dim() == 4 and mask.shape[1] == 1 and mask.shape[2] == 1)
            mask = ops.geo_bias(geo, h, key_mask=mask if per_key else None) if per_key else ops.geo_bias(geo, h) + mask
            geo = None
        pp = cfg.get("prob_dropout", 0.0) if cfg["training"] else 0.0          # dropout on the probabilities (HF BERT layers): its own counter range
        prob_dropout = None
        if pp > 0.0:
            ps, po = dropout_stream.next(B * h * Lq * Lk)
            prob_dropout = (pp, ps, po, dropout_stream.base(queries.device))
        att_ctx, lse, _ = ops.sdpa_fwd(q, k, v, h=h, d_k=dk, d_v=dv, mask=mask, geo=geo, scale=cfg["scale"], need_lse=need_grad, prob_dropout=prob_dropout)
        seed, off = dropout_stream.next(B * Lq * D) if p > 0.0 else (0, 0)
        rng_base = dropout_stream.base(queries.device) if p > 0.0 else None
        q_res = queries.reshape(-1, D)
        q_res = q_res if (q_res.dtype == torch.float32 and q_res.is_contiguous()) else q_res.float().contiguous()
        s, y, y16, mean, rstd = _proj_residual_ln(att_ctx.view(B * Lq, hv), weights.get([wo], cdt), fb(bo), q_res, gamma.detach(), beta.detach(),
                                                  eps=cfg["eps"], p=p, seed=seed, off=off, rng_base=rng_base, mode=mode, need_grad=need_grad)
        y = y.view(B, Lq, D)
        if need_grad:
            ctx.save_for_backward(xq, xk, xv, mask, *proj, att_ctx, lse, s, mean, rstd, gamma, wq, wk, wv, wo,
                                  boxes, geo_w, geo_b)
            ctx.n_proj = len(proj)
            ctx.cfg, ctx.p, ctx.seed, ctx.off, ctx.rng_base = cfg, p, seed, off, rng_base
            ctx.flags = (self_att, kv_same)
            ctx.geo_as_bias = geo_as_bias
            ctx.prob_dropout = prob_dropout
            ctx.shapes = (B, Lq, Lk, D)
        y16 = y16.view(B, Lq, D) if y16 is not None else y.new_empty(0)
        ctx.mark_non_differentiable(y16)
        return y, y16

    @staticmethod
    def backward(ctx, dy, _dy16):
        if dy is None:
            return (None,) * 18
        cfg = ctx.cfg
        mode, h, dk, dv = cfg["mode"], cfg["h"], cfg["d_k"], cfg["d_v"]
        cdt = _cdt(mode)
        sv = ctx.saved_tensors
        xq, xk, xv, mask = sv[:4]
        proj = sv[4:4 + ctx.n_proj]
        att_ctx, lse, s, mean, rstd, gamma, wq, wk, wv, wo, boxes, geo_w, geo_b = sv[4 + ctx.n_proj:]
        self_att, kv_same = ctx.flags
        B, Lq, Lk, D = ctx.shapes
        hk, hv = h * dk, h * dv
        p, seed, off, rng_base = ctx.p, ctx.seed, ctx.off, ctx.rng_base
        # every accumulator of this backward pass (split-K weight gradients, column sums) comes from one zero-filled slab
        if self_att:
            w_shapes = [(2 * hk + hv, D)]
        elif kv_same:
            w_shapes = [(hk, D), (hk + hv, D)]
        else:
            w_shapes = [(hk, D), (hk, D), (hv, D)]
        dgamma, dbeta, dbo, dbias, dwo, *dWs = _zero_slab(dy.device, (D,), (D,), (D,), (2 * hk,), (D, hv), *w_shapes)
        # LayerNorm backward; the bf16 copy already carries the output-projection dropout mask
        ds, ds16, *_ = ops.layernorm_bwd(dy.reshape(B * Lq, D), s, mean, rstd, gamma, want_bf16=(mode == "bf16"), dropout_p=p, seed=seed,
                                         offset=off, offset_base=rng_base, dgamma=dgamma, dbeta=dbeta, want_dbias=True, dbias=dbo)
        d_o = _masked_operand(ds, ds16, mode, p, seed, off, rng_base)             # grad of the fc_o output (its column sums: dbo)
        ctx2 = att_ctx.view(B * Lq, hv)
        _wgrad(d_o, ctx2, out=dwo)
        # projection bias gradients without a pass over dq/dk/dv:  d(bv) = sum_rows dV = sum_rows dctx (softmax rows sum to 1)
        # = (sum_rows d_o) Wo = dbo Wo, a 512-vector times the fp32 master weight;  d(bk) == 0 analytically (softmax is shift
        # invariant);  d(bq) = column sums of dq.
        dbq, dbk = dbias[:hk], dbias[hk:]
        dbv = torch.mv(wo.detach().float().t(), dbo)
        dctx = ops.gemm(d_o, weights.get([wo], cdt), b_op="T", out_dtype=cdt).view(B, Lq, hv)
        geo = ops.GeoParams(boxes, geo_w.detach(), geo_b.detach(), cfg["geo_trig"]) if boxes is not None else None
        gdw = torch.zeros_like(geo_w, dtype=torch.float32) if geo is not None else None
        gdb = torch.zeros_like(geo_b, dtype=torch.float32) if geo is not None else None
        if self_att:
            (qkv,) = proj
            dqkv = torch.empty_like(qkv)
            q, k, v = qkv[..., :hk], qkv[..., hk:2 * hk], qkv[..., 2 * hk:]
            dq, dk_, dv_ = dqkv[..., :hk], dqkv[..., hk:2 * hk], dqkv[..., 2 * hk:]
        elif kv_same:
            qb, kvb = proj
            dqb, dkvb = torch.empty_like(qb), torch.empty_like(kvb)
            q, k, v = qb, kvb[..., :hk], kvb[..., hk:]
            dq, dk_, dv_ = dqb, dkvb[..., :hk], dkvb[..., hk:]
        else:
            q, k, v = proj
            dq, dk_, dv_ = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        if ctx.geo_as_bias:                                      # `mask` is the saved (B,h,N,N) bias: get dS back, reduce the fc_g gradients from it
            dbias = torch.empty_like(mask)
            ops.sdpa_bwd(q, k, v, att_ctx, dctx, lse, dq, dk_, dv_, h=h, d_k=dk, d_v=dv, mask=mask, scale=cfg["scale"], dbias=dbias)
            ops.geo_bias_bwd(geo, h, dbias, gdw, gdb)
        else:
            ops.sdpa_bwd(q, k, v, att_ctx, dctx, lse, dq, dk_, dv_, h=h, d_k=dk, d_v=dv, mask=mask, geo=geo, geo_dw=gdw,
                         geo_db=gdb, scale=cfg["scale"], prob_dropout=ctx.prob_dropout)
        if ctx.prob_dropout is not None:
            # dropped probabilities do not sum to 1 over the keys: the value-bias shortcut above does not hold, take the column sums of dV
            dbv = ops.colsum(dv_.reshape(-1, dv_.shape[-1]) if dv_.is_contiguous() else dv_.view(B * Lk, dv_.shape[-1]))
        dkeys = dvalues = None
        if self_att:
            d2 = dqkv.view(B * Lq, dqkv.shape[-1])
            dW = _wgrad(d2, xq, out=dWs[0])
            ops.colsum(d2[:, :hk], out=dbq, accumulate=True)
            dwq, dwk, dwv = dW[:hk], dW[hk:2 * hk], dW[2 * hk:]
            dqueries = ops.gemm(d2, weights.get([wq, wk, wv], cdt), b_op="T", out_dtype=torch.float32, residual=ds.view(B * Lq, D)).view(B, Lq, D)
        else:
            dq2 = dq.reshape(B * Lq, hk)
            dwq = _wgrad(dq2, xq, out=dWs[0])
            ops.colsum(dq2, out=dbq, accumulate=True)
            dqueries = ops.gemm(dq2, weights.get([wq], cdt), b_op="T", out_dtype=torch.float32, residual=ds.view(B * Lq, D)).view(B, Lq, D)
            if kv_same:
                d2 = dkvb.view(B * Lk, dkvb.shape[-1])
                dW = _wgrad(d2, xk, out=dWs[1])
                dwk, dwv = dW[:hk], dW[hk:]
                dkeys = ops.gemm(d2, weights.get([wk, wv], cdt), b_op="T", out_dtype=torch.float32).view(B, Lk, D)
            else:
                dk2, dv2 = dk_.reshape(B * Lk, hk), dv_.reshape(B * Lk, hv)
                dwk, dwv = _wgrad(dk2, xk, out=dWs[1]), _wgrad(dv2, xv, out=dWs[2])
                dkeys = ops.gemm(dk2, weights.get([wk], cdt), b_op="T", out_dtype=torch.float32).view(B, Lk, D)
                dvalues = ops.gemm(dv2, weights.get([wv], cdt), b_op="T", out_dtype=torch.float32).view(B, Lk, D)
        return (dqueries, dkeys, dvalues, None, None, dwq, dbq, dwk, dbk, dwv, dbv, dwo, dbo, dgamma, dbeta, gdw, gdb, None)


ONE_CALL_INFERENCE = True      # eval-mode blocks without autograd go through ONE library call (ovqa_mha_block_infer); bench.py switches it off
                               # while it meters the individual kernels of a step
_block_plans = {}


def _block_plan(wq, bq, wk, bk, wv, bv, wo, bo, gamma, beta, mode):
    key = (id(wq), id(wo), mode)
    e = _block_plans.get(key)
    if e is not None and e[1]() is wq and e[0].valid():
        return e[0]
    plan = mha_decode_plan(wq=wq, bq=bq, wk=wk, bk=bk, wv=wv, bv=bv, wo=wo, bo=bo, gamma=gamma, beta=beta, mode=mode)
    if len(_block_plans) > 1024:
        for k in [k for k, v in _block_plans.items() if v[1]() is None]:
            del _block_plans[k]
    _block_plans[key] = (plan, weakref.ref(wq))
    return plan


def _mha_block_infer(queries, keys, values, mask, plan, *, h, d_k, d_v, eps, scale):
    """The block in eval mode from one library call; None when that path does not apply (the caller runs the per-op block)."""
    if queries.dim() != 3 or not queries.is_cuda:
        return None
    mode = plan.mode
    kind = 0 if (keys is queries and values is queries) else (1 if values is keys else 2)
    xq = _f32c(queries)
    xk = None if kind == 0 else _f32c(keys)
    xv = None if kind != 2 else _f32c(values)
    sh = lambda t: None if t is None else _shadow_of(t, mode)
    res = ops.mha_block_infer(xq, sh(queries), xk, None if kind == 0 else sh(keys), xv, None if kind != 2 else sh(values), kind, plan.ptrs, dt=plan.dt,
                              eps=eps, h=h, d_k=d_k, d_v=d_v, mask=mask, scale=scale)
    if res is None:
        return None
    y, y16 = res
    if y16 is not None:
        y._ovqa_bf16 = (y._version, y16)
    return y


def mha_block(queries, keys, values, mask, *, wq, bq, wk, bk, wv, bv, wo, bo, gamma, beta, h, d_k, d_v, mode="fp32",
              dropout=0.0, training=False, eps=1e-5, boxes=None, geo_w=None, geo_b=None, geo_trig=True, scale=None, prob_dropout=0.0):
    """``prob_dropout``: dropout on the attention PROBABILITIES (HF ``BertSelfAttention``; the reference's own cores have none)."""
    if ONE_CALL_INFERENCE and boxes is None and not torch.is_grad_enabled() and not (training and (dropout > 0.0 or prob_dropout > 0.0)):
        y = _mha_block_infer(queries, keys, values, mask, _block_plan(wq, bq, wk, bk, wv, bv, wo, bo, gamma, beta, mode), h=h, d_k=d_k, d_v=d_v,
                             eps=float(eps), scale=(1.0 / math.sqrt(d_k)) if scale is None else float(scale))
        if y is not None:
            return y
    cfg = dict(grad=torch.is_grad_enabled(), mode=mode, h=h, d_k=d_k, d_v=d_v, dropout=float(dropout), prob_dropout=float(prob_dropout),
               training=bool(training), eps=float(eps),
               geo_trig=bool(geo_trig), scale=(1.0 / math.sqrt(d_k)) if scale is None else float(scale))
    y, y16 = _MHAFn.apply(queries, keys, values, mask, boxes, wq, bq, wk, bk, wv, bv, wo, bo, gamma, beta, geo_w, geo_b, cfg)
    _attach_shadow(y, y16 if mode == "bf16" else None)
    return y


# ----------------------------------------------------------------------------- incremental decode (projected K/V cache)
def project_kv_fused(keys, values, *, wk, bk, wv, bv, mode="fp32"):
    """[K | V] = [fc_k(keys) | fc_v(values)] as one (B, L, h*d_k + h*d_v) buffer in the compute dtype (the layout of the decode cache)."""
    cdt = _cdt(mode)
    B, Lk = keys.shape[0], keys.shape[1]
    hk, hv = wk.shape[0], wv.shape[0]
    fb = lambda *bs: weights.get(list(bs), torch.float32)
    with torch.no_grad():
        xk = _operand(keys, mode)
        if values is keys:
            return ops.gemm(xk, weights.get([wk, wv], cdt), bias=fb(bk, bv), out_dtype=cdt).view(B, Lk, hk + hv)
        kv = torch.empty((B * Lk, hk + hv), device=keys.device, dtype=cdt)
        ops.gemm(xk, weights.get([wk], cdt), bias=fb(bk), out=kv[:, :hk])
        ops.gemm(_operand(values, mode), weights.get([wv], cdt), bias=fb(bv), out=kv[:, hk:])
        return kv.view(B, Lk, hk + hv)


def _f32c(x):
    return x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous()


def _shadow_of(x, mode):
    sh = getattr(x, "_ovqa_bf16", None) if mode == "bf16" else None
    return sh[1] if (sh is not None and sh[0] == x._version and sh[1].shape == x.shape and sh[1].is_contiguous()) else None


class InferPlan:
    """Staged operands of one block resolved ONCE (fused / bf16 weight copies, fp32 biases, LayerNorm affine) and their device
    pointers, so a decode step does no cache lookups: at beam-search batch sizes the step is bounded by host dispatch, not by the GPU.
    ``valid()`` re-checks the parameters' version counters / storage and the optimiser-step epoch (a handful of attribute reads)."""
    __slots__ = ("params", "stamp", "keep", "ptrs", "dt", "mode")

    def __init__(self, mode, groups, affine):
        """``groups``: list of (parameter list, "w" | "b") -> one staged operand each (weights in the compute dtype, biases fp32);
        ``affine``: (gamma, beta) passed as they are."""
        cdt = _cdt(mode)
        self.mode, self.dt = mode, (0 if mode == "fp32" else 1)
        self.params = [p for g, _ in groups for p in g] + list(affine)
        self.keep = [weights.get(list(g), cdt if kind == "w" else torch.float32) for g, kind in groups] + [a.detach() for a in affine]
        for t in self.keep:
            if not t.is_contiguous():
                raise ValueError("InferPlan: staged operands must be contiguous")
        self.ptrs = tuple(t.data_ptr() for t in self.keep)
        self.stamp = self._stamp()

    def _stamp(self):
        return (weights.epoch,) + tuple((p._version, p.data_ptr()) for p in self.params)

    def valid(self):
        return self.stamp == self._stamp()


def mha_decode_plan(*, wq, bq, wk, bk, wv, bv, wo, bo, gamma, beta, mode="fp32") -> InferPlan:
    return InferPlan(mode, [((wq, wk, wv), "w"), ((bq, bk, bv), "b"), ((wo,), "w"), ((bo,), "b")], (gamma, beta))


def pwff_plan(*, w1, b1, w2, b2, gamma, beta, mode="fp32") -> InferPlan:
    return InferPlan(mode, [((w1,), "w"), ((b1,), "b"), ((w2,), "w"), ((b2,), "b")], (gamma, beta))


def mha_decode_block(plan: InferPlan, queries, kv_src, self_attention, kv_cache, length, mask, *, h, d_k, d_v, eps=1e-5, scale=None):
    """Inference-only MHA block of a decode step over the projected-K/V cache (ONE library call, see csrc/decode.cu): fc_q (fused with
    fc_k / fc_v of the new rows) -> attention over cache + new rows (which it files into the cache) -> fc_o + residual -> LayerNorm.
    Same arithmetic as ``mha_block`` in eval mode (attentions.py:319-331); no autograd graph.  ``kv_src``: (B, Lnew, D) rows to append
    (None: none), ignored when ``self_attention`` (the query rows are appended)."""
    mode = plan.mode
    x = _f32c(queries)
    src = None
    if kv_src is not None and not self_attention:
        src = _operand(kv_src, mode).view(kv_src.shape)
    y, y16 = ops.mha_decode_step(x, _shadow_of(queries, mode), src, self_attention, plan.ptrs, dt=plan.dt, eps=eps, h=h, d_k=d_k, d_v=d_v,
                                 kv_cache=kv_cache, length=length, mask=mask, scale=(1.0 / math.sqrt(d_k)) if scale is None else scale)
    if y16 is not None:
        y._ovqa_bf16 = (y._version, y16)
    return y


def pwff_infer_block(plan: InferPlan, x, *, d_ff, eps=1e-5):
    """Inference-only feed-forward block from ONE library call (positionwise_feed_forward.py:23-27 with dropout off)."""
    y, y16 = ops.pwff_infer(_f32c(x), _shadow_of(x, plan.mode), plan.ptrs, dt=plan.dt, d_ff=d_ff, eps=eps)
    if y16 is not None:
        y._ovqa_bf16 = (y._version, y16)
    return y


# ----------------------------------------------------------------------------- fused PWFF block
class _PWFFFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, gamma, beta, cfg):
        ctx.set_materialize_grads(False)
        mode = cfg["mode"]
        p = cfg["dropout"] if cfg["training"] else 0.0
        p1 = cfg["dropout_inner"] if cfg["training"] else 0.0               # after the activation (0 for an HF BertIntermediate)
        cdt = _cdt(mode)
        D = x.shape[-1]
        x2 = _operand(x, mode)
        M, dff = x2.shape[0], w1.shape[0]
        need_grad = cfg["grad"] and any(ctx.needs_input_grad)
        pre = torch.empty((M, dff), device=x.device, dtype=cdt) if need_grad else None
        s1, o1 = dropout_stream.next(M * dff) if p1 > 0.0 else (0, 0)
        rng_base = dropout_stream.base(x.device) if (p > 0.0 or p1 > 0.0) else None
        hact = ops.gemm(x2, weights.get([w1], cdt), bias=weights.get([b1], torch.float32), out_dtype=cdt, gelu=True,
                        aux_out=pre, save_dgelu=need_grad, dropout_p=p1, seed=s1, offset=o1, offset_base=rng_base)   # pre := GELU'(fc1 x) * dropout scale
        s2, o2 = dropout_stream.next(M * D) if p > 0.0 else (0, 0)
        x_res = x.reshape(-1, D)
        x_res = x_res if (x_res.dtype == torch.float32 and x_res.is_contiguous()) else x_res.float().contiguous()
        s, y, y16, mean, rstd = _proj_residual_ln(hact, weights.get([w2], cdt), weights.get([b2], torch.float32), x_res, gamma.detach(), beta.detach(),
                                                  eps=cfg["eps"], p=p, seed=s2, off=o2, rng_base=rng_base, mode=mode, need_grad=need_grad)
        if need_grad:
            ctx.save_for_backward(x2, pre, hact, s, mean, rstd, gamma, w1, w2)
            ctx.cfg, ctx.p, ctx.rng, ctx.rng_base = cfg, p, (s1, o1, s2, o2), rng_base
            ctx.xshape = x.shape
        y16 = y16.view(x.shape) if y16 is not None else y.new_empty(0)
        ctx.mark_non_differentiable(y16)
        return y.view(x.shape), y16

    @staticmethod
    def backward(ctx, dy, _dy16):
        if dy is None:
            return (None,) * 8
        x2, pre, hact, s, mean, rstd, gamma, w1, w2 = ctx.saved_tensors
        mode, p = ctx.cfg["mode"], ctx.p
        s1, o1, s2, o2 = ctx.rng
        rng_base = ctx.rng_base
        cdt = _cdt(mode)
        D = x2.shape[1]
        dff = w1.shape[0]
        dgamma, dbeta, db2, db1, dw2, dw1 = _zero_slab(dy.device, (D,), (D,), (D,), (dff,), (D, dff), (dff, D))     # one fill for every accumulator
        ds, ds16, *_ = ops.layernorm_bwd(dy.reshape(-1, D), s, mean, rstd, gamma, want_bf16=(mode == "bf16"), dropout_p=p, seed=s2,
                                         offset=o2, offset_base=rng_base, dgamma=dgamma, dbeta=dbeta, want_dbias=True, dbias=db2)
        d2 = _masked_operand(ds, ds16, mode, p, s2, o2, rng_base)                   # grad of the fc2 output (its column sums: db2)
        _wgrad(d2, hact, out=dw2)
        # grad of the fc1 pre-activation: (d2 W2) * [dropout1 scale * gelu'(pre)] — the forward saved that product (the derivative
        # of its whole fused activation), so the dgrad epilogue is one multiply: no transcendental, no RNG
        dpre = ops.gemm(d2, weights.get([w2], cdt), b_op="T", out_dtype=cdt, mul_aux=pre, colsum=db1)     # + fc1 bias gradient (column sums)
        _wgrad(dpre, x2, out=dw1)
        dx = ops.gemm(dpre, weights.get([w1], cdt), b_op="T", out_dtype=torch.float32, residual=ds.reshape(-1, D)).view(ctx.xshape)
        return dx, dw1, db1, dw2, db2, dgamma, dbeta, None


def pwff_block(x, *, w1, b1, w2, b2, gamma, beta, mode="fp32", dropout=0.0, training=False, eps=1e-5, dropout_hidden=True):
    """``dropout_hidden=False``: no dropout after the activation (HF ``BertIntermediate``; the reference's own block drops there too)."""
    if ONE_CALL_INFERENCE and not torch.is_grad_enabled() and not (training and dropout > 0.0) and x.is_cuda:
        key = (id(w1), id(w2), mode)
        e = _block_plans.get(key)
        if e is None or e[1]() is not w1 or not e[0].valid():
            e = (pwff_plan(w1=w1, b1=b1, w2=w2, b2=b2, gamma=gamma, beta=beta, mode=mode), weakref.ref(w1))
            _block_plans[key] = e
        return pwff_infer_block(e[0], x, d_ff=w1.shape[0], eps=float(eps))
    cfg = dict(grad=torch.is_grad_enabled(), mode=mode, dropout=float(dropout), dropout_inner=float(dropout) if dropout_hidden else 0.0,
               training=bool(training), eps=float(eps))
    y, y16 = _PWFFFn.apply(x, w1, b1, w2, b2, gamma, beta, cfg)
    _attach_shadow(y, y16 if mode == "bf16" else None)
    return y


# ----------------------------------------------------------------------------- FeatureEmbedding front-end
class _FeatureEmbeddingFn(torch.autograd.Function):
    """``dropout(gelu(x W^T + b))`` and the zero-row padding mask of ``x`` (vision_embeddings.py:19-25), bf16 path: one pass over the fp32
    features (bf16 operand + mask), one GEMM with the whole activation fused (its saved derivative makes backward one multiply)."""

    @staticmethod
    def forward(ctx, x, w, b, cfg):
        ctx.set_materialize_grads(False)
        p = cfg["dropout"] if cfg["training"] else 0.0
        K, D = x.shape[-1], w.shape[0]
        x16, mask = ops.cast_rows_padmask(x.reshape(-1, K), cfg["mask_value"])
        need_grad = cfg["grad"] and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        aux = torch.empty((x16.shape[0], D), device=x.device, dtype=torch.float32) if need_grad else None
        seed, off = dropout_stream.next(x16.shape[0] * D) if p > 0.0 else (0, 0)
        y = ops.gemm(x16, weights.get([w], torch.bfloat16), bias=weights.get([b], torch.float32), out_dtype=torch.float32, gelu=True,
                     aux_out=aux, save_dgelu=need_grad, dropout_p=p, seed=seed, offset=off,
                     offset_base=dropout_stream.base(x.device) if p > 0.0 else None)
        if need_grad:
            ctx.save_for_backward(x16, aux)
        ctx.mark_non_differentiable(mask)
        return y.view(*x.shape[:-1], D), mask.view(x.shape[0], 1, 1, -1) if x.dim() == 3 else mask

    @staticmethod
    def backward(ctx, dy, _dmask):
        if dy is None:
            return None, None, None, None
        x16, aux = ctx.saved_tensors
        D = aux.shape[1]
        dpre = ops.mul_to_bf16(dy.reshape(-1, D).float(), aux)             # grad of the pre-activation: dy * GELU'(pre) * dropout scale
        dw, db = _zero_slab(dy.device, (D, x16.shape[1]), (D,))
        _wgrad(dpre, x16, out=dw)
        ops.colsum(dpre, out=db, accumulate=True)
        return None, dw, db, None


def feature_embedding(features, w, b, *, dropout=0.0, training=False, mask_value=-10e4):
    """(B, N, K) fp32 region features -> ((B, N, D) fp32, (B,1,1,N) additive padding mask); bf16 tensor-core path only."""
    cfg = dict(grad=torch.is_grad_enabled(), dropout=float(dropout), training=bool(training), mask_value=float(mask_value))
    return _FeatureEmbeddingFn.apply(features, w, b, cfg)


# ----------------------------------------------------------------------------- attribute-reduce pooling head
class _AttrReduceFn(torch.autograd.Function):
    """``(x * softmax(fc2(dropout(relu(fc1 x))), dim=1)).sum(dim=1)`` (models/mcan.py:21-25,70-77): one GEMM with the ReLU + dropout epilogue,
    one pooling kernel (scores, softmax over the sequence and the weighted sum in one pass per sample); backward: one pooling kernel (dx,
    fc1's pre-activation gradient, fc2's gradients), the split-K weight gradient, the dgrad GEMM that also adds the pooling's dx."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, cfg):
        ctx.set_materialize_grads(False)
        mode = cfg["mode"]
        p = cfg["dropout"] if cfg["training"] else 0.0
        cdt = _cdt(mode)
        B, N, D = x.shape
        Dh = w1.shape[0]
        x2 = _operand(x, mode)
        xf = _f32c(x)
        seed, off = dropout_stream.next(B * N * Dh) if p > 0.0 else (0, 0)
        hid = ops.gemm(x2, weights.get([w1], cdt), bias=weights.get([b1], torch.float32), out_dtype=cdt, relu=True, dropout_p=p, seed=seed, offset=off,
                       offset_base=dropout_stream.base(x.device) if p > 0.0 else None).view(B, N, Dh)
        need_grad = cfg["grad"] and any(ctx.needs_input_grad)
        pooled, probs = ops.attr_pool_fwd(hid, w2.detach(), b2.detach(), xf, save_probs=need_grad)
        if need_grad:
            ctx.save_for_backward(x2, xf, hid, probs, w1, w2)
            ctx.cfg, ctx.p = cfg, p
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        if dpooled is None:
            return (None,) * 6
        x2, xf, hid, probs, w1, w2 = ctx.saved_tensors
        mode = ctx.cfg["mode"]
        B, N, Dh = hid.shape
        D = xf.shape[-1]
        dw1, db1, dw2, db2 = _zero_slab(dpooled.device, (Dh, D), (Dh,), tuple(w2.shape), (1,))
        dx_pool, dpre = ops.attr_pool_bwd(dpooled, hid, w2.detach(), xf, probs, ctx.p, dw2.view(-1), db2, need_dx=ctx.needs_input_grad[0])
        dpre2 = dpre.view(B * N, Dh)
        _wgrad(dpre2, x2, out=dw1)
        ops.colsum(dpre2, out=db1, accumulate=True)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dpre2, weights.get([w1], _cdt(mode)), b_op="T", out_dtype=torch.float32, residual=dx_pool.view(B * N, D)).view(B, N, D)
        return dx, dw1, db1, dw2, db2, None


def attr_reduce_pool(x, *, w1, b1, w2, b2, mode="fp32", dropout=0.0, training=False):
    """(B,N,D) -> (B,D): the attribute-reduce MLP, the softmax over the sequence and the weighted sum of models/mcan.py:70-77."""
    cfg = dict(grad=torch.is_grad_enabled(), mode=mode, dropout=float(dropout), training=bool(training))
    return _AttrReduceFn.apply(x, w1, b1, w2, b2, cfg)


def pool_and_classify(model, vision_features, text_features, mode, log_softmax=False):
    """The classifier tail shared by MCAN / CrossModalityTransformer / VisiolinguisticTransformer (models/mcan.py:70-81): both
    attribute-reduce poolings, ``layer_norm(vision_proj(.) + text_proj(.))`` and ``classify`` through the library.  ``model`` is any
    module with the reference's attribute names (``vision_attr_reduce`` / ``text_attr_reduce`` MLPs with fc1 / dropout / fc2,
    ``vision_proj``, ``text_proj``, ``layer_norm``, ``classify``)."""
    def pool(mlp, feats):
        return attr_reduce_pool(feats, w1=mlp.fc1.weight, b1=mlp.fc1.bias, w2=mlp.fc2.weight, b2=mlp.fc2.bias, mode=mode, dropout=mlp.dropout.p,
                                training=mlp.training)
    wv, wt = pool(model.vision_attr_reduce, vision_features), pool(model.text_attr_reduce, text_features)
    # B rows each: the two fusion projections and the classifier stay on the exact fp32 GEMM in every mode — 0.2 GFLOP per batch, and the
    # logits are where operand rounding shows (bf16 operands here alone moved MCAN's log-probabilities by 5e-3)
    s = linear(wv, model.vision_proj.weight, model.vision_proj.bias, "fp32") + linear(wt, model.text_proj.weight, model.text_proj.bias, "fp32")
    out = layer_norm(s, model.layer_norm.weight, model.layer_norm.bias, model.layer_norm.eps)
    logits = linear(out, model.classify.weight, model.classify.bias, "fp32")
    return torch.log_softmax(logits, dim=-1) if log_softmax else logits


# ----------------------------------------------------------------------------- LSTM text embedding
class _LstmFn(torch.autograd.Function):
    """``nn.LSTM(batch_first=True)`` of one layer called without an initial state (text_embeddings.py:234,244): the input projection of all
    steps is one tensor-core GEMM, the recurrence one cluster-resident kernel (``ovqa_lstm_fwd``: W_hh in registers, h exchanged through
    distributed shared memory); backward: one recurrence kernel producing the gate pre-activation gradients, then the input gradient, both
    weight gradients (split-K, operands read in place) and the bias gradient as GEMMs / a column sum over all steps."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh, grad):
        ctx.set_materialize_grads(False)
        B, T, In = x.shape
        H = w_hh.shape[1]
        x2 = _operand(x, "bf16")
        wih16, whh16 = weights.get([w_ih], torch.bfloat16), weights.get([w_hh], torch.bfloat16)
        xg = ops.gemm(x2, wih16, bias=weights.get([b_ih], torch.float32) if b_ih is not None else None, out_dtype=torch.float32).view(B, T, 4 * H)
        need_grad = grad and any(ctx.needs_input_grad)
        y, y16, hprev16, gates16, c = ops.lstm_fwd(xg, whh16, weights.get([b_hh], torch.float32) if b_hh is not None else None, save=need_grad)
        _attach_shadow(y, y16)
        if need_grad:
            ctx.save_for_backward(x2, wih16, whh16, hprev16, gates16, c)
            ctx.has_b, ctx.dims = (b_ih is not None, b_hh is not None), (B, T, In, H)
        return y

    @staticmethod
    def backward(ctx, dy):
        if dy is None:
            return (None,) * 6
        x2, wih16, whh16, hprev16, gates16, c = ctx.saved_tensors
        B, T, In, H = ctx.dims
        dg2 = ops.lstm_bwd(dy.float().contiguous(), whh16, gates16, c).view(B * T, 4 * H)
        dw_ih, dw_hh, db = _zero_slab(dy.device, (4 * H, In), (4 * H, H), (4 * H,))
        dx = ops.gemm(dg2, wih16, b_op="T", out_dtype=torch.float32).view(B, T, In) if ctx.needs_input_grad[0] else None
        if ctx.needs_input_grad[1]:
            _wgrad(dg2, x2, out=dw_ih)
        if ctx.needs_input_grad[2]:
            _wgrad(dg2, hprev16.view(B * T, H), out=dw_hh)
        if (ctx.has_b[0] and ctx.needs_input_grad[3]) or (ctx.has_b[1] and ctx.needs_input_grad[4]):
            ops.colsum(dg2, out=db, accumulate=True)
        return (dx, dw_ih if ctx.needs_input_grad[1] else None, dw_hh if ctx.needs_input_grad[2] else None,
                db if ctx.has_b[0] and ctx.needs_input_grad[3] else None, db.clone() if ctx.has_b[1] and ctx.needs_input_grad[4] else None, None)


def lstm_supported(module, x, mode) -> bool:
    """Whether :func:`lstm` can run ``module`` (an ``nn.LSTM``) on ``x``: bf16 compute mode, CUDA, one unidirectional batch-first layer without
    projection, hidden size one of the kernel's (``ovqa_lstm_supported``), input width a multiple of 8 (TMA rows of the projection GEMM)."""
    return (mode == "bf16" and x.is_cuda and x.dim() == 3 and isinstance(module, torch.nn.LSTM) and module.num_layers == 1 and not module.bidirectional
            and module.batch_first and getattr(module, "proj_size", 0) == 0 and module.input_size % 8 == 0 and ops.lstm_supported(module.hidden_size))


def lstm(module, x):
    """``module(x)[0]`` for a supported ``nn.LSTM`` (see :func:`lstm_supported`), through the library; parameters are read from ``module`` so
    checkpoints and optimisers see the reference's names (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0)."""
    b_ih = getattr(module, "bias_ih_l0", None) if module.bias else None
    b_hh = getattr(module, "bias_hh_l0", None) if module.bias else None
    return _LstmFn.apply(x, module.weight_ih_l0, module.weight_hh_l0, b_ih, b_hh, torch.is_grad_enabled())


# ----------------------------------------------------------------------------- stand-alone pieces (direct core calls)
class _SdpaFn(torch.autograd.Function):
    """The attention core as a differentiable function of q, k, v, the geometry weights AND through both outputs: a direct caller may
    consume the probabilities (``MMF_LoRRA`` builds its features from them, mmf_lorra.py:127-146), so ``att`` carries a gradient too:
    dP = dout v^T + datt inside the same backward kernels."""

    @staticmethod
    def forward(ctx, q, k, v, mask, boxes, geo_w, geo_b, cfg):
        ctx.set_materialize_grads(False)
        h, dk, dv = cfg["h"], cfg["d_k"], cfg["d_v"]
        need_grad = cfg["grad"] and any(ctx.needs_input_grad)
        geo = ops.GeoParams(boxes, geo_w.detach(), geo_b.detach(), cfg["geo_trig"]) if boxes is not None else None
        out, lse, att = ops.sdpa_fwd(q, k, v, h=h, d_k=dk, d_v=dv, mask=mask, geo=geo, scale=cfg["scale"],
                                     need_lse=need_grad, need_att=cfg["need_att"])
        if need_grad:
            ctx.save_for_backward(q, k, v, mask, out, lse, att, boxes, geo_w, geo_b)
            ctx.cfg = cfg
        return out, (att if att is not None else torch.empty(0, device=q.device))

    @staticmethod
    def backward(ctx, dout, datt):
        if dout is None and datt is None:
            return (None,) * 8
        q, k, v, mask, out, lse, att, boxes, geo_w, geo_b = ctx.saved_tensors
        cfg = ctx.cfg
        if att is None:
            datt = None
        dout = torch.zeros_like(out) if dout is None else dout.to(out.dtype).contiguous()
        dq, dk_, dv_ = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        geo = ops.GeoParams(boxes, geo_w.detach(), geo_b.detach(), cfg["geo_trig"]) if boxes is not None else None
        want_geo = geo is not None and (ctx.needs_input_grad[5] or ctx.needs_input_grad[6])
        gdw = torch.zeros_like(geo_w, dtype=torch.float32) if want_geo else None
        gdb = torch.zeros_like(geo_b, dtype=torch.float32) if want_geo else None
        ops.sdpa_bwd(q, k, v, out, dout, lse, dq, dk_, dv_, h=cfg["h"], d_k=cfg["d_k"], d_v=cfg["d_v"], mask=mask, geo=geo,
                     geo_dw=gdw, geo_db=gdb, scale=cfg["scale"], att=att, datt=datt)
        return dq, dk_, dv_, None, None, gdw, gdb, None


def sdpa(q, k, v, mask, *, h, d_k, d_v, scale=None, need_att=False, boxes=None, geo_w=None, geo_b=None, geo_trig=True):
    """Attention core on already projected (B,L,h*d) tensors (fp32 or bf16).  Returns (out, att|None).  Gradients flow to q, k, v and to
    the geometry weights ``geo_w`` (h,d_g) / ``geo_b`` (h) — through ``out`` and, when requested, through the probabilities ``att``."""
    cfg = dict(grad=torch.is_grad_enabled(), h=h, d_k=d_k, d_v=d_v, scale=(1.0 / math.sqrt(d_k)) if scale is None else float(scale), need_att=need_att,
               geo_trig=bool(geo_trig))
    out, att = _SdpaFn.apply(q.contiguous(), k.contiguous(), v.contiguous(), mask, boxes, geo_w, geo_b, cfg)
    return out, (att if need_att else None)


class _PtrScoresFn(torch.autograd.Function):
    """Pointer-network scores ``q k^T * scale + mask`` (mmf_m4c.py:394-396 / m4c.py:27-33) with their backward."""

    @staticmethod
    def forward(ctx, q, k, mask, scale, mask_is_bool):
        scores = torch.empty((q.shape[0], q.shape[1], k.shape[1]), device=q.device, dtype=torch.float32)
        ops.ptr_scores(q, k, mask, scores, 0, scale, mask_is_bool=mask_is_bool)
        ctx.save_for_backward(q, k, mask if mask_is_bool else None)
        ctx.scale = scale
        return scores

    @staticmethod
    def backward(ctx, ds):
        q, k, bool_mask = ctx.saved_tensors
        dq, dk = ops.ptr_scores_bwd(ds.float(), q, k, ctx.scale, bool_mask=bool_mask)
        return dq, dk, None, None, None


def ptr_scores(q, k, mask, scale, mask_is_bool=False):
    """Differentiable pointer scores (B,T,L) fp32 of projected q (B,T,d) and k (B,L,d)."""
    return _PtrScoresFn.apply(q, k, mask, float(scale), bool(mask_is_bool))


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, post_add, grad_on):
        need_grad = grad_on and any(ctx.needs_input_grad)
        y, _, _, mean, rstd = ops.layernorm_fwd(x.float(), gamma.detach(), beta.detach(), eps=eps, post_add=post_add, save_stats=need_grad)
        if need_grad:
            ctx.save_for_backward(x, mean, rstd, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, gamma = ctx.saved_tensors
        dx, _, dg, db = ops.layernorm_bwd(dy, x.float().contiguous(), mean, rstd, gamma)
        return dx, dg, db, None, None, None


def layer_norm(x, gamma, beta, eps=1e-5, post_add=None):
    """``LN(x) [+ post_add]`` (e.g. the encoders' ``layer_norm(features) + pos_embedding``, encoders.py:113)."""
    return _LayerNormFn.apply(x, gamma, beta, float(eps), post_add, torch.is_grad_enabled())
