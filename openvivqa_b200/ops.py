"""Thin PyTorch-facing wrappers over the C ABI: shape/dtype/device validation, then one library call.

Every function enqueues on ``torch.cuda.current_stream()`` and returns device tensors.  CPU tensors are an
error (SURVEY.md §8b "Errors": no CPU fallback).
"""
from __future__ import annotations

import os

import math
from typing import Optional

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("openvivqa_b200 kernels need CUDA tensors (no CPU fallback); got a CPU tensor")


def _rowmajor2d(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() != 2:
        raise ValueError(f"{name}: expected a 2-D tensor, got {tuple(t.shape)}")
    if t.stride(1) != 1 and t.shape[1] != 1:
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def set_sm_reserve(sms: int) -> int:
    """Leave ``sms`` SMs free of the library's persistent kernels (for a co-running collective); returns the previous reserve."""
    return int(L.lib().ovqa_set_sm_reserve(int(sms)))


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_op: str = "N", b_op: str = "N", out: Optional[torch.Tensor] = None,
         out_dtype: Optional[torch.dtype] = None, bias: Optional[torch.Tensor] = None, alpha: float = 1.0,
         beta: float = 0.0, gelu: bool = False, dgelu_aux: Optional[torch.Tensor] = None,
         aux_out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         c2: Optional[torch.Tensor] = None, dropout_p: float = 0.0, seed: int = 0, offset: int = 0,
         offset_base: Optional[torch.Tensor] = None, split_k: int = 1, save_dgelu: bool = False,
         mul_aux: Optional[torch.Tensor] = None, colsum: Optional[torch.Tensor] = None, relu: bool = False) -> torch.Tensor:
    """``C[m,n] = epilogue(sum_k A[m,k] B[n,k])`` — see ``ovqa_gemm`` in include/ovqa.h.

    ``a_op``/``b_op``: "N" = operand stored [rows, k]; "T" = stored [k, rows].
    ``save_dgelu`` (with ``gelu`` + ``aux_out``): ``aux_out`` receives the derivative of the whole fused activation,
    GELU'(pre-activation) times the element's dropout scale (0 or 1/(1-p); 1 without dropout); the backward GEMM multiplies
    by it with ``mul_aux=`` — no transcendental and no RNG in the backward epilogue.
    ``colsum``: fp32 [n], += the column sums of the result (zero-filled by the caller): the bias gradient when the result is the
    gradient of a Linear's output, at no extra pass over the rows."""
    if mul_aux is not None:
        if dgelu_aux is not None:
            raise ValueError("gemm: dgelu_aux and mul_aux are exclusive")
        dgelu_aux = mul_aux
    _need_cuda(a, b, out, bias, dgelu_aux, aux_out, residual, c2, colsum)
    a, b = _rowmajor2d(a, "a"), _rowmajor2d(b, "b")
    if a.dtype != b.dtype or a.dtype not in _DT:
        raise ValueError(f"gemm: operand dtypes must match and be fp32/bf16, got {a.dtype}, {b.dtype}")
    m, k = (a.shape if a_op == "N" else (a.shape[1], a.shape[0]))
    n, kb = (b.shape if b_op == "N" else (b.shape[1], b.shape[0]))
    if k != kb:
        raise ValueError(f"gemm: reduction sizes differ ({k} vs {kb})")
    if out is None:
        out = torch.empty((m, n), device=a.device, dtype=out_dtype or a.dtype) if split_k == 1 and beta == 0.0 \
            else torch.zeros((m, n), device=a.device, dtype=out_dtype or torch.float32)
    if out.shape != (m, n) or out.dtype not in _DT or (out.stride(1) != 1 and n != 1):
        raise ValueError("gemm: bad `out`")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != n or not bias.is_contiguous()):
        raise ValueError("gemm: bias must be contiguous fp32 [n]")
    for t, nm in ((residual, "residual"),):
        if t is not None and (t.dtype != torch.float32 or t.shape != (m, n) or t.stride(1) != 1):
            raise ValueError(f"gemm: {nm} must be fp32 [m,n] with unit column stride")
    for t, nm in ((dgelu_aux, "dgelu_aux"), (aux_out, "aux_out")):
        if t is not None and (t.dtype != out.dtype or t.shape != (m, n) or t.stride(1) != 1):
            raise ValueError(f"gemm: {nm} must have out's dtype and shape")
    if dgelu_aux is not None and aux_out is not None and _ld(dgelu_aux) != _ld(aux_out):
        raise ValueError("gemm: aux_in/aux_out must share a leading dimension")
    if c2 is not None and (c2.dtype != torch.bfloat16 or c2.shape != (m, n) or c2.stride(1) != 1):
        raise ValueError("gemm: c2 must be bf16 [m,n]")
    if colsum is not None and (colsum.dtype != torch.float32 or colsum.numel() != n or not colsum.is_contiguous()):
        raise ValueError("gemm: colsum must be contiguous fp32 [n]")
    flags = (L.EPI_GELU if gelu else 0) | ((L.EPI_MUL_AUX if mul_aux is not None else L.EPI_DGELU) if dgelu_aux is not None else 0) | \
        (L.EPI_SAVE_DGELU if save_dgelu else 0) | (L.EPI_RELU if relu else 0)
    aux = dgelu_aux if dgelu_aux is not None else aux_out
    if m == 0 or n == 0:                     # empty batch: nothing to launch (an empty tensor has no device pointer to validate)
        return out
    if k == 0 and flags == 0 and residual is None and c2 is None and colsum is None:
        # empty reduction (a weight gradient over zero rows): the product is exactly zero
        out.zero_() if beta == 0.0 else out.mul_(beta)
        return out if bias is None else out.add_(bias.to(out.dtype))
    L.call("ovqa_gemm", _p(a), _p(b), _p(out), m, n, k, _ld(a), _ld(b), _ld(out),
           L.OP_N if a_op == "N" else L.OP_T, L.OP_N if b_op == "N" else L.OP_T, _DT[a.dtype], _DT[out.dtype],
           float(alpha), float(beta), _p(bias), flags, _p(dgelu_aux), _p(aux_out), _ld(aux) if aux is not None else 0,
           _p(residual), _ld(residual) if residual is not None else 0, _p(c2), _ld(c2) if c2 is not None else 0,
           float(dropout_p), int(seed), int(offset), _p(offset_base), _p(colsum), int(split_k), _stream())
    return out


def gemm_bf16_simt(a, b, *, a_op="N", b_op="N", out_dtype=torch.float32, bias=None, alpha=1.0, gelu=False):
    """CUDA-core twin of the tcgen05 path (tests only)."""
    _need_cuda(a, b)
    a, b = _rowmajor2d(a, "a"), _rowmajor2d(b, "b")
    m, k = (a.shape if a_op == "N" else (a.shape[1], a.shape[0]))
    n = b.shape[0] if b_op == "N" else b.shape[1]
    out = torch.empty((m, n), device=a.device, dtype=out_dtype)
    L.call("ovqa_gemm_bf16_simt", _p(a), _p(b), _p(out), m, n, k, _ld(a), _ld(b), _ld(out),
           L.OP_N if a_op == "N" else L.OP_T, L.OP_N if b_op == "N" else L.OP_T, _DT[out_dtype], float(alpha), _p(bias),
           L.EPI_GELU if gelu else 0, _stream())
    return out


def _mask_strides(mask: Optional[torch.Tensor], B: int, h: int, Lq: int, Lk: int):
    """Additive fp32 mask broadcastable to (B,h,Lq,Lk) -> (tensor, sb, sh, sq) with unit key stride."""
    if mask is None:
        return None, 0, 0, 0
    if mask.dtype != torch.float32:
        mask = mask.float()
    while mask.dim() < 4:
        mask = mask.unsqueeze(0)
    if mask.dim() != 4 or mask.shape[-1] != Lk:
        raise ValueError(f"attention_mask {tuple(mask.shape)} is not broadcastable to (B,h,Lq,{Lk})")
    for d, full in zip(range(3), (B, h, Lq)):
        if mask.shape[d] not in (1, full):
            raise ValueError(f"attention_mask {tuple(mask.shape)} is not broadcastable to ({B},{h},{Lq},{Lk})")
    if mask.stride(-1) != 1 and Lk > 1:
        mask = mask.contiguous()
    st = [0 if mask.shape[d] == 1 else mask.stride(d) for d in range(3)]
    return mask, st[0], st[1], st[2]


class GeoParams:
    """Box-geometry arguments of the fused core: boxes (B,N,4) fp32, w (h,d_g), b (h), freq table."""

    def __init__(self, boxes, w, b, trig: bool):
        self.boxes = boxes.float().contiguous()
        self.w = w.float().contiguous()
        self.b = b.float().contiguous()
        self.trig = bool(trig)
        self.d_g = self.w.shape[1]
        if self.trig:
            # models/utils.py:146-148, computed with the same torch ops so the table is bit-identical
            fr = torch.arange(self.d_g / 8, device=self.boxes.device)
            self.freq = (1.0 / torch.pow(1000, fr / (self.d_g / 8))).float().contiguous()
        else:
            self.freq = None


def sdpa_fwd(q, k, v, *, h: int, d_k: int, d_v: int, mask=None, geo: Optional[GeoParams] = None,
             scale: Optional[float] = None, need_lse: bool = False, need_att: bool = False, out=None, prob_dropout=None):
    """Attention core on projected tensors.  q:(B,Lq,>=h*d_k) k:(B,Lk,..) v:(B,Lk,..) — last-dim-contiguous views
    (e.g. slices of a fused QKV buffer).  Returns (out (B,Lq,h*d_v), lse|None, att|None)."""
    _need_cuda(q, k, v, mask)
    if q.dtype not in _DT or k.dtype != q.dtype or v.dtype != q.dtype:
        raise ValueError("sdpa: q,k,v must share a dtype in {fp32, bf16}")
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    for t, L_, nm in ((q, Lq, "q"), (k, Lk, "k"), (v, Lk, "v")):
        if t.dim() != 3 or t.stride(2) != 1 or (t.shape[0] > 1 and t.stride(0) != L_ * t.stride(1)):
            raise ValueError(f"sdpa: {nm} must be (B,L,*) with dense batches and unit last stride")
    if out is None:
        out = torch.empty((B, Lq, h * d_v), device=q.device, dtype=q.dtype)
    lse = torch.empty((B, h, Lq), device=q.device, dtype=torch.float32) if need_lse else None
    att = torch.empty((B, h, Lq, Lk), device=q.device, dtype=torch.float32) if need_att else None
    if scale is None:
        scale = 1.0 / math.sqrt(d_k)
    if B == 0 or Lq == 0:
        return out, lse, att
    # prefix-LM structure hint (harness.m4c.mmt_mask): per-key row + causal tail instead of the dense (B,1,L,L) tensor
    pd_p, pd_seed, pd_off, pd_base = prob_dropout if prob_dropout is not None else (0.0, 0, 0, None)
    hint = getattr(mask, "_ovqa_prefix_lm", None)
    if hint is not None and q.dtype == torch.bfloat16 and geo is None and not need_att and pd_p == 0.0:
        key_row, tail, value = hint
        rc = L.lib().ovqa_sdpa_fwd_prefix_lm(_p(q), _p(k), _p(v), _p(out), B, h, Lq, Lk, d_k, d_v, q.stride(1), k.stride(1), v.stride(1),
                                             out.stride(1), _p(key_row), key_row.stride(0) if key_row.shape[0] > 1 else 0, int(tail),
                                             float(value), float(scale), _DT[q.dtype], _p(lse), _stream())
        if rc == 0:
            return out, lse, att
        if rc != L.ERR_BAD_SHAPE:
            L.check(rc, "ovqa_sdpa_fwd_prefix_lm")                 # (BAD_SHAPE: this shape has no structured kernel -> dense mask below)
    mask, sb, sh, sq = _mask_strides(mask, B, h, Lq, Lk)
    g = geo
    L.call("ovqa_sdpa_fwd", _p(q), _p(k), _p(v), _p(out), B, h, Lq, Lk, d_k, d_v,
           q.stride(1), k.stride(1), v.stride(1), out.stride(1), _p(mask), sb, sh, sq,
           _p(g.boxes) if g else None, _p(g.w) if g else None, _p(g.b) if g else None,
           _p(g.freq) if (g and g.freq is not None) else None, g.d_g if g else 0, int(g.trig) if g else 0,
           float(scale), _DT[q.dtype], _p(lse), _p(att), float(pd_p), int(pd_seed), int(pd_off), _p(pd_base), _stream())
    return out, lse, att


def sdpa_bwd(q, k, v, out, dout, lse, dq, dk, dv, *, h: int, d_k: int, d_v: int, mask=None,
             geo: Optional[GeoParams] = None, geo_dw=None, geo_db=None, scale: Optional[float] = None, att=None, datt=None, dbias=None,
             prob_dropout=None):
    """``att`` / ``datt`` (B,h,Lq,Lk) fp32: the probabilities the forward returned and the gradient arriving through them (a direct
    caller of the core consumed ``att``, mmf_lorra.py:127-146); both or neither."""
    _need_cuda(q, k, v, out, dout, lse, dq, dk, dv, att, datt, dbias)
    if dbias is not None and (dbias.dtype != torch.float32 or not dbias.is_contiguous() or dbias.shape != (q.shape[0], h, q.shape[1], k.shape[1])):
        raise ValueError("sdpa_bwd: dbias must be a contiguous fp32 (B,h,Lq,Lk) tensor")
    if datt is not None:
        if att is None or att.shape != datt.shape or att.dtype != torch.float32:
            raise ValueError("sdpa_bwd: datt needs the saved fp32 probabilities `att` of the same shape")
        att, datt = att.contiguous(), datt.float().contiguous()
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    mask, sb, sh, sq = _mask_strides(mask, B, h, Lq, Lk)
    if dout.stride(2) != 1:
        dout = dout.contiguous()
    delta = torch.empty((B, h, Lq), device=q.device, dtype=torch.float32)
    if scale is None:
        scale = 1.0 / math.sqrt(d_k)
    g = geo
    if B == 0 or Lq == 0:                    # no query rows: every gradient is zero
        for t in (dq, dk, dv):
            t.zero_()
        return
    L.call("ovqa_sdpa_bwd", _p(q), _p(k), _p(v), _p(out), _p(dout), _p(dq), _p(dk), _p(dv), B, h, Lq, Lk, d_k, d_v,
           q.stride(1), k.stride(1), v.stride(1), out.stride(1), dout.stride(1), dq.stride(1), dk.stride(1), dv.stride(1),
           _p(mask), sb, sh, sq,
           _p(g.boxes) if g else None, _p(g.w) if g else None, _p(g.b) if g else None,
           _p(g.freq) if (g and g.freq is not None) else None, g.d_g if g else 0, int(g.trig) if g else 0,
           _p(geo_dw), _p(geo_db), float(scale), _DT[q.dtype], _p(lse), _p(delta), _p(att) if datt is not None else None, _p(datt), _p(dbias),
           *((float(prob_dropout[0]), int(prob_dropout[1]), int(prob_dropout[2]), _p(prob_dropout[3])) if prob_dropout is not None else (0.0, 0, 0, None)), _stream())


def layernorm_fwd(x, gamma, beta, *, residual=None, post_add=None, eps: float = 1e-5, want_f32: bool = True,
                  want_bf16: bool = False, save_stats: bool = False, save_sum: bool = False):
    """y = LN(x [+ residual]) [+ post_add];  returns (y_f32|None, y_bf16|None, sum|None, mean|None, rstd|None)."""
    _need_cuda(x, gamma, beta, residual, post_add)
    D = x.shape[-1]
    x2 = x.reshape(-1, D)
    if x2.dtype != torch.float32:
        raise ValueError("layernorm: x must be fp32")
    x2 = x2.contiguous()
    rows = x2.shape[0]
    r2 = None if residual is None else residual.reshape(-1, D).contiguous()
    pa, pa_rows = None, 0
    if post_add is not None:
        pa = post_add.reshape(-1, D).contiguous().float()
        pa_rows = pa.shape[0]
        if rows % pa_rows != 0:
            raise ValueError("layernorm: post_add rows must divide the row count")
    y = torch.empty_like(x2) if want_f32 else None
    y16 = torch.empty(x2.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    s = torch.empty_like(x2) if save_sum else None
    mean = torch.empty((rows,), device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty((rows,), device=x.device, dtype=torch.float32) if save_stats else None
    if rows > 0:
        L.call("ovqa_layernorm_fwd", _p(x2), _p(r2), _p(gamma), _p(beta), _p(pa), pa_rows, _p(y), _p(y16), _p(s), _p(mean),
               _p(rstd), rows, D, float(eps), _stream())
    shp = x.shape
    return (None if y is None else y.view(shp), None if y16 is None else y16.view(shp),
            None if s is None else s.view(shp), mean, rstd)


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dx_add=None, want_bf16: bool = False, dropout_p: float = 0.0,
                  seed: int = 0, offset: int = 0, offset_base=None, dgamma=None, dbeta=None, want_dbias: bool = False, dbias=None):
    """Returns (dx, dx_bf16|None, dgamma, dbeta[, dbias]).  ``want_dbias``: also the column sums of the (dropout-masked) dx —
    the bias gradient of the GEMM that produced this LayerNorm's input.  ``dgamma`` / ``dbeta`` / ``dbias``, when given, are
    ZERO-FILLED fp32 [D] accumulators owned by the caller (e.g. views of one slab zeroed by a single fill)."""
    _need_cuda(dy, x, mean, rstd, gamma)
    D = x.shape[-1]
    dy2, x2 = dy.reshape(-1, D).contiguous().float(), x.reshape(-1, D)
    rows = x2.shape[0]
    dx = torch.empty_like(x2)
    dx16 = torch.empty(x2.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    dbias_given = dbias
    dbias = None
    if dgamma is None and dbeta is None:
        acc = torch.zeros((3 if want_dbias else 2, D), device=x.device, dtype=torch.float32)     # one fill for all accumulators
        dgamma, dbeta = acc[0], acc[1]
        dbias = acc[2] if want_dbias else None
    else:
        if dgamma is None:
            dgamma = torch.zeros((D,), device=x.device, dtype=torch.float32)
        if dbeta is None:
            dbeta = torch.zeros((D,), device=x.device, dtype=torch.float32)
        dbias = (dbias_given if dbias_given is not None else torch.zeros((D,), device=x.device, dtype=torch.float32)) if want_dbias else None
    if rows > 0:
        L.call("ovqa_layernorm_bwd", _p(dy2), _p(x2), _p(mean), _p(rstd), _p(gamma),
               None if dx_add is None else _p(dx_add.reshape(-1, D).contiguous()), _p(dx), _p(dx16), _p(dgamma), _p(dbeta), _p(dbias),
               rows, D, float(dropout_p), int(seed), int(offset), _p(offset_base), _stream())
    out = (dx.view(x.shape), (None if dx16 is None else dx16.view(x.shape)), dgamma, dbeta)
    return out + (dbias,) if want_dbias else out


def cast_rows_padmask(x: torch.Tensor, mask_value: float):
    """fp32 (rows, K) -> (bf16 copy (rows, K), additive padding mask (rows,) fp32: ``mask_value`` where the row sums to 0) in one pass."""
    _need_cuda(x)
    x = _rowmajor2d(x, "x")
    if x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0:
        # features staged as bf16 (input feed in pinned bf16 host buffers): they ARE the GEMM operand, only the mask is computed
        x = x if _ld(x) % 8 == 0 else x.contiguous()
        mask = torch.empty((x.shape[0],), device=x.device, dtype=torch.float32)
        if x.shape[0]:
            L.call("ovqa_rows_padmask_bf16", _p(x), _p(mask), x.shape[0], x.shape[1], _ld(x), float(mask_value), _stream())
        return x, mask
    if x.dtype != torch.float32 or x.shape[1] % 8 != 0:
        raise ValueError("cast_rows_padmask: x must be fp32 (or bf16) with a width that is a multiple of 8")
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    mask = torch.empty((x.shape[0],), device=x.device, dtype=torch.float32)
    if x.shape[0] == 0:
        return y, mask
    L.call("ovqa_cast_rows_padmask", _p(x), _p(y), _p(mask), x.shape[0], x.shape[1], _ld(x), float(mask_value), _stream())
    return y, mask


def mul_to_bf16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16(a * b), a fp32, b fp32 or bf16 (same shape, contiguous)."""
    _need_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    if a.dtype != torch.float32 or b.dtype not in _DT or a.shape != b.shape:
        raise ValueError("mul_to_bf16: a fp32, b fp32/bf16, equal shapes")
    y = torch.empty(a.shape, device=a.device, dtype=torch.bfloat16)
    if a.numel() == 0:
        return y
    L.call("ovqa_mul_to_bf16", _p(a), _p(b), _DT[b.dtype], _p(y), a.numel(), _stream())
    return y


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    if x.dtype == torch.bfloat16:
        return x
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    if x.numel() == 0:
        return y
    L.call("ovqa_cast_f32_to_bf16", _p(x), _p(y), x.numel(), _stream())
    return y


def to_f32(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    if x.dtype == torch.float32:
        return x
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    if x.numel() == 0:
        return y
    L.call("ovqa_cast_bf16_to_f32", _p(x), _p(y), x.numel(), _stream())
    return y


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    _need_cuda(x)
    x = _rowmajor2d(x, "x")
    if out is None:
        out = torch.empty((x.shape[1],), device=x.device, dtype=torch.float32)
        accumulate = False
    if x.shape[0] == 0 or x.shape[1] == 0:
        return out if accumulate else out.zero_()
    L.call("ovqa_colsum", _p(x), _p(out), x.shape[0], x.shape[1], _ld(x), _DT[x.dtype], int(accumulate), _stream())
    return out


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _need_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    y = torch.empty_like(a)
    if a.numel() == 0:
        return y
    L.call("ovqa_add_f32", _p(a), _p(b), _p(y), a.numel(), _stream())
    return y


def dropout_apply(x: torch.Tensor, p: float, seed: int, offset: int, out_dtype=torch.float32, offset_base=None) -> torch.Tensor:
    """y = dropout(x) with the library's counter-based mask over the flattened index space (p=0: cast)."""
    _need_cuda(x)
    x = x.contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    if x.numel() == 0:
        return y
    L.call("ovqa_dropout_apply", _p(x), _p(y), x.numel(), float(p), int(seed), int(offset), _p(offset_base), _DT[out_dtype], _stream())
    return y


def dropout_keep_mask(n: int, p: float, seed: int, offset: int, device, offset_base=None) -> torch.Tensor:
    out = torch.empty((n,), device=device, dtype=torch.float32)
    if n == 0:
        return out
    L.call("ovqa_dropout_mask", _p(out), n, float(p), int(seed), int(offset), _p(offset_base), _stream())
    return out


def stage_many(table: torch.Tensor, n_desc: int):
    """One launch for ``n_desc`` fp32 -> {fp32, bf16} copies; ``table`` is a device int64 tensor (n_desc, 4) laid out as
    ``ovqa_stage_desc`` {src ptr, dst ptr, n, dst_dtype | reserved<<32}."""
    _need_cuda(table)
    if table.dtype != torch.int64 or not table.is_contiguous() or table.numel() < 4 * n_desc:
        raise ValueError("stage_many: table must be a contiguous int64 tensor of shape (n_desc, 4)")
    L.call("ovqa_stage_many", _p(table), int(n_desc), _stream())


def advance_u64(word: torch.Tensor, delta: int):
    """``word[0] += delta`` on the device (word: int64 tensor of one element; the library treats it as u64)."""
    _need_cuda(word)
    L.call("ovqa_advance_u64", _p(word), int(delta), _stream())


def box_geometry_bias(geo: GeoParams, h: int) -> torch.Tensor:
    B, N = geo.boxes.shape[:2]
    out = torch.empty((B, h, N, N), device=geo.boxes.device, dtype=torch.float32)
    L.call("ovqa_box_geometry_bias", _p(geo.boxes), _p(geo.w), _p(geo.b), _p(geo.freq) if geo.freq is not None else None,
           _p(out), B, N, h, geo.d_g, int(geo.trig), _stream())
    return out


def _dense_rows(t: torch.Tensor) -> torch.Tensor:
    """(B, L, d) with unit last stride and batch stride L * row pitch (what the pointer kernels address), else a contiguous copy."""
    if t.stride(2) != 1 or (t.shape[0] > 1 and t.stride(0) != t.shape[1] * t.stride(1)):
        return t.contiguous()
    return t


def ptr_scores(q, k, mask, scores, col_offset: int, scale: float, mask_is_bool: bool = False):
    """Writes ``scores[:, :, col_offset:col_offset+L]`` in place (scores: fp32 (B,T,W) contiguous)."""
    _need_cuda(q, k, mask, scores)
    B, T, d = q.shape
    Lk = k.shape[1]
    q, k = _dense_rows(q), _dense_rows(k)
    if not scores.is_contiguous() or scores.dtype != torch.float32:
        raise ValueError("ptr_scores: bad layout")
    if mask is not None:
        mask = mask.reshape(B, Lk).contiguous()
        mask = mask.to(torch.uint8) if mask_is_bool else mask.float()
    L.call("ovqa_ptr_scores", _p(q), _p(k), _p(mask), _p(scores), B, T, Lk, d, q.stride(1), k.stride(1), scores.shape[-1],
           col_offset, float(scale), _DT[q.dtype], int(mask_is_bool), _stream())
    return scores


def gemm_ln_supported(a: torch.Tensor, w: torch.Tensor) -> bool:
    """Shapes the fused projection + LayerNorm kernel takes (``ovqa_gemm_ln``): bf16 CUDA operands, 512 output columns, 16-byte aligned rows."""
    return (a.is_cuda and a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2 and w.shape[0] == 512
            and a.shape[1] == w.shape[1] and a.stride(1) == 1 and w.stride(1) == 1 and a.stride(0) % 8 == 0 and w.stride(0) % 8 == 0
            and a.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0)


def gemm_ln_enabled() -> bool:
    """Whether the blocks route their fc_o / fc2 + LayerNorm tail through :func:`gemm_ln`.  OFF by default: the kernel is parity-green
    (tests/test_gpu_gemm_ln.py) but measured SLOWER than the two-kernel form it would replace (MCAN step 11.00 vs 10.65 ms: its epilogue's
    per-chunk TMA round trips are serialised and, with all of tensor memory taken by one accumulator, do not overlap the next tile's main
    loop — DESIGN.md §9).  ``OVQA_GEMM_LN=1`` turns it on."""
    return os.environ.get("OVQA_GEMM_LN", "0") == "1"


def gemm_ln(a, w, bias, residual, gamma, beta, *, eps: float = 1e-5, dropout_p: float = 0.0, seed: int = 0, offset: int = 0, offset_base=None,
            save: bool = True, want_bf16: bool = True):
    """``s = dropout(a w^T + bias) + residual;  y = LayerNorm(s) * gamma + beta`` in one kernel (N = 512).  Returns (s | None, y, y16 | None,
    mean | None, rstd | None): ``s``, ``mean``, ``rstd`` (what :func:`layernorm_bwd` needs) only when ``save``."""
    _need_cuda(a, w, bias, residual, gamma, beta)
    m, n = a.shape[0], w.shape[0]
    if residual.dtype != torch.float32 or tuple(residual.shape) != (m, n) or not residual.is_contiguous():
        raise ValueError("gemm_ln: residual must be contiguous fp32 [m, n]")
    dev = a.device
    y = torch.empty((m, n), device=dev, dtype=torch.float32)
    y16 = torch.empty((m, n), device=dev, dtype=torch.bfloat16) if want_bf16 else None
    s = torch.empty((m, n), device=dev, dtype=torch.float32) if save else None
    mean = torch.empty((m,), device=dev, dtype=torch.float32) if save else None
    rstd = torch.empty((m,), device=dev, dtype=torch.float32) if save else None
    if m > 0:
        L.call("ovqa_gemm_ln", _p(a), _p(w), _p(bias), _p(residual), _p(gamma), _p(beta), _p(s), _p(y), _p(y16), _p(mean), _p(rstd), m, n, a.shape[1],
               a.stride(0), w.stride(0), float(eps), float(dropout_p), int(seed), int(offset), _p(offset_base), _stream())
    return s, y, y16, mean, rstd


def lstm_supported(H: int) -> bool:
    """Hidden sizes the cluster-resident LSTM kernels are built for (``ovqa_lstm_supported``)."""
    return bool(L.lib().ovqa_lstm_supported(int(H)))


def lstm_fwd(xg: torch.Tensor, w_hh16: torch.Tensor, b_hh: Optional[torch.Tensor], save: bool = True):
    """Recurrence of a single-layer batch-first LSTM with zero initial state (text_embeddings.py:234,244).  ``xg`` (B,L,4H) fp32 = the input
    projection of every step (x W_ih^T + b_ih), ``w_hh16`` (4H,H) bf16, ``b_hh`` (4H) fp32 | None.  Returns (y (B,L,H) fp32, y16 (B,L,H) bf16,
    hprev16 | None, gates16 | None, c | None): the last three are what :func:`lstm_bwd` and the weight-gradient GEMMs need."""
    _need_cuda(xg, w_hh16, b_hh)
    B, T, G = xg.shape
    H = G // 4
    if xg.dtype != torch.float32 or not xg.is_contiguous() or w_hh16.dtype != torch.bfloat16 or tuple(w_hh16.shape) != (4 * H, H) or not w_hh16.is_contiguous():
        raise ValueError("lstm_fwd: xg (B,L,4H) fp32 contiguous, w_hh (4H,H) bf16 contiguous")
    if b_hh is not None and (b_hh.dtype != torch.float32 or b_hh.numel() != 4 * H or not b_hh.is_contiguous()):
        raise ValueError("lstm_fwd: b_hh must be contiguous fp32 [4H]")
    dev = xg.device
    y = torch.empty((B, T, H), device=dev, dtype=torch.float32)
    y16 = torch.empty((B, T, H), device=dev, dtype=torch.bfloat16)
    hprev16 = torch.empty((B, T, H), device=dev, dtype=torch.bfloat16) if save else None
    gates16 = torch.empty((B, T, 4 * H), device=dev, dtype=torch.bfloat16) if save else None
    c = torch.empty((B, T, H), device=dev, dtype=torch.float32) if save else None
    if B > 0 and T > 0:
        L.call("ovqa_lstm_fwd", _p(xg), _p(w_hh16), _p(b_hh), _p(y), _p(y16), _p(hprev16), _p(gates16), _p(c), B, T, H, _stream())
    return y, y16, hprev16, gates16, c


def lstm_bwd(dy: torch.Tensor, w_hh16: torch.Tensor, gates16: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """Backward of the recurrence: ``dy`` (B,L,H) fp32 -> gradient of the gate pre-activations (B,L,4H) bf16."""
    _need_cuda(dy, w_hh16, gates16, c)
    B, T, H = dy.shape
    if dy.dtype != torch.float32 or not dy.is_contiguous() or gates16.shape != (B, T, 4 * H) or c.shape != (B, T, H):
        raise ValueError("lstm_bwd: dy (B,L,H) fp32 contiguous; gates (B,L,4H) bf16 and c (B,L,H) fp32 as saved by lstm_fwd")
    dgates16 = torch.empty((B, T, 4 * H), device=dy.device, dtype=torch.bfloat16)
    if B > 0 and T > 0:
        L.call("ovqa_lstm_bwd", _p(dy), _p(w_hh16), _p(gates16), _p(c), _p(dgates16), B, T, H, _stream())
    return dgates16


def attr_pool_fwd(hidden: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor, x: torch.Tensor, save_probs: bool = True):
    """softmax-over-the-sequence pooling of ``x`` (B,N,D) fp32 by the scores ``hidden . w2 + b2`` (hidden (B,N,Dh) fp32 / bf16, the
    attribute-reduce MLP's first layer).  Returns (pooled (B,D) fp32, probs (B,N) fp32 | None)."""
    _need_cuda(hidden, w2, b2, x)
    B, N, Dh = hidden.shape
    D = x.shape[-1]
    if hidden.dtype not in _DT or x.dtype != torch.float32 or x.shape[:2] != (B, N) or not hidden.is_contiguous() or not x.is_contiguous():
        raise ValueError("attr_pool_fwd: hidden (B,N,Dh) fp32/bf16 and x (B,N,D) fp32, both contiguous")
    w2, b2 = w2.reshape(-1).float().contiguous(), b2.reshape(-1).float().contiguous()
    if w2.numel() != Dh or b2.numel() != 1:
        raise ValueError("attr_pool_fwd: w2 must have Dh elements, b2 one")
    pooled = torch.empty((B, D), device=x.device, dtype=torch.float32)
    probs = torch.empty((B, N), device=x.device, dtype=torch.float32) if save_probs else None
    if B > 0:
        L.call("ovqa_attr_pool_fwd", _p(hidden), _p(w2), _p(b2), _p(x), _p(pooled), _p(probs), B, N, Dh, D, _DT[hidden.dtype], _stream())
    return pooled, probs


def attr_pool_bwd(dpooled, hidden, w2, x, probs, dropout_p: float, dw2: torch.Tensor, db2: torch.Tensor, need_dx: bool = True):
    """Backward of :func:`attr_pool_fwd`: returns (dx (B,N,D) fp32 | None, dpre (B,N,Dh) in hidden's dtype — the gradient of the MLP's fc1
    PRE-activation); accumulates into ``dw2`` (Dh) / ``db2`` (1) fp32."""
    _need_cuda(dpooled, hidden, w2, x, probs, dw2, db2)
    B, N, Dh = hidden.shape
    D = x.shape[-1]
    dpooled = dpooled.float().contiguous()
    w2 = w2.reshape(-1).float().contiguous()
    dx = torch.empty((B, N, D), device=x.device, dtype=torch.float32) if need_dx else None
    dpre = torch.empty_like(hidden)
    if B > 0:
        L.call("ovqa_attr_pool_bwd", _p(dpooled), _p(hidden), _p(w2), _p(x), _p(probs), _p(dx), _p(dpre), _p(dw2), _p(db2), B, N, Dh, D, float(dropout_p),
               _DT[hidden.dtype], _stream())
    return dx, dpre


def geo_bias(geo: GeoParams, h: int, key_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,h,N,N) fp32 additive attention bias ``log(clamp(relu(w_h . emb(box_i, box_j) + b_h), 1e-6)) [+ key_mask[b, j]]`` — the pair
    embedding is evaluated once for all heads.  ``key_mask``: additive per-key mask (B,1,1,N) / (B,N) / (1,..,N) or None."""
    _need_cuda(geo.boxes, key_mask)
    B, N = geo.boxes.shape[:2]
    out = torch.empty((B, h, N, N), device=geo.boxes.device, dtype=torch.float32)
    if B * N == 0:
        return out
    km, sb = None, 0
    if key_mask is not None:
        km = key_mask.reshape(-1, N).float().contiguous()
        if km.shape[0] not in (1, B):
            raise ValueError("geo_bias: key_mask must be a per-key mask broadcastable over heads and queries")
        sb = N if km.shape[0] == B else 0
    L.call("ovqa_geo_bias_fwd", _p(geo.boxes), _p(geo.w), _p(geo.b), _p(geo.freq) if geo.freq is not None else None, _p(km), sb, _p(out),
           B, N, h, geo.d_g, int(geo.trig), _stream())
    return out


def geo_bias_bwd(geo: GeoParams, h: int, dbias: torch.Tensor, geo_dw: torch.Tensor, geo_db: torch.Tensor):
    """Accumulates the geometry weight gradients (geo_dw (h,d_g), geo_db (h), fp32) from ``dbias`` = d loss / d bias (B,h,N,N)."""
    _need_cuda(geo.boxes, dbias, geo_dw, geo_db)
    B, N = geo.boxes.shape[:2]
    if dbias.shape != (B, h, N, N) or dbias.dtype != torch.float32 or not dbias.is_contiguous():
        raise ValueError("geo_bias_bwd: dbias must be contiguous fp32 (B,h,N,N)")
    if B * N == 0:
        return
    L.call("ovqa_geo_bias_bwd", _p(geo.boxes), _p(geo.w), _p(geo.b), _p(geo.freq) if geo.freq is not None else None, _p(dbias), _p(geo_dw), _p(geo_db),
           B, N, h, geo.d_g, int(geo.trig), _stream())


def ptr_scores_bwd(dscores, q, k, scale: float, bool_mask=None):
    """(dq, dk) of ``ptr_scores``: dq = scale * dS k, dk = scale * dS^T q (dscores: fp32 (B,T,L) view with unit last stride; keys a boolean
    mask filled with -inf carry no gradient)."""
    _need_cuda(dscores, q, k, bool_mask)
    B, T, d = q.shape
    Lk = k.shape[1]
    if dscores.dtype != torch.float32 or dscores.shape != (B, T, Lk):
        raise ValueError("ptr_scores_bwd: dscores must be fp32 (B,T,L)")
    if dscores.stride(2) != 1 or (B > 1 and dscores.stride(0) != T * dscores.stride(1)):
        dscores = dscores.contiguous()
    q, k = _dense_rows(q), _dense_rows(k)
    dq, dk = torch.empty((B, T, d), device=q.device, dtype=q.dtype), torch.empty((B, Lk, d), device=k.device, dtype=k.dtype)
    if B * T * Lk == 0:
        return dq.zero_(), dk.zero_()
    if bool_mask is not None:
        bool_mask = bool_mask.reshape(B, Lk).to(torch.uint8).contiguous()
    L.call("ovqa_ptr_scores_bwd", _p(dscores), _p(q), _p(k), _p(bool_mask), _p(dq), _p(dk), B, T, Lk, d, dscores.stride(1), q.stride(1), k.stride(1),
           d, d, float(scale), _DT[q.dtype], _stream())
    return dq, dk


def prev_pred_embed(ans_emb, ocr_emb, prev_inds, pos_emb, type_emb, ans_ln, ocr_ln, emb_ln, eps: float, out=None):
    """ans_ln/ocr_ln/emb_ln: (gamma, beta) pairs.  out: fp32 view (B,T,D) with unit last stride (may be a slice of the
    joint encoder input)."""
    _need_cuda(ans_emb, ocr_emb, prev_inds)
    B, T = prev_inds.shape
    V, D = ans_emb.shape
    Locr = ocr_emb.shape[1]
    if out is None:
        out = torch.empty((B, T, D), device=ans_emb.device, dtype=torch.float32)
    L.call("ovqa_prev_pred_embed", _p(ans_emb.contiguous()), _p(ocr_emb.contiguous()), _p(prev_inds.contiguous()),
           _p(pos_emb), _p(type_emb), _p(ans_ln[0]), _p(ans_ln[1]), _p(ocr_ln[0]), _p(ocr_ln[1]), _p(emb_ln[0]),
           _p(emb_ln[1]), _p(out), B, T, V, Locr, D, out.stride(1), out.stride(0) if B > 1 else T * out.stride(1),
           float(eps), _stream())
    return out


def greedy_step(scores, prev_inds, last_ids, argmax_out, all_eos, step: int, eos_idx: int):
    _need_cuda(scores, prev_inds, last_ids, argmax_out, all_eos)
    B, T, W = scores.shape
    L.call("ovqa_greedy_step", _p(scores.contiguous()), _p(prev_inds), _p(last_ids), _p(argmax_out), _p(all_eos), B, T, W,
           int(step), int(eos_idx), _stream())


# ----------------------------------------------------------------------------- inference blocks enqueued by one call (decode.cu)
_WS = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream): kernels of one stream run in order, so consecutive calls may reuse it."""
    if torch.cuda.is_current_stream_capturing():          # a captured graph keeps its own buffer (its memory pool keeps it alive)
        return torch.empty((nbytes,), device=device, dtype=torch.uint8)
    key = (device, _stream())
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty((max(nbytes, 1 << 20),), device=device, dtype=torch.uint8)
        _WS[key] = ws
    return ws


def _mask_args(mask, B: int, h: int, Lq: int, Lk: int):
    """(tensor kept alive, pointer, sb, sh, sq) of an additive mask; the common (·,·,·,Lk) fp32 layout without building views."""
    if mask is None:
        return None, None, 0, 0, 0
    if mask.dtype is torch.float32 and mask.dim() == 4:
        shp, st = mask.shape, mask.stride()
        if shp[3] == Lk and (st[3] == 1 or Lk == 1) and shp[0] in (1, B) and shp[1] in (1, h) and shp[2] in (1, Lq):
            return mask, mask.data_ptr(), (0 if shp[0] == 1 else st[0]), (0 if shp[1] == 1 else st[1]), (0 if shp[2] == 1 else st[2])
    m, sb, sh, sq = _mask_strides(mask, B, h, Lq, Lk)
    return m, m.data_ptr(), sb, sh, sq


def mha_decode_step(x, x_op, kv_src, self_attention: bool, wptrs, *, dt: int, eps: float, h: int, d_k: int, d_v: int, kv_cache, length: int,
                    mask=None, scale: float):
    """One MultiHeadAttention block of an inference decode step over the projected-K/V cache ``kv_cache`` (B, cap, h*d_k + h*d_v), of which
    ``length`` positions are valid; appends the new rows in place (the caller advances its length).  x (B,Lq,D) fp32 contiguous, ``x_op``
    the same rows in bf16 or None; ``wptrs`` = device pointers (w_qkv, b_qkv, w_o, b_o, gamma, beta) of tensors the caller keeps alive,
    ``dt`` their compute dtype code.  Returns (y fp32, y bf16 | None).  Kept lean on purpose: at beam-search batch sizes the step is
    bounded by host dispatch."""
    B, Lq, D = x.shape
    Lnew = Lq if self_attention else (0 if kv_src is None else kv_src.shape[1])
    cshape = kv_cache.shape
    if not x.is_cuda or not kv_cache.is_cuda or x.dtype is not torch.float32 or not x.is_contiguous() or not kv_cache.is_contiguous() or _DT.get(kv_cache.dtype) != dt:
        raise ValueError("mha_decode_step: x must be a contiguous fp32 CUDA tensor, the cache contiguous in the compute dtype")
    if cshape[0] != B or cshape[2] != h * (d_k + d_v) or length + Lnew > cshape[1]:
        raise ValueError(f"mha_decode_step: cache {tuple(cshape)} cannot take {length}+{Lnew} positions of batch {B}")
    if kv_src is not None and not self_attention and (not kv_src.is_cuda or not kv_src.is_contiguous() or _DT.get(kv_src.dtype) != dt
                                                      or kv_src.shape[0] != B):
        raise ValueError("mha_decode_step: kv_src must be a contiguous CUDA tensor (B, Lnew, D) in the compute dtype")
    _need_cuda(x_op, mask)
    mkeep, mptr, sb, sh, sq = _mask_args(mask, B, h, Lq, length + Lnew)
    y = torch.empty_like(x)
    y16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if dt == L.BF16 else None
    if B == 0 or Lq == 0:
        return y, y16
    l = L.lib()
    st = _stream()
    nws = l.ovqa_mha_decode_workspace_bytes(B, Lq, Lnew, D, h, d_k, d_v, dt)
    ws = _WS.get((x.device, st))
    if ws is None or ws.numel() < nws or torch.cuda.is_current_stream_capturing():
        ws = _workspace(x.device, nws)
    rc = l.ovqa_mha_decode_step(x.data_ptr(), None if x_op is None else x_op.data_ptr(), None if kv_src is None else kv_src.data_ptr(),
                                1 if self_attention else 0, B, Lq, Lnew, D, h, d_k, d_v, wptrs[0], wptrs[1], wptrs[2], wptrs[3], wptrs[4],
                                wptrs[5], eps, kv_cache.data_ptr(), cshape[1], length, mptr, sb, sh, sq, scale, dt, ws.data_ptr(), ws.numel(),
                                y.data_ptr(), None if y16 is None else y16.data_ptr(), st)
    if rc:
        L.check(rc, "ovqa_mha_decode_step")
    return y, y16


def mha_block_infer(xq, xq_op, xk, xk_op, xv, xv_op, kind: int, wptrs, *, dt: int, eps: float, h: int, d_k: int, d_v: int, mask=None, scale: float,
                    max_workspace: int = 1 << 30):
    """One MultiHeadAttention block in eval mode from ONE library call (ovqa_mha_block_infer).  xq (B,Lq,D) fp32 contiguous (queries and
    residual); xk / xv (B,Lk,D) fp32 contiguous for kind 1 / 2; ``*_op``: the same rows in bf16 or None.  ``wptrs`` = device pointers
    (w_qkv, b_qkv, w_o, b_o, gamma, beta).  Returns (y fp32, y bf16 | None), or None when the scratch space would exceed ``max_workspace``
    (huge batches: host dispatch is irrelevant there, the caller takes the per-op path)."""
    B, Lq, D = xq.shape
    Lk = Lq if kind == 0 else xk.shape[1]
    for t in (xq, xk, xv):
        if t is not None and (not t.is_cuda or t.dtype is not torch.float32 or not t.is_contiguous()):
            raise ValueError("mha_block_infer: inputs must be contiguous fp32 CUDA tensors")
    for t in (xq_op, xk_op, xv_op):
        if t is not None and (not t.is_cuda or t.dtype is not torch.bfloat16 or not t.is_contiguous()):
            raise ValueError("mha_block_infer: operand shadows must be contiguous bf16 CUDA tensors")
    _need_cuda(mask)
    l = L.lib()
    nws = l.ovqa_mha_block_infer_workspace_bytes(B, Lq, Lk, D, h, d_k, d_v, kind, dt)
    if nws > max_workspace:
        return None
    y = torch.empty_like(xq)
    y16 = torch.empty(xq.shape, device=xq.device, dtype=torch.bfloat16) if dt == L.BF16 else None
    if B == 0 or Lq == 0:
        return y, y16
    hint = getattr(mask, "_ovqa_prefix_lm", None) if dt == L.BF16 else None
    mkeep, mptr, sb, sh, sq = _mask_args(mask, B, h, Lq, Lk)
    pk, psb, ptail, pval = None, 0, 0, 0.0
    if hint is not None:
        key_row, ptail, pval = hint
        pk, psb = key_row.data_ptr(), (key_row.stride(0) if key_row.shape[0] > 1 else 0)
    st = _stream()
    ws = _WS.get((xq.device, st))
    if ws is None or ws.numel() < nws or torch.cuda.is_current_stream_capturing():
        ws = _workspace(xq.device, nws)
    p = lambda t: None if t is None else t.data_ptr()
    rc = l.ovqa_mha_block_infer(xq.data_ptr(), p(xq_op), p(xk), p(xk_op), p(xv), p(xv_op), kind, B, Lq, Lk, D, h, d_k, d_v, wptrs[0], wptrs[1],
                                wptrs[2], wptrs[3], wptrs[4], wptrs[5], eps, mptr, sb, sh, sq, pk, psb, int(ptail), float(pval), scale, dt,
                                ws.data_ptr(), ws.numel(), y.data_ptr(), p(y16), st)
    if rc:
        L.check(rc, "ovqa_mha_block_infer")
    return y, y16


def pwff_infer(x, x_op, wptrs, *, dt: int, d_ff: int, eps: float):
    """LN(x + fc2(gelu(fc1 x))) in eval mode from one call.  x (…, D) fp32 contiguous; ``wptrs`` = device pointers (w1, b1, w2, b2, gamma,
    beta) of tensors the caller keeps alive.  Returns (y fp32, y bf16 | None)."""
    D = x.shape[-1]
    rows = x.numel() // D
    if not x.is_cuda or x.dtype is not torch.float32 or not x.is_contiguous():
        raise ValueError("pwff_infer: x must be a contiguous fp32 CUDA tensor")
    _need_cuda(x_op)
    y = torch.empty_like(x)
    y16 = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if dt == L.BF16 else None
    if rows == 0:
        return y, y16
    l = L.lib()
    st = _stream()
    nws = l.ovqa_pwff_infer_workspace_bytes(rows, D, d_ff, dt)
    ws = _WS.get((x.device, st))
    if ws is None or ws.numel() < nws or torch.cuda.is_current_stream_capturing():
        ws = _workspace(x.device, nws)
    rc = l.ovqa_pwff_infer(x.data_ptr(), None if x_op is None else x_op.data_ptr(), rows, D, d_ff, wptrs[0], wptrs[1], wptrs[2], wptrs[3],
                           wptrs[4], wptrs[5], eps, dt, ws.data_ptr(), ws.numel(), y.data_ptr(), None if y16 is None else y16.data_ptr(), st)
    if rc:
        L.check(rc, "ovqa_pwff_infer")
    return y, y16


def topk_rows(x: torch.Tensor, k: int):
    """The k largest entries of every row of x (rows, n) fp32, descending: (values (rows, k) fp32, indices (rows, k) int64).
    NaN sorts first (as torch.sort(descending=True)); equal values come in ascending index order."""
    _need_cuda(x)
    x = _rowmajor2d(x, "x")
    if x.dtype != torch.float32:
        raise ValueError("topk_rows: x must be fp32")
    rows, n = x.shape
    if not (1 <= k <= min(16, n)):
        raise ValueError(f"topk_rows: k must be in 1..min(16, {n})")
    val = torch.empty((rows, k), device=x.device, dtype=torch.float32)
    idx = torch.empty((rows, k), device=x.device, dtype=torch.int64)
    if rows == 0:
        return val, idx
    l = L.lib()
    ws = _workspace(x.device, l.ovqa_topk_rows_workspace_bytes(rows, n, k))
    L.check(l.ovqa_topk_rows(_p(x), rows, n, _ld(x), k, _p(val), _p(idx), _p(ws), ws.numel(), _stream()), "ovqa_topk_rows")
    return val, idx
