"""TEST INFRASTRUCTURE: plain-torch CPU restatements of every call in ``openvivqa_b200.ops`` (one per C-ABI entry point).

The product has no CPU path (``ops.*`` raise on CPU tensors).  These stand-ins let the CPU suite exercise the HOST logic that
sits on top of the library — the hand-written backward passes in ``functional.py``, the module / session / cache bookkeeping in
``modules.py`` and ``m4c.py``, ``install()`` on the reference's own classes — with the arithmetic each entry point is documented to
perform (include/ovqa.h, csrc/common.cuh ``epilogue_value``).  Use through the ``cpu_ops`` fixture (tests/conftest.py); nothing
under ``openvivqa_b200/`` imports this file.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _keep_scale(n, p, seed, offset, offset_base=None):
    """Deterministic stand-in for the Philox keep-scale stream of (seed, offset): 1/(1-p) or 0 per flat element index."""
    if p <= 0.0:
        return torch.ones(n)
    base = int(offset_base.item()) if offset_base is not None else 0
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(offset) + base) & 0x7FFFFFFFFFFF)
    return (torch.rand(n, generator=g) >= p).float() / (1.0 - p)


def _op(t, op):
    return t.float() if op == "N" else t.float().t()


def gemm(a, b, *, a_op="N", b_op="N", out=None, out_dtype=None, bias=None, alpha=1.0, beta=0.0, gelu=False, dgelu_aux=None,
         aux_out=None, residual=None, c2=None, dropout_p=0.0, seed=0, offset=0, offset_base=None, split_k=1, save_dgelu=False,
         mul_aux=None, colsum=None, relu=False):
    A, B = _op(a, a_op), _op(b, b_op)
    m, n = A.shape[0], B.shape[0]
    v = (A @ B.t()) * alpha
    if split_k != 1:
        if out is None:
            out = torch.zeros((m, n), dtype=out_dtype or torch.float32)
        out += v.to(out.dtype)
        return out
    if bias is not None:
        v = v + bias
    ds = _keep_scale(m * n, dropout_p, seed, offset, offset_base).view(m, n)
    if aux_out is not None:
        if save_dgelu:
            cdf = 0.5 * (1.0 + torch.erf(v / math.sqrt(2.0)))
            pdf = torch.exp(-0.5 * v * v) / math.sqrt(2.0 * math.pi)
            aux_out.copy_(((cdf + v * pdf) * ds).to(aux_out.dtype))
        else:
            aux_out.copy_(v.to(aux_out.dtype))
    if gelu:
        v = F.gelu(v)
    if relu:
        v = torch.relu(v)
    if dgelu_aux is not None:
        x = dgelu_aux.float()
        cdf = 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))
        v = v * (cdf + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi))
    if mul_aux is not None:
        v = v * mul_aux.float()
    v = v * ds
    if residual is not None:
        v = v + residual
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype or a.dtype)
    if beta != 0.0:
        v = v + beta * out.float()
    if colsum is not None:
        colsum += v.sum(0)
    out.copy_(v.to(out.dtype))
    if c2 is not None:
        c2.copy_(v.to(torch.bfloat16))
    return out


def _scores(q, k, h, d_k, mask, geo, scale):
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    qh = q[..., :h * d_k].float().reshape(B, Lq, h, d_k).permute(0, 2, 1, 3)
    kh = k[..., :h * d_k].float().reshape(B, Lk, h, d_k).permute(0, 2, 3, 1)
    s = (qh @ kh) * scale
    if mask is not None:
        m = mask.float()
        while m.dim() < 4:
            m = m.unsqueeze(0)
        s = s + m
    if geo is not None:
        s = geo_log_bias(geo, h) + s
    return s


def geo_raw(geo, h):
    """(B,h,N,N) pre-relu geometry weights  w_h . emb(i,j) + b_h  (models/utils.py:102-162, attentions.py:113-119)."""
    boxes = geo.boxes.float()
    B, N = boxes.shape[:2]
    x0, y0, x1, y1 = boxes.unbind(-1)
    cx, cy, w, hh = (x0 + x1) * 0.5, (y0 + y1) * 0.5, (x1 - x0) + 1.0, (y1 - y0) + 1.0
    dx = torch.log(torch.clamp(torch.abs((cx[:, :, None] - cx[:, None, :]) / w[:, :, None]), min=1e-3))
    dy = torch.log(torch.clamp(torch.abs((cy[:, :, None] - cy[:, None, :]) / hh[:, :, None]), min=1e-3))
    dw = torch.log(w[:, :, None] / w[:, None, :])
    dh = torch.log(hh[:, :, None] / hh[:, None, :])
    pos = torch.stack((dx, dy, dw, dh), dim=-1)
    if geo.trig:
        mul = (100.0 * pos.unsqueeze(-1)) * geo.freq.view(1, 1, 1, 1, -1)
        mul = mul.reshape(B, N, N, -1)
        emb = torch.cat((torch.sin(mul), torch.cos(mul)), dim=-1)
    else:
        emb = pos
    return torch.einsum("bijg,hg->bhij", emb, geo.w.float()) + geo.b.float().view(1, -1, 1, 1), emb


def geo_log_bias(geo, h):
    raw, _ = geo_raw(geo, h)
    return torch.log(torch.clamp(torch.relu(raw), min=1e-6))


def box_geometry_bias(geo, h):
    return geo_log_bias(geo, h)


def gemm_ln_supported(a, w):
    return False          # the CPU restatement of the blocks keeps the two-step form (gemm + layernorm_fwd)


def gemm_ln_enabled():
    return False


def lstm_supported(H):
    return H in (128, 256, 512) or H % 4 == 0      # (the CPU restatement has no shape limits; small sizes keep the host-logic tests fast)


def lstm_fwd(xg, w_hh16, b_hh, save=True):
    B, T, G = xg.shape
    H = G // 4
    w = w_hh16.float()
    h = xg.new_zeros(B, H)
    c = xg.new_zeros(B, H)
    ys, hp, gs, cs = [], [], [], []
    for t in range(T):
        hp.append(h.to(torch.bfloat16))
        pre = xg[:, t] + h.to(torch.bfloat16).float() @ w.t() + (b_hh if b_hh is not None else 0.0)
        i, f, g, o = pre.chunk(4, dim=-1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c = f * c + i * g
        h = o * torch.tanh(c)
        ys.append(h); gs.append(torch.cat((i, f, g, o), -1).to(torch.bfloat16)); cs.append(c)
    y = torch.stack(ys, 1)
    if not save:
        return y, y.to(torch.bfloat16), None, None, None
    return y, y.to(torch.bfloat16), torch.stack(hp, 1), torch.stack(gs, 1), torch.stack(cs, 1)


def lstm_bwd(dy, w_hh16, gates16, c):
    B, T, H = dy.shape
    w = w_hh16.float()
    dc_next = dy.new_zeros(B, H)
    rec = dy.new_zeros(B, H)
    out = [None] * T
    for t in range(T - 1, -1, -1):
        i, f, g, o = gates16[:, t].float().chunk(4, dim=-1)
        ct = c[:, t]
        cp = c[:, t - 1] if t > 0 else torch.zeros_like(ct)
        dh = dy[:, t] + rec
        tc = torch.tanh(ct)
        dc = dc_next + dh * o * (1 - tc * tc)
        dc_next = dc * f
        da = torch.cat((dc * g * i * (1 - i), dc * cp * f * (1 - f), dc * i * (1 - g * g), dh * tc * o * (1 - o)), -1).to(torch.bfloat16)
        out[t] = da
        rec = da.float() @ w
    return torch.stack(out, 1)


def attr_pool_fwd(hidden, w2, b2, x, save_probs=True):
    logits = hidden.float() @ w2.reshape(-1).float() + b2.reshape(-1).float()
    p = torch.softmax(logits, dim=1)
    return (x * p.unsqueeze(-1)).sum(1), (p if save_probs else None)


def attr_pool_bwd(dpooled, hidden, w2, x, probs, dropout_p, dw2, db2, need_dx=True):
    dp = (x * dpooled.unsqueeze(1)).sum(-1)
    dl = probs * (dp - (probs * dp).sum(1, keepdim=True))
    dw2 += (dl.unsqueeze(-1) * hidden.float()).sum((0, 1))
    db2 += dl.sum()
    dpre = torch.where(hidden.float() > 0, dl.unsqueeze(-1) * w2.reshape(1, 1, -1).float() / (1.0 - dropout_p), torch.zeros(())).to(hidden.dtype)
    return (probs.unsqueeze(-1) * dpooled.unsqueeze(1) if need_dx else None), dpre


def geo_bias(geo, h, key_mask=None):
    out = geo_log_bias(geo, h)
    if key_mask is not None:
        out = out + key_mask.reshape(-1, 1, 1, geo.boxes.shape[1]).float()
    return out.contiguous()


def geo_bias_bwd(geo, h, dbias, geo_dw, geo_db):
    raw, emb = geo_raw(geo, h)
    coef = torch.where(raw > 1e-6, dbias / raw, torch.zeros_like(dbias))
    geo_dw += torch.einsum("bhij,bijg->hg", coef, emb)
    geo_db += coef.sum(dim=(0, 2, 3))


def _prob_keep(prob_dropout, shape):
    if prob_dropout is None or prob_dropout[0] <= 0.0:
        return None
    p, seed, off, base = prob_dropout
    return _keep_scale(math.prod(shape), p, seed, off, base).view(shape)


def sdpa_fwd(q, k, v, *, h, d_k, d_v, mask=None, geo=None, scale=None, need_lse=False, need_att=False, out=None, prob_dropout=None):
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    scale = 1.0 / math.sqrt(d_k) if scale is None else scale
    s = _scores(q, k, h, d_k, mask, geo, scale)
    p = torch.softmax(s, dim=-1)
    keep = _prob_keep(prob_dropout, p.shape)
    if keep is not None:
        p = p * keep
    vh = v[..., :h * d_v].float().reshape(B, Lk, h, d_v).permute(0, 2, 1, 3)
    o = (p @ vh).permute(0, 2, 1, 3).reshape(B, Lq, h * d_v).to(q.dtype)
    if out is not None:
        out.copy_(o)
        o = out
    return o, (torch.logsumexp(s, dim=-1) if need_lse else None), (p.contiguous() if need_att else None)


def sdpa_bwd(q, k, v, out, dout, lse, dq, dk, dv, *, h, d_k, d_v, mask=None, geo=None, geo_dw=None, geo_db=None, scale=None,
             att=None, datt=None, dbias=None, prob_dropout=None):
    B, Lq, Lk = q.shape[0], q.shape[1], k.shape[1]
    scale = 1.0 / math.sqrt(d_k) if scale is None else scale
    s = _scores(q, k, h, d_k, mask, geo, scale)
    p = torch.exp(s - lse.unsqueeze(-1))
    do = dout.float().reshape(B, Lq, h, d_v).permute(0, 2, 1, 3)
    vh = v[..., :h * d_v].float().reshape(B, Lk, h, d_v).permute(0, 2, 1, 3)
    qh = q[..., :h * d_k].float().reshape(B, Lq, h, d_k).permute(0, 2, 1, 3)
    kh = k[..., :h * d_k].float().reshape(B, Lk, h, d_k).permute(0, 2, 1, 3)
    dp = do @ vh.transpose(-1, -2)
    if datt is not None:
        dp = dp + datt.float()
    keep = _prob_keep(prob_dropout, p.shape)
    if keep is not None:
        dp = dp * keep
    ds = p * (dp - (p * dp).sum(-1, keepdim=True))
    if dbias is not None:
        dbias.copy_(ds)
    if keep is not None:
        p = p * keep
    dv.copy_((p.transpose(-1, -2) @ do).permute(0, 2, 1, 3).reshape(B, Lk, h * d_v).to(dv.dtype))
    dq.copy_(((ds @ kh) * scale).permute(0, 2, 1, 3).reshape(B, Lq, h * d_k).to(dq.dtype))
    dk.copy_(((ds.transpose(-1, -2) @ qh) * scale).permute(0, 2, 1, 3).reshape(B, Lk, h * d_k).to(dk.dtype))
    if geo is not None and geo_dw is not None:
        raw, emb = geo_raw(geo, h)
        coef = torch.where(raw > 1e-6, ds / raw, torch.zeros_like(ds))
        geo_dw += torch.einsum("bhij,bijg->hg", coef, emb)
        geo_db += coef.sum(dim=(0, 2, 3))


def layernorm_fwd(x, gamma, beta, *, residual=None, post_add=None, eps=1e-5, want_f32=True, want_bf16=False, save_stats=False,
                  save_sum=False):
    D = x.shape[-1]
    s = x.float() if residual is None else x.float() + residual.float()
    mean = s.mean(-1, keepdim=True)
    var = s.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + eps)
    y = (s - mean) * rstd * gamma.float() + beta.float()
    if post_add is not None:
        pa = post_add.reshape(-1, D).float()
        y = (y.reshape(-1, pa.shape[0], D) + pa).reshape(x.shape)
    return (y if want_f32 else None, y.to(torch.bfloat16) if want_bf16 else None, s if save_sum else None,
            mean.reshape(-1) if save_stats else None, rstd.reshape(-1) if save_stats else None)


def layernorm_bwd(dy, x, mean, rstd, gamma, *, dx_add=None, want_bf16=False, dropout_p=0.0, seed=0, offset=0, offset_base=None,
                  dgamma=None, dbeta=None, want_dbias=False, dbias=None):
    D = x.shape[-1]
    dy2, x2 = dy.reshape(-1, D).float(), x.reshape(-1, D).float()
    xh = (x2 - mean.view(-1, 1)) * rstd.view(-1, 1)
    g = dy2 * gamma.float()
    dx = (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True)) * rstd.view(-1, 1)
    if dx_add is not None:
        dx = dx + dx_add.reshape(-1, D)
    dgamma = torch.zeros(D) if dgamma is None else dgamma
    dbeta = torch.zeros(D) if dbeta is None else dbeta
    dgamma += (dy2 * xh).sum(0)
    dbeta += dy2.sum(0)
    masked = dx * _keep_scale(dx.numel(), dropout_p, seed, offset, offset_base).view_as(dx)
    outs = (dx.view(x.shape), masked.to(torch.bfloat16).view(x.shape) if want_bf16 else None, dgamma, dbeta)
    if want_dbias:
        dbias = torch.zeros(D) if dbias is None else dbias
        dbias += masked.sum(0)
        outs = outs + (dbias,)
    return outs


def cast_rows_padmask(x, mask_value):
    return x.to(torch.bfloat16), (x.float().sum(-1) == 0).float() * mask_value


def mul_to_bf16(a, b):
    return (a.float() * b.float()).to(torch.bfloat16)


def to_bf16(x):
    return x.to(torch.bfloat16)


def to_f32(x):
    return x.float()


def colsum(x, out=None, accumulate=False):
    s = x.float().sum(0)
    if out is None:
        return s
    if accumulate:
        out += s
    else:
        out.copy_(s)
    return out


def add(a, b):
    return a + b


def dropout_apply(x, p, seed, offset, out_dtype=torch.float32, offset_base=None):
    return (x.float() * _keep_scale(x.numel(), p, seed, offset, offset_base).view_as(x)).to(out_dtype)


def dropout_keep_mask(n, p, seed, offset, device, offset_base=None):
    return (_keep_scale(n, p, seed, offset, offset_base) > 0).float()


def stage_many(table, n_desc):
    raise AssertionError("stage_many is CUDA-only (the weight cache only stages CUDA parameters)")


def advance_u64(word, delta):
    word += delta


def ptr_scores(q, k, mask, scores, col_offset, scale, mask_is_bool=False):
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * scale
    if mask is not None:
        m = mask.reshape(q.shape[0], 1, k.shape[1])
        s = s.masked_fill(m.bool(), float("-inf")) if mask_is_bool else s + m.float()
    scores[..., col_offset:col_offset + k.shape[1]] = s
    return scores


def ptr_scores_bwd(ds, q, k, scale, bool_mask=None):
    dsf = ds.float()
    if bool_mask is not None:
        dsf = dsf.masked_fill(bool_mask.reshape(q.shape[0], 1, k.shape[1]).bool(), 0.0)
    return (torch.matmul(dsf, k.float()) * scale).to(q.dtype), (torch.matmul(dsf.transpose(-1, -2), q.float()) * scale).to(k.dtype)


def prev_pred_embed(ans_emb, ocr_emb, prev_inds, pos_emb, type_emb, ans_ln, ocr_ln, emb_ln, eps, out=None):
    B, T = prev_inds.shape
    V, D = ans_emb.shape
    ln = lambda x, p: F.layer_norm(x, (D,), p[0], p[1], eps)
    is_ocr = prev_inds >= V
    ans_rows = ln(ans_emb[prev_inds.clamp(max=V - 1)], ans_ln)
    ocr_rows = torch.gather(ln(ocr_emb, ocr_ln), 1, (prev_inds - V).clamp(min=0).unsqueeze(-1).expand(B, T, D))
    raw = torch.where(is_ocr.unsqueeze(-1), ocr_rows, ans_rows)
    emb = ln(pos_emb[:T].unsqueeze(0) + type_emb[is_ocr.long()], emb_ln)
    res = raw + emb
    if out is None:
        return res
    out.copy_(res)
    return out


def greedy_step(scores, prev_inds, last_ids, argmax_out, all_eos, step, eos_idx):
    am = scores.argmax(dim=-1)
    argmax_out.copy_(am)
    prev_inds[:, 1:] = am[:, :-1]
    last_ids.copy_(torch.where(last_ids == eos_idx, last_ids, am[:, step]))
    all_eos.fill_(int(bool((last_ids == eos_idx).all())))


def topk_rows(x, k):
    v, i = torch.sort(x, dim=-1, descending=True, stable=True)
    return v[:, :k].contiguous(), i[:, :k].contiguous()


def mha_decode_step(x, x_op, kv_src, self_attention, wptrs, **kw):
    raise AssertionError("mha_decode_step takes raw device pointers: stub functional.mha_decode_block instead")


def pwff_infer(x, x_op, wptrs, **kw):
    raise AssertionError("pwff_infer takes raw device pointers: stub functional.pwff_infer_block instead")


# blocks that hand raw device pointers to the library: restated one level up (functional.*)
def mha_decode_block(plan, queries, kv_src, self_att, cache, length, mask, *, h, d_k, d_v, eps=1e-5, scale=None):
    w_qkv, b_qkv, w_o, b_o, gamma, beta = plan.keep
    hk = h * d_k
    B, Lq, D = queries.shape
    q = F.linear(queries, w_qkv[:hk].float(), b_qkv[:hk])
    src = queries if self_att else kv_src
    if src is not None:
        cache[:, length:length + src.shape[1]] = F.linear(src.float(), w_qkv[hk:].float(), b_qkv[hk:]).to(cache.dtype)
        length += src.shape[1]
    k, v = cache[:, :length, :hk].float(), cache[:, :length, hk:].float()
    sc = 1.0 / math.sqrt(d_k) if scale is None else scale
    att = q.view(B, Lq, h, d_k).permute(0, 2, 1, 3) @ k.reshape(B, length, h, d_k).permute(0, 2, 3, 1) * sc
    att = torch.softmax(att if mask is None else att + mask, -1)
    ctx = (att @ v.reshape(B, length, h, d_v).permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(B, Lq, h * d_v)
    return F.layer_norm(queries + F.linear(ctx, w_o.float(), b_o), (D,), gamma, beta, eps)


def pwff_infer_block(plan, x, *, d_ff, eps=1e-5):
    w1, b1, w2, b2, g, b = plan.keep
    return F.layer_norm(x + F.linear(F.gelu(F.linear(x, w1.float(), b1)), w2.float(), b2), (x.shape[-1],), g, b, eps)


OPS = ("gemm", "sdpa_fwd", "sdpa_bwd", "layernorm_fwd", "layernorm_bwd", "cast_rows_padmask", "mul_to_bf16", "to_bf16", "to_f32", "colsum",
       "add", "dropout_apply", "dropout_keep_mask", "stage_many", "advance_u64", "box_geometry_bias", "geo_bias", "geo_bias_bwd", "attr_pool_fwd", "attr_pool_bwd", "lstm_supported", "lstm_fwd", "lstm_bwd", "gemm_ln_supported", "gemm_ln_enabled", "ptr_scores", "ptr_scores_bwd",
       "prev_pred_embed", "greedy_step", "topk_rows", "mha_decode_step", "pwff_infer")


def install(monkeypatch):
    """Replace every library-backed call with its CPU restatement for the duration of a test."""
    from openvivqa_b200 import functional as Fn, ops
    g = globals()
    for name in OPS:
        monkeypatch.setattr(ops, name, g[name], raising=False)
    monkeypatch.setattr(Fn, "mha_decode_block", mha_decode_block)
    monkeypatch.setattr(Fn, "pwff_infer_block", pwff_infer_block)
