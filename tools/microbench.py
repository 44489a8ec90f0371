#!/usr/bin/env python
"""Kernel-level timings on the GPU box (CUDA events, L2 flushed between iterations), each against its roofline:

    python tools/microbench.py [gemm] [sdpa] [ln] [mha] > gpurun_out/microbench.jsonl

* gemm : ovqa_gemm (tcgen05) on the MCAN/M4C training shapes (fwd, dgrad, wgrad forms) next to torch.matmul (cuBLAS)
* sdpa : attention core forward/backward at the §8a shape classes -> GB/s of algorithmic bytes
* ln   : LayerNorm forward/backward -> GB/s
* mha  : BASELINE configs[4] sweep — fused MHA block (eval) B x Lq=30 x Lk, D 512/768 vs roofline
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import openvivqa_b200 as ovqa  # noqa: E402
from openvivqa_b200 import harness as H, ops  # noqa: E402

dev = torch.device("cuda:0")
PEAKS = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2


def timeit(fn, iters=20, warmup=3, flush=True):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]        # median, microseconds


def emit(**kw):
    print(json.dumps(kw), flush=True)


def bench_gemm():
    B = 256
    shapes = [  # (tag, m, n, k, a_op, b_op, out dtype)
        ("qkv_self_vision  NN bf16", B * 100, 1536, 512, "N", "N", torch.bfloat16),
        ("fc1_vision       NN bf16", B * 100, 2048, 512, "N", "N", torch.bfloat16),
        ("fc2_vision       NN f32 ", B * 100, 512, 2048, "N", "N", torch.float32),
        ("fc_o_vision      NN f32 ", B * 100, 512, 512, "N", "N", torch.float32),
        ("qkv_self_text    NN bf16", B * 30, 1536, 512, "N", "N", torch.bfloat16),
        ("dgrad fc2 (dpre) NT bf16", B * 100, 2048, 512, "N", "T", torch.bfloat16),
        ("dgrad fc1 (dx)   NT f32 ", B * 100, 512, 2048, "N", "T", torch.float32),
        ("dgrad qkv (dx)   NT f32 ", B * 100, 512, 1536, "N", "T", torch.float32),
        ("wgrad fc1        TT f32 ", 2048, 512, B * 100, "T", "T", torch.float32),
        ("wgrad fc2        TT f32 ", 512, 2048, B * 100, "T", "T", torch.float32),
        ("wgrad qkv        TT f32 ", 1536, 512, B * 100, "T", "T", torch.float32),
        ("wgrad fc_o       TT f32 ", 512, 512, B * 100, "T", "T", torch.float32),
        ("m4c mmt qkv      NN bf16", 64 * 182, 2304, 768, "N", "N", torch.bfloat16),
        ("m4c mmt ffn1     NN bf16", 64 * 182, 3072, 768, "N", "N", torch.bfloat16),
        ("square 8192      NN bf16", 8192, 8192, 8192, "N", "N", torch.bfloat16),
        ("square 4096      NN bf16", 4096, 4096, 4096, "N", "N", torch.bfloat16),
    ]
    from openvivqa_b200 import functional as Fn
    for tag, m, n, k, ao, bo, odt in shapes:
        a = torch.randn((m, k) if ao == "N" else (k, m), device=dev).bfloat16()
        b = torch.randn((n, k) if bo == "N" else (k, n), device=dev).bfloat16()
        sk = 0 if ao == "T" else 1      # 0: library picks split + tile width
        out = torch.zeros((m, n), device=dev, dtype=odt)
        t_ours = timeit(lambda: ops.gemm(a, b, a_op=ao, b_op=bo, out=out, split_k=sk))
        am = a if ao == "N" else a.t()
        bm = b.t() if bo == "N" else b
        ref_out = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
        t_ref = timeit(lambda: torch.matmul(am, bm, out=ref_out))
        fl = 2.0 * m * n * k
        emit(kind="gemm", tag=tag, m=m, n=n, k=k, split_k=sk, us=t_ours, tflops=fl / t_ours / 1e6, cublas_us=t_ref,
             cublas_tflops=fl / t_ref / 1e6, frac_of_peak=fl / t_ours / 1e6 / PEAKS["bf16_tflops"])


def bench_gemm_epi():
    """The fused-epilogue GEMMs of one MCAN training step (vision stream, batch 256): microseconds next to the HBM and tensor floors."""
    B = 256
    M = B * 100
    for tag, m, n, k, bo, odt, feat in [
            ("fc_o  bias+drop+res  f32 ", M, 512, 512, "N", torch.float32, "bdr"),
            ("fc2   bias+drop+res  f32 ", M, 512, 2048, "N", torch.float32, "bdr"),
            ("fc_o  bias+res eval  f32 ", M, 512, 512, "N", torch.float32, "br"),
            ("fc1   bias+gelu+save+drop", M, 2048, 512, "N", torch.bfloat16, "bgsd"),
            ("fc1   bias+gelu eval     ", M, 2048, 512, "N", torch.bfloat16, "bg"),
            ("qkv   bias           bf16", M, 1536, 512, "N", torch.bfloat16, "b"),
            ("dpre  mulaux         bf16", M, 2048, 512, "T", torch.bfloat16, "m"),
            ("dpre  mulaux+colsum  bf16", M, 2048, 512, "T", torch.bfloat16, "mc"),
            ("dctx  colsum         bf16", M, 512, 512, "T", torch.bfloat16, "c"),
            ("dx    res            f32 ", M, 512, 512, "T", torch.float32, "r"),
            ("dx    res k=2048     f32 ", M, 512, 2048, "T", torch.float32, "r"),
            ("plain                bf16", M, 512, 512, "T", torch.bfloat16, ""),
            ("plain                f32 ", M, 512, 512, "N", torch.float32, ""),
            ("fc_o text bias+drop+res  ", B * 30, 512, 512, "N", torch.float32, "bdr")]:
        a = torch.randn((m, k), device=dev).bfloat16()
        b = torch.randn((n, k) if bo == "N" else (k, n), device=dev).bfloat16()
        out = torch.empty((m, n), device=dev, dtype=odt)
        kw = {}
        bytes_ = m * k * 2 + n * k * 2 + m * n * out.element_size()
        if "b" in feat:
            kw["bias"] = torch.randn(n, device=dev)
        if "d" in feat:
            kw.update(dropout_p=0.1, seed=3, offset=7)
        if "r" in feat:
            kw["residual"] = torch.randn((m, n), device=dev)
            bytes_ += m * n * 4
        if "g" in feat:
            kw["gelu"] = True
        if "s" in feat:
            kw.update(aux_out=torch.empty((m, n), device=dev, dtype=odt), save_dgelu=True)
            bytes_ += m * n * out.element_size()
        if "m" in feat:
            kw["mul_aux"] = torch.randn((m, n), device=dev).to(odt)
            bytes_ += m * n * out.element_size()
        if "c" in feat:
            kw["colsum"] = torch.zeros(n, device=dev)
        t = timeit(lambda: ops.gemm(a, b, b_op=bo, out=out, **kw))
        fl = 2.0 * m * n * k
        emit(kind="gemm_epi", tag=tag, m=m, n=n, k=k, us=t, tflops=fl / t / 1e6, hbm_floor_us=bytes_ / PEAKS["hbm_gbs"] / 1e3,
             tensor_floor_us=fl / PEAKS["bf16_tflops"] / 1e6, gbs=bytes_ / t / 1e3)


def bench_wgrad_split():
    """Split-K sweep for the weight-gradient GEMMs (dW = dy^T x, k = B*L rows): atomics per output element grow with the
    split, SM occupancy with the number of work units."""
    B = 256
    for tag, m, n, k in [("wgrad fc_o 512x512", 512, 512, B * 100), ("wgrad qkv 1536x512", 1536, 512, B * 100), ("wgrad fc1 2048x512", 2048, 512, B * 100),
                         ("wgrad fc2 512x2048", 512, 2048, B * 100), ("wgrad fc_o text 512x512", 512, 512, B * 30), ("wgrad fc1 text", 2048, 512, B * 30)]:
        a = torch.randn((k, m), device=dev).bfloat16()
        b = torch.randn((k, n), device=dev).bfloat16()
        out = torch.zeros((m, n), device=dev, dtype=torch.float32)
        for sk in (0, 1, 2, 3, 4, 6, 9, 12, 18, 24, 37):
            t = timeit(lambda: ops.gemm(a, b, a_op="T", b_op="T", out=out, split_k=sk), iters=10)
            emit(kind="wgrad_split", tag=tag, m=m, n=n, k=k, split_k=sk, us=t, tflops=2.0 * m * n * k / t / 1e6)


def bench_sdpa():
    for (B, h, Lq, Lk, d) in [(256, 8, 100, 100, 64), (256, 8, 30, 30, 64), (256, 8, 100, 30, 64), (256, 8, 30, 100, 64),
                              (1024, 8, 30, 100, 64), (256, 8, 30, 600, 64), (64, 8, 182, 182, 96), (4096, 8, 30, 100, 64)]:
        fused = Lq == Lk
        if fused:
            buf = torch.randn((B, Lq, 3 * h * d), device=dev).bfloat16()
            q, k, v = buf[..., :h * d], buf[..., h * d:2 * h * d], buf[..., 2 * h * d:]
            dbuf = torch.empty_like(buf)
            dq, dk, dv = dbuf[..., :h * d], dbuf[..., h * d:2 * h * d], dbuf[..., 2 * h * d:]
        else:
            q = torch.randn((B, Lq, h * d), device=dev).bfloat16()
            kv = torch.randn((B, Lk, 2 * h * d), device=dev).bfloat16()
            k, v = kv[..., :h * d], kv[..., h * d:]
            dq, dkv = torch.empty_like(q), torch.empty_like(kv)
            dk, dv = dkv[..., :h * d], dkv[..., h * d:]
        mask = torch.zeros((B, 1, 1, Lk), device=dev)
        mask[:, :, :, int(Lk * 0.8):] = -10e4
        out = torch.empty((B, Lq, h * d), device=dev, dtype=torch.bfloat16)
        dout = torch.randn((B, Lq, h * d), device=dev).bfloat16()
        lse = ops.sdpa_fwd(q, k, v, h=h, d_k=d, d_v=d, mask=mask, need_lse=True, out=out)[1]
        t_f = timeit(lambda: ops.sdpa_fwd(q, k, v, h=h, d_k=d, d_v=d, mask=mask, need_lse=True, out=out))
        t_b = timeit(lambda: ops.sdpa_bwd(q, k, v, out, dout, lse, dq, dk, dv, h=h, d_k=d, d_v=d, mask=mask))
        e = 2
        y_f = B * (e * h * d * (2 * Lq + 2 * Lk) + 4 * Lk + 4 * h * Lq)                    # q,k,v in; out, lse out; mask
        y_b = B * (e * h * d * (3 * Lq + 2 * Lk) + e * h * d * (Lq + 2 * Lk) + 4 * Lk + 4 * h * Lq)   # q,k,v,o,do in; dq,dk,dv out
        fl = 4.0 * B * h * Lq * Lk * d
        emit(kind="sdpa", B=B, h=h, Lq=Lq, Lk=Lk, d=d, fwd_us=t_f, fwd_gbs=y_f / t_f / 1e3, fwd_frac_hbm=y_f / t_f / 1e3 / PEAKS["hbm_gbs"],
             fwd_tflops=fl / t_f / 1e6, bwd_us=t_b, bwd_gbs=y_b / t_b / 1e3, bwd_frac_hbm=y_b / t_b / 1e3 / PEAKS["hbm_gbs"],
             bwd_tflops=2.5 * fl / t_b / 1e6)


def bench_ln():
    for rows, D in [(25600, 512), (7680, 512), (11648, 768)]:
        x = torch.randn((rows, D), device=dev)
        g, b = torch.randn(D, device=dev), torch.randn(D, device=dev)
        y, y16, _, mean, rstd = ops.layernorm_fwd(x, g, b, want_bf16=True, save_stats=True)
        t_f = timeit(lambda: ops.layernorm_fwd(x, g, b, want_bf16=True, save_stats=True))
        dy = torch.randn_like(x)
        t_b = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, want_bf16=True))
        y_f, y_b = rows * D * (4 + 4 + 2), rows * D * (4 + 4 + 4 + 2)
        emit(kind="layernorm", rows=rows, D=D, fwd_us=t_f, fwd_gbs=y_f / t_f / 1e3, fwd_frac_hbm=y_f / t_f / 1e3 / PEAKS["hbm_gbs"],
             bwd_us=t_b, bwd_gbs=y_b / t_b / 1e3, bwd_frac_hbm=y_b / t_b / 1e3 / PEAKS["hbm_gbs"])


def _torch_mha(q, kv, mask, w, h):
    """The reference block's arithmetic in plain PyTorch ops (attentions.py:46-60, 319-339, eval): what the unmodified module launches."""
    import torch.nn.functional as F
    B, Lq, D = q.shape
    Lk, d = kv.shape[1], D // h
    Q = F.linear(q, w["q"][0], w["q"][1]).view(B, Lq, h, d).permute(0, 2, 1, 3)
    K = F.linear(kv, w["k"][0], w["k"][1]).view(B, Lk, h, d).permute(0, 2, 3, 1)
    V = F.linear(kv, w["v"][0], w["v"][1]).view(B, Lk, h, d).permute(0, 2, 1, 3)
    att = torch.matmul(Q, K) / math.sqrt(d)
    att = torch.softmax(att + mask, dim=-1)
    out = torch.matmul(att, V).permute(0, 2, 1, 3).contiguous().view(B, Lq, D)
    out = F.linear(out, w["o"][0], w["o"][1])
    return F.layer_norm(q + out, (D,), w["ln"][0], w["ln"][1])


def bench_mha():
    """BASELINE configs[4]: fused MHA block, eval, k=v!=q, padding mask with 20% trailing keys masked; next to the same block in plain
    PyTorch ops on the same GPU (fp32 eager = the reference as shipped, and bf16 autocast)."""
    ovqa.set_compute_dtype("bf16")
    for D, dk in [(512, 64), (768, 96)]:
        m = ovqa.MultiHeadAttention(H.attention_cfg(D, 8, dk, dk, dropout=0.0, compute_dtype="bf16")).to(dev).eval()
        for Lk in (100, 250, 400, 600):
            for B in (16, 64, 256, 1024, 4096):
                if B * Lk * D * 4 > 6e9:
                    continue
                Lq = 30
                q = torch.randn((B, Lq, D), device=dev)
                kv = torch.randn((B, Lk, D), device=dev)
                mask = torch.zeros((B, 1, 1, Lk), device=dev)
                mask[:, :, :, int(Lk * 0.8):] = -10e4
                a = m.attention
                w = {"q": (a.fc_q.weight, a.fc_q.bias), "k": (a.fc_k.weight, a.fc_k.bias), "v": (a.fc_v.weight, a.fc_v.bias),
                     "o": (a.fc_o.weight, a.fc_o.bias), "ln": (m.layer_norm.weight, m.layer_norm.bias)}
                with torch.no_grad():
                    t = timeit(lambda: m(q, kv, kv, mask), iters=10, warmup=2)
                    t32 = t16 = None
                    if B * Lk * 8 * Lq * 4 < 3e9:        # the reference materialises (B, h, Lq, Lk) fp32 three times
                        t32 = timeit(lambda: _torch_mha(q, kv, mask, w, 8), iters=5, warmup=1)
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            t16 = timeit(lambda: _torch_mha(q, kv, mask, w, 8), iters=5, warmup=1)
                F_mha = 4.0 * D * D * (Lq + Lk) + 4.0 * Lq * Lk * D
                Y_mha = 4 * D * (Lq + Lk + Lq) + 4 * Lk            # fp32 activations in/out as the module sees them
                t_roof = max(B * F_mha / (PEAKS["bf16_tflops"] * 1e12), B * Y_mha / (PEAKS["hbm_gbs"] * 1e9)) * 1e6
                emit(kind="mha_block", D=D, B=B, Lq=Lq, Lk=Lk, us=t, samples_per_s=B / t * 1e6, tflops=B * F_mha / t / 1e6,
                     roofline_us=t_roof, frac_of_roofline=t_roof / t, torch_fp32_us=t32, torch_bf16_autocast_us=t16,
                     speedup_vs_torch_fp32=(t32 / t if t32 else None), speedup_vs_torch_bf16=(t16 / t if t16 else None))
    ovqa.set_compute_dtype("fp32")


def bench_decode():
    """T-step incremental decode through 3 decoder layers (decoders.py:13-26) at batch x beam = 320 rows, 100
    encoder regions.  (a) projected-K/V cache (the inference path), (b) the same kernels with the reference's bookkeeping (re-project
    the running sequence and the encoder output at every step), (c) the reference arithmetic in plain PyTorch ops, fp32 and bf16 autocast."""
    import torch.nn.functional as F
    from openvivqa_b200 import modules as M
    B, T, N, D, h, dff, nl = 320, 20, 100, 512, 8, 2048, 3
    att = lambda st: H.attention_cfg(D, h, d_ff=dff, dropout=0.0, stateful=st, compute_dtype="bf16")
    layers = [H.DecoderLayer(H.NS(SELF_ATTENTION=att(True), ENC_ATTENTION=att(False))).to(dev).eval() for _ in range(nl)]
    xs = [torch.randn((B, 1, D), device=dev) for _ in range(T)]
    enc = torch.randn((B, N, D), device=dev)
    enc_mask = torch.zeros((B, 1, 1, N), device=dev)
    enc_mask[..., 80:] = -10e4

    def ours():
        for l in layers:
            l.enable_statefulness(B)
        mask = torch.zeros((B, 1, 1, 0), device=dev)
        for t in range(T):
            mask = torch.cat([mask, torch.zeros((B, 1, 1, 1), device=dev)], -1)
            y = xs[t]
            for l in layers:
                y = l(y, enc, enc, mask, enc_mask)
        for l in layers:
            l.disable_statefulness()
        return y

    def wts(m):
        a = m.attention
        return {"q": (a.fc_q.weight, a.fc_q.bias), "k": (a.fc_k.weight, a.fc_k.bias), "v": (a.fc_v.weight, a.fc_v.bias),
                "o": (a.fc_o.weight, a.fc_o.bias), "ln": (m.layer_norm.weight, m.layer_norm.bias)}
    W = [(wts(l.self_attn), wts(l.enc_attn), l.pwff) for l in layers]

    def torch_ref():
        running = [None] * nl
        mask = torch.zeros((B, 1, 1, 0), device=dev)
        for t in range(T):
            mask = torch.cat([mask, torch.zeros((B, 1, 1, 1), device=dev)], -1)
            y = xs[t]
            for i, (ws, we, ff) in enumerate(W):
                running[i] = y if running[i] is None else torch.cat([running[i], y], 1)         # attentions.py:321-325
                y = _torch_mha(y, running[i], mask, ws, h)
                y = _torch_mha(y, enc, enc_mask, we, h)
                o = F.linear(F.gelu(F.linear(y, ff.fc1.weight, ff.fc1.bias)), ff.fc2.weight, ff.fc2.bias)
                y = F.layer_norm(y + o, (D,), ff.layer_norm.weight, ff.layer_norm.bias)
        return y

    with torch.no_grad():
        t_cache = timeit(ours, iters=5, warmup=2)
        M.DECODE_KV_CACHE = False
        t_nocache = timeit(ours, iters=5, warmup=2)
        M.DECODE_KV_CACHE = True
        # the same 20 steps captured into ONE CUDA graph (every step's shapes are known in advance): what the GPU alone needs
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ours()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            y_graph = ours()
        t_graph = timeit(graph.replay, iters=5, warmup=2)
        t32 = timeit(torch_ref, iters=3, warmup=1)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            t16 = timeit(torch_ref, iters=3, warmup=1)
        a, b = ours(), torch_ref()
    emit(kind="decode_3_layers", rows=B, steps=T, regions=N, D=D, cached_us=t_cache, cached_one_graph_us=t_graph, recompute_us=t_nocache, torch_fp32_us=t32,
         torch_bf16_autocast_us=t16, rows_per_s_cached=B / t_cache * 1e6, speedup_vs_recompute=t_nocache / t_cache,
         speedup_vs_torch_fp32=t32 / t_cache, speedup_vs_torch_bf16=t16 / t_cache, max_abs_diff_vs_torch_fp32=(a - b).abs().max().item())


def bench_beam_search():
    """The whole search as the reference drives it (base_transformer.py:30-54 + beam_search.py): 64 samples x 5 beams, 20 steps, 3 decoder
    layers, 100 encoder regions, 10 000-word vocabulary; decode path on / off (off = the reference's bookkeeping on the same kernels:
    raw running keys re-projected every step, encoder keys re-projected every step)."""
    from openvivqa_b200 import modules as M
    B, beam, T, N, D, V = 64, 5, 20, 100, 512, 10000
    att = lambda st: H.attention_cfg(D, 8, d_ff=2048, dropout=0.0, stateful=st, compute_dtype="bf16")
    cfg = H.NS(D_MODEL=D, LAYERS=3, ATTENTION=H.NS(SELF_ATTENTION=att(True), ENC_ATTENTION=att(False)))
    m = H.Captioner(cfg, H.StubVocab(V, max_answer_length=T)).to(dev).eval()
    m.eos_idx = -1                                         # never finishes: every step does full work
    enc = torch.randn((B, N, D), device=dev)
    mask = torch.zeros((B, 1, 1, N), device=dev)
    mask[..., 80:] = -10e4
    run = lambda: m.beam_search(enc, mask, B, beam)
    with torch.no_grad():
        t_on = timeit(run, iters=5, warmup=2)
        M.DECODE_KV_CACHE = False
        t_off = timeit(run, iters=5, warmup=2)
        M.DECODE_KV_CACHE = True
        cand = torch.randn((B, beam * V), device=dev)
        t_topk = timeit(lambda: ops.topk_rows(cand, beam), iters=20, warmup=3)
        t_sort = timeit(lambda: torch.sort(cand, -1, descending=True), iters=20, warmup=3)
    emit(kind="beam_search_3_layers", samples=B, beam=beam, steps=T, regions=N, D=D, vocab=V, decode_path_us=t_on, reference_bookkeeping_us=t_off,
         speedup=t_off / t_on, samples_per_s=B / t_on * 1e6, select_topk_us=t_topk, select_sort_us=t_sort)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "epi", "sdpa", "ln", "mha"]
    emit(kind="meta", gpu=torch.cuda.get_device_name(0), peaks=PEAKS)
    for w in which:
        {"gemm": bench_gemm, "wgrad": bench_wgrad_split, "epi": bench_gemm_epi, "sdpa": bench_sdpa, "ln": bench_ln, "mha": bench_mha, "decode": bench_decode, "beam": bench_beam_search}[w]()
