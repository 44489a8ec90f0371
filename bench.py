#!/usr/bin/env python
"""bench.py — the BASELINE.json metric ("VQA samples/sec (MCAN train, M4C decode) ... attention % roofline") through the B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|mcan_train|m4c_decode|xmt_train|mha_sweep]

ONE JSON line (rank 0).  The top level is the headline workload, BASELINE configs[1]: the MCAN training step (30 question tokens, 100
regions x 2048-d, d_model 512, 8 heads, 6 SA + 6 SGA layers, batch 256 per GPU, bf16 operands / fp32 accumulate, dropout 0.1 active,
NLL loss, Adam), synthetic ViVQA-shaped data, random-init weights, one process per GPU, NCCL gradient all-reduce overlapped with backward,
weak scaling.  The rest of the metric rides in the same line as nested objects measured in the same process at the same N:

    "m4c_decode": configs[2]  MMF_M4C 12-step greedy decode (value, e2e, roofline, step_breakdown)
    "xmt_train" : configs[3]  cross-modality transformer with the box-geometry attention core, batch 128, train step
    "mha_sweep" : configs[4]  fused MHA block sweep (B x Lk x D) against its roofline and against the same block in plain PyTorch ops

`value` = device-resident whole-job samples/s (CUDA events, max over ranks); `e2e` = the same step fed from pinned HOST buffers every
step (H2D of the next batch on a copy stream, D2H read of the result); `roofline` = the dominant kernel (tcgen05 GEMM) timed live with
event-record nodes inside an instrumented replay of the step; `cpu_baseline` = the reference's CPU path timed on the host cores
(bounded sample).  `--impl reference` prints the CPU arm as its own line: the REFERENCE's own modules through oracle/ref_loader.py when a
checkout is on the box (`kind: "reference"`), else the restated oracle port (`kind: "port"`), at the workload's batch.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

MCAN = dict(name="mcan_train", B=256, Lt=30, Nv=100, d_feature=2048, d_model=512, h=8, layers=6, d_ff=2048, vocab=4096, answers=353, dropout=0.1)
XMT = dict(name="xmt_train", B=128, Lt=30, Nv=100, d_feature=2048, d_model=512, h=8, layers=3, d_ff=2048, vocab=4096, answers=353, dropout=0.1)
M4C = dict(name="m4c_decode", B=128, Lt=20, Nobj=100, Nocr=50, d_obj=1024, hidden=768, heads=8, mmt_layers=4, text_layers=3, vocab=3000, T=12)

MCAN_WORKLOAD_NAME = ("mcan_train: MCAN 6SA+6SGA d512 h8 dff2048, 30 tok, 100x2048 regions, batch 256/GPU, dropout 0.1, "
                      "NLL+Adam (BASELINE configs[1])")
XMT_WORKLOAD_NAME = ("xmt_train: CrossModalityTransformer 3 layers d512 h8 dff2048, vision self-attention = box-geometry core (trig embedding), 30 tok, "
                     "100x2048 regions + boxes, batch 128/GPU, dropout 0.1, CE+Adam (BASELINE configs[3])")


# ----------------------------------------------------------------------------------------------- helpers
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md recipe), sampled from the first warm-up step to the end of the last timed
    region; ``window(t0, t1)`` summarises the samples of one region, ``window()`` all of them."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()

    def window(self, t0=None, t1=None):
        rows = [r for (t, r) in self.rows if len(r) >= 8 and (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        if not rows:
            return None
        num = lambda s: float(s) if s.replace(".", "").isdigit() else None
        sm = [num(r[1]) for r in rows if num(r[1]) is not None]
        pw = [num(r[3]) for r in rows if num(r[3]) is not None]
        busy = [num(r[1]) for r in rows if num(r[1]) is not None and num(r[3]) is not None and num(r[3]) > 300.0]       # sampled while the GPU was working
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return dict(sm_mhz=statistics.median(busy or sm) if sm else None, sm_mhz_min=min(busy or sm) if sm else None, sm_max_mhz=num(rows[0][2]),
                    power_w_max=max(pw) if pw else None, reasons=reasons, samples=len(rows), samples_under_load=len(busy))


def synth_batches(n, wl, seed=1, boxes=False, feature_dtype=torch.float32):
    """n distinct synthetic ViVQA-shaped batches in pinned host memory: (features, tokens, answers[, boxes]).
    ``feature_dtype=torch.bfloat16``: the region features are STAGED as bf16 (pre-padded pinned bf16 staging buffers) —
    the vision embedding's GEMM rounds its operand to bf16 anyway, so the step's results are bit-identical and the feed is half the bytes."""
    g = torch.Generator().manual_seed(seed)
    B, Nv, Lt = wl["B"], wl["Nv"], wl["Lt"]
    out = []
    for _ in range(n):
        feats = torch.randn((B, Nv, wl["d_feature"]), generator=g).abs_()
        nv = torch.randint(36, Nv + 1, (B,), generator=g)
        feats[torch.arange(Nv).unsqueeze(0) >= nv.unsqueeze(1)] = 0.0
        feats = feats.to(feature_dtype)
        toks = torch.randint(4, wl["vocab"], (B, Lt), generator=g)
        nt = torch.randint(5, Lt + 1, (B,), generator=g)
        toks[torch.arange(Lt).unsqueeze(0) >= nt.unsqueeze(1)] = 0
        ans = torch.randint(0, wl["answers"], (B,), generator=g)
        ts = [feats, toks, ans]
        if boxes:
            a = torch.rand((B, Nv, 2, 2), generator=g) * 600.0                    # pixel-like range, models/utils.py:122-123 (+1 convention)
            ts.append(torch.cat([a.min(dim=2).values, a.max(dim=2).values], dim=-1))
        out.append(tuple(t.pin_memory() if torch.cuda.is_available() else t for t in ts))
    return out


class Ctx:
    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = None
        self.sampler = None
        self.peaks = load_peaks()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        import torch.distributed as dist
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()


# ----------------------------------------------------------------------------------------------- reference / CPU arm
def _reference_available():
    from oracle import ref_loader as RL
    return RL.available()


def cpu_mcan_step_fn(wl, B, seed=0):
    """(step function, kind): the reference MCAN training step (fwd + NLL + bwd + Adam, dropout active) on the host cores — the REFERENCE's
    own classes imported through oracle/ref_loader.py when a checkout is present, else the restated oracle port of the same arithmetic."""
    from openvivqa_b200 import harness as H
    from oracle import restated as R, synth as S

    feats, toks, ans = synth_batches(1, dict(wl, B=B))[0]
    if _reference_available():
        from oracle import ref_loader as RL
        RL.load()
        import models.mcan  # noqa: F401
        from builders.model_builder import META_ARCHITECTURE

        torch.manual_seed(seed)
        model = META_ARCHITECTURE.get("MCAN")(RL.mcan_cfg(wl["d_model"], wl["d_feature"], wl["layers"], wl["h"], wl["d_ff"], wl["dropout"], device="cpu"),
                                              RL.StubVocab(wl["vocab"], wl["answers"])).train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.98))
        batch = H.Batch(region_features=feats, question_tokens=toks)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = F.nll_loss(model(batch), ans)
            loss.backward()
            opt.step()
            return loss.item()

        return step, "reference"
    spec_model = H.MCAN(H.mcan_cfg(wl["d_model"], wl["d_feature"], wl["layers"], wl["h"], wl["d_ff"]), H.StubVocab(wl["vocab"], wl["answers"]))
    params = S.synth_params([(k, tuple(v.shape)) for k, v in spec_model.state_dict().items()], seed)
    params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-4, betas=(0.9, 0.98))

    def step():
        opt.zero_grad(set_to_none=True)
        logp = R.mcan_forward(params, feats, toks, d_model=wl["d_model"], h=wl["h"], layers=wl["layers"], dropout=wl["dropout"], training=True)
        loss = F.nll_loss(logp, ans)
        loss.backward()
        opt.step()
        return loss.item()

    return step, "port"


def time_cpu(wl, B, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    step, kind = cpu_mcan_step_fn(wl, B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps, kind


def synth_m4c_batches(n, c, B, seed=5):
    """n synthetic UIT-EVJVQA-shaped batches in pinned host memory."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        toks = torch.randint(4, c["vocab"], (B, c["Lt"]), generator=g)
        nt = torch.randint(2, c["Lt"] + 1, (B,), generator=g)
        toks[torch.arange(c["Lt"]).unsqueeze(0) >= nt.unsqueeze(1)] = 0
        obj = torch.randn((B, c["Nobj"], c["d_obj"]), generator=g).abs_()
        no = torch.randint(10, c["Nobj"] + 1, (B,), generator=g)
        obj[torch.arange(c["Nobj"]).unsqueeze(0) >= no.unsqueeze(1)] = 0.0

        def boxes(N):
            a = torch.rand((B, N, 2, 2), generator=g)
            return torch.cat([a.min(dim=2).values, a.max(dim=2).values], dim=-1)

        nocr = torch.randint(0, c["Nocr"] + 1, (B,), generator=g)
        valid = (torch.arange(c["Nocr"]).unsqueeze(0) < nocr.unsqueeze(1)).unsqueeze(-1).float()
        b = dict(question_tokens=toks, region_features=obj, region_boxes=boxes(c["Nobj"]),
                 ocr_det_features=torch.randn((B, c["Nocr"], 256), generator=g) * valid,
                 ocr_rec_features=torch.randn((B, c["Nocr"], 256), generator=g) * valid,
                 ocr_fasttext_features=torch.randn((B, c["Nocr"], 300), generator=g) * valid, ocr_boxes=boxes(c["Nocr"]))
        out.append({k: (v.pin_memory() if torch.cuda.is_available() else v) for k, v in b.items()})
    return out


def m4c_workload_name(c, B):
    return (f"m4c_decode: MMF_M4C {c['Lt']} question tokens + {c['Nobj']} objects x{c['d_obj']} + {c['Nocr']} OCR tokens, hidden {c['hidden']}, "
            f"{c['heads']} heads, {c['mmt_layers']} MMT + {c['text_layers']} TextBert layers, vocab {c['vocab']}, {c['T']}-step greedy decode "
            f"(early exit off), batch {B}/GPU (BASELINE configs[2])")


def time_cpu_m4c(c, B, steps, warmup):
    """The reference greedy decode (mmf_m4c.py:231-256) on the host cores, fp32: the reference's own MMF_M4C when a checkout is present,
    else the restated oracle port."""
    from openvivqa_b200 import harness as H
    from oracle import restated as R, synth as S

    torch.set_num_threads(os.cpu_count() or 1)
    x = synth_m4c_batches(1, c, B)[0]
    if _reference_available():
        from oracle import ref_loader as RL
        RL.load_m4c()
        import models.mmf_m4c as M4

        cfg = RL.NS(DEVICE="cpu", MMT=RL.NS(HIDDEN_SIZE=c["hidden"], NUM_HIDDEN_LAYERS=c["mmt_layers"], NUM_ATTENTION_HEADS=c["heads"]),
                    TEXT_BERT=RL.NS(HIDDEN_SIZE=c["hidden"], NUM_HIDDEN_LAYERS=c["text_layers"], LOAD_PRETRAINED=False, PRETRAINED_NAME=None),
                    OBJECT_EMBEDDING=RL.NS(D_FEATURE=c["d_obj"], DROPOUT=0.1), OCR_EMBEDDING=RL.NS(D_FEATURE=812, DROPOUT=0.1),
                    OCR_PTR_NET=RL.NS(HIDDEN_SIZE=c["hidden"], QUERY_KEY_SIZE=c["hidden"]))
        torch.manual_seed(0)
        model = M4.MMF_M4C(cfg, RL.StubVocab(c["vocab"], max_answer_length=c["T"])).eval()
        model.vocab.eos_idx = -1                                   # early exit off (every sample runs all T steps), as in our arm
        items = H.Batch(batch_size=B, **x)
        kind = "reference"

        def run():
            with torch.no_grad():
                model(items)
    else:
        spec_model = H.MMF_M4C(H.m4c_cfg(c["hidden"], c["heads"], c["mmt_layers"], c["text_layers"], c["d_obj"]), H.StubVocab(c["vocab"], max_answer_length=c["T"]))
        params = S.synth_params([(k, tuple(v.shape)) for k, v in spec_model.state_dict().items()], 0)
        kind = "port"

        def run():
            with torch.no_grad():
                R.m4c_greedy_decode(params, x, max_iter=c["T"], mmt_layers=c["mmt_layers"], text_layers=c["text_layers"], heads=c["heads"],
                                    ptr_qk=c["hidden"], early_exit=False)

    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps, kind


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on this box's host cores, same config / metric as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = MCAN
    B = int(os.environ.get("OVQA_REFERENCE_ARM_BATCH", wl["B"]))   # the workload's own batch: same config as our arm (the knob is for the CPU test suite)
    warm = max(1, min(args.warmup, 1))
    sps, per, kind = time_cpu(wl, B, max(1, args.steps), warm)
    cores = torch.get_num_threads()
    m_sps, m_per, m_kind = time_cpu_m4c(M4C, int(os.environ.get("OVQA_REFERENCE_ARM_M4C_BATCH", "8")), 1, 0)
    line = dict(metric="VQA samples/sec (MCAN train)", value=sps, unit="samples/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=per * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=MCAN_WORKLOAD_NAME, global_batch=wl["B"] * max(1, args.gpus), parallelism=f"dp{max(1, args.gpus)}",
                            batch_per_step=B, note=("reference CPU path = the reference's own MCAN (models/mcan.py through oracle/ref_loader.py)" if kind == "reference"
                                                    else "reference CPU path = oracle port of the reference PyTorch step (no reference checkout on this box)")
                            + f" — fwd+NLL+bwd+Adam on all host threads, each timed step one batch of {B}; {warm} warm-up step(s) (a CPU step takes seconds)"),
                cpu_baseline=dict(value=sps, unit="samples/s", cores=cores, kind=kind, sample=f"{args.steps} training steps of batch {B}"),
                e2e=dict(value=sps, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                m4c_decode=dict(metric="VQA samples/sec (M4C decode)", value=m_sps, unit="samples/s", ms_per_step=m_per * 1e3, impl="reference",
                                config=dict(workload=m4c_workload_name(M4C, 8)),
                                cpu_baseline=dict(value=m_sps, unit="samples/s", cores=cores, kind=m_kind, sample="1 greedy decode (12 steps, early exit off) of batch 8")))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- our arm: instrumentation
class GemmMeter:
    """CUDA-event timing of every library launch of a step (events recorded on the launching stream): the tcgen05 GEMMs
    (flops per launch) and the attention / LayerNorm / reduction kernels (algorithmic bytes per launch).

    ``external=True`` makes the events graph-capturable: inside a ``torch.cuda.graph`` capture each ``record()`` becomes an
    event-record NODE, so replaying the instrumented graph times every kernel exactly where it runs in the step, back to back
    with its neighbours (eager events would also measure the host's launch gaps)."""

    def __init__(self):
        self.records, self.other, self.enabled, self.external = [], [], False, False

    def reset(self):
        self.records, self.other, self.enabled, self.external = [], [], False, False

    def _pair(self):
        return (torch.cuda.Event(enable_timing=True, external=self.external), torch.cuda.Event(enable_timing=True, external=self.external))

    def install(self):
        from openvivqa_b200 import ops
        orig = ops.gemm
        meter = self

        def gemm(a, b, **kw):
            if not meter.enabled or a.dtype != torch.bfloat16:
                return orig(a, b, **kw)
            m, k = a.shape if kw.get("a_op", "N") == "N" else (a.shape[1], a.shape[0])
            n = b.shape[0] if kw.get("b_op", "N") == "N" else b.shape[1]
            e0, e1 = meter._pair()
            e0.record()
            out = orig(a, b, **kw)
            e1.record()
            epi = "+".join(t for t, on in (("bias", kw.get("bias") is not None), ("gelu", kw.get("gelu")), ("dgelu", kw.get("dgelu_aux") is not None), ("mulaux", kw.get("mul_aux") is not None),
                                           ("save", kw.get("aux_out") is not None), ("drop", kw.get("dropout_p", 0.0) > 0), ("res", kw.get("residual") is not None),
                                           ("splitk", kw.get("split_k", 1) != 1)) if on) or "plain"
            tag = f"{kw.get('a_op', 'N')}{kw.get('b_op', 'N')} m={m} n={n} k={k} {epi} -> {str(out.dtype).replace('torch.', '')}"
            meter.records.append((e0, e1, 2.0 * m * n * k, tag))
            return out

        ops.gemm = gemm
        orig_ln = ops.gemm_ln

        def gemm_ln(a, w, *rest, **kw):          # the fused fc_o / fc2 + LayerNorm kernel: a GEMM launch like the others (2 m n k flops)
            if not meter.enabled:
                return orig_ln(a, w, *rest, **kw)
            e0, e1 = meter._pair()
            e0.record()
            out = orig_ln(a, w, *rest, **kw)
            e1.record()
            m, k, n = a.shape[0], a.shape[1], w.shape[0]
            tag = f"NN m={m} n={n} k={k} bias{'+drop' if kw.get('dropout_p', 0.0) > 0 else ''}+res+layernorm -> float32"
            meter.records.append((e0, e1, 2.0 * m * n * k, tag))
            return out

        ops.gemm_ln = gemm_ln

        def nbytes(*ts):
            return float(sum(t.numel() * t.element_size() for t in ts if t is not None))

        # algorithmic HBM bytes of the memory-bound kernels: every operand once, every result once
        def y_sdpa_fwd(a, kw, out):
            q, k, v = a[:3]
            h, dk, dv = kw["h"], kw["d_k"], kw["d_v"]
            B, Lq, Lk, e = q.shape[0], q.shape[1], k.shape[1], q.element_size()
            geo = " geo" if kw.get("geo") is not None else ""
            return e * B * h * (Lq * dk + Lk * dk + Lk * dv + Lq * dv) + 4.0 * B * Lk + 4.0 * B * h * Lq, f"sdpa_fwd{geo} B={B} Lq={Lq} Lk={Lk} h={h} d={dk}"

        def y_sdpa_bwd(a, kw, out):
            q, k, v = a[:3]
            h, dk, dv = kw["h"], kw["d_k"], kw["d_v"]
            B, Lq, Lk, e = q.shape[0], q.shape[1], k.shape[1], q.element_size()
            geo = " geo" if kw.get("geo") is not None else ""
            # q, k, v, o, do read; dq, dk, dv written (the tensor-memory core does not even read o)
            return e * B * h * (2 * Lq * dk + 2 * Lk * dk + 2 * Lk * dv + 2 * Lq * dv) + 4.0 * B * Lk + 4.0 * B * h * Lq, f"sdpa_bwd{geo} B={B} Lq={Lq} Lk={Lk} h={h} d={dk}"

        def y_ln_fwd(a, kw, out):
            x = a[0]
            return x.numel() * (4.0 + 4.0 + (2.0 if kw.get("want_bf16") else 0.0) + (4.0 if kw.get("residual") is not None else 0.0)), f"layernorm_fwd rows={x.numel() // x.shape[-1]} D={x.shape[-1]}"

        def y_ln_bwd(a, kw, out):
            x = a[1]
            return x.numel() * (4.0 + 4.0 + 4.0 + (2.0 if kw.get("want_bf16") else 0.0)), f"layernorm_bwd rows={x.numel() // x.shape[-1]} D={x.shape[-1]}"

        def y_in(name):
            return lambda a, kw, out: (nbytes(a[0]), f"{name} {tuple(a[0].shape)} {str(a[0].dtype).replace('torch.', '')}")

        def y_io(name):
            return lambda a, kw, out: (nbytes(a[0]) + nbytes(out if torch.is_tensor(out) else None), f"{name} {tuple(a[0].shape)}")

        table = dict(sdpa_fwd=y_sdpa_fwd, sdpa_bwd=y_sdpa_bwd, layernorm_fwd=y_ln_fwd, layernorm_bwd=y_ln_bwd, colsum=y_in("colsum"),
                     to_bf16=y_io("to_bf16"), to_f32=y_io("to_f32"), add=y_io("add"), dropout_apply=y_io("dropout_apply"),
                     ptr_scores=lambda a, kw, out: (nbytes(a[0], a[1]) + 4.0 * a[0].shape[0] * a[0].shape[1] * a[1].shape[1], "ptr_scores"),
                     prev_pred_embed=lambda a, kw, out: (nbytes(out if torch.is_tensor(out) else None) * 2, "prev_pred_embed"),
                     greedy_step=lambda a, kw, out: (nbytes(a[0]), "greedy_step"),
                     # geometry bias: boxes in, (B,h,N,N) fp32 out (forward) / in (backward)
                     geo_bias=lambda a, kw, out: (nbytes(a[0].boxes) + nbytes(out if torch.is_tensor(out) else None), f"geo_bias {tuple(out.shape)}"),
                     geo_bias_bwd=lambda a, kw, out: (nbytes(a[0].boxes) + nbytes(a[2]), f"geo_bias_bwd {tuple(a[2].shape)}"),
                     # LSTM recurrence (latency-bound by its 30 dependent steps, not by bytes: the GB/s column only says how little it moves):
                     # xg in; y, y16 (+ hprev, gates, c when training) out  /  dy, gates, c in; dgates out
                     lstm_fwd=lambda a, kw, out: (nbytes(a[0]) + nbytes(*[t for t in out if torch.is_tensor(t)]), f"lstm_fwd {tuple(out[0].shape)}"),
                     lstm_bwd=lambda a, kw, out: (nbytes(a[0], a[2], a[3]) + nbytes(out), f"lstm_bwd {tuple(a[0].shape)}"),
                     # attribute-reduce pooling: hidden + x in (pooled out is B x D)  /  + dx, dpre out
                     attr_pool_fwd=lambda a, kw, out: (nbytes(a[0], a[3]), f"attr_pool_fwd {tuple(a[3].shape)}"),
                     attr_pool_bwd=lambda a, kw, out: (nbytes(a[1], a[3]) + nbytes(*[t for t in out if torch.is_tensor(t)]), f"attr_pool_bwd {tuple(a[3].shape)}"))
        for name, fn_bytes in table.items():
            if not hasattr(ops, name):
                continue

            def wrap(orig_fn, fn_bytes, name):
                def w(*a, **kw):
                    if not meter.enabled:
                        return orig_fn(*a, **kw)
                    e0, e1 = meter._pair()
                    e0.record()
                    out = orig_fn(*a, **kw)
                    e1.record()
                    try:
                        y, tag = fn_bytes(a, kw, out)
                    except Exception:  # noqa: BLE001
                        y, tag = 0.0, name
                    meter.other.append((e0, e1, y, name, tag))
                    return out
                return w
            setattr(ops, name, wrap(getattr(ops, name), fn_bytes, name))

    def summary(self, steps=1):
        if not self.records:
            return None
        per = {}
        ms = fl = 0.0
        for e0, e1, f, tag in self.records:
            t = e0.elapsed_time(e1)
            ms += t
            fl += f
            c = per.setdefault(tag, [0, 0.0, 0.0])
            c[0] += 1; c[1] += t; c[2] += f
        table = sorted(({"gemm": k, "launches": v[0], "ms": round(v[1], 4), "tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1)} for k, v in per.items()), key=lambda r: -r["ms"])
        return dict(launches=len(self.records), total_ms=ms, flops=fl, tflops=fl / (ms * 1e-3) / 1e12, avg_us=ms * 1e3 / len(self.records),
                    steps=steps, table=table)

    def summary_other(self, hbm_gbs):
        """Per-kernel table of the memory-bound library launches: ms, algorithmic GB/s, fraction of the measured HBM peak."""
        per, ops_ = {}, {}
        for e0, e1, y, name, tag in self.other:
            t = e0.elapsed_time(e1)
            c = per.setdefault(tag, [0, 0.0, 0.0]); c[0] += 1; c[1] += t; c[2] += y
            o = ops_.setdefault(name, [0, 0.0, 0.0]); o[0] += 1; o[1] += t; o[2] += y
        row = lambda k, v: {"kernel": k, "launches": v[0], "ms": round(v[1], 4), "gbs": round(v[2] / (v[1] * 1e-3) / 1e9, 1) if v[1] > 0 else None,
                            "frac_hbm": round(v[2] / (v[1] * 1e-3) / 1e9 / hbm_gbs, 4) if v[1] > 0 else None}
        return (sorted((row(k, v) for k, v in per.items()), key=lambda r: -r["ms"]), {k: row(k, v) for k, v in ops_.items()})


METER = GemmMeter()


def meter_step(meter, body, use_graph, prepare=None):
    """Run ``body`` once more with an event pair around every library launch; returns (ms of the instrumented step, how, graph).
    With graphs: a second, instrumented capture (event-record nodes), replayed twice; otherwise eager events."""
    from openvivqa_b200 import functional as Fn
    meter.reset()
    meter.enabled = True
    g2 = None
    one_call, Fn.ONE_CALL_INFERENCE = Fn.ONE_CALL_INFERENCE, False      # the same kernels, enqueued op by op so that each launch can be bracketed
    try:
        if not use_graph:
            raise RuntimeError("graphs disabled")
        meter.external = True
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            body()
        torch.cuda.synchronize()
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g2.replay()
        e_a.record()
        g2.replay()
        e_b.record()
        torch.cuda.synchronize()
        ms, how = e_a.elapsed_time(e_b), ("event-record nodes around every library launch inside a second, instrumented CUDA-graph capture of the same "
                                          "step, replayed after the timed regions (per-launch times of the last replay)")
    except Exception as e:  # noqa: BLE001
        meter.records, meter.other, meter.external = [], [], False
        torch.cuda.synchronize()
        if prepare is not None:
            prepare()
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_a.record()
        body()
        e_b.record()
        torch.cuda.synchronize()
        ms, how = e_a.elapsed_time(e_b), ("CUDA events around every library launch during 1 eager step after the timed regions (host launch gaps "
                                          f"included; graph metering unavailable: {repr(e)[:80]})")
    meter.enabled = False
    Fn.ONE_CALL_INFERENCE = one_call
    return ms, how, g2


def rooflines(ctx, meter, meter_ms, meter_how, what):
    """`roofline` (dominant kernel: the tcgen05 GEMM, tensor-bound), `roofline_attention` (HBM-bound) and `step_breakdown` of one metered step."""
    peaks = ctx.peaks
    gm = meter.summary()
    roof = None
    if gm:
        burst, sust = peaks["bf16_tflops"], peaks["bf16_tflops_sustained"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = dict(dram_bytes_per_launch=tj.get("dram_bytes_per_launch"), kind="static: ncu --set full capture of one fc1 launch (dram__bytes_read.sum + "
                           "dram__bytes_write.sum), committed under profiles/ — not a counter of this run", source=tj.get("source", "profiles/gemm_traffic.json"))
        roof = dict(bound="tensor", kernel=f"gemm_tcgen05_kernel (all bf16 GEMM launches of {what})", achieved=gm["tflops"], peak=burst, unit="TFLOP/s",
                    frac=gm["tflops"] / burst, frac_burst=gm["tflops"] / burst, frac_sustained=gm["tflops"] / sust, peak_sustained=sust,
                    peak_source=peaks["source"] + " (cuBLAS bf16 8192^3: `peak` = burst figure — the timed regions last well under a second at max clocks; "
                    "`frac_sustained` uses the seconds-long power-capped figure)",
                    traffic=traffic, launches=gm["launches"], avg_launch_us=gm["avg_us"], share_of_step=gm["total_ms"] / meter_ms,
                    instrumented_ms_per_step=meter_ms, how=meter_how, by_gemm=gm["table"][:40])
    by_kernel, by_op = meter.summary_other(peaks["hbm_gbs"])
    att = None
    if "sdpa_fwd" in by_op:
        names = [n for n in ("sdpa_fwd", "sdpa_bwd") if n in by_op]
        a_ms = sum(by_op[n]["ms"] for n in names)
        a_y = sum(by_op[n]["gbs"] * by_op[n]["ms"] for n in names) * 1e6     # bytes
        att = dict(bound="hbm", kernel="attention core (sdpa_fwd_* / sdpa_bwd_* kernels), all launches of " + what,
                   achieved=a_y / (a_ms * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s", frac=a_y / (a_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                   traffic=None, share_of_step=a_ms / meter_ms, launches=sum(by_op[n]["launches"] for n in names),
                   how="same instrumented replay as `roofline`; algorithmic bytes = q,k,v,out(,dout,dq,dk,dv),lse,mask once each")
    lib_ms = (gm["total_ms"] if gm else 0.0) + sum(v["ms"] for v in by_op.values())
    breakdown = dict(instrumented_ms_per_step=meter_ms, gemm_ms=gm["total_ms"] if gm else None, by_op=by_op, library_ms=lib_ms, other_ms=meter_ms - lib_ms,
                     other="torch-native kernels around the path + gaps", by_kernel=by_kernel[:14])
    return roof, att, breakdown


# ----------------------------------------------------------------------------------------------- training workloads (MCAN, XMT)
def run_train(ctx, wl, workload_name, metric, make_model, forward_loss, boxes=False, comm_breakdown=True):
    """One data-parallel training workload: eager warm-up, whole-step CUDA graph, device-resident timing, end-to-end timing from pinned
    host buffers, instrumented replay for the rooflines.  Returns the result dict (rank 0) / None."""
    import torch.distributed as dist

    from openvivqa_b200 import _lib, dist as D, functional as Fn

    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    torch.manual_seed(0)
    model = make_model().to(dev).train()
    use_graph = not args.no_graph and (world == 1 or os.environ.get("OVQA_DDP_GRAPH", "1") == "1")
    if world > 1 and use_graph:
        # capturing DDP's bucketed NCCL all-reduce: build DDP and warm up (>= 11 iterations: bucket rebuild, logging) on a side stream
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        side0 = torch.cuda.Stream()
        side0.wait_stream(torch.cuda.current_stream())
        torch.cuda.set_stream(side0)
    net = D.data_parallel(model, device_ids=[ctx.local], bucket_cap_mb=int(os.environ.get("OVQA_DDP_BUCKET_MB", "25")), grad_dtype=os.environ.get("OVQA_DDP_GRAD_DTYPE", "bf16"),
                          sm_reserve=int(os.environ.get("OVQA_DDP_SM_RESERVE", "0")))
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.98), fused=True, capturable=use_graph)
    feat_dt = torch.float32 if os.environ.get("OVQA_FEATURE_FEED", "bf16") == "fp32" else torch.bfloat16
    host = synth_batches(3, wl, seed=1 + rank, boxes=boxes, feature_dtype=feat_dt)
    dev_batches = [tuple(t.to(dev) for t in b) for b in host]
    static = tuple(torch.empty_like(t) for t in dev_batches[0])           # the step's input buffers (graph-stable addresses)
    meter = METER

    def step_body():
        """One training step on the static input buffers: fwd + loss + bwd + Adam.  Capturable."""
        Fn.dropout_stream.advance_on_device(dev)                          # fresh dropout masks every replay (device-side offset)
        loss = forward_loss(net, static)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    def load(batch, non_blocking=False):
        for dst, src in zip(static, batch):
            dst.copy_(src, non_blocking=non_blocking)

    # ---- eager warm-up (also what a CUDA graph capture needs); the last 2 steps are timed as the eager figure
    n_eager = max(3, args.warmup) if not (world > 1 and use_graph) else max(11, args.warmup)
    for i in range(n_eager):
        if i == n_eager - 2:
            ctx.barrier()
            e_m0, e_m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_m0.record()
        load(dev_batches[i % len(dev_batches)])
        step_body()
    e_m1.record()
    ctx.barrier()
    eager_ms_per_step = e_m0.elapsed_time(e_m1) / 2.0

    # ---- capture the whole step into one CUDA graph (launch-bound otherwise: ~1.4k kernels per step)
    graph, graph_launches, graph_err = None, None, None
    if use_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                load(dev_batches[0])
                step_body()                                                # allocator / cuDNN / NCCL warm-up on the capture stream
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                static_loss = step_body()
            graph_launches = _lib.launch_count() - l0                      # our kernel nodes per replay
        except Exception as e:  # noqa: BLE001
            graph, graph_err = None, repr(e)[:200]
            torch.cuda.synchronize()

    def step(batch, non_blocking=False):
        load(batch, non_blocking)
        if graph is not None:
            graph.replay()
            return static_loss
        return step_body()

    # ---- device-resident timing (`value`)
    for i in range(args.warmup):
        step(dev_batches[i % len(dev_batches)])
    ctx.barrier()
    launches0 = _lib.launch_count()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(dev_batches[i % len(dev_batches)])
    e1.record()
    ctx.barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = (_lib.launch_count() - launches0) if graph is None else graph_launches * args.steps
    final_loss = loss.item()
    ms = ctx.max_over_ranks(ms)
    value = wl["B"] * world * args.steps / (ms * 1e-3)
    if args.no_e2e:
        del graph
        return dict(metric=metric, value=value, ms_per_step=ms / args.steps, gpu_launches=launches, note="profiling run: no e2e / roofline / cpu legs") if rank == 0 else None

    # ---- end-to-end timing: pinned host buffers -> device every step (prefetch on a copy stream), loss read back
    copy_stream = torch.cuda.Stream()
    h2d_bytes = sum(t.numel() * t.element_size() for t in host[0])
    staging = [tuple(torch.empty_like(t) for t in static) for _ in range(2)]   # double-buffered device landing zone

    def fetch(i):
        with torch.cuda.stream(copy_stream):
            b = staging[i % 2]
            for dst, src in zip(b, host[i % len(host)]):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return b, ev

    for i in range(2):
        b, ev = fetch(i)
        torch.cuda.current_stream().wait_event(ev)
        step(b)
    ctx.barrier()
    t0 = time.perf_counter()
    nxt = fetch(0)
    prev_freed = None
    for i in range(args.steps):
        b, ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        load(b)                                          # landing zone -> the step's static input buffers (compute stream)
        freed = torch.cuda.Event()
        freed.record()
        if graph is not None:
            graph.replay()
            loss = static_loss
        else:
            loss = step_body()
        if i + 1 < args.steps:
            if prev_freed is not None:                   # landing zone (i+1)%2 was last read by step i-1's load
                copy_stream.wait_event(prev_freed)
            nxt = fetch(i + 1)                           # H2D of the next batch overlaps this step
        prev_freed = freed
        _ = loss.item()                                  # D2H read of the step's result
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    t_wall2 = time.time()
    e2e = wl["B"] * world * args.steps / e2e_s

    # ---- roofline leg (kept out of the timed regions above)
    meter_ms, meter_how, g2 = meter_step(meter, step_body, graph is not None, lambda: load(dev_batches[0]))
    roof, att, breakdown = rooflines(ctx, meter, meter_ms, meter_how, "a training step") if rank == 0 else (None, None, None)

    # ---- multi-GPU: what the gradient all-reduce costs (the same step captured WITHOUT communication, and the collective alone)
    comm = None
    if world > 1 and comm_breakdown and graph is not None:
        try:
            g3 = torch.cuda.CUDAGraph()
            with net.no_sync():
                with torch.cuda.graph(g3):
                    step_body()

            def timed(fn, n=5):
                fn()
                ctx.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n):
                    fn()
                b.record()
                ctx.barrier()
                return ctx.max_over_ranks(a.elapsed_time(b) / n)

            t_sync, t_nosync = timed(graph.replay), timed(g3.replay)
            gdt = D.grad_dtype_of(net)
            nbytes_grad = sum(p.numel() for p in model.parameters() if p.requires_grad) * (2 if gdt == "bf16" else 4)
            flat = torch.zeros((nbytes_grad // 2,), device=dev, dtype=torch.bfloat16)
            t_ar = timed(lambda: dist.all_reduce(flat))
            comm = dict(step_ms=t_sync, step_without_allreduce_ms=t_nosync, exposed_ms=t_sync - t_nosync, allreduce_alone_ms=t_ar,
                        overlap=max(0.0, 1.0 - (t_sync - t_nosync) / t_ar) if t_ar > 0 else None, gradient_bytes_per_step=nbytes_grad,
                        gradient_dtype=gdt, how="whole-step CUDA graph replayed with the bucketed NCCL all-reduce vs the same step captured under "
                        "no_sync() (no collective); `allreduce_alone_ms` = one all-reduce of the same byte count with nothing else running; max over ranks")
            del g3, flat
        except Exception as e:  # noqa: BLE001
            comm = dict(error=repr(e)[:200])
            torch.cuda.synchronize()
    res = None
    if rank == 0:
        res = dict(metric=metric, value=value, unit="samples/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                   ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
                   config=dict(workload=workload_name, global_batch=wl["B"] * world, parallelism=f"dp{world}",
                               l2=f"inputs larger than L2 ({h2d_bytes / 1e6:.0f} MB per step, 3 rotating batches)", final_loss=final_loss,
                               feature_feed=("region features staged as bf16 in pinned host memory and in HBM (the vision embedding's GEMM rounds its operand to "
                                             "bf16 anyway: bit-identical step, half the feed bytes)" if feat_dt == torch.bfloat16 else "fp32 region features"),
                               cuda_graph=graph is not None, cuda_graph_error=graph_err, eager_ms_per_step=eager_ms_per_step,
                               numerics="bf16 operands / fp32 accumulate in the path's blocks; library ops around them fp32 with TF32 allowed (harness.bench_library_math)"),
                   e2e=dict(value=e2e, unit="samples/s", h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=4),
                   gpu_launches=launches, roofline=roof, roofline_attention=att, step_breakdown=breakdown,
                   clocks=ctx.sampler.window(t_wall0, t_wall2) if ctx.sampler else None,
                   clocks_timed_region=ctx.sampler.window(t_wall0, t_wall1) if ctx.sampler else None)
        if comm is not None:
            res["step_breakdown"]["allreduce"] = comm
    # graphs hold NCCL work and the optimiser's capturable state: release them before the next workload / communicator teardown
    del graph, g2
    if world > 1 and use_graph:
        torch.cuda.set_stream(torch.cuda.default_stream())
    torch.cuda.synchronize()
    return res


def run_mcan(ctx):
    from openvivqa_b200 import harness as H
    wl = MCAN

    def make_model():
        return H.MCAN(H.mcan_cfg(wl["d_model"], wl["d_feature"], wl["layers"], wl["h"], wl["d_ff"], dropout=wl["dropout"], compute_dtype="bf16"),
                      H.StubVocab(wl["vocab"], wl["answers"]))

    def forward_loss(net, static):
        feats, toks, ans = static
        logp = H.bench_forward(net, H.Batch(region_features=feats, question_tokens=toks))
        return F.nll_loss(logp.float(), ans)

    res = run_train(ctx, wl, MCAN_WORKLOAD_NAME, "VQA samples/sec (MCAN train)", make_model, forward_loss)
    if res is not None and ctx.world == 1 and not ctx.args.no_cpu_baseline and not ctx.args.no_e2e:
        sps, _, kind = time_cpu(wl, 32, 2, 1)
        res["cpu_baseline"] = dict(value=sps, unit="samples/s", cores=torch.get_num_threads(), kind=kind,
                                   sample="2 training steps (fwd+NLL+bwd+Adam) of batch 32 after 1 warm-up on all host threads "
                                          "(the --impl reference arm times the full batch of 256)")
    return res


def run_xmt(ctx):
    from openvivqa_b200 import harness as H
    wl = XMT

    def make_model():
        return H.CrossModalTransformer(H.xmt_cfg("cross", wl["d_model"], wl["d_feature"], wl["layers"], wl["h"], wl["d_ff"], dropout=wl["dropout"],
                                                 geometry=True, compute_dtype="bf16"), H.StubVocab(wl["vocab"], wl["answers"]))

    def forward_loss(net, static):
        feats, toks, ans, boxes = static
        logits = H.bench_forward(net, H.Batch(region_features=feats, question_tokens=toks, region_boxes=boxes))
        return F.cross_entropy(logits.float(), ans)

    return run_train(ctx, wl, XMT_WORKLOAD_NAME, "VQA samples/sec (XMT geometry train)", make_model, forward_loss, boxes=True, comm_breakdown=False)


# ----------------------------------------------------------------------------------------------- M4C decode (configs[2])
def run_m4c(ctx):
    from openvivqa_b200 import _lib, harness as H

    args, world, rank, dev = ctx.args, ctx.world, ctx.rank, ctx.dev
    c = M4C
    metric = "VQA samples/sec (M4C decode)"
    B = c["B"]
    torch.manual_seed(0)
    model = H.MMF_M4C(H.m4c_cfg(c["hidden"], c["heads"], c["mmt_layers"], c["text_layers"], c["d_obj"], compute_dtype="bf16"),
                      H.StubVocab(c["vocab"], max_answer_length=c["T"])).to(dev).eval()
    host = synth_m4c_batches(3, c, B, seed=5 + rank)
    dev_batches = [{k: v.to(dev) for k, v in b.items()} for b in host]
    static = {k: torch.empty_like(v) for k, v in dev_batches[0].items()}
    meter = METER

    def load(batch, non_blocking=False):
        for k, v in batch.items():
            static[k].copy_(v, non_blocking=non_blocking)

    def decode_body():
        return model(H.Batch(static), early_exit=False)["prev_inds"]

    for i in range(3):
        if i == 1:
            ctx.barrier()
            e_m0, e_m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_m0.record()
        load(dev_batches[i % 3])
        decode_body()
    e_m1.record()
    ctx.barrier()
    eager_ms = e_m0.elapsed_time(e_m1) / 2.0
    graph, graph_launches, graph_err = None, None, None
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                decode_body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(graph):
                static_out = decode_body()
            graph_launches = _lib.launch_count() - l0
        except Exception as e:  # noqa: BLE001
            graph, graph_err = None, repr(e)[:200]
            torch.cuda.synchronize()

    def step(batch, non_blocking=False):
        load(batch, non_blocking)
        if graph is not None:
            graph.replay()
            return static_out
        return decode_body()

    for i in range(args.warmup):
        step(dev_batches[i % 3])
    ctx.barrier()
    launches0 = _lib.launch_count()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = step(dev_batches[i % 3])
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = (_lib.launch_count() - launches0) if graph is None else graph_launches * args.steps
    value = B * world * args.steps / (ms * 1e-3)
    if args.no_e2e:
        del graph
        return dict(metric=metric, value=value, ms_per_step=ms / args.steps, gpu_launches=launches) if rank == 0 else None
    # end to end: host inputs -> device each step, decoded token ids back to the host
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        out = step(host[i % 3], non_blocking=True)
        _toks = out.cpu()                                    # D2H of the (B, T) int64 answers
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    t_wall2 = time.time()
    meter_ms, meter_how, g2 = meter_step(meter, decode_body, graph is not None)
    res = None
    if rank == 0:
        roof, att, breakdown = rooflines(ctx, meter, meter_ms, meter_how, "a 12-step decode")
        res = dict(metric=metric, value=value, unit="samples/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                   higher_is_better=True, scaling="weak", dtype="bf16", data="synthetic",
                   config=dict(workload=m4c_workload_name(c, B), global_batch=B * world, parallelism=f"dp{world} (batch partition, no collective)",
                               l2="3 rotating batches; weights + activations of a decode exceed L2", cuda_graph=graph is not None, cuda_graph_error=graph_err,
                               eager_ms_per_step=eager_ms, hoisted="TextBert + input encodings run once per batch (loop-invariant); last MMT layer computes only the OCR+decoder rows"),
                   e2e=dict(value=B * world * args.steps / e2e_s, unit="samples/s", h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=B * c["T"] * 8),
                   gpu_launches=launches, roofline=roof, roofline_attention=att, step_breakdown=breakdown,
                   clocks=ctx.sampler.window(t_wall0, t_wall2) if ctx.sampler else None)
        if world == 1 and not args.no_cpu_baseline:
            cs, _, kind = time_cpu_m4c(c, 4, 1, 1)
            res["cpu_baseline"] = dict(value=cs, unit="samples/s", cores=torch.get_num_threads(), kind=kind,
                                       sample="1 greedy decode (12 steps) of batch 4 after 1 warm-up on all host threads")
    del graph, g2
    torch.cuda.synchronize()
    return res


# ----------------------------------------------------------------------------------------------- fused-MHA microbench sweep (configs[4])
def _torch_mha(q, kv, mask, w, h):
    """The reference block's arithmetic in plain PyTorch ops (attentions.py:46-60, 319-339, eval): what the unmodified module launches."""
    import math
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


def run_sweep(ctx):
    """BASELINE configs[4] (compact): fused MHA block, eval, Lq 30, k = v != q, padding mask with 20 % trailing keys masked, batch 16..4096,
    Lk 100 / 600, d_model 512 / 768 — microseconds per call (CUDA events, L2 flushed), fraction of the block's roofline
    (max(F / bf16 peak, Y / HBM peak)), and the same block in plain PyTorch ops on the same GPU (fp32 = the reference as
    shipped, and bf16 autocast).  Rank 0 only (a kernel-level measurement)."""
    import openvivqa_b200 as ovqa
    from openvivqa_b200 import harness as H
    if ctx.rank != 0:
        return None
    dev, peaks = ctx.dev, ctx.peaks
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def timeit(fn, iters=10, warmup=2):
        for _ in range(warmup):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    rows = []
    for D, dk in [(512, 64), (768, 96)]:
        m = ovqa.MultiHeadAttention(H.attention_cfg(D, 8, dk, dk, dropout=0.0, compute_dtype="bf16")).to(dev).eval()
        a = m.attention
        w = {"q": (a.fc_q.weight, a.fc_q.bias), "k": (a.fc_k.weight, a.fc_k.bias), "v": (a.fc_v.weight, a.fc_v.bias),
             "o": (a.fc_o.weight, a.fc_o.bias), "ln": (m.layer_norm.weight, m.layer_norm.bias)}
        for Lk in (100, 600):
            for B in (16, 64, 256, 1024, 4096):
                if B * Lk * D * 4 > 6e9:
                    continue
                Lq = 30
                q = torch.randn((B, Lq, D), device=dev)
                kv = torch.randn((B, Lk, D), device=dev)
                mask = torch.zeros((B, 1, 1, Lk), device=dev)
                mask[:, :, :, int(Lk * 0.8):] = -10e4
                with torch.no_grad():
                    t = timeit(lambda: m(q, kv, kv, mask))
                    t32 = t16 = None
                    if B * Lk * 8 * Lq * 4 < 1.5e9:        # the reference materialises (B, h, Lq, Lk) fp32 three times
                        t32 = timeit(lambda: _torch_mha(q, kv, mask, w, 8), iters=5, warmup=1)
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            t16 = timeit(lambda: _torch_mha(q, kv, mask, w, 8), iters=5, warmup=1)
                F_mha = 4.0 * D * D * (Lq + Lk) + 4.0 * Lq * Lk * D
                Y_mha = 4 * D * (Lq + Lk + Lq) + 4 * Lk            # fp32 activations in/out as the module sees them
                t_roof = max(B * F_mha / (peaks["bf16_tflops"] * 1e12), B * Y_mha / (peaks["hbm_gbs"] * 1e9)) * 1e6
                rows.append(dict(D=D, B=B, Lk=Lk, us=round(t, 1), frac_of_roofline=round(t_roof / t, 3), torch_fp32_us=None if t32 is None else round(t32, 1),
                                 torch_bf16_autocast_us=None if t16 is None else round(t16, 1),
                                 speedup_vs_torch_bf16=None if t16 is None else round(t16 / t, 2)))
                del q, kv, mask
    big = [r for r in rows if r["B"] >= 256]
    cmp_ = [r for r in rows if r["speedup_vs_torch_bf16"] is not None]
    return dict(workload="mha_sweep: fused MultiHeadAttention block (eval), Lq 30, 8 heads, B x Lk x D sweep (BASELINE configs[4], compact)",
                points=len(rows), min_frac_of_roofline_B_ge_256=min(r["frac_of_roofline"] for r in big), max_frac_of_roofline=max(r["frac_of_roofline"] for r in rows),
                min_speedup_vs_torch_bf16_autocast=min(r["speedup_vs_torch_bf16"] for r in cmp_) if cmp_ else None,
                points_slower_than_torch_bf16_autocast=sum(1 for r in cmp_ if r["speedup_vs_torch_bf16"] < 1.0),
                roofline="max(B*F_mha / burst bf16 peak, B*Y_mha / HBM peak), F_mha = 4D^2(Lq+Lk) + 4 Lq Lk D, Y_mha = 4D(2Lq+Lk) + 4Lk (fp32 module I/O)",
                rows=rows)


# ----------------------------------------------------------------------------------------------- main
def run_ours(args):
    import torch.distributed as dist

    import openvivqa_b200 as ovqa
    from openvivqa_b200 import _lib, harness as H

    ctx = Ctx(args)
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(ctx.local)
    ctx.dev = torch.device("cuda", ctx.local)
    if ctx.world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)
    _lib.lib()                                   # fail loudly if the CUDA library is missing
    ovqa.set_compute_dtype("bf16")
    H.bench_library_math().__enter__()           # numerical setup shared with tests/test_gpu_bench_configs.py (TF32 library ops around the path, no autocast)
    METER.install()
    if ctx.rank == 0:
        ctx.sampler = ClockSampler(ctx.local).start()
    which = ["mcan_train", "m4c_decode", "xmt_train", "mha_sweep"] if args.workload == "all" else [args.workload]
    runners = dict(mcan_train=run_mcan, m4c_decode=run_m4c, xmt_train=run_xmt, mha_sweep=run_sweep)
    results = {}
    for name in which:
        try:
            results[name] = runners[name](ctx)
        except Exception as e:  # noqa: BLE001
            if name == which[0]:
                raise
            results[name] = dict(error=repr(e)[:300]) if ctx.rank == 0 else None
            torch.cuda.synchronize()
        torch.cuda.empty_cache()
    if ctx.sampler:
        ctx.sampler.stop()
    if ctx.rank == 0:
        line = results[which[0]]
        for name in which[1:]:
            line[name] = results[name]
        if ctx.sampler:
            line["clocks_whole_run"] = ctx.sampler.window()
        print(json.dumps(line), flush=True)
    finish(ctx.world)


def finish(world):
    """End of a multi-rank run: every CUDA graph that holds NCCL work has been released by its workload; synchronise, meet at a barrier,
    destroy the process group.  A watchdog ends the process if communicator teardown still stalls (seen once in round 1 with captured
    all-reduce graphs alive) — the JSON line is already printed and flushed by then."""
    if world <= 1:
        return
    import torch.distributed as dist
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    guard = threading.Timer(20.0, lambda: os._exit(0))
    guard.daemon = True
    guard.start()
    dist.destroy_process_group()
    guard.cancel()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end / roofline / cpu legs (profiling runs only)")
    ap.add_argument("--no-graph", action="store_true", help="do not capture the step into a CUDA graph (eager launches)")
    ap.add_argument("--workload", default="all", choices=["all", "mcan_train", "m4c_decode", "xmt_train", "mha_sweep"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
