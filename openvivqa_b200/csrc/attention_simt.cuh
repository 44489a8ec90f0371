// Shared pieces of the attention-core kernels: mask addressing and the on-the-fly box-geometry log-bias.
#pragma once
#include "common.cuh"

namespace ovqa {

struct MaskArg {
    const float* ptr;
    int64_t sb, sh, sq;  // element strides over batch / head / query row (0 = broadcast); keys are unit stride
    // prefix-LM structure (ovqa_sdpa_fwd_prefix_lm): on top of a per-key row, the last `causal_tail` queries x keys form a causal block
    // whose entries above the diagonal get `causal_value` added
    int64_t causal_tail;
    float causal_value;
};

struct GeoArg {
    const float* boxes;  // [B, N, 4]
    const float* w;      // [h, d_g]
    const float* b;      // [h]
    const float* freq;   // [d_g/8] or NULL
    int d_g;
    int trig;
};

struct SdpaArgs {
    const void *q, *k, *v;
    void* out;
    int64_t B, h, Lq, Lk, dk, dv;
    int64_t ldq, ldk, ldv, ldo;
    MaskArg mask;
    GeoArg geo;
    float scale;
    float* lse;
    float* att;
    // dropout on the PROBABILITIES (HF BertSelfAttention: attention_probs_dropout_prob; the repo's own cores have none): element
    // ((b*h + head)*Lq + i)*Lk + j of the library's counter-based keep stream.  CUDA-core kernels only.
    float pdrop;
    uint64_t pdrop_seed, pdrop_offset;
    const uint64_t* pdrop_offset_base;
};

// models/utils.py:118-137 per box: centre, and (extent + 1)
struct BoxGeom { float cx, cy, w, h; };
__device__ __forceinline__ BoxGeom box_geom(const float* b4) {
    const float x0 = b4[0], y0 = b4[1], x1 = b4[2], y1 = b4[3];
    return {(x0 + x1) * 0.5f, (y0 + y1) * 0.5f, (x1 - x0) + 1.0f, (y1 - y0) + 1.0f};
}
// the four log-deltas of pair (i, j)  (models/utils.py:126-137): i indexes rows (queries), j columns (keys)
__device__ __forceinline__ void box_deltas(const BoxGeom& bi, const BoxGeom& bj, float d[4]) {
    d[0] = logf(fmaxf(fabsf((bi.cx - bj.cx) / bi.w), 1e-3f));
    d[1] = logf(fmaxf(fabsf((bi.cy - bj.cy) / bi.h), 1e-3f));
    d[2] = logf(bi.w / bj.w);
    d[3] = logf(bi.h / bj.h);
}
__device__ __forceinline__ float geo_freq(const GeoArg& g, int kf, int n_freq) {
    return g.freq ? __ldg(g.freq + kf) : 1.0f / powf(1000.0f, (float)kf / (float)n_freq);
}
// raw (pre-relu) geometry weight  w_h . emb(i,j) + b_h   (attentions.py:116, models/utils.py:146-158).
// Accurate sinf/cosf on purpose: arguments reach |100*log(1e-3)| ~ 690.
__device__ __forceinline__ float geo_raw(const GeoArg& g, int head, const BoxGeom& bi, const BoxGeom& bj) {
    float d[4];
    box_deltas(bi, bj, d);
    const float* w = g.w + (int64_t)head * g.d_g;
    float acc = 0.0f;
    if (g.trig) {
        const int nf = g.d_g / 8;
        for (int t = 0; t < 4; ++t) {
            const float base = 100.0f * d[t];
            for (int kf = 0; kf < nf; ++kf) {
                float s, c;
                sincosf(base * geo_freq(g, kf, nf), &s, &c);
                acc = fmaf(__ldg(w + t * nf + kf), s, acc);
                acc = fmaf(__ldg(w + 4 * nf + t * nf + kf), c, acc);
            }
        }
    } else {
        for (int t = 0; t < 4; ++t) acc = fmaf(__ldg(w + t), d[t], acc);
    }
    return acc + __ldg(g.b + head);
}
__device__ __forceinline__ float geo_log_bias(float raw) { return logf(fmaxf(fmaxf(raw, 0.0f), 1e-6f)); }

}  // namespace ovqa
