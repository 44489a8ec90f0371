// Box-geometry attention bias as its own pair of kernels (attentions.py:113-119,131-132 + models/utils.py:102-162):
//     bias[b,h,i,j] = log(clamp(relu(w_h . emb(box_i, box_j) + b_h), 1e-6))  [+ key_mask[b,j]]
// The 64-d trigonometric pair embedding does not depend on the head or the layer's weights, so it is evaluated ONCE per pair and dotted
// against all h weight rows; the (B,h,N,N) fp32 result goes to the tensor-core attention kernels as a per-(head, query, key) additive
// mask.  Backward: the attention kernel returns dS (= d bias); geo_bias_bwd re-evaluates the embedding per pair and reduces
//     dw[h,:] += sum_pairs dS/raw * emb,   db[h] += sum_pairs dS/raw        (where raw > 1e-6)
// inside the CTA (shared-memory tile of embeddings x coefficients, register accumulators per (h, g)), one atomic per CTA and output.
// Round 1 recomputed the embedding per head INSIDE the CUDA-core attention kernel and accumulated the weight gradient with one atomic
// per (pair, head, frequency): 51 ms per backward launch at batch 128 x 100 regions (profiles/r02c); this path costs ~0.1 ms.
#include "attention_simt.cuh"

namespace ovqa {
namespace {

constexpr int kMaxDg = 64, kMaxHeads = 16;

// embedding of one pair into registers: trig: [sin(100 d_t f_k)]_{t,k} ++ [cos(...)]_{t,k}; raw: the four log-deltas
__device__ __forceinline__ void pair_embedding(const GeoArg& g, const BoxGeom& bi, const BoxGeom& bj, float (&e)[kMaxDg]) {
    float d[4];
    box_deltas(bi, bj, d);
    if (g.trig) {
        const int nf = g.d_g / 8;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float base = 100.0f * d[t];
#pragma unroll
            for (int kf = 0; kf < kMaxDg / 8; ++kf) {
                if (kf < nf) {
                    float s, c;
                    sincosf(base * geo_freq(g, kf, nf), &s, &c);          // accurate: arguments reach ~690 rad
                    e[t * nf + kf] = s;
                    e[4 * nf + t * nf + kf] = c;
                }
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) e[t] = d[t];
    }
}

__global__ void __launch_bounds__(256) geo_bias_fwd_kernel(GeoArg g, const float* __restrict__ key_mask, int64_t mask_sb, float* __restrict__ out,
                                                           int64_t B, int N, int h) {
    __shared__ float sw[kMaxHeads * kMaxDg + kMaxHeads];
    for (int i = threadIdx.x; i < h * g.d_g; i += blockDim.x) sw[i] = g.w[i];
    for (int i = threadIdx.x; i < h; i += blockDim.x) sw[h * g.d_g + i] = g.b[i];
    __syncthreads();
    const int64_t pairs = B * N * N;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(p % N), i = (int)((p / N) % N);
        const int64_t b = p / ((int64_t)N * N);
        float e[kMaxDg];
        pair_embedding(g, box_geom(g.boxes + (b * N + i) * 4), box_geom(g.boxes + (b * N + j) * 4), e);
        const float mk = key_mask ? __ldg(key_mask + b * mask_sb + j) : 0.f;
        for (int hd = 0; hd < h; ++hd) {
            const float* w = sw + hd * g.d_g;
            float raw = 0.f;
#pragma unroll
            for (int c = 0; c < kMaxDg; ++c)
                if (c < g.d_g) raw = fmaf(w[c], e[c], raw);
            raw += sw[h * g.d_g + hd];
            out[((b * h + hd) * N + i) * N + j] = geo_log_bias(raw) + mk;
        }
    }
}

// tile of TP pairs per iteration: phase 1 (thread = pair) embedding + per-head coefficient into shared memory, phase 2 (thread = a few
// (h, g) outputs) accumulates coef^T . emb over the tile's pairs in registers.
constexpr int TP = 256;
__global__ void __launch_bounds__(TP) geo_bias_bwd_kernel(GeoArg g, const float* __restrict__ dbias, float* __restrict__ dw, float* __restrict__ db,
                                                          int64_t B, int N, int h) {
    extern __shared__ float sm[];
    const int dg = g.d_g, pitch = dg + 1;
    float* sE = sm;                          // [TP][dg + 1]
    float* sC = sE + TP * pitch;             // [TP][h]
    float* sw = sC + TP * h;                 // [h][dg] + [h]
    for (int i = threadIdx.x; i < h * dg; i += blockDim.x) sw[i] = g.w[i];
    for (int i = threadIdx.x; i < h; i += blockDim.x) sw[h * dg + i] = g.b[i];
    const int nout = h * dg + h;             // outputs: dw[h][dg] then db[h] (the bias pairs with a constant-1 embedding column)
    float acc[(kMaxHeads * kMaxDg + kMaxHeads + TP - 1) / TP];
#pragma unroll
    for (int r = 0; r < (int)(sizeof(acc) / sizeof(float)); ++r) acc[r] = 0.f;
    const int64_t pairs = B * N * N, tiles = (pairs + TP - 1) / TP;
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t p = tile * TP + threadIdx.x;
        if (p < pairs) {
            const int j = (int)(p % N), i = (int)((p / N) % N);
            const int64_t b = p / ((int64_t)N * N);
            float e[kMaxDg];
            pair_embedding(g, box_geom(g.boxes + (b * N + i) * 4), box_geom(g.boxes + (b * N + j) * 4), e);
#pragma unroll
            for (int c = 0; c < kMaxDg; ++c)
                if (c < dg) sE[threadIdx.x * pitch + c] = e[c];
            for (int hd = 0; hd < h; ++hd) {
                const float* w = sw + hd * dg;
                float raw = 0.f;
#pragma unroll
                for (int c = 0; c < kMaxDg; ++c)
                    if (c < dg) raw = fmaf(w[c], e[c], raw);
                raw += sw[h * dg + hd];
                // d log(clamp(relu(raw), 1e-6)) / d raw = 1 / raw where raw > 1e-6, else 0
                sC[threadIdx.x * h + hd] = raw > 1e-6f ? __ldg(dbias + ((b * h + hd) * N + i) * N + j) / raw : 0.f;
            }
        } else {
            for (int hd = 0; hd < h; ++hd) sC[threadIdx.x * h + hd] = 0.f;
        }
        __syncthreads();
        const int np = (int)min((int64_t)TP, pairs - tile * TP);
#pragma unroll
        for (int r = 0; r < (int)(sizeof(acc) / sizeof(float)); ++r) {
            const int o = threadIdx.x + r * TP;
            if (o < nout) {
                float a = acc[r];
                if (o < h * dg) {
                    const int hd = o / dg, c = o % dg;
                    for (int q = 0; q < np; ++q) a = fmaf(sC[q * h + hd], sE[q * pitch + c], a);
                } else {
                    const int hd = o - h * dg;
                    for (int q = 0; q < np; ++q) a += sC[q * h + hd];
                }
                acc[r] = a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < (int)(sizeof(acc) / sizeof(float)); ++r) {
        const int o = threadIdx.x + r * TP;
        if (o < h * dg) atomicAdd(dw + o, acc[r]);
        else if (o < nout) atomicAdd(db + (o - h * dg), acc[r]);
    }
}

}  // namespace
}  // namespace ovqa

using namespace ovqa;

static int check_geo(const float* boxes, const float* w, const float* b, int64_t B, int64_t N, int64_t h, int64_t d_g, int trig) {
    if (!boxes || !w || !b || B < 0 || N < 0 || h <= 0) return OVQA_ERR_BAD_ARG;
    if (h > kMaxHeads || d_g > kMaxDg) return OVQA_ERR_BAD_SHAPE;
    if (trig ? (d_g % 8 != 0 || d_g <= 0) : d_g != 4) return OVQA_ERR_BAD_SHAPE;
    return OVQA_OK;
}

extern "C" int ovqa_geo_bias_fwd(const float* boxes, const float* geo_w, const float* geo_b, const float* geo_freq, const float* key_mask,
                                 int64_t key_mask_sb, float* out, int64_t B, int64_t N, int64_t h, int64_t d_g, int geo_trig, void* stream) {
    int rc = check_geo(boxes, geo_w, geo_b, B, N, h, d_g, geo_trig);
    if (rc) return rc;
    if (!out) return OVQA_ERR_BAD_ARG;
    const int64_t pairs = B * N * N;
    if (pairs == 0) return OVQA_OK;
    const int64_t want = (pairs + 255) / 256, cap = (int64_t)sm_count() * 8;
    geo_bias_fwd_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>({boxes, geo_w, geo_b, geo_freq, (int)d_g, geo_trig}, key_mask,
                                                                                             key_mask_sb, out, B, (int)N, (int)h);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

extern "C" int ovqa_geo_bias_bwd(const float* boxes, const float* geo_w, const float* geo_b, const float* geo_freq, const float* dbias,
                                 float* geo_dw, float* geo_db, int64_t B, int64_t N, int64_t h, int64_t d_g, int geo_trig, void* stream) {
    int rc = check_geo(boxes, geo_w, geo_b, B, N, h, d_g, geo_trig);
    if (rc) return rc;
    if (!dbias || !geo_dw || !geo_db) return OVQA_ERR_BAD_ARG;
    const int64_t pairs = B * N * N;
    if (pairs == 0) return OVQA_OK;
    const size_t smem = sizeof(float) * ((size_t)TP * (d_g + 1) + (size_t)TP * h + (size_t)h * d_g + h);
    static OncePerDevice attr_set;
    if (attr_set.pending()) { OVQA_CUDA_OK(cudaFuncSetAttribute(geo_bias_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr_set.mark(); }
    const int64_t tiles = (pairs + TP - 1) / TP, cap = (int64_t)sm_count() * 2;
    geo_bias_bwd_kernel<<<(unsigned)(tiles < cap ? tiles : cap), TP, smem, (cudaStream_t)stream>>>({boxes, geo_w, geo_b, geo_freq, (int)d_g, geo_trig}, dbias,
                                                                                                  geo_dw, geo_db, B, (int)N, (int)h);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
