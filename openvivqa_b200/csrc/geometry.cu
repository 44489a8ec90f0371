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

// embedding of one pair into registers: trig: [sin(100 d_t f_k)]_{t,k} ++ [cos(...)]_{t,k}; raw: the four log-deltas.  NF = d_g / 8 is a
// template parameter so that every index into e[] is a compile-time constant (a run-time NF put the 64-float array in local memory:
// 0.21 ms per forward launch at 128 x 100 x 100 pairs, profiles/r02h)
template <int NF>
__device__ __forceinline__ void pair_embedding(const GeoArg& g, const BoxGeom& bi, const BoxGeom& bj, float (&e)[NF > 0 ? 8 * NF : 4]) {
    float d[4];
    box_deltas(bi, bj, d);
    if constexpr (NF > 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float base = 100.0f * d[t];
#pragma unroll
            for (int kf = 0; kf < NF; ++kf) {
                float s, c;
                sincosf(base * geo_freq(g, kf, NF), &s, &c);          // accurate: arguments reach ~690 rad
                e[t * NF + kf] = s;
                e[4 * NF + t * NF + kf] = c;
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) e[t] = d[t];
    }
}

template <int NF>
__global__ void __launch_bounds__(256) geo_bias_fwd_kernel(GeoArg g, const float* __restrict__ key_mask, int64_t mask_sb, float* __restrict__ out,
                                                           int64_t B, int N, int h) {
    constexpr int DG = NF > 0 ? 8 * NF : 4;
    __shared__ __align__(16) float sw[kMaxHeads * kMaxDg + kMaxHeads];
    for (int i = threadIdx.x; i < h * DG; i += blockDim.x) sw[i] = g.w[i];
    for (int i = threadIdx.x; i < h; i += blockDim.x) sw[h * DG + i] = g.b[i];
    __syncthreads();
    const int64_t pairs = B * N * N;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(p % N), i = (int)((p / N) % N);
        const int64_t b = p / ((int64_t)N * N);
        float e[DG];
        pair_embedding<NF>(g, box_geom(g.boxes + (b * N + i) * 4), box_geom(g.boxes + (b * N + j) * 4), e);
        const float mk = key_mask ? __ldg(key_mask + b * mask_sb + j) : 0.f;
        for (int hd = 0; hd < h; ++hd) {
            const float4* w4 = reinterpret_cast<const float4*>(sw + hd * DG);      // DG % 4 == 0: broadcast 16-byte reads of the weight row
            float raw = 0.f;
#pragma unroll
            for (int c = 0; c < DG / 4; ++c) {
                const float4 w = w4[c];
                raw = fmaf(w.x, e[4 * c], fmaf(w.y, e[4 * c + 1], fmaf(w.z, e[4 * c + 2], fmaf(w.w, e[4 * c + 3], raw))));
            }
            raw += sw[h * DG + hd];
            out[((b * h + hd) * N + i) * N + j] = geo_log_bias(raw) + mk;
        }
    }
}

// tile of TP pairs per iteration: phase 1 (thread = pair) embedding + per-head coefficient into shared memory; phase 2: thread (c, part)
// owns embedding column c for ALL heads (h accumulators) over a quarter of the tile's pairs — one shared-memory read of the embedding and
// two broadcast 16-byte reads of the coefficients per 8 multiply-adds.  The parts are summed through shared memory once, at the end.
constexpr int TP = 256;
template <int NF, int H>
__global__ void __launch_bounds__(TP) geo_bias_bwd_kernel(GeoArg g, const float* __restrict__ dbias, float* __restrict__ dw, float* __restrict__ db,
                                                          int64_t B, int N) {
    constexpr int DG = NF > 0 ? 8 * NF : 4, PITCH = DG + 1, COLS = DG + 1 /* + the constant-1 column that pairs with the bias */, PARTS = TP / 64;
    static_assert(COLS <= 64 + 1, "one thread per embedding column");
    extern __shared__ __align__(16) float sm[];
    float* sE = sm;                          // [TP][DG + 1]
    float* sC = sE + TP * PITCH;             // [TP][H]
    float* sw = sC + TP * H;                 // [H][DG] + [H]
    for (int i = threadIdx.x; i < H * DG; i += blockDim.x) sw[i] = g.w[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) sw[H * DG + i] = g.b[i];
    const int col = threadIdx.x % 64, part = threadIdx.x / 64;      // phase-2 role: column (col < DG: embedding column; DG: the bias), quarter of the pairs
    // the bias column is handled by the threads with col == 0 in a second accumulator set (DG may be 64: no spare thread column)
    float acc[H], accb[H];
#pragma unroll
    for (int r = 0; r < H; ++r) acc[r] = accb[r] = 0.f;
    const int64_t pairs = B * N * N, tiles = (pairs + TP - 1) / TP;
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t p = tile * TP + threadIdx.x;
        if (p < pairs) {
            const int j = (int)(p % N), i = (int)((p / N) % N);
            const int64_t b = p / ((int64_t)N * N);
            float e[DG];
            pair_embedding<NF>(g, box_geom(g.boxes + (b * N + i) * 4), box_geom(g.boxes + (b * N + j) * 4), e);
#pragma unroll
            for (int c = 0; c < DG; ++c) sE[threadIdx.x * PITCH + c] = e[c];
#pragma unroll
            for (int hd = 0; hd < H; ++hd) {
                const float4* w4 = reinterpret_cast<const float4*>(sw + hd * DG);
                float raw = 0.f;
#pragma unroll
                for (int c = 0; c < DG / 4; ++c) {
                    const float4 w = w4[c];
                    raw = fmaf(w.x, e[4 * c], fmaf(w.y, e[4 * c + 1], fmaf(w.z, e[4 * c + 2], fmaf(w.w, e[4 * c + 3], raw))));
                }
                raw += sw[H * DG + hd];
                // d log(clamp(relu(raw), 1e-6)) / d raw = 1 / raw where raw > 1e-6, else 0
                sC[threadIdx.x * H + hd] = raw > 1e-6f ? __ldg(dbias + ((b * H + hd) * N + i) * N + j) / raw : 0.f;
            }
        } else {
#pragma unroll
            for (int hd = 0; hd < H; ++hd) sC[threadIdx.x * H + hd] = 0.f;
#pragma unroll
            for (int c = 0; c < DG; ++c) sE[threadIdx.x * PITCH + c] = 0.f;
        }
        __syncthreads();
        if (col < DG) {
            for (int q = part; q < TP; q += PARTS) {
                const float ev = sE[q * PITCH + col];
#pragma unroll
                for (int hd = 0; hd < H; ++hd) acc[hd] = fmaf(sC[q * H + hd], ev, acc[hd]);
            }
        }
        if (col == 0) {
            for (int q = part; q < TP; q += PARTS)
#pragma unroll
                for (int hd = 0; hd < H; ++hd) accb[hd] += sC[q * H + hd];
        }
        __syncthreads();
    }
    if (col < DG) {
#pragma unroll
        for (int hd = 0; hd < H; ++hd) atomicAdd(dw + hd * DG + col, acc[hd]);
    }
    if (col == 0) {
#pragma unroll
        for (int hd = 0; hd < H; ++hd) atomicAdd(db + hd, accb[hd]);
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
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const GeoArg g{boxes, geo_w, geo_b, geo_freq, (int)d_g, geo_trig};
    cudaStream_t st = (cudaStream_t)stream;
#define OVQA_GEO_FWD(NF_) geo_bias_fwd_kernel<NF_><<<grid, 256, 0, st>>>(g, key_mask, key_mask_sb, out, B, (int)N, (int)h)
    switch (geo_trig ? (int)d_g / 8 : 0) {
        case 0: OVQA_GEO_FWD(0); break;
        case 1: OVQA_GEO_FWD(1); break;
        case 2: OVQA_GEO_FWD(2); break;
        case 4: OVQA_GEO_FWD(4); break;
        case 8: OVQA_GEO_FWD(8); break;
        default: return OVQA_ERR_BAD_SHAPE;          // d_g = D / h: 8, 16, 32 or 64 with the trigonometric embedding
    }
#undef OVQA_GEO_FWD
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
    const int64_t tiles = (pairs + TP - 1) / TP, cap = (int64_t)sm_count() * 2;
    const unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
    const GeoArg g{boxes, geo_w, geo_b, geo_freq, (int)d_g, geo_trig};
    cudaStream_t st = (cudaStream_t)stream;
    // instantiated for the head counts / embedding widths of the reference's configs (8 heads of d_g = 64; 4 heads and the raw 4-d variant for tests)
#define OVQA_GEO_BWD(NF_, H_)                                                                                                        \
    do {                                                                                                                             \
        static OncePerDevice attr_set;                                                                                               \
        if (attr_set.pending()) { OVQA_CUDA_OK(cudaFuncSetAttribute(geo_bias_bwd_kernel<NF_, H_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr_set.mark(); } \
        geo_bias_bwd_kernel<NF_, H_><<<grid, TP, smem, st>>>(g, dbias, geo_dw, geo_db, B, (int)N);                                   \
    } while (0)
    const int nf = geo_trig ? (int)d_g / 8 : 0;
    if (nf == 8 && h == 8) OVQA_GEO_BWD(8, 8);
    else if (nf == 8 && h == 4) OVQA_GEO_BWD(8, 4);
    else if (nf == 4 && h == 4) OVQA_GEO_BWD(4, 4);
    else if (nf == 4 && h == 8) OVQA_GEO_BWD(4, 8);
    else if (nf == 8 && h == 12) OVQA_GEO_BWD(8, 12);
    else if (nf == 8 && h == 16) OVQA_GEO_BWD(8, 16);
    else if (nf == 0 && h == 4) OVQA_GEO_BWD(0, 4);
    else if (nf == 0 && h == 8) OVQA_GEO_BWD(0, 8);
    else return OVQA_ERR_BAD_SHAPE;
#undef OVQA_GEO_BWD
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
