// Attribute-reduce pooling head (models/mcan.py:70-77, the same lines in cross_modality_transformer.py:67-75 and
// visiolinguistic_transformer.py:67-75):
//     logits[b,n] = hidden[b,n,:] . w2 + b2        (hidden = dropout(relu(fc1 x)), the MLP's first layer: one tensor-core GEMM with the
//                                                   ReLU + dropout epilogue)
//     p[b,:]      = softmax over the sequence (no padding mask, as in the reference)
//     pooled[b,:] = sum_n p[b,n] x[b,n,:]
// One CTA per sample: the N logits, their softmax and the weighted sum happen in one pass over the sample's rows; the reference runs
// fc2 (a 512 -> 1 GEMV), softmax, a broadcast multiply that materialises (B,N,D) and a sum — five kernels and three (B,N,D) fp32 round trips.
// Backward: dx = p dpooled^T,  dlogit = p (x.dpooled - sum p x.dpooled),  d pre-activation of fc1 = dlogit w2^T relu'/dropout mask
// (recovered from hidden itself: hidden > 0 <=> the unit was active AND kept), dw2 / db2 reduced per CTA.
#include "common.cuh"

namespace ovqa {
namespace {

constexpr int kPoolThreads = 256;

template <typename TH>
__global__ void __launch_bounds__(kPoolThreads) attr_pool_fwd_kernel(const TH* __restrict__ hid, const float* __restrict__ w2, const float* __restrict__ b2,
                                                                     const float* __restrict__ x, float* __restrict__ pooled, float* __restrict__ probs,
                                                                     int N, int Dh, int D) {
    extern __shared__ float sp[];                   // [N] logits -> probabilities
    __shared__ float red[kPoolThreads / 32];
    const int64_t b = blockIdx.x;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = kPoolThreads / 32;
    const TH* hb = hid + b * (int64_t)N * Dh;
    for (int n = warp; n < N; n += nw) {
        float s = 0.f;
        for (int c = lane; c < Dh; c += 32) s = fmaf(to_f32<TH>(hb[(int64_t)n * Dh + c]), __ldg(w2 + c), s);
        s = warp_sum(s);
        if (lane == 0) sp[n] = s + __ldg(b2);
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) mx = fmaxf(mx, sp[n]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < nw; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) { const float e = expf(sp[n] - mx); sp[n] = e; sum += e; }
    sum = warp_sum(sum);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    sum = 0.f;
    for (int i = 0; i < nw; ++i) sum += red[i];
    const float inv = 1.f / sum;
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += kPoolThreads) { const float p = sp[n] * inv; sp[n] = p; if (probs) probs[b * N + n] = p; }
    __syncthreads();
    const float* xb = x + b * (int64_t)N * D;
    for (int c = threadIdx.x; c < D; c += kPoolThreads) {          // consecutive threads -> consecutive columns: every row read is coalesced
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;           // four rows in flight per thread: the loop is latency-bound otherwise
        int n = 0;
        for (; n + 4 <= N; n += 4) {
            const float x0 = __ldg(xb + (int64_t)n * D + c), x1 = __ldg(xb + (int64_t)(n + 1) * D + c), x2 = __ldg(xb + (int64_t)(n + 2) * D + c),
                        x3 = __ldg(xb + (int64_t)(n + 3) * D + c);
            a0 = fmaf(sp[n], x0, a0); a1 = fmaf(sp[n + 1], x1, a1); a2 = fmaf(sp[n + 2], x2, a2); a3 = fmaf(sp[n + 3], x3, a3);
        }
        for (; n < N; ++n) a0 = fmaf(sp[n], __ldg(xb + (int64_t)n * D + c), a0);
        pooled[b * D + c] = (a0 + a1) + (a2 + a3);
    }
}

template <typename TH>
__global__ void __launch_bounds__(kPoolThreads) attr_pool_bwd_kernel(const float* __restrict__ dpooled, const TH* __restrict__ hid, const float* __restrict__ w2,
                                                                     const float* __restrict__ x, const float* __restrict__ probs, float* __restrict__ dx,
                                                                     TH* __restrict__ dpre, float* __restrict__ dw2, float* __restrict__ db2, int N, int Dh,
                                                                     int D, float inv_keep) {
    extern __shared__ float sd[];                   // [N] x.dpooled -> dlogit
    __shared__ float red[kPoolThreads / 32];
    const int64_t b = blockIdx.x;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = kPoolThreads / 32;
    const float* xb = x + b * (int64_t)N * D;
    const float* dpb = dpooled + b * D;
    const float* pb = probs + b * N;
    for (int n = warp; n < N; n += nw) {
        float s = 0.f;
        for (int c = lane; c < D; c += 32) s = fmaf(__ldg(xb + (int64_t)n * D + c), __ldg(dpb + c), s);
        s = warp_sum(s);
        if (lane == 0) sd[n] = s;
    }
    __syncthreads();
    float dot = 0.f;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) dot = fmaf(__ldg(pb + n), sd[n], dot);
    dot = warp_sum(dot);
    if (lane == 0) red[warp] = dot;
    __syncthreads();
    dot = 0.f;
    for (int i = 0; i < nw; ++i) dot += red[i];
    __syncthreads();
    float dl_sum = 0.f;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) { const float dl = __ldg(pb + n) * (sd[n] - dot); sd[n] = dl; dl_sum += dl; }
    dl_sum = warp_sum(dl_sum);
    if (lane == 0) red[warp] = dl_sum;
    __syncthreads();
    if (threadIdx.x == 0 && db2) { float t = 0.f; for (int i = 0; i < nw; ++i) t += red[i]; atomicAdd(db2, t); }
    // dx[b,n,:] = p[n] dpooled
    if (dx) {
        float* dxb = dx + b * (int64_t)N * D;
        for (int c = threadIdx.x; c < D; c += kPoolThreads) {
            const float g = __ldg(dpb + c);
#pragma unroll 8
            for (int n = 0; n < N; ++n) dxb[(int64_t)n * D + c] = __ldg(pb + n) * g;
        }
    }
    // d pre-activation = dlogit[n] w2[c] where the unit was active and kept (x the dropout keep scale);  dw2[c] += sum_n dlogit[n] hidden[b,n,c]
    const TH* hb = hid + b * (int64_t)N * Dh;
    TH* db = dpre + b * (int64_t)N * Dh;
    for (int c = threadIdx.x; c < Dh; c += kPoolThreads) {
        const float w = __ldg(w2 + c) * inv_keep;
        float acc = 0.f;
#pragma unroll 8
        for (int n = 0; n < N; ++n) {
            const float hv = to_f32<TH>(hb[(int64_t)n * Dh + c]), dl = sd[n];
            acc = fmaf(dl, hv, acc);
            db[(int64_t)n * Dh + c] = from_f32<TH>(hv > 0.f ? dl * w : 0.f);
        }
        if (dw2) atomicAdd(dw2 + c, acc);
    }
}

}  // namespace
}  // namespace ovqa

using namespace ovqa;

extern "C" int ovqa_attr_pool_fwd(const void* hidden, const float* w2, const float* b2, const float* x, float* pooled, float* probs, int64_t B, int64_t N,
                                  int64_t Dh, int64_t D, int hidden_dtype, void* stream) {
    if (!hidden || !w2 || !b2 || !x || !pooled || B < 0 || N <= 0 || Dh <= 0 || D <= 0) return OVQA_ERR_BAD_ARG;
    if (hidden_dtype != OVQA_F32 && hidden_dtype != OVQA_BF16) return OVQA_ERR_BAD_ARG;
    if (N * sizeof(float) > 48 * 1024 || B > 0x7fffffff) return OVQA_ERR_BAD_SHAPE;
    if (B == 0) return OVQA_OK;
    const size_t smem = (size_t)N * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (hidden_dtype == OVQA_F32)
        attr_pool_fwd_kernel<float><<<(unsigned)B, kPoolThreads, smem, st>>>((const float*)hidden, w2, b2, x, pooled, probs, (int)N, (int)Dh, (int)D);
    else
        attr_pool_fwd_kernel<__nv_bfloat16><<<(unsigned)B, kPoolThreads, smem, st>>>((const __nv_bfloat16*)hidden, w2, b2, x, pooled, probs, (int)N, (int)Dh, (int)D);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

extern "C" int ovqa_attr_pool_bwd(const float* dpooled, const void* hidden, const float* w2, const float* x, const float* probs, float* dx, void* dpre,
                                  float* dw2, float* db2, int64_t B, int64_t N, int64_t Dh, int64_t D, float dropout_p, int hidden_dtype, void* stream) {
    if (!dpooled || !hidden || !w2 || !x || !probs || !dpre || B < 0 || N <= 0 || Dh <= 0 || D <= 0) return OVQA_ERR_BAD_ARG;
    if (hidden_dtype != OVQA_F32 && hidden_dtype != OVQA_BF16) return OVQA_ERR_BAD_ARG;
    if (dropout_p < 0.f || dropout_p >= 1.f) return OVQA_ERR_BAD_ARG;
    if (N * sizeof(float) > 48 * 1024 || B > 0x7fffffff) return OVQA_ERR_BAD_SHAPE;
    if (B == 0) return OVQA_OK;
    const size_t smem = (size_t)N * sizeof(float);
    const float inv_keep = 1.0f / (1.0f - dropout_p);
    cudaStream_t st = (cudaStream_t)stream;
    if (hidden_dtype == OVQA_F32)
        attr_pool_bwd_kernel<float><<<(unsigned)B, kPoolThreads, smem, st>>>(dpooled, (const float*)hidden, w2, x, probs, dx, (float*)dpre, dw2, db2, (int)N,
                                                                             (int)Dh, (int)D, inv_keep);
    else
        attr_pool_bwd_kernel<__nv_bfloat16><<<(unsigned)B, kPoolThreads, smem, st>>>(dpooled, (const __nv_bfloat16*)hidden, w2, x, probs, dx,
                                                                                     (__nv_bfloat16*)dpre, dw2, db2, (int)N, (int)Dh, (int)D, inv_keep);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
