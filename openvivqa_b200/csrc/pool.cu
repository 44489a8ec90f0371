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
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
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
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
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


// ---------------------------------------------------------------------------------------------------------------- vector kernels
// The kernels above touch one 2- or 4-byte element per thread and load (profiles/r02_kernels_ncu.md: 80 us for 42 MB, 79 % of the stall
// samples on outstanding loads, 12 % occupancy — 256 CTAs cannot hide DRAM latency with 4 bytes per request).  These read 16 bytes per
// request with 4-8 requests in flight per thread; shapes they need: D % 4 == 0, Dh % 8 == 0 (bf16) / % 4 (fp32), 16-byte aligned rows.
template <typename TH> struct HVec;
template <> struct HVec<float> {
    static constexpr int kN = 4;
    static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct HVec<__nv_bfloat16> {
    static constexpr int kN = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
};

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = kPoolThreads / 32;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();                                 // `red` may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

template <typename TH>
__global__ void __launch_bounds__(kPoolThreads) attr_pool_fwd_vec_kernel(const TH* __restrict__ hid, const float* __restrict__ w2, const float* __restrict__ b2,
                                                                         const float* __restrict__ x, float* __restrict__ pooled, float* __restrict__ probs,
                                                                         int N, int Dh, int D) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ __align__(16) float sp[];      // [N] logits -> probabilities | [kPoolThreads * 4] partial weighted sums
    __shared__ float red[kPoolThreads / 32];
    constexpr int HN = HVec<TH>::kN;
    const int64_t b = blockIdx.x;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = kPoolThreads / 32;
    const TH* hb = hid + b * (int64_t)N * Dh;
    float* part = sp + ((N + 3) & ~3);
    // logits: one warp per row, two rows at a time (their loads are all issued before the first FMA)
    const int hvecs = Dh / HN;
    for (int n = warp; n < N; n += 2 * nw) {
        const int n1 = n + nw;
        float s0 = 0.f, s1 = 0.f;
        for (int c = lane; c < hvecs; c += 32) {
            float h0[HN], h1[HN], w[HN];
            HVec<TH>::load(hb + (int64_t)n * Dh + c * HN, h0);
            if (n1 < N) HVec<TH>::load(hb + (int64_t)n1 * Dh + c * HN, h1);
#pragma unroll
            for (int j = 0; j < HN; j += 4) {
                const float4 ww = __ldg(reinterpret_cast<const float4*>(w2 + c * HN + j));
                w[j] = ww.x; w[j + 1] = ww.y; w[j + 2] = ww.z; w[j + 3] = ww.w;
            }
#pragma unroll
            for (int j = 0; j < HN; ++j) { s0 = fmaf(h0[j], w[j], s0); if (n1 < N) s1 = fmaf(h1[j], w[j], s1); }
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { sp[n] = s0 + __ldg(b2); if (n1 < N) sp[n1] = s1 + __ldg(b2); }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) mx = fmaxf(mx, sp[n]);
    mx = block_reduce(mx, red, true);
    float sum = 0.f;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) { const float e = expf(sp[n] - mx); sp[n] = e; sum += e; }
    sum = block_reduce(sum, red, false);
    const float inv = 1.f / sum;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) { const float p = sp[n] * inv; sp[n] = p; if (probs) probs[b * N + n] = p; }
    __syncthreads();
    // pooled: thread = (row partition, float4 column group); eight rows in flight per thread, partitions combined through shared memory
    const float* xb = x + b * (int64_t)N * D;
    const int dvecs = D / 4;
    for (int c0 = 0; c0 < dvecs; c0 += kPoolThreads) {
        const int cols = min(dvecs - c0, kPoolThreads);          // column groups of this pass
        const int parts = kPoolThreads / cols > 0 ? kPoolThreads / cols : 1;   // row partitions (threads beyond parts * cols idle)
        const int c = c0 + (int)threadIdx.x % cols, part_id = (int)threadIdx.x / cols;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (part_id < parts) {
            int n = part_id;
            for (; n + 3 * parts < N; n += 4 * parts) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)(n + u * parts) * D) + c);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float p = sp[n + u * parts];
                    acc.x = fmaf(p, v[u].x, acc.x); acc.y = fmaf(p, v[u].y, acc.y); acc.z = fmaf(p, v[u].z, acc.z); acc.w = fmaf(p, v[u].w, acc.w);
                }
            }
            for (; n < N; n += parts) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)n * D) + c);
                const float p = sp[n];
                acc.x = fmaf(p, v.x, acc.x); acc.y = fmaf(p, v.y, acc.y); acc.z = fmaf(p, v.z, acc.z); acc.w = fmaf(p, v.w, acc.w);
            }
        }
        __syncthreads();
        reinterpret_cast<float4*>(part)[threadIdx.x] = acc;
        __syncthreads();
        if ((int)threadIdx.x < cols) {
            float4 t = reinterpret_cast<float4*>(part)[threadIdx.x];
            for (int q = 1; q < parts; ++q) {
                const float4 o = reinterpret_cast<float4*>(part)[q * cols + threadIdx.x];
                t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
            }
            *reinterpret_cast<float4*>(pooled + b * D + 4 * (c0 + (int)threadIdx.x)) = t;
        }
    }
}

template <typename TH>
__global__ void __launch_bounds__(kPoolThreads) attr_pool_bwd_vec_kernel(const float* __restrict__ dpooled, const TH* __restrict__ hid, const float* __restrict__ w2,
                                                                         const float* __restrict__ x, const float* __restrict__ probs, float* __restrict__ dx,
                                                                         TH* __restrict__ dpre, float* __restrict__ dw2, float* __restrict__ db2, int N, int Dh,
                                                                         int D, float inv_keep) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ __align__(16) float sd[];      // [N] x.dpooled -> dlogit | [N] p | [kPoolThreads * 8] partial dw2
    __shared__ float red[kPoolThreads / 32];
    constexpr int HN = HVec<TH>::kN;
    const int64_t b = blockIdx.x;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nw = kPoolThreads / 32;
    const int Np = (N + 3) & ~3;
    float* sprob = sd + Np;
    float* part = sprob + Np;
    const float* xb = x + b * (int64_t)N * D;
    const float* dpb = dpooled + b * D;
    const float* pb = probs + b * N;
    const int dvecs = D / 4;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) sprob[n] = __ldg(pb + n);
    // x . dpooled per row: one warp per row, two rows at a time
    for (int n = warp; n < N; n += 2 * nw) {
        const int n1 = n + nw;
        float s0 = 0.f, s1 = 0.f;
        for (int c = lane; c < dvecs; c += 32) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(dpb) + c);
            const float4 a = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)n * D) + c);
            float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n1 < N) a1 = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)n1 * D) + c);
            s0 = fmaf(a.x, g.x, fmaf(a.y, g.y, fmaf(a.z, g.z, fmaf(a.w, g.w, s0))));
            s1 = fmaf(a1.x, g.x, fmaf(a1.y, g.y, fmaf(a1.z, g.z, fmaf(a1.w, g.w, s1))));
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        if (lane == 0) { sd[n] = s0; if (n1 < N) sd[n1] = s1; }
    }
    __syncthreads();
    float dot = 0.f;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) dot = fmaf(sprob[n], sd[n], dot);
    dot = block_reduce(dot, red, false);
    float dl_sum = 0.f;
    for (int n = threadIdx.x; n < N; n += kPoolThreads) { const float dl = sprob[n] * (sd[n] - dot); sd[n] = dl; dl_sum += dl; }
    dl_sum = block_reduce(dl_sum, red, false);       // (its barriers also publish sd[] = dlogit)
    if (threadIdx.x == 0 && db2) atomicAdd(db2, dl_sum);
    // dx[b,n,:] = p[n] dpooled: 16-byte stores, consecutive threads -> consecutive column groups
    if (dx) {
        float* dxb = dx + b * (int64_t)N * D;
        for (int idx = threadIdx.x; idx < N * dvecs; idx += kPoolThreads) {
            const int n = idx / dvecs, c = idx % dvecs;
            const float4 g = __ldg(reinterpret_cast<const float4*>(dpb) + c);
            const float p = sprob[n];
            *reinterpret_cast<float4*>(dxb + (int64_t)n * D + 4 * c) = make_float4(p * g.x, p * g.y, p * g.z, p * g.w);
        }
    }
    // d pre-activation and dw2: thread = (row partition, HN-column group), four rows in flight
    const TH* hb = hid + b * (int64_t)N * Dh;
    TH* db = dpre + b * (int64_t)N * Dh;
    const int hvecs = Dh / HN;
    for (int c0 = 0; c0 < hvecs; c0 += kPoolThreads) {
        const int cols = min(hvecs - c0, kPoolThreads);
        const int parts = kPoolThreads / cols > 0 ? kPoolThreads / cols : 1;
        const int c = c0 + (int)threadIdx.x % cols, part_id = (int)threadIdx.x / cols;
        float acc[HN], w[HN];
#pragma unroll
        for (int j = 0; j < HN; ++j) acc[j] = 0.f;
        if (part_id < parts) {
#pragma unroll
            for (int j = 0; j < HN; j += 4) {
                const float4 ww = __ldg(reinterpret_cast<const float4*>(w2 + c * HN + j));
                w[j] = ww.x * inv_keep; w[j + 1] = ww.y * inv_keep; w[j + 2] = ww.z * inv_keep; w[j + 3] = ww.w * inv_keep;
            }
            int n = part_id;
            for (; n + 3 * parts < N; n += 4 * parts) {
                float h[4][HN];
#pragma unroll
                for (int u = 0; u < 4; ++u) HVec<TH>::load(hb + (int64_t)(n + u * parts) * Dh + c * HN, h[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float dl = sd[n + u * parts];
                    float o[HN];
#pragma unroll
                    for (int j = 0; j < HN; ++j) { acc[j] = fmaf(dl, h[u][j], acc[j]); o[j] = h[u][j] > 0.f ? dl * w[j] : 0.f; }
                    HVec<TH>::store(db + (int64_t)(n + u * parts) * Dh + c * HN, o);
                }
            }
            for (; n < N; n += parts) {
                float h[HN], o[HN];
                HVec<TH>::load(hb + (int64_t)n * Dh + c * HN, h);
                const float dl = sd[n];
#pragma unroll
                for (int j = 0; j < HN; ++j) { acc[j] = fmaf(dl, h[j], acc[j]); o[j] = h[j] > 0.f ? dl * w[j] : 0.f; }
                HVec<TH>::store(db + (int64_t)n * Dh + c * HN, o);
            }
        }
        if (dw2) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < HN; ++j) part[j * kPoolThreads + threadIdx.x] = acc[j];
            __syncthreads();
            if ((int)threadIdx.x < cols) {
#pragma unroll
                for (int j = 0; j < HN; ++j) {
                    float t = 0.f;
                    for (int q = 0; q < parts; ++q) t += part[j * kPoolThreads + q * cols + threadIdx.x];
                    atomicAdd(dw2 + (c0 + (int)threadIdx.x) * HN + j, t);
                }
            }
        }
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace ovqa

using namespace ovqa;

extern "C" int ovqa_attr_pool_fwd(const void* hidden, const float* w2, const float* b2, const float* x, float* pooled, float* probs, int64_t B, int64_t N,
                                  int64_t Dh, int64_t D, int hidden_dtype, void* stream) {
    if (!hidden || !w2 || !b2 || !x || !pooled || B < 0 || N <= 0 || Dh <= 0 || D <= 0) return OVQA_ERR_BAD_ARG;
    if (hidden_dtype != OVQA_F32 && hidden_dtype != OVQA_BF16) return OVQA_ERR_BAD_ARG;
    if (N * sizeof(float) > 48 * 1024 || B > 0x7fffffff) return OVQA_ERR_BAD_SHAPE;
    if (B == 0) return OVQA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int hn = hidden_dtype == OVQA_F32 ? 4 : 8;
    const size_t smem_v = (size_t)(((N + 3) & ~3) + kPoolThreads * 4) * sizeof(float);
    if (D % 4 == 0 && Dh % hn == 0 && al16(hidden) && al16(w2) && al16(x) && al16(pooled) && smem_v <= 48 * 1024) {
        if (hidden_dtype == OVQA_F32)
            OVQA_CUDA_OK(launch_pdl(attr_pool_fwd_vec_kernel<float>, dim3((unsigned)B), dim3(kPoolThreads), (size_t)(smem_v), st, (const float*)hidden, w2, b2, x, pooled, probs, (int)N, (int)Dh, (int)D));
        else
            OVQA_CUDA_OK(launch_pdl(attr_pool_fwd_vec_kernel<__nv_bfloat16>, dim3((unsigned)B), dim3(kPoolThreads), (size_t)(smem_v), st, (const __nv_bfloat16*)hidden, w2, b2, x, pooled, probs, (int)N, (int)Dh, (int)D));
        OVQA_LAUNCH_OK();
        return OVQA_OK;
    }
    const size_t smem = (size_t)N * sizeof(float);
    if (hidden_dtype == OVQA_F32)
        OVQA_CUDA_OK(launch_pdl(attr_pool_fwd_kernel<float>, dim3((unsigned)B), dim3(kPoolThreads), (size_t)(smem), st, (const float*)hidden, w2, b2, x, pooled, probs, (int)N, (int)Dh, (int)D));
    else
        OVQA_CUDA_OK(launch_pdl(attr_pool_fwd_kernel<__nv_bfloat16>, dim3((unsigned)B), dim3(kPoolThreads), (size_t)(smem), st, (const __nv_bfloat16*)hidden, w2, b2, x, pooled, probs, (int)N, (int)Dh, (int)D));
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
    const int hn = hidden_dtype == OVQA_F32 ? 4 : 8;
    const size_t smem_v = (size_t)(2 * ((N + 3) & ~3) + kPoolThreads * 8) * sizeof(float);
    if (D % 4 == 0 && Dh % hn == 0 && al16(hidden) && al16(w2) && al16(x) && al16(dpooled) && al16(dpre) && (!dx || al16(dx)) && smem_v <= 48 * 1024) {
        if (hidden_dtype == OVQA_F32)
            OVQA_CUDA_OK(launch_pdl(attr_pool_bwd_vec_kernel<float>, dim3((unsigned)B), dim3(kPoolThreads), (size_t)(smem_v), st, dpooled, (const float*)hidden, w2, x, probs, dx, (float*)dpre, dw2, db2,
                                                                                     (int)N, (int)Dh, (int)D, inv_keep));
        else
            OVQA_CUDA_OK(launch_pdl(attr_pool_bwd_vec_kernel<__nv_bfloat16>, dim3((unsigned)B), dim3(kPoolThreads), (size_t)(smem_v), st, dpooled, (const __nv_bfloat16*)hidden, w2, x, probs, dx,
                                                                                             (__nv_bfloat16*)dpre, dw2, db2, (int)N, (int)Dh, (int)D, inv_keep));
        OVQA_LAUNCH_OK();
        return OVQA_OK;
    }
    if (hidden_dtype == OVQA_F32)
        OVQA_CUDA_OK(launch_pdl(attr_pool_bwd_kernel<float>, dim3((unsigned)B), dim3(kPoolThreads), (size_t)(smem), st, dpooled, (const float*)hidden, w2, x, probs, dx, (float*)dpre, dw2, db2, (int)N,
                                                                             (int)Dh, (int)D, inv_keep));
    else
        OVQA_CUDA_OK(launch_pdl(attr_pool_bwd_kernel<__nv_bfloat16>, dim3((unsigned)B), dim3(kPoolThreads), (size_t)(smem), st, dpooled, (const __nv_bfloat16*)hidden, w2, x, probs, dx,
                                                                                     (__nv_bfloat16*)dpre, dw2, db2, (int)N, (int)Dh, (int)D, inv_keep));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
