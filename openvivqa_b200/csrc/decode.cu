// Inference-time incremental decoding: one MultiHeadAttention block of a decode step over a cache of
// PROJECTED keys / values, and the feed-forward block that follows it, each enqueued by ONE call.
//
// The reference (attentions.py:321-325 -> :48-50, decoders.py:21-26, beam_search.py:41-58) concatenates the raw decoder inputs of
// every step into running_keys / running_values and re-projects the whole sequence — and the unchanged encoder output — at every
// step: O(T^2) projection work for a T-token answer and ~45 small kernel launches per layer-step, which at beam-search batch sizes
// (a few hundred rows) is bounded by launch latency and the Python dispatch between the launches, not by arithmetic.  Here:
//   * the cache holds the projected rows [K | V] per (batch row, position): a step projects only its new rows;
//   * the attention kernel reads positions < len from the cache and the new rows from the projection output, and files the new rows
//     into the cache itself (no separate append pass, no torch.cat);
//   * the host enqueues GEMM -> attention -> GEMM(+bias+residual) -> LayerNorm back to back from C: 4 launches, one FFI call.
// All memory (cache, workspace, outputs) is caller-owned; nothing here synchronises.
#include <algorithm>

#include "common.cuh"

namespace ovqa {

constexpr int kDecWarps = 4;

template <typename T> struct Vec16;   // 16-byte vector of T
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ void load16(const T* p, float (&f)[Vec16<T>::N]);
template <> __device__ __forceinline__ void load16<float>(const float* p, float (&f)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load16<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = unpack_bf16x2(u[i]);
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}

struct DecodeAttn {
    const void* q;        // [B*Lq, ldq], head hd at column hd*dk
    const void* kv_new;   // [B*Lnew, ldn]: K at column k_off + hd*dk, V at column v_off + hd*dv   (NULL when Lnew == 0)
    void* cache;          // [B, cap, ldc]: K at column hd*dk, V at column hk + hd*dv
    void* out;            // [B*Lq, ldo]
    int64_t B, h, Lq, Lnew, len, cap, dk, dv;
    int64_t ldq, ldn, k_off, v_off, ldc, ldo;
    const float* mask;    // additive, element (b, hd, lq, j) at b*sb + hd*sh + lq*sq + j;  NULL = none
    int64_t sb, sh, sq;
    float scale;
};

// One warp per (batch row, query row, head); K / V rows are read with one 16-byte load per lane, a group of lanes per row and several rows
// per pass (coalesced, independent loads in flight); softmax in fp32 over the warp.
template <typename T>
__global__ void __launch_bounds__(kDecWarps * 32) sdpa_decode_kernel(DecodeAttn a) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * kDecWarps + warp;
    const int64_t n_items = a.B * a.Lq * a.h;
    if (item >= n_items) return;
    const int hd = (int)(item % a.h);
    const int64_t row = item / a.h;            // b*Lq + lq
    const int lq = (int)(row % a.Lq);
    const int64_t b = row / a.Lq;
    const int Lk = (int)(a.len + a.Lnew), dk = (int)a.dk, dv = (int)a.dv;
    const int hk = (int)(a.h * a.dk);
    const int slab = max(Lk, 32 * (int)Vec16<T>::N);
    float* sq_ = smem + (size_t)warp * slab;
    float* sp = sq_;                                        // scores, then probabilities; afterwards the partial output rows
    constexpr int VN = Vec16<T>::N;

    const T* q = reinterpret_cast<const T*>(a.q) + row * a.ldq + (int64_t)hd * dk;
    const T* cacheK = reinterpret_cast<const T*>(a.cache) + b * a.cap * a.ldc + (int64_t)hd * dk;
    const T* cacheV = reinterpret_cast<const T*>(a.cache) + b * a.cap * a.ldc + hk + (int64_t)hd * dv;
    const T* newK = a.kv_new ? reinterpret_cast<const T*>(a.kv_new) + b * a.Lnew * a.ldn + a.k_off + (int64_t)hd * dk : nullptr;
    const T* newV = a.kv_new ? reinterpret_cast<const T*>(a.kv_new) + b * a.Lnew * a.ldn + a.v_off + (int64_t)hd * dv : nullptr;
    const float* mrow = a.mask ? a.mask + b * a.sb + (int64_t)hd * a.sh + (int64_t)lq * a.sq : nullptr;
    const int len = (int)a.len;

    // scores: dk / VN lanes cover one key row with a 16-byte load each (one coalesced row per group), padded to a power of two so the
    // partial dot products meet by xor shuffles; the warp takes 32 / P keys per pass
    const int lpk = dk / VN;
    int P = 1;
    while (P < lpk) P <<= 1;
    const int Gk = 32 / P, gk = lane / P, ck = (lane % P) * VN;
    const bool live = (lane % P) < lpk;
    float qf[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) qf[i] = 0.f;
    if (live) {
        load16<T>(q + ck, qf);
#pragma unroll
        for (int i = 0; i < VN; ++i) qf[i] *= a.scale;
    }
    float mx = -INFINITY;
#pragma unroll 2
    for (int j0 = 0; j0 < Lk; j0 += Gk) {
        const int j = j0 + gk;
        float acc = 0.f;
        if (live && j < Lk) {
            const T* kr = j < len ? cacheK + (int64_t)j * a.ldc : newK + (int64_t)(j - len) * a.ldn;
            float f[VN];
            load16<T>(kr + ck, f);
#pragma unroll
            for (int i = 0; i < VN; ++i) acc = fmaf(f[i], qf[i], acc);
        }
        for (int o = P >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (j < Lk) {
            if (mrow) acc += mrow[j];
            if ((lane % P) == 0) sp[j] = acc;
            mx = fmaxf(mx, acc);
        }
    }
    __syncwarp();
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
        const float e = __expf(sp[j] - mx);
        sp[j] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.0f / sum;
    // output: dv / VN lanes cover the head dimension with one 16-byte load per key row; the warp's lanes form G such groups that take the
    // keys in turn (independent loads in flight instead of one dependent chain per lane); the G partial rows meet in shared memory
    const int lpg = dv / VN, G = 32 / lpg, grp = lane / lpg, c0 = (lane % lpg) * VN;
    float acc[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) acc[i] = 0.f;
    if (grp < G) {
#pragma unroll 4
        for (int j = grp; j < Lk; j += G) {
            const T* vr = j < len ? cacheV + (int64_t)j * a.ldc : newV + (int64_t)(j - len) * a.ldn;
            float f[VN];
            load16<T>(vr + c0, f);
            const float p = sp[j];
#pragma unroll
            for (int i = 0; i < VN; ++i) acc[i] = fmaf(p, f[i], acc[i]);
        }
    }
    __syncwarp();                       // everyone is done reading sp[]: reuse it (and the q row before it) for the partial rows
    float* red = sq_;                   // G * dv <= 32 * VN floats: the host sizes the slab as max(Lk, 32 * VN)
    if (grp < G) {
#pragma unroll
        for (int i = 0; i < VN; ++i) red[grp * dv + c0 + i] = acc[i];
    }
    __syncwarp();
    T* o = reinterpret_cast<T*>(a.out) + row * a.ldo + (int64_t)hd * dv;
    for (int c = lane; c < dv; c += 32) {
        float t = 0.f;
        for (int g2 = 0; g2 < G; ++g2) t += red[g2 * dv + c];
        o[c] = from_f32<T>(t * inv);
    }
    // file this head's slice of the new rows into the cache (once per batch row and head: the lq == 0 warp does it)
    if (lq == 0 && a.Lnew > 0) {
        T* dK = reinterpret_cast<T*>(a.cache) + (b * a.cap + len) * a.ldc + (int64_t)hd * dk;
        T* dV = reinterpret_cast<T*>(a.cache) + (b * a.cap + len) * a.ldc + hk + (int64_t)hd * dv;
        for (int r = 0; r < (int)a.Lnew; ++r) {
            for (int c = lane * VN; c < dk; c += 32 * VN)
                *reinterpret_cast<uint4*>(dK + (int64_t)r * a.ldc + c) = *reinterpret_cast<const uint4*>(newK + (int64_t)r * a.ldn + c);
            for (int c = lane * VN; c < dv; c += 32 * VN)
                *reinterpret_cast<uint4*>(dV + (int64_t)r * a.ldc + c) = *reinterpret_cast<const uint4*>(newV + (int64_t)r * a.ldn + c);
        }
    }
}

template <typename T>
static int launch_decode_attn(const DecodeAttn& a, cudaStream_t st) {
    const int64_t items = a.B * a.Lq * a.h;
    if (items == 0) return OVQA_OK;
    const size_t smem = (size_t)kDecWarps * (size_t)std::max<int64_t>(a.len + a.Lnew, 32 * Vec16<T>::N) * sizeof(float);
    if (smem > 200 * 1024) return OVQA_ERR_BAD_SHAPE;
    if (smem > 48 * 1024) OVQA_CUDA_OK(cudaFuncSetAttribute(sdpa_decode_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = (items + kDecWarps - 1) / kDecWarps;
    sdpa_decode_kernel<T><<<(unsigned)blocks, kDecWarps * 32, smem, st>>>(a);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

// ----------------------------------------------------------------------------- beam selection: the k largest entries of every row
// BeamSearch.select (beam_search.py:35-38) sorts all beam x vocabulary candidates of every sample and keeps the first `beam_size`.
// Here: grid (chunks, rows); every thread keeps the k best of its strided share in registers (sorted insertion), the block merges the
// 256 sorted lists by k rounds of a block-wide arg-max over the list heads; a second tiny launch merges the chunks' results the same way.
// Order: descending value, NaN greatest (as torch.sort), equal values by ascending index (torch's order among ties is unspecified).
constexpr int kTopkMax = 16, kTopkThreads = 256;

__device__ __forceinline__ bool topk_before(float a, int64_t ia, float b, int64_t ib) {      // does (a, ia) come before (b, ib)?
    const bool na = a != a, nb = b != b;
    if (na != nb) return na;
    if (!na && a != b) return a > b;
    return ia < ib;
}

// block-wide selection of the k best among the threads' sorted lists (v[], ix[] of length k, sentinel index -1 = empty slot)
template <int K>
__device__ void topk_block_merge(float (&v)[K], int64_t (&ix)[K], int k, float* out_v, int64_t* out_i) {
    __shared__ float s_v[kTopkThreads / 32];
    __shared__ int64_t s_i[kTopkThreads / 32];
    __shared__ int s_t[kTopkThreads / 32];
    __shared__ int s_win;
    int head = 0;
    for (int r = 0; r < k; ++r) {
        // candidate of this thread: its list head
        float cv = 0.f; int64_t ci = -1; int ct = threadIdx.x;
        if (head < k) {
#pragma unroll
            for (int j = 0; j < K; ++j) if (j == head) { cv = v[j]; ci = ix[j]; }
        }
        // warp arg-best
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, cv, o);
            const int64_t oi = __shfl_xor_sync(0xffffffffu, ci, o);
            const int ot = __shfl_xor_sync(0xffffffffu, ct, o);
            const bool take = oi >= 0 && (ci < 0 || topk_before(ov, oi, cv, ci));
            if (take) { cv = ov; ci = oi; ct = ot; }
        }
        if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = cv; s_i[threadIdx.x >> 5] = ci; s_t[threadIdx.x >> 5] = ct; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float bv = s_v[0]; int64_t bi = s_i[0]; int bt = s_t[0];
            for (int w = 1; w < kTopkThreads / 32; ++w)
                if (s_i[w] >= 0 && (bi < 0 || topk_before(s_v[w], s_i[w], bv, bi))) { bv = s_v[w]; bi = s_i[w]; bt = s_t[w]; }
            out_v[r] = bv; out_i[r] = bi; s_win = bi >= 0 ? bt : -1;
        }
        __syncthreads();
        if (s_win == (int)threadIdx.x) ++head;
        __syncthreads();
    }
}

template <int K>
__device__ __forceinline__ void topk_insert(float (&v)[K], int64_t (&ix)[K], int k, float x, int64_t i) {
    if (ix[k - 1] >= 0 && !topk_before(x, i, v[k - 1], ix[k - 1])) return;      // not better than the current k-th
    // bubble up from the end (k is small)
    float cv = x; int64_t ci = i;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        if (j < k) {
            const bool first = ix[j] < 0 || topk_before(cv, ci, v[j], ix[j]);
            if (first) { const float tv = v[j]; const int64_t ti = ix[j]; v[j] = cv; ix[j] = ci; cv = tv; ci = ti; }
            if (ci < 0) break;
        }
    }
}

// phase 1: x [rows, n] (row pitch ldx); partial results [rows, chunks, k]
template <int K>
__global__ void __launch_bounds__(kTopkThreads) topk_partial_kernel(const float* __restrict__ x, int64_t n, int64_t ldx, int k, int chunks,
                                                                    float* __restrict__ pv, int64_t* __restrict__ pi) {
    const int64_t row = blockIdx.y;
    const int chunk = blockIdx.x;
    const int64_t per = (n + chunks - 1) / chunks, lo = chunk * per, hi = min(n, lo + per);
    float v[K]; int64_t ix[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { v[j] = 0.f; ix[j] = -1; }
    const float* xr = x + row * ldx;
    for (int64_t i = lo + threadIdx.x; i < hi; i += kTopkThreads) topk_insert<K>(v, ix, k, __ldg(xr + i), i);
    __shared__ float o_v[kTopkMax];
    __shared__ int64_t o_i[kTopkMax];
    topk_block_merge<K>(v, ix, k, o_v, o_i);
    if ((int)threadIdx.x < k) {
        pv[(row * chunks + chunk) * k + threadIdx.x] = o_v[threadIdx.x];
        pi[(row * chunks + chunk) * k + threadIdx.x] = o_i[threadIdx.x];
    }
}

// phase 2: merge the chunks' lists of a row (chunks * k <= kTopkThreads * K candidates)
template <int K>
__global__ void __launch_bounds__(kTopkThreads) topk_merge_kernel(const float* __restrict__ pv, const int64_t* __restrict__ pi, int k, int chunks,
                                                                  float* __restrict__ out_v, int64_t* __restrict__ out_i) {
    const int64_t row = blockIdx.x;
    float v[K]; int64_t ix[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { v[j] = 0.f; ix[j] = -1; }
    const int total = chunks * k;
    for (int c = threadIdx.x; c < total; c += kTopkThreads) {
        const int64_t i = pi[row * total + c];
        if (i >= 0) topk_insert<K>(v, ix, k, pv[row * total + c], i);
    }
    __shared__ float o_v[kTopkMax];
    __shared__ int64_t o_i[kTopkMax];
    topk_block_merge<K>(v, ix, k, o_v, o_i);
    if ((int)threadIdx.x < k) { out_v[row * k + threadIdx.x] = o_v[threadIdx.x]; out_i[row * k + threadIdx.x] = o_i[threadIdx.x]; }
}

static inline int topk_chunks(int64_t rows, int64_t n) {
    // enough blocks to fill the SMs, at least ~4 elements per thread and chunk
    int64_t c = (2 * (int64_t)sm_count() + rows - 1) / (rows > 0 ? rows : 1);
    const int64_t cmax = (n + 4 * kTopkThreads - 1) / (4 * kTopkThreads);
    if (c > cmax) c = cmax;
    if (c < 1) c = 1;
    if (c > 64) c = 64;
    return (int)c;
}

static inline int64_t up256(int64_t v) { return (v + 255) / 256 * 256; }

}  // namespace ovqa

using namespace ovqa;

static int plain_gemm(const void* A, const void* W, void* C, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldw, int64_t ldc, int dtype,
                      int c_dtype, const float* bias, int flags, const float* residual, int64_t ld_res, void* stream) {
    return ovqa_gemm(A, W, C, m, n, k, lda, ldw, ldc, OVQA_OP_N, OVQA_OP_N, dtype, c_dtype, 1.0f, 0.0f, bias, flags, nullptr, nullptr, 0,
                     residual, ld_res, nullptr, 0, 0.0f, 0, 0, nullptr, nullptr, 1, stream);
}

extern "C" int64_t ovqa_mha_decode_workspace_bytes(int64_t B, int64_t Lq, int64_t Lnew, int64_t D, int64_t h, int64_t dk, int64_t dv, int dtype) {
    const int64_t e = dtype == OVQA_BF16 ? 2 : 4, hk = h * dk, hv = h * dv, rows = B * Lq;
    return up256(rows * D * e) + up256(rows * (2 * hk + hv) * e) + up256(B * Lnew * (hk + hv) * e) + up256(rows * hv * e) + up256(rows * D * 4);
}

extern "C" int ovqa_mha_decode_step(const float* x_f32, const void* x_op, const void* kv_src_op, int self_attention, int64_t B, int64_t Lq,
                                    int64_t Lnew, int64_t D, int64_t h, int64_t dk, int64_t dv, const void* w_qkv, const float* b_qkv,
                                    const void* w_o, const float* b_o, const float* gamma, const float* beta, float eps, void* kv_cache,
                                    int64_t cap, int64_t len, const float* mask, int64_t mask_sb, int64_t mask_sh, int64_t mask_sq,
                                    float scale, int dtype, void* workspace, int64_t workspace_bytes, float* y_f32, void* y_op,
                                    void* stream) {
    if (!x_f32 || !w_qkv || !b_qkv || !w_o || !b_o || !gamma || !beta || !y_f32) return OVQA_ERR_BAD_ARG;
    if (B < 0 || Lq < 0 || Lnew < 0 || len < 0 || D <= 0 || h <= 0 || dk <= 0 || dv <= 0) return OVQA_ERR_BAD_ARG;
    if (dtype != OVQA_F32 && dtype != OVQA_BF16) return OVQA_ERR_BAD_ARG;
    if (self_attention ? (Lnew != Lq) : ((Lnew > 0) != (kv_src_op != nullptr))) return OVQA_ERR_BAD_ARG;
    if (len + Lnew > cap || (cap > 0 && !kv_cache) || len + Lnew <= 0) return OVQA_ERR_BAD_ARG;
    if (B == 0 || Lq == 0) return OVQA_OK;
    const int64_t e = dtype == OVQA_BF16 ? 2 : 4, vn = 16 / e, hk = h * dk, hv = h * dv, rows = B * Lq, W = hk + hv;
    if (dk % vn || dv % vn || dk > 32 * vn || dv > 32 * vn) return OVQA_ERR_BAD_SHAPE;
    if (!workspace || workspace_bytes < ovqa_mha_decode_workspace_bytes(B, Lq, Lnew, D, h, dk, dv, dtype)) return OVQA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(kv_cache)) & 15) return OVQA_ERR_ALIGNMENT;
    char* ws = reinterpret_cast<char*>(workspace);
    void* ws_x = ws;          ws += up256(rows * D * e);
    void* ws_q = ws;          ws += up256(rows * (2 * hk + hv) * e);
    void* ws_kv = ws;         ws += up256(B * Lnew * W * e);
    void* ws_ctx = ws;        ws += up256(rows * hv * e);
    float* ws_s = reinterpret_cast<float*>(ws);
    int rc;
    // operand in the compute dtype
    const void* xq = dtype == OVQA_F32 ? (const void*)x_f32 : x_op;
    if (!xq) {
        if ((rc = ovqa_cast_f32_to_bf16(x_f32, ws_x, rows * D, stream))) return rc;
        xq = ws_x;
    }
    DecodeAttn a{};
    a.B = B; a.h = h; a.Lq = Lq; a.Lnew = Lnew; a.len = len; a.cap = cap; a.dk = dk; a.dv = dv;
    a.cache = kv_cache; a.ldc = W; a.out = ws_ctx; a.ldo = hv;
    a.mask = mask; a.sb = mask_sb; a.sh = mask_sh; a.sq = mask_sq; a.scale = scale;
    if (self_attention) {          // one GEMM over [Wq; Wk; Wv]
        if ((rc = plain_gemm(xq, w_qkv, ws_q, rows, 2 * hk + hv, D, D, D, 2 * hk + hv, dtype, dtype, b_qkv, 0, nullptr, 0, stream))) return rc;
        a.q = ws_q; a.ldq = 2 * hk + hv;
        a.kv_new = ws_q; a.ldn = 2 * hk + hv; a.k_off = hk; a.v_off = 2 * hk;
    } else {
        if ((rc = plain_gemm(xq, w_qkv, ws_q, rows, hk, D, D, D, hk, dtype, dtype, b_qkv, 0, nullptr, 0, stream))) return rc;
        a.q = ws_q; a.ldq = hk;
        if (Lnew > 0) {            // rows [hk, 2hk+hv) of the fused weight are [Wk; Wv]
            const char* w_kv = reinterpret_cast<const char*>(w_qkv) + hk * D * e;
            if ((rc = plain_gemm(kv_src_op, w_kv, ws_kv, B * Lnew, W, D, D, D, W, dtype, dtype, b_qkv + hk, 0, nullptr, 0, stream))) return rc;
            a.kv_new = ws_kv; a.ldn = W; a.k_off = 0; a.v_off = hk;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = dtype == OVQA_F32 ? launch_decode_attn<float>(a, st) : launch_decode_attn<__nv_bfloat16>(a, st);
    if (rc) return rc;
    // fc_o + bias + residual, then LayerNorm   (attentions.py:58, :330-331 with dropout off)
    if ((rc = plain_gemm(ws_ctx, w_o, ws_s, rows, D, hv, hv, hv, D, dtype, OVQA_F32, b_o, 0, x_f32, D, stream))) return rc;
    return ovqa_layernorm_fwd(ws_s, nullptr, gamma, beta, nullptr, 0, y_f32, dtype == OVQA_BF16 ? y_op : nullptr, nullptr, nullptr, nullptr, rows,
                              D, eps, stream);
}

// ----------------------------------------------------------------------------- whole MHA block, inference, one call
// MultiHeadAttention.forward in eval mode (attentions.py:319-331 with :46-60): projection GEMM(s) over the fused [Wq; Wk; Wv] operand ->
// attention core -> fc_o + bias + residual -> LayerNorm, enqueued back to back from C.  The Python block (functional.mha_block) costs
// ~190 us of host dispatch (six library calls, each through ctypes, with their allocations): at batch 16-64 that is more than the GPU
// work and made the drop-in slower than eager PyTorch (profiles/r01l_microbench.jsonl); this entry leaves one ctypes call.
//   kind 0: self-attention (keys = values = queries): one GEMM over the whole fused weight
//   kind 1: keys = values != queries: fc_q over the query rows, [fc_k | fc_v] over the key rows in one GEMM
//   kind 2: three different inputs: three GEMMs
extern "C" int64_t ovqa_mha_block_infer_workspace_bytes(int64_t B, int64_t Lq, int64_t Lk, int64_t D, int64_t h, int64_t dk, int64_t dv, int kind,
                                                        int dtype) {
    const int64_t e = dtype == OVQA_BF16 ? 2 : 4, hk = h * dk, hv = h * dv, rq = B * Lq, rk = B * Lk;
    int64_t n = up256(rq * D * e) + up256(rq * hv * e) + up256(rq * D * 4);                      // query operand, context, fc_o + residual sum
    if (kind == 0) n += up256(rq * (2 * hk + hv) * e);
    else n += up256(rq * hk * e) + up256(rk * (hk + hv) * e) + (kind == 1 ? 1 : 2) * up256(rk * D * e);
    return n;
}

extern "C" int ovqa_mha_block_infer(const float* xq_f32, const void* xq_op, const void* xk, const void* xk_op, const void* xv, const void* xv_op,
                                    int kind, int64_t B, int64_t Lq, int64_t Lk, int64_t D, int64_t h, int64_t dk, int64_t dv, const void* w_qkv,
                                    const float* b_qkv, const void* w_o, const float* b_o, const float* gamma, const float* beta, float eps,
                                    const float* mask, int64_t mask_sb, int64_t mask_sh, int64_t mask_sq, const float* prefix_key_mask,
                                    int64_t prefix_sb, int64_t prefix_tail, float prefix_value, float scale, int dtype, void* workspace,
                                    int64_t workspace_bytes, float* y_f32, void* y_op, void* stream) {
    if (!xq_f32 || !w_qkv || !b_qkv || !w_o || !b_o || !gamma || !beta || !y_f32) return OVQA_ERR_BAD_ARG;
    if (B < 0 || Lq < 0 || Lk <= 0 || D <= 0 || h <= 0 || dk <= 0 || dv <= 0 || kind < 0 || kind > 2) return OVQA_ERR_BAD_ARG;
    if (dtype != OVQA_F32 && dtype != OVQA_BF16) return OVQA_ERR_BAD_ARG;
    if ((kind == 0 && Lk != Lq) || (kind >= 1 && !xk) || (kind == 2 && !xv)) return OVQA_ERR_BAD_ARG;
    if (B == 0 || Lq == 0) return OVQA_OK;
    if (!workspace || workspace_bytes < ovqa_mha_block_infer_workspace_bytes(B, Lq, Lk, D, h, dk, dv, kind, dtype)) return OVQA_ERR_BAD_ARG;
    const int64_t e = dtype == OVQA_BF16 ? 2 : 4, hk = h * dk, hv = h * dv, rq = B * Lq, rk = B * Lk;
    char* ws = reinterpret_cast<char*>(workspace);
    void* ws_xq = ws;        ws += up256(rq * D * e);
    void* ws_ctx = ws;       ws += up256(rq * hv * e);
    float* ws_s = reinterpret_cast<float*>(ws);   ws += up256(rq * D * 4);
    int rc;
    // operands in the compute dtype: fp32 inputs are used as they are; bf16 takes the caller's shadow copy or casts here
    auto operand = [&](const void* f32, const void* op, void* scratch, int64_t rows, const void** out) -> int {
        if (dtype == OVQA_F32) { *out = f32; return OVQA_OK; }
        if (op) { *out = op; return OVQA_OK; }
        *out = scratch;
        return ovqa_cast_f32_to_bf16(reinterpret_cast<const float*>(f32), scratch, rows * D, stream);
    };
    const void* q_in;
    if ((rc = operand(xq_f32, xq_op, ws_xq, rq, &q_in))) return rc;
    const char* wq = reinterpret_cast<const char*>(w_qkv);
    const void *q, *k, *v;
    int64_t ldq, ldk, ldv;
    if (kind == 0) {
        void* qkv = ws;      ws += up256(rq * (2 * hk + hv) * e);
        if ((rc = plain_gemm(q_in, w_qkv, qkv, rq, 2 * hk + hv, D, D, D, 2 * hk + hv, dtype, dtype, b_qkv, 0, nullptr, 0, stream))) return rc;
        q = qkv; k = reinterpret_cast<char*>(qkv) + hk * e; v = reinterpret_cast<char*>(qkv) + 2 * hk * e;
        ldq = ldk = ldv = 2 * hk + hv;
    } else {
        void* qb = ws;       ws += up256(rq * hk * e);
        void* kvb = ws;      ws += up256(rk * (hk + hv) * e);
        void* ws_xk = ws;    ws += up256(rk * D * e);
        if ((rc = plain_gemm(q_in, w_qkv, qb, rq, hk, D, D, D, hk, dtype, dtype, b_qkv, 0, nullptr, 0, stream))) return rc;
        const void* k_in;
        if ((rc = operand(xk, xk_op, ws_xk, rk, &k_in))) return rc;
        if (kind == 1) {
            if ((rc = plain_gemm(k_in, wq + hk * D * e, kvb, rk, hk + hv, D, D, D, hk + hv, dtype, dtype, b_qkv + hk, 0, nullptr, 0, stream))) return rc;
        } else {
            void* ws_xv = ws;    ws += up256(rk * D * e);
            const void* v_in;
            if ((rc = operand(xv, xv_op, ws_xv, rk, &v_in))) return rc;
            if ((rc = plain_gemm(k_in, wq + hk * D * e, kvb, rk, hk, D, D, D, hk + hv, dtype, dtype, b_qkv + hk, 0, nullptr, 0, stream))) return rc;
            if ((rc = plain_gemm(v_in, wq + 2 * hk * D * e, reinterpret_cast<char*>(kvb) + hk * e, rk, hv, D, D, D, hk + hv, dtype, dtype, b_qkv + 2 * hk, 0,
                                 nullptr, 0, stream))) return rc;
        }
        q = qb; k = kvb; v = reinterpret_cast<char*>(kvb) + hk * e;
        ldq = hk; ldk = ldv = hk + hv;
    }
    // attention core: the structured prefix-LM form when the caller described one and this shape has that kernel, else the dense mask
    rc = OVQA_ERR_BAD_SHAPE;
    if (prefix_key_mask)
        rc = ovqa_sdpa_fwd_prefix_lm(q, k, v, ws_ctx, B, h, Lq, Lk, dk, dv, ldq, ldk, ldv, hv, prefix_key_mask, prefix_sb, prefix_tail, prefix_value,
                                     scale, dtype, nullptr, stream);
    if (rc == OVQA_ERR_BAD_SHAPE)
        rc = ovqa_sdpa_fwd(q, k, v, ws_ctx, B, h, Lq, Lk, dk, dv, ldq, ldk, ldv, hv, mask, mask_sb, mask_sh, mask_sq, nullptr, nullptr, nullptr, nullptr, 0,
                           0, scale, dtype, nullptr, nullptr, 0.f, 0, 0, nullptr, stream);
    if (rc) return rc;
    // fc_o + bias + residual, then LayerNorm   (attentions.py:58, :330-331 with dropout off)
    if ((rc = plain_gemm(ws_ctx, w_o, ws_s, rq, D, hv, hv, hv, D, dtype, OVQA_F32, b_o, 0, xq_f32, D, stream))) return rc;
    return ovqa_layernorm_fwd(ws_s, nullptr, gamma, beta, nullptr, 0, y_f32, dtype == OVQA_BF16 ? y_op : nullptr, nullptr, nullptr, nullptr, rq, D,
                              eps, stream);
}

extern "C" int64_t ovqa_pwff_infer_workspace_bytes(int64_t rows, int64_t D, int64_t d_ff, int dtype) {
    const int64_t e = dtype == OVQA_BF16 ? 2 : 4;
    return up256(rows * D * e) + up256(rows * d_ff * e) + up256(rows * D * 4);
}

// positionwise_feed_forward.py:23-27 with dropout off: LN(x + fc2(gelu(fc1 x))) — three launches from one call
extern "C" int ovqa_pwff_infer(const float* x_f32, const void* x_op, int64_t rows, int64_t D, int64_t d_ff, const void* w1, const float* b1,
                               const void* w2, const float* b2, const float* gamma, const float* beta, float eps, int dtype, void* workspace,
                               int64_t workspace_bytes, float* y_f32, void* y_op, void* stream) {
    if (!x_f32 || !w1 || !b1 || !w2 || !b2 || !gamma || !beta || !y_f32 || rows < 0 || D <= 0 || d_ff <= 0) return OVQA_ERR_BAD_ARG;
    if (dtype != OVQA_F32 && dtype != OVQA_BF16) return OVQA_ERR_BAD_ARG;
    if (rows == 0) return OVQA_OK;
    if (!workspace || workspace_bytes < ovqa_pwff_infer_workspace_bytes(rows, D, d_ff, dtype)) return OVQA_ERR_BAD_ARG;
    const int64_t e = dtype == OVQA_BF16 ? 2 : 4;
    char* ws = reinterpret_cast<char*>(workspace);
    void* ws_x = ws;   ws += up256(rows * D * e);
    void* ws_h = ws;   ws += up256(rows * d_ff * e);
    float* ws_s = reinterpret_cast<float*>(ws);
    int rc;
    const void* x = dtype == OVQA_F32 ? (const void*)x_f32 : x_op;
    if (!x) {
        if ((rc = ovqa_cast_f32_to_bf16(x_f32, ws_x, rows * D, stream))) return rc;
        x = ws_x;
    }
    if ((rc = plain_gemm(x, w1, ws_h, rows, d_ff, D, D, D, d_ff, dtype, dtype, b1, OVQA_EPI_GELU, nullptr, 0, stream))) return rc;
    if ((rc = plain_gemm(ws_h, w2, ws_s, rows, D, d_ff, d_ff, d_ff, D, dtype, OVQA_F32, b2, 0, x_f32, D, stream))) return rc;
    return ovqa_layernorm_fwd(ws_s, nullptr, gamma, beta, nullptr, 0, y_f32, dtype == OVQA_BF16 ? y_op : nullptr, nullptr, nullptr, nullptr, rows,
                              D, eps, stream);
}

extern "C" int64_t ovqa_topk_rows_workspace_bytes(int64_t rows, int64_t n, int k) {
    if (rows <= 0 || n <= 0 || k <= 0) return 0;
    return up256(rows * topk_chunks(rows, n) * k * 4) + up256(rows * topk_chunks(rows, n) * k * 8);
}

extern "C" int ovqa_topk_rows(const float* x, int64_t rows, int64_t n, int64_t ldx, int k, float* out_val, int64_t* out_idx, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    if (!x || !out_val || !out_idx || rows < 0 || n <= 0 || ldx < n || k <= 0) return OVQA_ERR_BAD_ARG;
    if (k > kTopkMax || k > n || rows > 65535) return OVQA_ERR_BAD_SHAPE;      // (rows ride on gridDim.y)
    if (rows == 0) return OVQA_OK;
    if (!workspace || workspace_bytes < ovqa_topk_rows_workspace_bytes(rows, n, k)) return OVQA_ERR_BAD_ARG;
    const int chunks = topk_chunks(rows, n);
    float* pv = reinterpret_cast<float*>(workspace);
    int64_t* pi = reinterpret_cast<int64_t*>(reinterpret_cast<char*>(workspace) + up256(rows * chunks * k * 4));
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid((unsigned)chunks, (unsigned)rows);
    if (k <= 8) topk_partial_kernel<8><<<grid, kTopkThreads, 0, st>>>(x, n, ldx, k, chunks, pv, pi);
    else topk_partial_kernel<16><<<grid, kTopkThreads, 0, st>>>(x, n, ldx, k, chunks, pv, pi);
    OVQA_LAUNCH_OK();
    if (k <= 8) topk_merge_kernel<8><<<(unsigned)rows, kTopkThreads, 0, st>>>(pv, pi, k, chunks, out_val, out_idx);
    else topk_merge_kernel<16><<<(unsigned)rows, kTopkThreads, 0, st>>>(pv, pi, k, chunks, out_val, out_idx);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
