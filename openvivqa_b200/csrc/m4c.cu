// M4C dynamic pointer network, previous-prediction embedding and the greedy decode step (HBM/latency-bound
// integer+float glue of mmf_m4c.py:236-256, :379-396, :416-459), fused so that no (B, V+Locr, D) table,
// torch.cat or per-step host sync is needed.
#include "common.cuh"

namespace ovqa {
namespace {

constexpr int kMaxVec = 8;

// one warp per (b, t, j): scores[b, t, col_offset + j] = q[b,t,:].k[b,j,:] * scale + mask[b,j]
template <typename T>
__global__ void ptr_scores_kernel(const T* __restrict__ q, const T* __restrict__ k, const void* __restrict__ mask, float* __restrict__ scores,
                                  int64_t B, int64_t Tn, int64_t L, int64_t d, int64_t ldq, int64_t ldk, int64_t ld_scores,
                                  int64_t col_offset, float scale, int mask_is_bool) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (w >= B * Tn * L) return;
    const int64_t j = w % L, t = (w / L) % Tn, b = w / (L * Tn);
    const T* qr = q + (b * Tn + t) * ldq;
    const T* kr = k + (b * L + j) * ldk;
    float s = 0.f;
    for (int64_t c = lane; c < d; c += 32) s = fmaf(to_f32<T>(qr[c]), to_f32<T>(kr[c]), s);
    s = warp_sum(s);
    if (lane == 0) {
        s *= scale;
        if (mask) {
            if (mask_is_bool) { if (reinterpret_cast<const uint8_t*>(mask)[b * L + j]) s = -INFINITY; }
            else s += reinterpret_cast<const float*>(mask)[b * L + j];
        }
        scores[(b * Tn + t) * ld_scores + col_offset + j] = s;
    }
}

// Vector form for the decode step (bf16, d % 8 == 0, 16-byte aligned rows, T <= 16): one warp per KEY row (b, j).  The key row is read
// once with 16-byte loads and stays in registers while the T query rows of the sample (L1/L2 resident) are dotted against it; the
// scalar kernel above spent one warp and 2-byte loads on every (b, t, j) — 41 us for 128 x 12 x 50 scores of 768 elements.
constexpr int kPtrMaxT = 16;
__global__ void __launch_bounds__(256) ptr_scores_vec_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                             const void* __restrict__ mask, float* __restrict__ scores, int64_t B, int Tn,
                                                             int64_t L, int d, int64_t ldq, int64_t ldk, int64_t ld_scores, int64_t col_offset,
                                                             float scale, int mask_is_bool) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (w >= B * L) return;
    const int64_t j = w % L, b = w / L;
    const uint4* kr = reinterpret_cast<const uint4*>(k + (b * L + j) * ldk);
    const int nv = d / 8;                                     // 16-byte vectors per row; lane owns vectors lane, lane + 32, ...
    float acc[kPtrMaxT];
#pragma unroll
    for (int t = 0; t < kPtrMaxT; ++t) acc[t] = 0.f;
    for (int c = lane; c < nv; c += 32) {
        const uint4 kv = __ldg(kr + c);
        const float2 k0 = unpack_bf16x2(kv.x), k1 = unpack_bf16x2(kv.y), k2 = unpack_bf16x2(kv.z), k3 = unpack_bf16x2(kv.w);
        // every query row's vector is requested before the first is used (rows beyond Tn: row 0 again, result unused): requested one by
        // one inside the predicated loop they were 13 serial L2 round trips per vector — 46 us for 128 x 12 x 50 scores (gpurun_out/r02x_ptr)
        uint4 qv[kPtrMaxT];
#pragma unroll
        for (int t = 0; t < kPtrMaxT; ++t) qv[t] = __ldg(reinterpret_cast<const uint4*>(q + (b * Tn + (t < Tn ? t : 0)) * ldq) + c);
#pragma unroll
        for (int t = 0; t < kPtrMaxT; ++t) {
            const float2 q0 = unpack_bf16x2(qv[t].x), q1 = unpack_bf16x2(qv[t].y), q2 = unpack_bf16x2(qv[t].z), q3 = unpack_bf16x2(qv[t].w);
            acc[t] += (q0.x * k0.x + q0.y * k0.y) + (q1.x * k1.x + q1.y * k1.y) + (q2.x * k2.x + q2.y * k2.y) + (q3.x * k3.x + q3.y * k3.y);
        }
    }
    float add = 0.f;
    bool kill = false;
    if (mask) {
        if (mask_is_bool) kill = reinterpret_cast<const uint8_t*>(mask)[b * L + j] != 0;
        else add = reinterpret_cast<const float*>(mask)[b * L + j];
    }
#pragma unroll
    for (int t = 0; t < kPtrMaxT; ++t) {
        if (t < Tn) {
            const float s = warp_sum(acc[t]);
            if (lane == 0) scores[(b * Tn + t) * ld_scores + col_offset + j] = kill ? -INFINITY : s * scale + add;
        }
    }
}

// Backward of the pointer scores (autograd of mmf_m4c.py:394-396):  dq[b,t,:] = scale * sum_j dS[b,t,j] k[b,j,:],
// dk[b,j,:] = scale * sum_t dS[b,t,j] q[b,t,:].  The additive mask is a constant; positions a boolean mask filled with -inf carry no
// gradient.  CTA = (sample, 128-column slab of d): dS of the sample sits in shared memory, every thread owns one column and keeps 16
// accumulators (rows of dq, then rows of dk, 16 at a time); q / k column reads are coalesced across the CTA.
constexpr int kPtrBwdThreads = 128, kPtrBwdAcc = 16;
template <typename T>
__global__ void __launch_bounds__(kPtrBwdThreads) ptr_scores_bwd_kernel(const float* __restrict__ ds, const T* __restrict__ q, const T* __restrict__ k,
                                                                       const uint8_t* __restrict__ bool_mask, T* __restrict__ dq, T* __restrict__ dk,
                                                                       int Tn, int L, int d, int64_t ld_ds, int64_t ldq, int64_t ldk, int64_t lddq,
                                                                       int64_t lddk, float scale) {
    extern __shared__ float sds[];                       // [Tn][L]
    const int64_t b = blockIdx.x;
    for (int i = threadIdx.x; i < Tn * L; i += blockDim.x) {
        const int t = i / L, j = i % L;
        float v = ds[(b * Tn + t) * ld_ds + j];
        if (bool_mask && bool_mask[b * L + j]) v = 0.f;
        sds[i] = v;
    }
    __syncthreads();
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= d) return;
    const T* qb = q + b * Tn * ldq + c;
    const T* kb = k + b * L * ldk + c;
    float acc[kPtrBwdAcc];
    for (int t0 = 0; t0 < Tn; t0 += kPtrBwdAcc) {
#pragma unroll
        for (int i = 0; i < kPtrBwdAcc; ++i) acc[i] = 0.f;
        for (int j = 0; j < L; ++j) {
            const float kv = to_f32<T>(kb[(int64_t)j * ldk]);
#pragma unroll
            for (int i = 0; i < kPtrBwdAcc; ++i)
                if (t0 + i < Tn) acc[i] = fmaf(sds[(t0 + i) * L + j], kv, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < kPtrBwdAcc; ++i)
            if (t0 + i < Tn) dq[(b * Tn + t0 + i) * lddq + c] = from_f32<T>(acc[i] * scale);
    }
    for (int j0 = 0; j0 < L; j0 += kPtrBwdAcc) {
#pragma unroll
        for (int i = 0; i < kPtrBwdAcc; ++i) acc[i] = 0.f;
        for (int t = 0; t < Tn; ++t) {
            const float qv = to_f32<T>(qb[(int64_t)t * ldq]);
#pragma unroll
            for (int i = 0; i < kPtrBwdAcc; ++i)
                if (j0 + i < L) acc[i] = fmaf(sds[t * L + j0 + i], qv, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < kPtrBwdAcc; ++i)
            if (j0 + i < L) dk[(b * L + j0 + i) * lddk + c] = from_f32<T>(acc[i] * scale);
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// LayerNorm of a register-resident row (float4 per lane), fp32 two-pass
__device__ __forceinline__ void ln_row(float4 (&v)[kMaxVec], int nvec, int lane, int D, float eps, const float* g, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxVec; ++t) if (lane + 32 * t < nvec) s += (v[t].x + v[t].y) + (v[t].z + v[t].w);
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxVec; ++t) if (lane + 32 * t < nvec) {
        const float a = v[t].x - mean, bb = v[t].y - mean, c = v[t].z - mean, dd = v[t].w - mean;
        q += (a * a + bb * bb) + (c * c + dd * dd);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
    for (int t = 0; t < kMaxVec; ++t) {
        const int c = lane + 32 * t;
        if (c < nvec) {
            const float4 gg = ld4(g + 4 * c), bb = ld4(b + 4 * c);
            v[t] = make_float4((v[t].x - mean) * rstd * gg.x + bb.x, (v[t].y - mean) * rstd * gg.y + bb.y,
                               (v[t].z - mean) * rstd * gg.z + bb.z, (v[t].w - mean) * rstd * gg.w + bb.w);
        }
    }
}

// one warp per (b, t)
__global__ void prev_pred_embed_kernel(const float* __restrict__ ans_emb, const float* __restrict__ ocr_emb, const int64_t* __restrict__ prev_inds,
                                       const float* __restrict__ pos_emb, const float* __restrict__ type_emb,
                                       const float* ans_g, const float* ans_b, const float* ocr_g, const float* ocr_b,
                                       const float* emb_g, const float* emb_b, float* __restrict__ out,
                                       int64_t B, int64_t Tn, int64_t V, int64_t Locr, int D, int64_t ld_out, int64_t out_bs, float eps) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (w >= B * Tn) return;
    const int64_t t = w % Tn, b = w / Tn;
    const int nvec = D / 4;
    int64_t idx = prev_inds[b * Tn + t];
    idx = idx < 0 ? 0 : (idx >= V + Locr ? V + Locr - 1 : idx);
    const bool is_ocr = idx >= V;
    const float* src = is_ocr ? ocr_emb + (b * Locr + (idx - V)) * D : ans_emb + idx * D;
    float4 a[kMaxVec], e[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) {
            a[i] = ld4(src + 4 * c);
            const float4 p = ld4(pos_emb + t * D + 4 * c), ty = ld4(type_emb + (is_ocr ? 1 : 0) * D + 4 * c);
            e[i] = make_float4(p.x + ty.x, p.y + ty.y, p.z + ty.z, p.w + ty.w);
        }
    }
    ln_row(a, nvec, lane, D, eps, is_ocr ? ocr_g : ans_g, is_ocr ? ocr_b : ans_b);
    ln_row(e, nvec, lane, D, eps, emb_g, emb_b);
    float* o = out + b * out_bs + t * ld_out;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
        const int c = lane + 32 * i;
        if (c < nvec) *reinterpret_cast<float4*>(o + 4 * c) = make_float4(a[i].x + e[i].x, a[i].y + e[i].y, a[i].z + e[i].z, a[i].w + e[i].w);
    }
}

// one warp per (b, t): first index of the row maximum (torch.argmax tie rule)
__global__ void argmax_rows_kernel(const float* __restrict__ scores, int64_t* __restrict__ out, int64_t rows, int64_t W) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (w >= rows) return;
    const float* r = scores + w * W;
    float best = -INFINITY;
    int64_t bi = W;  // sentinel larger than any index
    int64_t c = lane;
    for (; c + 96 < W; c += 128) {                    // four loads in flight per lane; compared in ascending column order
        const float v0 = __ldg(r + c), v1 = __ldg(r + c + 32), v2 = __ldg(r + c + 64), v3 = __ldg(r + c + 96);
        if (bi == W || v0 > best) { best = v0; bi = c; }
        if (v1 > best) { best = v1; bi = c + 32; }
        if (v2 > best) { best = v2; bi = c + 64; }
        if (v3 > best) { best = v3; bi = c + 96; }
    }
    for (; c < W; c += 32) {
        const float v = __ldg(r + c);
        if (bi == W || v > best) { best = v; bi = c; }  // c ascends per lane: strict > keeps the first maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi < W && (bi == W || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) out[w] = bi == W ? 0 : bi;
}

__global__ void greedy_update_kernel(const int64_t* __restrict__ argmax, int64_t* __restrict__ prev_inds, int64_t* __restrict__ last_ids,
                                     int32_t* __restrict__ all_eos, int64_t B, int64_t Tn, int64_t step, int64_t eos) {
    int ok = 1;
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) {
        for (int64_t t = 1; t < Tn; ++t) prev_inds[b * Tn + t] = argmax[b * Tn + t - 1];
        int64_t l = last_ids[b];
        if (l != eos) l = argmax[b * Tn + step];
        last_ids[b] = l;
        ok &= (l == eos);
    }
    ok = __syncthreads_and(ok);
    if (threadIdx.x == 0 && all_eos) *all_eos = ok;
}

}  // namespace
}  // namespace ovqa

using namespace ovqa;

extern "C" int ovqa_ptr_scores(const void* q, const void* k, const void* mask, float* scores, int64_t B, int64_t T, int64_t L,
                               int64_t d, int64_t ldq, int64_t ldk, int64_t ld_scores, int64_t col_offset, float scale, int dtype,
                               int mask_is_bool, void* stream) {
    if (!q || !k || !scores || B < 0 || T < 0 || L < 0 || d <= 0) return OVQA_ERR_BAD_ARG;
    const int64_t warps = B * T * L;
    if (warps == 0) return OVQA_OK;
    const unsigned grid = (unsigned)((warps + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OVQA_BF16 && T <= kPtrMaxT && d % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k)) & 15) == 0) {
        ptr_scores_vec_kernel<<<(unsigned)((B * L + 7) / 8), 256, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, mask, scores, B, (int)T, L,
                                                                          (int)d, ldq, ldk, ld_scores, col_offset, scale, mask_is_bool);
        OVQA_LAUNCH_OK();
        return OVQA_OK;
    }
    if (dtype == OVQA_F32) ptr_scores_kernel<float><<<grid, 256, 0, st>>>((const float*)q, (const float*)k, mask, scores, B, T, L, d, ldq, ldk, ld_scores, col_offset, scale, mask_is_bool);
    else if (dtype == OVQA_BF16) ptr_scores_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)q, (const __nv_bfloat16*)k, mask, scores, B, T, L, d, ldq, ldk, ld_scores, col_offset, scale, mask_is_bool);
    else return OVQA_ERR_BAD_ARG;
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

extern "C" int ovqa_ptr_scores_bwd(const float* dscores, const void* q, const void* k, const void* bool_mask, void* dq, void* dk, int64_t B,
                                   int64_t T, int64_t L, int64_t d, int64_t ld_dscores, int64_t ldq, int64_t ldk, int64_t lddq, int64_t lddk,
                                   float scale, int dtype, void* stream) {
    if (!dscores || !q || !k || !dq || !dk || B < 0 || T < 0 || L < 0 || d <= 0) return OVQA_ERR_BAD_ARG;
    if (dtype != OVQA_F32 && dtype != OVQA_BF16) return OVQA_ERR_BAD_ARG;
    if (B == 0 || T == 0 || L == 0) return OVQA_OK;                 // (the caller zero-fills whichever of dq / dk has rows then)
    const size_t smem = (size_t)T * (size_t)L * sizeof(float);
    if (smem > 48 * 1024 || B > 0x7fffffff) return OVQA_ERR_BAD_SHAPE;
    dim3 grid((unsigned)B, (unsigned)((d + kPtrBwdThreads - 1) / kPtrBwdThreads));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OVQA_F32)
        ptr_scores_bwd_kernel<float><<<grid, kPtrBwdThreads, smem, st>>>(dscores, (const float*)q, (const float*)k, (const uint8_t*)bool_mask, (float*)dq,
                                                                         (float*)dk, (int)T, (int)L, (int)d, ld_dscores, ldq, ldk, lddq, lddk, scale);
    else
        ptr_scores_bwd_kernel<__nv_bfloat16><<<grid, kPtrBwdThreads, smem, st>>>(dscores, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                                                 (const uint8_t*)bool_mask, (__nv_bfloat16*)dq, (__nv_bfloat16*)dk, (int)T,
                                                                                 (int)L, (int)d, ld_dscores, ldq, ldk, lddq, lddk, scale);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

extern "C" int ovqa_prev_pred_embed(const float* ans_emb, const float* ocr_emb, const int64_t* prev_inds, const float* pos_emb,
                                    const float* type_emb, const float* ans_g, const float* ans_b, const float* ocr_g,
                                    const float* ocr_b, const float* emb_g, const float* emb_b, float* out, int64_t B, int64_t T,
                                    int64_t V, int64_t Locr, int64_t D, int64_t ld_out, int64_t out_batch_stride, float eps,
                                    void* stream) {
    if (!ans_emb || !ocr_emb || !prev_inds || !pos_emb || !type_emb || !ans_g || !ans_b || !ocr_g || !ocr_b || !emb_g || !emb_b || !out) return OVQA_ERR_BAD_ARG;
    if (B < 0 || T < 0 || V <= 0 || Locr < 0 || D <= 0) return OVQA_ERR_BAD_ARG;
    if (D % 4 != 0 || D > kMaxVec * 128 || ld_out % 4 != 0 || out_batch_stride % 4 != 0) return OVQA_ERR_BAD_SHAPE;
    if (B * T == 0) return OVQA_OK;
    prev_pred_embed_kernel<<<(unsigned)((B * T + 3) / 4), 128, 0, (cudaStream_t)stream>>>(ans_emb, ocr_emb, prev_inds, pos_emb, type_emb, ans_g, ans_b, ocr_g, ocr_b, emb_g, emb_b, out, B, T, V, Locr, (int)D, ld_out, out_batch_stride, eps);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

extern "C" int ovqa_greedy_step(const float* scores, int64_t* prev_inds, int64_t* last_ids, int64_t* argmax_out, int32_t* all_eos,
                                int64_t B, int64_t T, int64_t W, int64_t step, int64_t eos_idx, void* stream) {
    if (!scores || !prev_inds || !last_ids || !argmax_out || B < 0 || T <= 0 || W <= 0 || step < 0 || step >= T) return OVQA_ERR_BAD_ARG;
    if (B == 0) return OVQA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    argmax_rows_kernel<<<(unsigned)((B * T + 7) / 8), 256, 0, st>>>(scores, argmax_out, B * T, W);
    OVQA_LAUNCH_OK();
    greedy_update_kernel<<<1, 1024, 0, st>>>(argmax_out, prev_inds, last_ids, all_eos, B, T, step, eos_idx);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
