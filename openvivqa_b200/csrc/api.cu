// extern "C" entry points of libovqa_b200.so (see include/ovqa.h): argument validation + dispatch.
#include <cstdlib>

#include "attention_simt.cuh"

namespace ovqa {

thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;

// SMs the persistent kernels size their grids for: the device's count minus a reserve (ovqa_set_sm_reserve) left free for kernels of
// OTHER streams that must co-run — NCCL's all-reduce during a data-parallel backward pass.  A persistent grid of one CTA per SM whose last
// CTAs cannot become resident (their SMs hold the collective) runs those CTAs' tiles as a second wave: the whole launch takes ~1.7x
// (profiles/r02k_ddp_timeline.json).
static int g_sm_reserve = [] { const char* e = getenv("OVQA_SM_RESERVE"); const int v = e ? atoi(e) : 0; return v < 0 ? 0 : (v + 1) / 2 * 2; }();
int sm_count() {
    static int cache[64] = {};
    const int dev = current_device() & 63;
    int v = __atomic_load_n(&cache[dev], __ATOMIC_RELAXED);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMsB200;
        __atomic_store_n(&cache[dev], v, __ATOMIC_RELAXED);
    }
    const int r = __atomic_load_n(&g_sm_reserve, __ATOMIC_RELAXED);
    return v - r >= 8 ? v - r : v;
}

template <typename T> int sdpa_fwd_simt_launch(const SdpaArgs&, cudaStream_t);
template <typename T> int sdpa_bwd_simt_launch(const SdpaArgs&, const void*, void*, void*, void*, int64_t, int64_t, int64_t, int64_t, float*, float*, float*, const float*, const float*, cudaStream_t);
int box_geometry_bias_launch(const GeoArg&, float*, int64_t, int64_t, int64_t, cudaStream_t);
int sdpa_fwd_mma_bf16(const SdpaArgs&, cudaStream_t);   // attention_mma.cu
bool sdpa_fwd_mma_supported(const SdpaArgs&);
int sdpa_fwd_tc5_bf16(const SdpaArgs&, cudaStream_t);   // attention_tc5.cu (tcgen05 / TMEM / TMA: head dim 64, Lq, Lk <= 128, per-key mask)
bool sdpa_fwd_tc5_supported(const SdpaArgs&);
int sdpa_fwd_tc5w_bf16(const SdpaArgs&, cudaStream_t);  // attention_tc5.cu, wide heads (d = 96 / 128), any Lq, Lk <= 256
bool sdpa_fwd_tc5w_supported(const SdpaArgs&);
int sdpa_bwd_tc5_bf16(const SdpaArgs&, const void*, void*, void*, void*, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
bool sdpa_bwd_tc5_supported(const SdpaArgs&, const void*, const void*, const void*, const void*, int64_t, int64_t, int64_t, int64_t);
int sdpa_bwd_mma_bf16(const SdpaArgs&, const void*, void*, void*, void*, int64_t, int64_t, int64_t, int64_t, float*, cudaStream_t);
bool sdpa_bwd_mma_supported(const SdpaArgs&, const void*, const void*, const void*, const void*, int64_t, int64_t, int64_t, int64_t);

}  // namespace ovqa

using namespace ovqa;

extern "C" int ovqa_version(void) { return 100; }

extern "C" const char* ovqa_error_string(int code) {
    switch (code) {
        case OVQA_OK: return "ok";
        case OVQA_ERR_BAD_ARG: return "bad argument (null pointer, negative size or unsupported dtype/flag)";
        case OVQA_ERR_BAD_SHAPE: return "shape outside the supported range of this kernel";
        case OVQA_ERR_ALIGNMENT: return "pointer or leading dimension not aligned as required (16 bytes)";
        case OVQA_ERR_CUDA: return "CUDA call failed (see ovqa_last_cuda_error)";
        case OVQA_ERR_NO_DEVICE: return "no usable sm_100 device / driver entry point";
        default: return "unknown error";
    }
}
extern "C" int ovqa_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" int ovqa_device_sm_count(void) { return sm_count(); }
extern "C" int ovqa_set_sm_reserve(int sms) {
    if (sms < 0) sms = 0;
    return __atomic_exchange_n(&g_sm_reserve, (sms + 1) / 2 * 2, __ATOMIC_RELAXED);
}
extern "C" uint64_t ovqa_launch_count(void) { return __atomic_load_n(&g_launch_count, __ATOMIC_RELAXED); }

extern "C" int ovqa_gemm(const void* A, const void* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                         int64_t ldc, int a_op, int b_op, int ab_dtype, int c_dtype, float alpha, float beta, const float* bias,
                         int epi_flags, const void* aux_in, void* aux_out, int64_t ld_aux, const float* residual, int64_t ld_res,
                         void* c2_bf16, int64_t ldc2, float dropout_p, uint64_t seed, uint64_t offset, const uint64_t* offset_base,
                         float* colsum_out, int split_k, void* stream) {
    if (!A || !B || !C || m < 0 || n < 0 || k < 0) return OVQA_ERR_BAD_ARG;
    if ((a_op != OVQA_OP_N && a_op != OVQA_OP_T) || (b_op != OVQA_OP_N && b_op != OVQA_OP_T)) return OVQA_ERR_BAD_ARG;
    if ((ab_dtype != OVQA_F32 && ab_dtype != OVQA_BF16) || (c_dtype != OVQA_F32 && c_dtype != OVQA_BF16)) return OVQA_ERR_BAD_ARG;
    if (dropout_p < 0.f || dropout_p >= 1.f) return OVQA_ERR_BAD_ARG;
    if ((epi_flags & (OVQA_EPI_DGELU | OVQA_EPI_MUL_AUX)) && !aux_in) return OVQA_ERR_BAD_ARG;
    if ((epi_flags & OVQA_EPI_SAVE_DGELU) && !((epi_flags & OVQA_EPI_GELU) && aux_out)) return OVQA_ERR_BAD_ARG;
    if (beta != 0.f && c_dtype != OVQA_F32) return OVQA_ERR_BAD_ARG;
    if (split_k != 1 && (bias || epi_flags || aux_out || residual || c2_bf16 || dropout_p > 0.f || beta != 0.f || colsum_out)) return OVQA_ERR_BAD_ARG;
    if (m == 0 || n == 0) return OVQA_OK;
    Epilogue e{alpha, beta, bias, epi_flags, aux_in, aux_out, ld_aux, residual, ld_res, c2_bf16, ldc2, dropout_p, seed, offset, offset_base, c_dtype, colsum_out};
    cudaStream_t st = (cudaStream_t)stream;
    if (ab_dtype == OVQA_F32) return gemm_simt_f32((const float*)A, (const float*)B, C, m, n, k, lda, ldb, ldc, a_op, b_op, e, split_k, st);
    // bf16: tensor cores.  Operand pitches must allow TMA (checked there); the epilogue picks 16-byte vector accesses
    // when every pointer/pitch allows them and falls back to scalar stores otherwise.
    // debugging aid only: OVQA_BF16_GEMM=simt routes bf16 GEMMs to the CUDA-core twin (to bisect a tensor-core bug)
    static const bool force_simt = [] { const char* e = getenv("OVQA_BF16_GEMM"); return e && e[0] == 's'; }();
    if (force_simt) return gemm_simt_bf16((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, C, m, n, k, lda, ldb, ldc, a_op, b_op, e, split_k, st);
    // TMA needs 16-byte aligned bases and row pitches; operands that are not (a width that is no multiple of 8, an odd column slice)
    // go to the CUDA-core twin — same operand forms and epilogue — instead of failing (split_k == 0 "auto" means 1 there)
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda % 8) || (ldb % 8))
        return gemm_simt_bf16((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, C, m, n, k, lda, ldb, ldc, a_op, b_op, e, split_k, st);
    return gemm_tcgen05_bf16((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, C, m, n, k, lda, ldb, ldc, a_op, b_op, e, split_k, st);
}

// Debug/verification twin of ovqa_gemm's bf16 path on CUDA cores (same operand forms and epilogue); used by
// the GPU tests to cross-check the tcgen05 kernel element by element.  Not part of the product path.
extern "C" int ovqa_gemm_bf16_simt(const void* A, const void* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                                   int64_t ldc, int a_op, int b_op, int c_dtype, float alpha, const float* bias, int epi_flags,
                                   void* stream) {
    if (!A || !B || !C) return OVQA_ERR_BAD_ARG;
    Epilogue e{alpha, 0.f, bias, epi_flags, nullptr, nullptr, 0, nullptr, 0, nullptr, 0, 0.f, 0, 0, nullptr, c_dtype, nullptr};
    return gemm_simt_bf16((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, C, m, n, k, lda, ldb, ldc, a_op, b_op, e, 1, (cudaStream_t)stream);
}

static int check_sdpa(const void* q, const void* k, const void* v, int64_t B, int64_t h, int64_t Lq, int64_t Lk, int64_t dk,
                      int64_t dv, const float* boxes, const float* geo_w, const float* geo_b, int64_t d_g, int geo_trig, int dtype) {
    if (!q || !k || !v || B < 0 || h <= 0 || Lq < 0 || Lk <= 0 || dk <= 0 || dv <= 0) return OVQA_ERR_BAD_ARG;
    if (dtype != OVQA_F32 && dtype != OVQA_BF16) return OVQA_ERR_BAD_ARG;
    if (dk > 128 || dv > 128) return OVQA_ERR_BAD_SHAPE;
    if (boxes) {
        if (!geo_w || !geo_b || Lq != Lk) return OVQA_ERR_BAD_ARG;     // attentions.py:116-118,132: self-attention over the boxed set
        if (geo_trig ? (d_g % 8 != 0 || d_g <= 0) : d_g != 4) return OVQA_ERR_BAD_SHAPE;
    }
    return OVQA_OK;
}

extern "C" int ovqa_sdpa_fwd(const void* q, const void* k, const void* v, void* out, int64_t B, int64_t h, int64_t Lq, int64_t Lk,
                             int64_t dk, int64_t dv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const float* mask,
                             int64_t mask_sb, int64_t mask_sh, int64_t mask_sq, const float* boxes, const float* geo_w,
                             const float* geo_b, const float* geo_freq, int64_t d_g, int geo_trig, float scale, int dtype,
                             float* lse, float* att, float prob_dropout_p, uint64_t prob_dropout_seed, uint64_t prob_dropout_offset,
                             const uint64_t* prob_dropout_offset_base, void* stream) {
    int rc = check_sdpa(q, k, v, B, h, Lq, Lk, dk, dv, boxes, geo_w, geo_b, d_g, geo_trig, dtype);
    if (rc) return rc;
    if (!out || prob_dropout_p < 0.f || prob_dropout_p >= 1.f) return OVQA_ERR_BAD_ARG;
    if (B == 0 || Lq == 0) return OVQA_OK;
    SdpaArgs a{q, k, v, out, B, h, Lq, Lk, dk, dv, ldq, ldk, ldv, ldo, {mask, mask_sb, mask_sh, mask_sq},
               {boxes, geo_w, geo_b, geo_freq, (int)d_g, geo_trig}, scale, lse, att, prob_dropout_p, prob_dropout_seed, prob_dropout_offset,
               prob_dropout_offset_base};
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == OVQA_F32) return sdpa_fwd_simt_launch<float>(a, st);
    if (prob_dropout_p > 0.f) return sdpa_fwd_simt_launch<__nv_bfloat16>(a, st);      // probability dropout: CUDA-core kernels (training of HF BERT layers)
    if (sdpa_fwd_tc5_supported(a)) return sdpa_fwd_tc5_bf16(a, st);
    if (sdpa_fwd_tc5w_supported(a)) return sdpa_fwd_tc5w_bf16(a, st);
    if (sdpa_fwd_mma_supported(a)) return sdpa_fwd_mma_bf16(a, st);
    return sdpa_fwd_simt_launch<__nv_bfloat16>(a, st);
}

// Attention core under M4C's prefix-LM mask (mmf_m4c.py:310-343: every query sees the unpadded keys; the last `causal_tail` positions —
// the decoder steps — see each other causally) WITHOUT the dense (B,1,L,L) tensor: a per-key additive row + the tail length.  Tensor-memory
// kernel only (bf16, d = 96 / 128, Lk <= 256); OVQA_ERR_BAD_SHAPE tells the caller to pass the dense mask to ovqa_sdpa_fwd instead.
extern "C" int ovqa_sdpa_fwd_prefix_lm(const void* q, const void* k, const void* v, void* out, int64_t B, int64_t h, int64_t Lq, int64_t Lk,
                                       int64_t dk, int64_t dv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, const float* key_mask,
                                       int64_t key_mask_sb, int64_t causal_tail, float causal_value, float scale, int dtype, float* lse,
                                       void* stream) {
    int rc = check_sdpa(q, k, v, B, h, Lq, Lk, dk, dv, nullptr, nullptr, nullptr, 0, 0, dtype);
    if (rc) return rc;
    if (!out || causal_tail < 0 || causal_tail > Lq || causal_tail > Lk) return OVQA_ERR_BAD_ARG;
    if (B == 0 || Lq == 0) return OVQA_OK;
    SdpaArgs a{q, k, v, out, B, h, Lq, Lk, dk, dv, ldq, ldk, ldv, ldo, {key_mask, key_mask_sb, 0, 0, causal_tail, causal_value},
               {nullptr, nullptr, nullptr, nullptr, 0, 0}, scale, lse, nullptr};
    if (dtype != OVQA_BF16 || !sdpa_fwd_tc5w_supported(a)) return OVQA_ERR_BAD_SHAPE;
    return sdpa_fwd_tc5w_bf16(a, (cudaStream_t)stream);
}

extern "C" int ovqa_sdpa_bwd(const void* q, const void* k, const void* v, const void* out, const void* dout, void* dq, void* dk,
                             void* dv, int64_t B, int64_t h, int64_t Lq, int64_t Lk, int64_t dk_, int64_t dv_, int64_t ldq,
                             int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv,
                             const float* mask, int64_t mask_sb, int64_t mask_sh, int64_t mask_sq, const float* boxes,
                             const float* geo_w, const float* geo_b, const float* geo_freq, int64_t d_g, int geo_trig,
                             float* geo_dw, float* geo_db, float scale, int dtype, const float* lse, float* delta, const float* att,
                             const float* datt, float* dbias, float prob_dropout_p, uint64_t prob_dropout_seed, uint64_t prob_dropout_offset,
                             const uint64_t* prob_dropout_offset_base, void* stream) {
    int rc = check_sdpa(q, k, v, B, h, Lq, Lk, dk_, dv_, boxes, geo_w, geo_b, d_g, geo_trig, dtype);
    if (rc) return rc;
    if (!out || !dout || !dq || !dk || !dv || !lse) return OVQA_ERR_BAD_ARG;
    if ((geo_dw == nullptr) != (geo_db == nullptr)) return OVQA_ERR_BAD_ARG;
    if (datt && !att) return OVQA_ERR_BAD_ARG;
    if (prob_dropout_p < 0.f || prob_dropout_p >= 1.f || (prob_dropout_p > 0.f && dbias)) return OVQA_ERR_BAD_ARG;
    if (B == 0 || Lq == 0) return OVQA_OK;
    SdpaArgs a{q, k, v, const_cast<void*>(out), B, h, Lq, Lk, dk_, dv_, ldq, ldk, ldv, ldo, {mask, mask_sb, mask_sh, mask_sq},
               {boxes, geo_w, geo_b, geo_freq, (int)d_g, geo_trig}, scale, const_cast<float*>(lse), nullptr, prob_dropout_p, prob_dropout_seed,
               prob_dropout_offset, prob_dropout_offset_base};
    const bool tensor_ok = prob_dropout_p == 0.f;         // the tensor-core kernels draw no probability dropout
    cudaStream_t st = (cudaStream_t)stream;
    // debugging aid only: OVQA_SDPA_BWD=simt keeps the bf16 backward on the CUDA-core kernels
    static const bool force_simt = [] { const char* e = getenv("OVQA_SDPA_BWD"); return e && e[0] == 's'; }();
    // a gradient THROUGH the probabilities (datt) is folded into dP by the CUDA-core kernels only
    if (dtype == OVQA_BF16 && tensor_ok && !force_simt && !datt && !dbias && sdpa_bwd_tc5_supported(a, dout, dq, dk, dv, lddo, lddq, lddk, lddv))
        return sdpa_bwd_tc5_bf16(a, dout, dq, dk, dv, lddo, lddq, lddk, lddv, st);
    if (dtype == OVQA_BF16 && tensor_ok && (!force_simt || dbias) && !datt && sdpa_bwd_mma_supported(a, dout, dq, dk, dv, lddo, lddq, lddk, lddv))
        return sdpa_bwd_mma_bf16(a, dout, dq, dk, dv, lddo, lddq, lddk, lddv, dbias, st);
    if (dbias) return OVQA_ERR_BAD_SHAPE;          // dS is returned by the warp-tile bf16 kernel only (d in {32, 64, 96}, 16-byte aligned rows)
    if (!delta) return OVQA_ERR_BAD_ARG;
    if (dtype == OVQA_F32) return sdpa_bwd_simt_launch<float>(a, dout, dq, dk, dv, lddo, lddq, lddk, lddv, delta, geo_dw, geo_db, att, datt, st);
    return sdpa_bwd_simt_launch<__nv_bfloat16>(a, dout, dq, dk, dv, lddo, lddq, lddk, lddv, delta, geo_dw, geo_db, att, datt, st);
}

extern "C" int ovqa_box_geometry_bias(const float* boxes, const float* geo_w, const float* geo_b, const float* geo_freq, float* out,
                                      int64_t B, int64_t N, int64_t h, int64_t d_g, int geo_trig, void* stream) {
    if (!boxes || !geo_w || !geo_b || !out || B < 0 || N < 0 || h <= 0) return OVQA_ERR_BAD_ARG;
    if (geo_trig ? (d_g % 8 != 0 || d_g <= 0) : d_g != 4) return OVQA_ERR_BAD_SHAPE;
    return box_geometry_bias_launch({boxes, geo_w, geo_b, geo_freq, (int)d_g, geo_trig}, out, B, N, h, (cudaStream_t)stream);
}
