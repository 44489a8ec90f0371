// Shared device/host helpers for libovqa_b200 (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ovqa.h"

namespace ovqa {

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return OVQA_ERR_CUDA;
}
#define OVQA_CUDA_OK(expr)                                    \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ::ovqa::cuda_fail(_e);  \
    } while (0)
// every kernel launch of the library goes through this check; it also feeds ovqa_launch_count()
extern unsigned long long g_launch_count;
#define OVQA_LAUNCH_OK()                                      \
    do {                                                      \
        __atomic_fetch_add(&::ovqa::g_launch_count, 1ull, __ATOMIC_RELAXED); \
        OVQA_CUDA_OK(cudaPeekAtLastError());                  \
    } while (0)

// ----------------------------------------------------------------------------- programmatic dependent launch
// The library's kernels are launched with programmaticStreamSerializationAllowed: a kernel's CTAs may be scheduled (and run their set-up)
// while the previous kernel of the stream is still draining; `pdl_wait()` — the first thing every such kernel does before it touches
// global memory — blocks until that kernel has completed and its writes are visible, `pdl_trigger()` lets the NEXT kernel be scheduled.
// Captured CUDA graphs keep the edges as programmatic dependencies.  OVQA_PDL=0 launches everything fully serialised.
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("OVQA_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&lc, kern, static_cast<KArgs>(args)...);
}

constexpr int kNumSMsB200 = 148;
int sm_count();                  // of the CURRENT device (cached per device)

// "once per device" latch for per-function attributes (cudaFuncSetAttribute applies to the current device only: a process that
// drives several GPUs must opt every one of them into > 48 KB of dynamic shared memory)
inline int current_device() { int d = 0; return cudaGetDevice(&d) == cudaSuccess ? d : 0; }
struct OncePerDevice {
    unsigned long long done = 0;
    unsigned long long bit() const { return 1ull << (current_device() & 63); }
    bool pending() const { return (__atomic_load_n(&done, __ATOMIC_RELAXED) & bit()) == 0; }
    void mark() { __atomic_fetch_or(&done, bit(), __ATOMIC_RELAXED); }
};

// ----------------------------------------------------------------------------- element access
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 t = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(t);
}

// ----------------------------------------------------------------------------- warp / block reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ----------------------------------------------------------------------------- exact GELU (erf), F.gelu default
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Standard normal CDF Phi(x) = 0.5 (1 + erf(x / sqrt 2)) for the tensor-core GEMM epilogues, where GELU runs in the
// shadow of the MMA pipe and instruction count matters: Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7 on erf, i.e.
// fp32-rounding level), evaluated on the erfc side so the negative tail has no 1+erf cancellation.  Exactly two
// MUFU operations per element (rcp.approx, ex2.approx — the IEEE-rounded intrinsics compile to call sequences), and
// GELU' reuses the same exponential: exp(-z^2) with z = x/sqrt 2 IS the Gaussian kernel exp(-x^2/2).
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float normal_cdf_fast(float x, float& gauss) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
    gauss = ex2_approx(x * x * -0.72134752044448170368f);           // exp(-x^2/2) = 2^(-x^2 / (2 ln 2))
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    const float h = 0.5f * poly * t * gauss;               // 0.5 erfc(|x|/sqrt 2)
    return x < 0.0f ? h : 1.0f - h;
}
__device__ __forceinline__ float gelu_fast(float x) { float g; return x * normal_cdf_fast(x, g); }
// GELU(x) and GELU'(x) from one CDF / exponential evaluation (the forward pass saves GELU' for backward)
__device__ __forceinline__ float gelu_and_grad_fast(float x, float& grad) {
    float g;
    const float cdf = normal_cdf_fast(x, g);
    grad = fmaf(x * 0.39894228040143267794f, g, cdf);
    return x * cdf;
}
// GELU for the specialised tensor-core epilogues: Phi(x) ~= 0.5 (1 + tanh(x (a + b x^2))) with (a, b) fitted to the erf
// definition on [-8, 8] (max |GELU error| 3.4e-4, max |GELU' error| 6.7e-4 — an order of magnitude below the bf16
// rounding of the stored result; the coefficients of the textbook tanh form give 4.7e-4 / 8.7e-4).  One MUFU (tanh.approx)
// per element instead of two, 6 instructions for GELU, 11 for GELU and GELU' together; `k` is an output scale folded into
// the constants (the dropout keep-scale).  x^2 is clamped so that huge inputs cannot produce 0 * inf in the derivative.
constexpr float kGeluA = 0.79991365f, kGeluB = 0.03497315f;
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
struct GeluK {          // per-thread constants for scale k
    float hk, ka, kb;
    __device__ __forceinline__ explicit GeluK(float k) : hk(0.5f * k), ka(0.5f * k * kGeluA), kb(1.5f * k * kGeluB) {}
};
__device__ __forceinline__ float gelu_tanhfit(float x, const GeluK& c) {
    const float x2 = fminf(x * x, 100.0f);
    const float t = tanh_approx(x * fmaf(kGeluB, x2, kGeluA));
    return x * fmaf(c.hk, t, c.hk);
}
__device__ __forceinline__ float gelu_and_grad_tanhfit(float x, const GeluK& c, float& grad) {
    const float x2 = fminf(x * x, 100.0f);
    const float t = tanh_approx(x * fmaf(kGeluB, x2, kGeluA));
    const float cdf = fmaf(c.hk, t, c.hk);
    grad = fmaf(fmaf(-t, t, 1.0f), x * fmaf(c.kb, x2, c.ka), cdf);
    return x * cdf;
}
__device__ __forceinline__ float dgelu_fast(float x) {
    float g;
    const float cdf = normal_cdf_fast(x, g);
    return fmaf(x * 0.39894228040143267794f, g, cdf);
}

// ----------------------------------------------------------------------------- counter-based dropout RNG
// Philox4x32-7 (the round count the Random123 authors give as the minimum passing BigCrush; the mask only has to be
// uncorrelated, and the generator runs inside GEMM epilogues) keyed by `seed`; one call serves EIGHT consecutive elements: counter = offset + element_index/8,
// element (index % 8) takes the 16-bit field (index % 8) of the 128 random bits and is kept iff field >= round(p * 65536)
// (drop probability quantised to 1/65536).  Stateless, so the backward pass regenerates the forward mask from
// (seed, offset) — no mask tensor is ever stored.
__device__ __forceinline__ uint4 philox4x32(uint64_t ctr, uint64_t key) {
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x243F6A88u, c3 = 0x85A308D3u;
    uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
    for (int r = 0; r < 7; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t dropout_threshold(float p) { return (uint32_t)(p * 65536.0f + 0.5f); }
__device__ __forceinline__ uint32_t dropout_threshold_hi(float p) {     // threshold << 16, clamped to the 16-bit field range
    const uint32_t t = dropout_threshold(p);
    return (t > 65535u ? 65535u : t) << 16;
}
__device__ __forceinline__ uint32_t philox_field16(const uint4& r, uint32_t lane8) {
    const uint32_t w = lane8 < 2 ? r.x : lane8 < 4 ? r.y : lane8 < 6 ? r.z : r.w;
    return (lane8 & 1) ? (w >> 16) : (w & 0xFFFFu);
}
// keep-scale for element `idx`: 1/(1-p) if kept else 0.
__device__ __forceinline__ float dropout_scale(uint64_t idx, float p, float inv_keep, uint64_t seed, uint64_t offset) {
    const uint4 r = philox4x32(offset + (idx >> 3), seed);
    return philox_field16(r, (uint32_t)(idx & 7)) >= (dropout_threshold_hi(p) >> 16) ? inv_keep : 0.0f;
}
// keep flags of the eight elements of one Philox call, without extracting the fields: with T = thr << 16 (thr <= 65535),
// high field >= thr  <=>  w >= T,   low field >= thr  <=>  (w << 16) >= T.
__device__ __forceinline__ void dropout_keep8(bool (&keep)[8], const uint4& r, uint32_t thr_hi) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        keep[2 * i] = (w[i] << 16) >= thr_hi;
        keep[2 * i + 1] = w[i] >= thr_hi;
    }
}
// eight consecutive elements starting at idx (idx % 8 == 0): one Philox call
__device__ __forceinline__ void dropout_scale8(float (&o)[8], uint64_t idx, float p, float inv_keep, uint64_t seed, uint64_t offset) {
    const uint4 r = philox4x32(offset + (idx >> 3), seed);
    bool keep[8];
    dropout_keep8(keep, r, dropout_threshold_hi(p));
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = keep[i] ? inv_keep : 0.0f;
}
// four consecutive elements starting at idx (idx % 4 == 0): the matching half of the group's Philox call
__device__ __forceinline__ float4 dropout_scale4(uint64_t idx, float p, float inv_keep, uint64_t seed, uint64_t offset) {
    const uint4 r = philox4x32(offset + (idx >> 3), seed);
    const uint32_t thr = dropout_threshold_hi(p) >> 16;
    const bool hi = (idx & 4) != 0;
    const uint32_t w0 = hi ? r.z : r.x, w1 = hi ? r.w : r.y;
    float4 o;
    o.x = (w0 & 0xFFFFu) >= thr ? inv_keep : 0.0f;
    o.y = (w0 >> 16) >= thr ? inv_keep : 0.0f;
    o.z = (w1 & 0xFFFFu) >= thr ? inv_keep : 0.0f;
    o.w = (w1 >> 16) >= thr ? inv_keep : 0.0f;
    return o;
}

// ----------------------------------------------------------------------------- GEMM epilogue description
struct Epilogue {
    float alpha, beta;
    const float* bias;
    int flags;
    const void* aux_in;   // c_dtype, ld_aux
    void* aux_out;        // c_dtype, ld_aux
    int64_t ld_aux;
    const float* residual;
    int64_t ld_res;
    void* c2;             // bf16 copy
    int64_t ldc2;
    float dropout_p;
    uint64_t seed, offset;
    const uint64_t* offset_base;   // device word added to `offset` (CUDA-graph replays), or NULL
    int c_dtype;
    float* colsum;                 // fp32 [n] += column sums of the final C (the bias gradient of the layer that consumes dC), or NULL
};
__device__ __forceinline__ uint64_t effective_offset(uint64_t offset, const uint64_t* base) { return base ? offset + __ldg(base) : offset; }

// Apply the epilogue to one accumulator value at (row, col); C addressing is left to the caller.
template <typename TC>
__device__ __forceinline__ float epilogue_value(const Epilogue& e, float acc, int64_t row, int64_t col, int64_t n_cols,
                                                const TC* c_old_ptr) {
    float v = acc * e.alpha;
    if (e.bias) v += __ldg(e.bias + col);
    const float ds = e.dropout_p > 0.0f ? dropout_scale((uint64_t)(row * n_cols + col), e.dropout_p, 1.0f / (1.0f - e.dropout_p), e.seed, effective_offset(e.offset, e.offset_base)) : 1.0f;
    if (e.aux_out) reinterpret_cast<TC*>(e.aux_out)[row * e.ld_aux + col] = from_f32<TC>((e.flags & OVQA_EPI_SAVE_DGELU) ? dgelu_erf(v) * ds : v);
    if (e.flags & OVQA_EPI_GELU) v = gelu_erf(v);
    if (e.flags & OVQA_EPI_RELU) v = fmaxf(v, 0.0f);
    if (e.flags & OVQA_EPI_DGELU) v *= dgelu_erf(to_f32<TC>(reinterpret_cast<const TC*>(e.aux_in)[row * e.ld_aux + col]));
    if (e.flags & OVQA_EPI_MUL_AUX) v *= to_f32<TC>(reinterpret_cast<const TC*>(e.aux_in)[row * e.ld_aux + col]);
    v *= ds;
    if (e.residual) v += __ldg(e.residual + row * e.ld_res + col);
    if (e.beta != 0.0f && c_old_ptr) v += e.beta * to_f32<TC>(*c_old_ptr);
    if (e.colsum) atomicAdd(e.colsum + col, v);
    return v;
}

// internal entry points (implemented in the per-kernel .cu files)
int gemm_simt_f32(const float* A, const float* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                  int64_t ldc, int a_op, int b_op, const Epilogue& epi, int split_k, cudaStream_t st);
int gemm_simt_bf16(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda,
                   int64_t ldb, int64_t ldc, int a_op, int b_op, const Epilogue& epi, int split_k, cudaStream_t st);
int gemm_tcgen05_bf16(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, int64_t m, int64_t n, int64_t k,
                      int64_t lda, int64_t ldb, int64_t ldc, int a_op, int b_op, const Epilogue& epi, int split_k,
                      cudaStream_t st);

}  // namespace ovqa
