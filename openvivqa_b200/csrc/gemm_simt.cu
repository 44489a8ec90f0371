// Exact-fp32 FFMA GEMM: the 1e-4 parity path of ovqa_gemm (and the on-device cross-check of the tcgen05
// kernel).  Plain smem-tiled 64x64x16 with 4x4 register blocking; operands addressed through generic
// (row, k) strides so N/T forms share one kernel.  Not the performance path: bf16 goes to gemm_tcgen05.cu.
#include <type_traits>

#include "common.cuh"

namespace ovqa {

namespace {

constexpr int BK = 16;

// Tile T x T with (T/4)^2 threads, each a 4 x 4 block of outputs.  T = 64 for ordinary sizes; T = 32 when a 64-wide tiling would leave most
// SMs idle (the B-row projections and the classifier of the pooling head, 256 x 512 x 512: 32 CTAs with a 32-step serial k loop took
// 44 us for 0.13 GFLOP) — four times the CTAs, same summation order per output, so results stay deterministic (no split-K atomics).
template <typename TIn, typename TC, int T>
__global__ void __launch_bounds__((T / 4) * (T / 4)) gemm_simt_kernel(const TIn* __restrict__ A, const TIn* __restrict__ B, TC* __restrict__ C,
                                                       int64_t m, int64_t n, int64_t k, int64_t sa_m, int64_t sa_k,
                                                       int64_t sb_n, int64_t sb_k, int64_t ldc, Epilogue epi, int split_k) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    constexpr int BM = T, BN = T, NT = (T / 4) * (T / 4);
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int tx = tid % (T / 4), ty = tid / (T / 4);  // (T/4) x (T/4) threads, each 4x4 outputs

    // k range of this split
    const int64_t k_per = ((k + split_k - 1) / split_k + BK - 1) / BK * BK;
    const int64_t k_begin = (int64_t)blockIdx.z * k_per;
    const int64_t k_end = min(k, k_begin + k_per);

    float acc[4][4] = {};
    // loader mapping: T*BK elements per operand tile over NT threads; consecutive threads follow the operand's contiguous dimension
    const bool a_kc = (sa_k == 1), b_kc = (sb_k == 1);
    for (int64_t kk = k_begin; kk < k_end; kk += BK) {
#pragma unroll
        for (int e = tid; e < T * BK; e += NT) {
            int r, c;  // r: m (or n) index in tile, c: k index in tile
            if (a_kc) { r = e / BK; c = e % BK; } else { c = e / T; r = e % T; }
            const int64_t gm = m0 + r, gk = kk + c;
            As[c][r] = (gm < m && gk < k_end) ? to_f32<TIn>(A[gm * sa_m + gk * sa_k]) : 0.0f;
            if (b_kc) { r = e / BK; c = e % BK; } else { c = e / T; r = e % T; }
            const int64_t gn = n0 + r, gk2 = kk + c;
            Bs[c][r] = (gn < n && gk2 < k_end) ? to_f32<TIn>(B[gn * sb_n + gk2 * sb_k]) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < BK; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[c][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[c][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = m0 + ty * 4 + i;
        if (row >= m) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t col = n0 + tx * 4 + j;
            if (col >= n) continue;
            TC* cp = C + row * ldc + col;
            if (split_k > 1) {
                if constexpr (sizeof(TC) == 4) atomicAdd(reinterpret_cast<float*>(cp), acc[i][j] * epi.alpha);
            } else {
                const float v = epilogue_value<TC>(epi, acc[i][j], row, col, n, cp);
                *cp = from_f32<TC>(v);
                if (epi.c2) reinterpret_cast<__nv_bfloat16*>(epi.c2)[row * epi.ldc2 + col] = __float2bfloat16_rn(v);
            }
        }
    }
}

// Small problems (a few thousand outputs, k in the hundreds: the B-row projections and the classifier of the pooling head and their
// gradients): one 32 x 32 output tile per CTA, the reduction split FOUR ways inside the CTA (256 threads = 4 k-groups x 64), partial tiles
// combined through shared memory in a fixed order — deterministic, and the serial k loop is a quarter as long (34 us -> single digits for
// 256 x 512 x 512).
template <typename TIn, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_small_kernel(const TIn* __restrict__ A, const TIn* __restrict__ B, TC* __restrict__ C, int64_t m,
                                                              int64_t n, int64_t k, int64_t sa_m, int64_t sa_k, int64_t sb_n, int64_t sb_k, int64_t ldc,
                                                              Epilogue epi) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    constexpr int T = 32, KG = 4;
    __shared__ float As[KG][BK][T + 4];
    __shared__ float Bs[KG][BK][T + 4];
    __shared__ float red[KG - 1][T][T + 1];
    const int tid = threadIdx.x, kg = tid / 64, t = tid % 64;
    const int64_t m0 = (int64_t)blockIdx.y * T, n0 = (int64_t)blockIdx.x * T;
    const int tx = t % 8, ty = t / 8;
    const int64_t k_per = ((k + KG - 1) / KG + BK - 1) / BK * BK;
    const int64_t k_begin = (int64_t)kg * k_per, k_end = min(k, k_begin + k_per);
    const int64_t iters = (k_per + BK - 1) / BK;                     // same trip count in every group (block-wide barriers below)
    float acc[4][4] = {};
    const bool a_kc = (sa_k == 1), b_kc = (sb_k == 1);
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t kk = k_begin + it * BK;
#pragma unroll
        for (int e = t; e < T * BK; e += 64) {
            int r, c;
            if (a_kc) { r = e / BK; c = e % BK; } else { c = e / T; r = e % T; }
            const int64_t gm = m0 + r, gk = kk + c;
            As[kg][c][r] = (gm < m && gk < k_end) ? to_f32<TIn>(A[gm * sa_m + gk * sa_k]) : 0.0f;
            if (b_kc) { r = e / BK; c = e % BK; } else { c = e / T; r = e % T; }
            const int64_t gn = n0 + r, gk2 = kk + c;
            Bs[kg][c][r] = (gn < n && gk2 < k_end) ? to_f32<TIn>(B[gn * sb_n + gk2 * sb_k]) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < BK; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kg][c][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kg][c][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (kg > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[kg - 1][ty * 4 + i][tx * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (kg == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t row = m0 + ty * 4 + i;
            if (row >= m) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t col = n0 + tx * 4 + j;
                if (col >= n) continue;
                float v = acc[i][j];
#pragma unroll
                for (int g = 0; g < KG - 1; ++g) v += red[g][ty * 4 + i][tx * 4 + j];
                TC* cp = C + row * ldc + col;
                const float o = epilogue_value<TC>(epi, v, row, col, n, cp);
                *cp = from_f32<TC>(o);
                if (epi.c2) reinterpret_cast<__nv_bfloat16*>(epi.c2)[row * epi.ldc2 + col] = __float2bfloat16_rn(o);
            }
        }
    }
}

template <typename TIn>
int launch(const TIn* A, const TIn* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
           int a_op, int b_op, const Epilogue& epi, int split_k, cudaStream_t st) {
    if (m == 0 || n == 0) return OVQA_OK;
    const int64_t sa_m = a_op == OVQA_OP_N ? lda : 1, sa_k = a_op == OVQA_OP_N ? 1 : lda;
    const int64_t sb_n = b_op == OVQA_OP_N ? ldb : 1, sb_k = b_op == OVQA_OP_N ? 1 : ldb;
    if (split_k < 1) split_k = 1;
    if (split_k > 1 && epi.c_dtype != OVQA_F32) return OVQA_ERR_BAD_ARG;
    const int64_t tiles64 = ((n + 63) / 64) * ((m + 63) / 64) * split_k;
    const bool small = tiles64 * 2 <= sm_count();
    if (small && split_k == 1 && k >= 256) {
        dim3 grid((unsigned)((n + 31) / 32), (unsigned)((m + 31) / 32));
        if (epi.c_dtype == OVQA_F32)
            OVQA_CUDA_OK(launch_pdl(gemm_simt_small_kernel<TIn, float>, dim3(grid), dim3(256), (size_t)(0), st, A, B, (float*)C, m, n, k, sa_m, sa_k, sb_n, sb_k, ldc, epi));
        else
            OVQA_CUDA_OK(launch_pdl(gemm_simt_small_kernel<TIn, __nv_bfloat16>, dim3(grid), dim3(256), (size_t)(0), st, A, B, (__nv_bfloat16*)C, m, n, k, sa_m, sa_k, sb_n, sb_k, ldc, epi));
        OVQA_LAUNCH_OK();
        return OVQA_OK;
    }
    auto go = [&](auto tile_tag) {
        constexpr int T = decltype(tile_tag)::value;
        dim3 grid((unsigned)((n + T - 1) / T), (unsigned)((m + T - 1) / T), (unsigned)split_k);
        if (epi.c_dtype == OVQA_F32)
            OVQA_CUDA_OK(launch_pdl(gemm_simt_kernel<TIn, float, T>, dim3(grid), dim3((T / 4) * (T / 4)), (size_t)(0), st, A, B, (float*)C, m, n, k, sa_m, sa_k, sb_n, sb_k, ldc, epi, split_k));
        else
            OVQA_CUDA_OK(launch_pdl(gemm_simt_kernel<TIn, __nv_bfloat16, T>, dim3(grid), dim3((T / 4) * (T / 4)), (size_t)(0), st, A, B, (__nv_bfloat16*)C, m, n, k, sa_m, sa_k, sb_n, sb_k, ldc, epi, split_k));
    };
    if (small) go(std::integral_constant<int, 32>{}); else go(std::integral_constant<int, 64>{});
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

}  // namespace

int gemm_simt_f32(const float* A, const float* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                  int64_t ldc, int a_op, int b_op, const Epilogue& epi, int split_k, cudaStream_t st) {
    return launch<float>(A, B, C, m, n, k, lda, ldb, ldc, a_op, b_op, epi, split_k, st);
}
int gemm_simt_bf16(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda,
                   int64_t ldb, int64_t ldc, int a_op, int b_op, const Epilogue& epi, int split_k, cudaStream_t st) {
    return launch<__nv_bfloat16>(A, B, C, m, n, k, lda, ldb, ldc, a_op, b_op, epi, split_k, st);
}

}  // namespace ovqa
