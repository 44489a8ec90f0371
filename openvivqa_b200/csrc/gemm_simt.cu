// Exact-fp32 FFMA GEMM: the 1e-4 parity path of ovqa_gemm (and the on-device cross-check of the tcgen05
// kernel).  Plain smem-tiled 64x64x16 with 4x4 register blocking; operands addressed through generic
// (row, k) strides so N/T forms share one kernel.  Not the performance path: bf16 goes to gemm_tcgen05.cu.
#include "common.cuh"

namespace ovqa {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename TIn, typename TC>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const TIn* __restrict__ A, const TIn* __restrict__ B, TC* __restrict__ C,
                                                       int64_t m, int64_t n, int64_t k, int64_t sa_m, int64_t sa_k,
                                                       int64_t sb_n, int64_t sb_k, int64_t ldc, Epilogue epi, int split_k) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int tx = tid % 16, ty = tid / 16;  // 16x16 threads, each 4x4 outputs

    // k range of this split
    const int64_t k_per = ((k + split_k - 1) / split_k + BK - 1) / BK * BK;
    const int64_t k_begin = (int64_t)blockIdx.z * k_per;
    const int64_t k_end = min(k, k_begin + k_per);

    float acc[4][4] = {};
    // loader mapping: 1024 elements per tile / 256 threads = 4 each
    const bool a_kc = (sa_k == 1), b_kc = (sb_k == 1);
    for (int64_t kk = k_begin; kk < k_end; kk += BK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int r, c;  // r: m (or n) index in tile, c: k index in tile
            if (a_kc) { r = tid / 4; c = (tid % 4) * 4 + i; } else { c = tid / 16; r = (tid % 16) * 4 + i; }
            const int64_t gm = m0 + r, gk = kk + c;
            As[c][r] = (gm < m && gk < k_end) ? to_f32<TIn>(A[gm * sa_m + gk * sa_k]) : 0.0f;
            if (b_kc) { r = tid / 4; c = (tid % 4) * 4 + i; } else { c = tid / 16; r = (tid % 16) * 4 + i; }
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

template <typename TIn>
int launch(const TIn* A, const TIn* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
           int a_op, int b_op, const Epilogue& epi, int split_k, cudaStream_t st) {
    if (m == 0 || n == 0) return OVQA_OK;
    const int64_t sa_m = a_op == OVQA_OP_N ? lda : 1, sa_k = a_op == OVQA_OP_N ? 1 : lda;
    const int64_t sb_n = b_op == OVQA_OP_N ? ldb : 1, sb_k = b_op == OVQA_OP_N ? 1 : ldb;
    if (split_k < 1) split_k = 1;
    if (split_k > 1 && epi.c_dtype != OVQA_F32) return OVQA_ERR_BAD_ARG;
    dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)split_k);
    if (epi.c_dtype == OVQA_F32)
        gemm_simt_kernel<TIn, float><<<grid, NT, 0, st>>>(A, B, (float*)C, m, n, k, sa_m, sa_k, sb_n, sb_k, ldc, epi, split_k);
    else
        gemm_simt_kernel<TIn, __nv_bfloat16><<<grid, NT, 0, st>>>(A, B, (__nv_bfloat16*)C, m, n, k, sa_m, sa_k, sb_n, sb_k, ldc, epi, split_k);
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
