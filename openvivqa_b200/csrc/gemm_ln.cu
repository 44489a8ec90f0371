// Output projection + bias + dropout + residual + LayerNorm in ONE kernel, for d_model = 512 (MCAN / cross-modality models):
//     s = dropout(A W^T + bias) + residual          (attentions.py:58,331: fc_o;  positionwise_feed_forward.py:25-26: fc2)
//     y = LayerNorm(s) * gamma + beta
// A 128-row tile keeps its FULL 512-column fp32 rows in tensor memory (128 lanes x 512 columns = all of TMEM), so the row statistics are
// taken where the accumulator lives: the separate LayerNorm launch, its 2 KB-per-row re-read of s and the launch gap are gone (33 LayerNorm
// launches = 0.69 ms of a 10.7 ms MCAN step, and each one re-read what the GEMM had just written).
//
// Warp roles (320 threads, 1 CTA/SM): warp 0 = TMA producer (A: one 64 x 128 box, W: two 64 x 256 boxes per k-block, SWIZZLE_128B, 2 stages of
// 80 KB), warp 1 = TMEM allocator + MMA issuer (two N = 256 tcgen05.mma per 16-wide k-step), warps 2-9 = epilogue: the warp's TMEM lane
// quadrant (32 rows) x one half of the columns (8 chunks of 32).
//   pass 1, per 32 x 32 chunk: accumulator -> registers (row per lane), + bias, dropout (the library's Philox stream: backward regenerates
//           the mask), + residual (TMA-loaded into the warp's staging buffer one chunk ahead), row sum / sum of squares; s goes BACK into
//           tensor memory (tcgen05.st) and, for training, to global memory through the staging buffer (TMA store);
//   the two column halves of a row exchange their partial sums through shared memory;
//   pass 2, per chunk: s from tensor memory, normalise, gamma / beta, y (fp32) and its bf16 copy out through the staging buffers (TMA stores).
// With all 512 columns taken by one accumulator there is no second buffer: the epilogue of a tile does not overlap the next tile's main
// loop (the producer still runs ahead by the two operand stages).
#include "tcgen05.cuh"

namespace ovqa {
namespace {

using namespace tc5;

constexpr int kLnN = 512, kLnBM = 128, kLnBK = 64, kLnStages = 2, kLnEpw = 8;
constexpr int kLnABytes = kLnBM * kLnBK * 2, kLnBBytes = kLnN * kLnBK * 2, kLnStageBytes = kLnABytes + kLnBBytes;
constexpr int kLnThreads = 64 + kLnEpw * 32;
constexpr int kLnScratch = kLnEpw * 2 * 4096;
constexpr int kLnBarBytes = 192;
constexpr int kLnUsed = kLnStages * kLnStageBytes + kLnScratch + 2 * 128 * 2 * 4 /*row statistics exchange*/ + kLnBarBytes;
constexpr int kLnSmem = 227 * 1024;            // everything an SM has: the operand stages need a 1024-byte aligned start, which leaves
static_assert(kLnUsed + 768 <= kLnSmem, "shared memory budget");   // 832 bytes for the alignment pad (checked at run time)
static_assert((2 * kLnStages + 2 + 2 * kLnEpw) * 8 + 8 <= kLnBarBytes, "barrier region");

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
          "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
          "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ln_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct GemmLnArgs {
    const float *bias, *gamma, *beta;
    float *mean, *rstd;
    int64_t m, k;
    float eps, dropout_p;
    uint64_t seed, offset;
    const uint64_t* offset_base;
    int store_s, store_y16;
};

__global__ void __launch_bounds__(kLnThreads, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_res,
               const __grid_constant__ CUtensorMap tma_s, const __grid_constant__ CUtensorMap tma_y, const __grid_constant__ CUtensorMap tma_y16,
               GemmLnArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kLnStages * kLnABytes;
    uint8_t* scratch = smem + kLnStages * kLnStageBytes;
    float* s_red = reinterpret_cast<float*>(scratch + kLnScratch);            // [2 column halves][128 rows][sum, sum of squares]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 2 * 128 * 2);
    uint64_t* full_bar = bars;                         // [stages]  TMA -> MMA
    uint64_t* empty_bar = bars + kLnStages;            // [stages]  MMA -> TMA
    uint64_t* tmem_full = bars + 2 * kLnStages;        // MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 1;              // epilogue -> MMA (8 arrivals)
    uint64_t* in_bar = tmem_empty + 1;                 // [8 warps][2 buffers]  TMA -> epilogue warp (residual chunk)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_bar + kLnEpw * 2);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    if ((int)(smem - smem_raw) + kLnUsed > kLnSmem) {       // (never on the configurations seen: the dynamic window starts 1024-aligned)
        if (threadIdx.x == 0) printf("ovqa gemm_ln: shared-memory window not aligned enough (pad %d)\n", (int)(smem - smem_raw));
        __trap();
    }
    const int num_m_tiles = (int)((g.m + kLnBM - 1) / kLnBM);
    const int num_k_blocks = (int)((g.k + kLnBK - 1) / kLnBK);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_res)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_y)) : "memory");
        for (int i = 0; i < kLnStages; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, kLnEpw);
        for (int i = 0; i < kLnEpw * 2; ++i) mbar_init(in_bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // programmatic dependent launch: nothing above touches global memory
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    mbar_expect_tx(full_bar + stage, kLnStageBytes);
                    tma_load_2d(smem_a + stage * kLnABytes, &tma_a, full_bar + stage, kb * kLnBK, tile * kLnBM);
                    tma_load_2d(smem_b + stage * kLnBBytes, &tma_w, full_bar + stage, kb * kLnBK, 0);
                    tma_load_2d(smem_b + stage * kLnBBytes + kLnBBytes / 2, &tma_w, full_bar + stage, kb * kLnBK, 256);
                    if (++stage == kLnStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(kLnBM, 256, false, false);
            int stage = 0, nt = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++nt) {
                mbar_wait(tmem_empty, ((uint32_t)nt & 1u) ^ 1u);          // the previous tile's epilogue has drained the accumulator
                tcgen05_fence_after();
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(full_bar + stage, phase);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * kLnABytes), b_addr = smem_u32(smem_b + stage * kLnBBytes);
#pragma unroll
                    for (int kk = 0; kk < kLnBK / 16; ++kk) {
                        const uint64_t a_desc = make_smem_desc(a_addr + kk * 32, 16, 1024);
                        const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
                        tcgen05_mma_bf16(tmem_base, a_desc, make_smem_desc(b_addr + kk * 32, 16, 1024), idesc, acc);
                        tcgen05_mma_bf16(tmem_base + 256u, a_desc, make_smem_desc(b_addr + kLnBBytes / 2 + kk * 32, 16, 1024), idesc, acc);
                    }
                    tcgen05_commit(empty_bar + stage);
                    if (++stage == kLnStages) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(tmem_full);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps 2..9
        const int ew = warp - 2, quad = warp % 4, colhalf = ew / 4;
        const uint32_t bufs = smem_u32(scratch) + (uint32_t)ew * 8192u;              // two staging buffers of 4 KB
        uint64_t* my_bar = in_bar + ew * 2;
        const uint32_t row_off = (uint32_t)lane * 128u, sx = (uint32_t)(lane & 7);   // fp32 chunk: 128-byte rows under SWIZZLE_128B
        const uint32_t row_off16 = (uint32_t)lane * 64u, sx16 = (uint32_t)((lane >> 1) & 3);   // bf16 chunk: 64-byte rows under SWIZZLE_64B
        const bool drop = g.dropout_p > 0.f;
        const float k_out = drop ? 1.0f / (1.0f - g.dropout_p) : 1.0f;
        const uint64_t rng_offset = drop ? effective_offset(g.offset, g.offset_base) : 0;
        const uint32_t thr_hi = dropout_threshold_hi(g.dropout_p);
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(256 * colhalf);
        uint32_t par[2] = {0u, 0u};
        int nt = 0;
        for (int tile = blockIdx.x; tile < num_m_tiles; tile += gridDim.x, ++nt) {
            const int row0 = tile * kLnBM + quad * 32;
            const int64_t row = (int64_t)row0 + lane;
            auto load_res = [&](int c) {            // lane 0: residual rows of chunk c -> buffer c & 1
                mbar_expect_tx(my_bar + (c & 1), 4096u);
                tma_load_2d_u32(bufs + (uint32_t)(c & 1) * 4096u, &tma_res, my_bar + (c & 1), 256 * colhalf + 32 * c, row0);
            };
            if (lane == 0) { bulk_wait_read<0>(); load_res(0); }      // (the previous tile's stores have read both buffers)
            mbar_wait(tmem_full, (uint32_t)nt & 1u);
            tcgen05_fence_after();
            float sum = 0.f, sq = 0.f;
            // ---------------------------------------------------------- pass 1
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                const int col0 = 256 * colhalf + 32 * c;
                const uint32_t buf = bufs + (uint32_t)(c & 1) * 4096u, my = buf + row_off;
                if (lane == 0 && c + 1 < 8) {
                    bulk_wait_read<0>();                                  // the store that last read buffer (c + 1) & 1 (chunk c - 1) is done with it
                    load_res(c + 1);
                }
                const float bias_l = g.bias ? __ldg(g.bias + col0 + lane) * k_out : 0.f;
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_lane + (uint32_t)(32 * c), r);
                mbar_wait(my_bar + (c & 1), par[c & 1]);
                par[c & 1] ^= 1u;
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), k_out, __shfl_sync(0xffffffffu, bias_l, j));
                if (drop) {
                    const uint64_t ctr = rng_offset + (uint64_t)((row * kLnN + col0) >> 3);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 rr = philox4x32(ctr + (uint64_t)q, g.seed);
                        bool keep[8];
                        dropout_keep8(keep, rr, thr_hi);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[8 * q + j] = keep[j] ? v[8 * q + j] : 0.0f;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t ga = my + (((uint32_t)q ^ sx) << 4);
                    float x0, x1, x2, x3;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(ga) : "memory");
                    v[4 * q] += x0; v[4 * q + 1] += x1; v[4 * q + 2] += x2; v[4 * q + 3] += x3;
                    if (g.store_s)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(ga), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3]) : "memory");
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) { sum += v[j]; sq = fmaf(v[j], v[j], sq); r[j] = __float_as_uint(v[j]); }
                tmem_st_32x32b_x32(t_lane + (uint32_t)(32 * c), r);
                tmem_st_wait();                                           // (the registers are reused by the next chunk)
                if (g.store_s) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { tma_store_2d_u32(&tma_s, buf, col0, row0); bulk_commit(); }
                } else {
                    __syncwarp();                                         // every lane has read its residual row before the buffer is reloaded
                }
            }
            // ---------------------------------------------------------- row statistics: the two column halves of a row meet in shared memory
            float* mine = s_red + ((colhalf * 128) + quad * 32 + lane) * 2;
            mine[0] = sum; mine[1] = sq;
            ln_bar_sync(1, kLnEpw * 32);
            const float* other = s_red + (((colhalf ^ 1) * 128) + quad * 32 + lane) * 2;
            const float mean = (sum + other[0]) * (1.0f / kLnN);
            const float var = fmaxf((sq + other[1]) * (1.0f / kLnN) - mean * mean, 0.0f);
            const float rstd = rsqrtf(var + g.eps);
            if (colhalf == 0 && row < g.m) {
                if (g.mean) g.mean[row] = mean;
                if (g.rstd) g.rstd[row] = rstd;
            }
            // ---------------------------------------------------------- pass 2
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                const int col0 = 256 * colhalf + 32 * c;
                const float gam = __ldg(g.gamma + col0 + lane), bet = __ldg(g.beta + col0 + lane);
                uint32_t r[32];
                tmem_ld_32x32b_x32(t_lane + (uint32_t)(32 * c), r);
                if (lane == 0) bulk_wait_read<0>();                       // the previous chunk's stores have read both buffers
                __syncwarp();
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = fmaf((__uint_as_float(r[j]) - mean) * rstd, __shfl_sync(0xffffffffu, gam, j), __shfl_sync(0xffffffffu, bet, j));
                const uint32_t my = bufs + row_off, my16 = bufs + 4096u + row_off16;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my + (((uint32_t)q ^ sx) << 4)), "f"(v[4 * q]), "f"(v[4 * q + 1]), "f"(v[4 * q + 2]),
                                 "f"(v[4 * q + 3]) : "memory");
                if (g.store_y16) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my16 + (((uint32_t)q ^ sx16) << 4)), "r"(pack_bf16x2(v[8 * q], v[8 * q + 1])),
                                     "r"(pack_bf16x2(v[8 * q + 2], v[8 * q + 3])), "r"(pack_bf16x2(v[8 * q + 4], v[8 * q + 5])), "r"(pack_bf16x2(v[8 * q + 6], v[8 * q + 7]))
                                     : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d_u32(&tma_y, bufs, col0, row0);
                    if (g.store_y16) tma_store_2d_u32(&tma_y16, bufs + 4096u, col0, row0);
                    bulk_commit();
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            ln_bar_sync(1, kLnEpw * 32);                                 // s_red is rewritten by the next tile
        }
        if (lane == 0) bulk_wait_all();                                   // shared memory stays valid until the last store has read it
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

inline bool ln_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace ovqa

using namespace ovqa;

// y = LayerNorm(dropout(A W^T + bias) + residual) * gamma + beta for N = 512 output columns; see the header of this file.
// A [M,K] bf16 (row pitch lda), W [512,K] bf16 (ldw), residual / s_out / y [M,512] fp32 contiguous, y_bf16 [M,512] bf16 contiguous.
// s_out (the LayerNorm input, needed by the backward pass), y_bf16, mean, rstd may be NULL.  OVQA_ERR_BAD_SHAPE for other widths: the
// caller then runs ovqa_gemm + ovqa_layernorm_fwd.
extern "C" int ovqa_gemm_ln(const void* A, const void* W, const float* bias, const float* residual, const float* gamma, const float* beta,
                            float* s_out, float* y, void* y_bf16, float* mean, float* rstd, int64_t M, int64_t N, int64_t K, int64_t lda,
                            int64_t ldw, float eps, float dropout_p, uint64_t seed, uint64_t offset, const uint64_t* offset_base, void* stream) {
    if (!A || !W || !residual || !gamma || !beta || !y || M < 0 || K <= 0) return OVQA_ERR_BAD_ARG;
    if (N != kLnN || M >= ((int64_t)1 << 30) || dropout_p < 0.f || dropout_p >= 1.f) return OVQA_ERR_BAD_SHAPE;
    if (!ln_al16(A) || !ln_al16(W) || !ln_al16(residual) || !ln_al16(y) || !ln_al16(s_out) || !ln_al16(y_bf16) || lda % 8 || ldw % 8) return OVQA_ERR_ALIGNMENT;
    if (M == 0) return OVQA_OK;
    CUtensorMap ma, mw, mr, ms, my, my16;
    int rc;
    if ((rc = make_tma_map_bf16(&ma, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, kLnBK, kLnBM))) return rc;
    if ((rc = make_tma_map_bf16(&mw, W, (uint64_t)K, (uint64_t)kLnN, (uint64_t)ldw, kLnBK, 256))) return rc;
    if ((rc = make_tma_map(&mr, residual, 4, kLnN, (uint64_t)M, kLnN, 32, 32, 128))) return rc;
    if ((rc = make_tma_map(&my, y, 4, kLnN, (uint64_t)M, kLnN, 32, 32, 128))) return rc;
    ms = my; my16 = my;
    if (s_out && (rc = make_tma_map(&ms, s_out, 4, kLnN, (uint64_t)M, kLnN, 32, 32, 128))) return rc;
    if (y_bf16 && (rc = make_tma_map(&my16, y_bf16, 2, kLnN, (uint64_t)M, kLnN, 32, 32, 64))) return rc;
    static OncePerDevice attr_set;
    if (attr_set.pending()) { OVQA_CUDA_OK(cudaFuncSetAttribute(gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLnSmem)); attr_set.mark(); }
    GemmLnArgs g{bias, gamma, beta, mean, rstd, M, K, eps, dropout_p, seed, offset, offset_base, s_out ? 1 : 0, y_bf16 ? 1 : 0};
    const int64_t tiles = (M + kLnBM - 1) / kLnBM;
    const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
    OVQA_CUDA_OK(launch_pdl(gemm_ln_kernel, dim3((unsigned)grid), dim3(kLnThreads), (size_t)kLnSmem, (cudaStream_t)stream, ma, mw, mr, ms, my, my16, g));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
