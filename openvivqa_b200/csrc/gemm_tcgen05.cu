// bf16 x bf16 -> fp32 GEMM on 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), operands staged
// in shared memory by TMA (cp.async.bulk.tensor, SWIZZLE_128B), persistent over output tiles with a
// double-buffered TMEM accumulator so the epilogue of tile i overlaps the mainloop of tile i+1.
//
// Warp roles (576 threads, 1 CTA/SM):  warp 0 = TMA producer (one elected lane), warp 1 = TMEM allocator +
// MMA issuer (one elected lane), warps 2..17 = epilogue: four warps per TMEM lane quadrant, each owning every fourth
// 32-column chunk (the fused epilogues are latency/issue bound: 16 warps hide what 8 could not — profiles/r01d).  A chunk goes TMEM -> registers (row per lane) -> XOR-swizzled shared-memory scratch -> registers
// (8 consecutive columns per lane, 4 lanes per row), so every global access of the fused epilogue (bias, saved
// pre-activation, residual, C, bf16 copy, split-K reductions) is a coalesced 16-byte vector covering whole 128-byte
// row segments, one Philox call serves a lane's 8 elements, and address arithmetic is amortised over 8 outputs.
// With K = 512 the mainloop of a 128x256 tile is only ~4k cycles: the epilogue (GELU + dropout in fc1) runs in
// its shadow on the other TMEM buffer and must not be the longer of the two.
//
// Operand forms (ovqa.h): OVQA_OP_N = reduction dim contiguous ("K-major" UMMA operand, one TMA box of
// 64(k) x ROWS); OVQA_OP_T = M/N dim contiguous ("MN-major" UMMA operand, ROWS/64 TMA boxes of 64(mn) x 64(k)).
// The latter lets dgrad (dy W) and wgrad (dy^T x) read the tensors where they lie, with no transposes.
#include <mutex>
#include <unordered_map>

#include "tcgen05.cuh"

namespace ovqa {

namespace {

using namespace tc5;

constexpr int BM = 128;          // UMMA M (cta_group::1)
constexpr int BK = 64;           // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
// Epilogue warps per CTA (template parameter EPW).  The specialised epilogues run on 8 warps (<= 168 registers per thread: 10 warps, 3 per
// sub-partition) and are fed by the TMA unit (below); the run-time-feature epilogue keeps 16 warps (<= 96 registers) to hide the latency of
// its per-element global accesses.
constexpr int epw_for(int mode, int feat) { return (mode == 1 /*EPI_FULL*/ && feat < 0) ? 16 : 8; }
constexpr int kAccStages = 2;

// CTAS = 1: one CTA per 128 x BN tile.  CTAS = 2: a CTA pair (cluster of 2, cta_group::2) per 256 x BN tile — each CTA stages its
// own 128 rows of A and HALF of B (the MMA reads the other half from the peer's shared memory): 32 KB instead of 48 KB of L2 -> smem
// traffic per SM and k-block at BN = 256 (profiles/r01d: the single-CTA 128 x 256 tile is bound by the L2 -> SM path, not by the tensor pipe).
//
// TMA-fed epilogue (8-warp kernels): every epilogue warp owns kEpiBufs staging buffers of 4 KB, one per 32 x 32 chunk in flight.  The
// chunk's input operand (residual rows, fp32; saved derivative, bf16) is brought in by cp.async.bulk.tensor two chunks ahead, the lane
// that holds accumulator row r (tcgen05.ld 32x32b: row per lane) finishes its 32 columns in registers, overwrites ITS OWN row of the
// buffer with the result (128B / 64B swizzle: conflict-free), and one lane hands the buffer to the TMA unit (tensor store, or fp32
// add-reduction for split-K).  No per-element global access, no transposition pass, ragged edges clipped by the tensor map.
// The pipeline depth and the number of staging buffers are chosen per LAUNCH (kernel arguments `stages`, `nb`; shared memory carved
// at run time): short reductions (K = 512) are bound by the epilogue and want 3 buffers per warp, long ones by the main loop and want
// the 6-deep operand ring (host: launch()).
template <int BN, int CTAS, int EPW> struct Cfg {
    static constexpr int kBarBytes = 512;
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = (BN / CTAS) * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTmemCols = kAccStages * BN;  // 256 or 512: power of two
    static constexpr int kMaxStages = 6, kMaxEpiBufs = EPW == 16 ? 1 : 3;   // (16-warp kernels: the run-time-feature epilogue, one scratch chunk per warp)
    static constexpr int kSmemLimit = 227 * 1024;
    static constexpr int scratch_bytes(int nb) { return EPW * nb * 4096; }                  // 32x32 fp32 chunks
    static constexpr int stages_for(int nb) {
        const int fit = (kSmemLimit - 1024 /*align slack*/ - kBarBytes - scratch_bytes(nb)) / kStageBytes;
        return fit > kMaxStages ? kMaxStages : fit;
    }
    static constexpr int smem_bytes(int stages, int nb) { return stages * kStageBytes + scratch_bytes(nb) + 1024 + kBarBytes; }
    static_assert(stages_for(1) >= 3, "shared memory budget");
    static_assert((2 * kMaxStages + 2 * kAccStages + EPW * kMaxEpiBufs) * 8 + 8 <= kBarBytes, "barrier region");
};

// ----------------------------------------------------------------------------- epilogue pieces
enum { EPI_PLAIN = 0, EPI_FULL = 1, EPI_SPLITK = 2 };   // alpha/bias only | every fused feature | fp32 atomic accumulation
// EPI_FULL comes in two flavours.  FEAT == F_RUNTIME tests the Epilogue fields at run time (any combination; the compiler
// keeps every feature's state live and serialises the per-element math: profiles/r01d, 45 instructions per element at an
// IPC of 0.09 per warp).  FEAT >= 0 is a compile-time feature set for the combinations the attention / feed-forward blocks
// actually launch: no feature tests, the operands a pass needs are requested before the accumulator is read, and the eight
// elements of a pass are independent instruction streams.
enum : int { F_RUNTIME = -1, F_BIAS = 1, F_GELU = 2, F_SAVE_D = 4, F_MULAUX = 8, F_DROP = 16, F_RES = 32, F_COLSUM = 64 };

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, float& a, float& b, float& c, float& d) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr) : "memory");
}
// A lane's 8 values of a chunk row are columns [col, col+4) and [col+HI, col+HI+4).  bf16 results: HI = 4 (8 consecutive columns, one
// 16-byte access).  fp32 results: HI = 16 — the four lanes of a row then cover 64 CONTIGUOUS bytes per access (two whole sectors);
// with 8 consecutive fp32 columns per lane every 16-byte access was half a sector at a 32-byte stride, i.e. twice the L2
// transactions for C, the residual and the split-K atomics.
template <typename TC, int HI> __device__ __forceinline__ void store8(TC* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float, 4>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<float, 16>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 16) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16, 4>(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16, 16>(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
    *reinterpret_cast<uint2*>(p + 16) = make_uint2(pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}
template <typename TC, int HI> __device__ __forceinline__ void load8(const TC* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float, 4>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<float, 16>(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 16));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16, 4>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
    v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16, 16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(p)), b = __ldg(reinterpret_cast<const uint2*>(p + 16));
    const float2 p0 = unpack_bf16x2(a.x), p1 = unpack_bf16x2(a.y), p2 = unpack_bf16x2(b.x), p3 = unpack_bf16x2(b.y);
    v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
}
// Column sums of a 32 x 32 chunk: every lane holds partial sums of its 8 columns over the rows it finished (rows rr, rr+8, ...);
// the 8 lanes with the same column group (lane % 4) are combined by shuffles, lanes 0..3 add the chunk's 32 sums to colsum[].
template <int HI>
__device__ __forceinline__ void colsum_flush(float* __restrict__ colsum, int64_t col, const float (&cs)[8], int lane) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        t[j] = cs[j];
        t[j] += __shfl_xor_sync(0xffffffffu, t[j], 4);
        t[j] += __shfl_xor_sync(0xffffffffu, t[j], 8);
        t[j] += __shfl_xor_sync(0xffffffffu, t[j], 16);
    }
    if (lane < 4) {
        atomicAdd(reinterpret_cast<float4*>(colsum + col), make_float4(t[0], t[1], t[2], t[3]));
        atomicAdd(reinterpret_cast<float4*>(colsum + col + HI), make_float4(t[4], t[5], t[6], t[7]));
    }
}
// slow path for ragged column tails / pitches that do not allow 16-byte vectors: one element, every feature
template <typename TC>
__device__ __noinline__ void epilogue_scalar(const Epilogue& epi, float acc, int64_t row, int64_t col, int64_t n, TC* C, int64_t ldc, int split_k) {
    TC* cp = C + row * ldc + col;
    if (split_k > 1) {
        if constexpr (sizeof(TC) == 4) atomicAdd(reinterpret_cast<float*>(cp), acc * epi.alpha);
        return;
    }
    const float o = epilogue_value<TC>(epi, acc, row, col, n, cp);
    *cp = from_f32<TC>(o);
    if (epi.c2) reinterpret_cast<__nv_bfloat16*>(epi.c2)[row * epi.ldc2 + col] = __float2bfloat16_rn(o);
}

// ----------------------------------------------------------------------------- kernel
template <int BN, bool A_MN, bool B_MN, typename TC, int MODE, int FEAT, int CTAS, int EPW>
__global__ void __launch_bounds__(EPW == 16 ? 640 : 384, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const __grid_constant__ CUtensorMap tma_c,
                    const __grid_constant__ CUtensorMap tma_x, TC* __restrict__ C, int64_t m, int64_t n, int64_t k, int64_t ldc, Epilogue epi,
                    int split_k, int use_tma, int nb, int stages) {
    using cfg = Cfg<BN, CTAS, EPW>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + stages * cfg::kABytes;
    float4* scratch = reinterpret_cast<float4*>(smem + stages * cfg::kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + stages * cfg::kStageBytes + cfg::scratch_bytes(nb));
    uint64_t* full_bar = bars;                              // [kStages]   TMA -> MMA
    uint64_t* empty_bar = bars + stages;                    // [stages]    MMA -> TMA
    uint64_t* tmem_full = bars + 2 * stages;                // [2]         MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + kAccStages;          // [2]         epilogue -> MMA
    uint64_t* epi_bar = tmem_empty + kAccStages;            // [EPW * kEpiBufs]  TMA -> epilogue warp (input operand of a chunk)
    uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(epi_bar + EPW * nb);

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    constexpr int TM = BM * CTAS;                            // rows of a work tile (a CTA pair shares one 256-row tile)
    const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
    const int num_m_tiles = (int)((m + TM - 1) / TM), num_n_tiles = (int)((n + BN - 1) / BN);
    const int num_k_blocks = (int)((k + BK - 1) / BK);
    const int num_tiles = num_m_tiles * num_n_tiles * split_k;
    const int tile0 = (int)blockIdx.x / CTAS, tile_step = (int)gridDim.x / CTAS;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
        if (use_tma) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_c)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_x)) : "memory");
        }
        // full: the leader's own arrive.expect_tx (+ the peer's remote arrive); tmem_empty: the epilogue warps of every CTA of the pair
        for (int i = 0; i < stages; ++i) { mbar_init(full_bar + i, CTAS); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, EPW * CTAS); }
        for (int i = 0; i < EPW * nb; ++i) mbar_init(epi_bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"((uint32_t)cfg::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"((uint32_t)cfg::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_slot;
    // Programmatic dependent launch: everything above (barriers, tensor memory, descriptor prefetch) may run while the previous kernel of
    // the stream drains; nothing below may touch global memory before that kernel's writes are visible.  (No-op without the launch attribute.)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // ... and the NEXT kernel of the stream may be scheduled from here on (it waits the same way before touching memory): its CTAs
    // take over SMs as this grid's CTAs retire instead of after the grid has drained.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // tile -> (m tile, n tile, k split); n fastest so concurrently running CTAs share the A (activation) tile in L2
    auto decode = [&](int tile, int& mt, int& nt, int& kb0, int& kb1) {
        const int ks = tile % split_k;
        const int t2 = tile / split_k;
        nt = t2 % num_n_tiles;
        mt = t2 / num_n_tiles;
        const int base = num_k_blocks / split_k, rem = num_k_blocks % split_k;
        kb0 = ks * base + min(ks, rem);
        kb1 = kb0 + base + (ks < rem ? 1 : 0);
    };

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                int mt, nt, kb0, kb1;
                decode(tile, mt, nt, kb0, kb1);
                const int a_row = mt * TM + (int)cta_rank * BM, b_row = nt * BN + (int)cta_rank * (BN / CTAS);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint64_t* fb = full_bar + stage;
                    if constexpr (CTAS == 2) {
                        if (cta_rank == 0) mbar_expect_tx(fb, 2 * cfg::kStageBytes);           // both CTAs' bytes land on the leader's barrier
                        else mbar_arrive_cluster(mapa_u32(smem_u32(fb), 0));
                    } else {
                        mbar_expect_tx(fb, cfg::kStageBytes);
                    }
                    auto load = [&](void* dst, const CUtensorMap* map, int c0, int c1) {
                        if constexpr (CTAS == 2) tma_load_2d_2sm(dst, map, fb, c0, c1); else tma_load_2d(dst, map, fb, c0, c1);
                    };
                    uint8_t* sa = smem_a + stage * cfg::kABytes;
                    uint8_t* sb = smem_b + stage * cfg::kBBytes;
                    if constexpr (!A_MN) {
                        load(sa, &tma_a, kb * BK, a_row);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) load(sa + j * 8192, &tma_a, a_row + j * 64, kb * BK);
                    }
                    if constexpr (!B_MN) {
                        load(sb, &tma_b, kb * BK, b_row);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / CTAS / 64; ++j) load(sb + j * 8192, &tma_b, b_row + j * 64, kb * BK);
                    }
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = make_idesc(TM, BN, A_MN, B_MN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                int mt, nt, kb0, kb1;
                decode(tile, mt, nt, kb0, kb1);
                mbar_wait(tmem_empty + acc, acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar + stage, phase);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * cfg::kABytes);
                    const uint32_t b_addr = smem_u32(smem_b + stage * cfg::kBBytes);
#pragma unroll
                    for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                        // K-major: 8-row groups are 1024 B apart (SBO), K step of 16 elems = +32 B inside the swizzled row.
                        // MN-major: 64-elem MN blocks are 8192 B apart (LBO), 8-k-row groups 1024 B apart (SBO), K step = +2048 B.
                        const uint64_t a_desc = A_MN ? make_smem_desc(a_addr + kk * 2048, 8192, 1024) : make_smem_desc(a_addr + kk * 32, 16, 1024);
                        const uint64_t b_desc = B_MN ? make_smem_desc(b_addr + kk * 2048, 8192, 1024) : make_smem_desc(b_addr + kk * 32, 16, 1024);
                        if constexpr (CTAS == 2) tcgen05_mma_bf16_2sm(d_tmem, a_desc, b_desc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                        else tcgen05_mma_bf16(d_tmem, a_desc, b_desc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
                    }
                    // frees the smem slot (in both CTAs of a pair) when these MMAs retire
                    if constexpr (CTAS == 2) tcgen05_commit_2sm(empty_bar + stage); else tcgen05_commit(empty_bar + stage);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                // accumulator ready for the epilogue warps (of both CTAs)
                if constexpr (CTAS == 2) tcgen05_commit_2sm(tmem_full + acc); else tcgen05_commit(tmem_full + acc);
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps 2..17
        const int quad = warp % 4;                        // tcgen05.ld: warp w may touch TMEM lanes [32*(w%4), +32)
        const int slot = (warp - 2) / 4;                  // which 32-column chunks of the tile this warp owns: slot, slot+4, ...
        const int rr = lane / 4, c8 = lane % 4;           // read-back mapping: row rr (+8 per pass); bf16 C: columns [8 c8, +8); fp32 C: [4 c8, +4) and [16 + 4 c8, +4)
        constexpr bool kSplit = sizeof(TC) == 4;
        constexpr int HI = kSplit ? 16 : 4;               // column distance between a lane's two groups of four
        // scratch: 32 rows x 8 granules of 16 B per warp, granule index XOR (row & 7)  (conflict-free both ways)
        const uint32_t sc = smem_u32(scratch) + (uint32_t)(warp - 2) * (uint32_t)(nb * 4096);
        const uint32_t st_row = sc + (uint32_t)lane * 128u, st_x = (uint32_t)(lane & 7);
        const uint32_t g_lo = kSplit ? (uint32_t)c8 : (uint32_t)(2 * c8), g_hi = kSplit ? (uint32_t)(4 + c8) : (uint32_t)(2 * c8 + 1);   // 16-byte granules of the row
        const uint32_t ld_lo = sc + (uint32_t)rr * 128u + ((g_lo ^ (uint32_t)rr) << 4), ld_hi = sc + (uint32_t)rr * 128u + ((g_hi ^ (uint32_t)rr) << 4);
        const float alpha = epi.alpha;
        const float* __restrict__ bias = (MODE == EPI_SPLITK) ? nullptr : epi.bias;
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        constexpr int kCVec = sizeof(TC) == 4 ? 4 : 8;
        // 16-byte vector accesses need every touched row start aligned: checked once
        const bool vec_ok = (ldc % kCVec == 0) && al16(C) && al16(bias) &&
                            (MODE != EPI_FULL ||
                             ((!epi.residual || (epi.ld_res % 4 == 0 && al16(epi.residual))) &&
                              (!(epi.aux_in || epi.aux_out) || (epi.ld_aux % kCVec == 0 && al16(epi.aux_in) && al16(epi.aux_out))) &&
                              (!epi.c2 || (epi.ldc2 % 8 == 0 && al16(epi.c2))) && (epi.dropout_p <= 0.f || n % 8 == 0) && al16(epi.colsum)));
        // FULL-mode features, hoisted (warp-uniform)
        const bool f_gelu = MODE == EPI_FULL && (epi.flags & OVQA_EPI_GELU), f_dgelu = MODE == EPI_FULL && (epi.flags & OVQA_EPI_DGELU);
        const bool f_save_d = MODE == EPI_FULL && (epi.flags & OVQA_EPI_SAVE_DGELU) && (epi.flags & OVQA_EPI_GELU) && epi.aux_out;
        const bool f_mul_aux = MODE == EPI_FULL && (epi.flags & OVQA_EPI_MUL_AUX);
        const bool f_relu = MODE == EPI_FULL && (epi.flags & OVQA_EPI_RELU);
        const bool f_drop = MODE == EPI_FULL && epi.dropout_p > 0.f;
        const float* __restrict__ residual = MODE == EPI_FULL ? epi.residual : nullptr;
        const TC* __restrict__ aux_in = MODE == EPI_FULL ? reinterpret_cast<const TC*>(epi.aux_in) : nullptr;
        TC* __restrict__ aux_out = MODE == EPI_FULL ? reinterpret_cast<TC*>(epi.aux_out) : nullptr;
        __nv_bfloat16* __restrict__ c2 = MODE == EPI_FULL ? reinterpret_cast<__nv_bfloat16*>(epi.c2) : nullptr;
        const float beta = MODE == EPI_FULL ? epi.beta : 0.f;
        float* __restrict__ colsum = (MODE == EPI_FULL && (FEAT < 0 || (FEAT & F_COLSUM))) ? epi.colsum : nullptr;
        const float drop_p = epi.dropout_p, inv_keep = f_drop ? 1.0f / (1.0f - epi.dropout_p) : 1.0f;
        const uint64_t seed = epi.seed, rng_offset = f_drop ? effective_offset(epi.offset, epi.offset_base) : 0;
        const uint32_t thr_hi = dropout_threshold_hi(drop_p);
        const int64_t ld_res = epi.ld_res, ld_aux = epi.ld_aux, ldc2 = epi.ldc2;
        int acc = 0;
        uint32_t acc_phase = 0;
        // ------------------------------------------------------------------ TMA-fed epilogue (see Cfg)
        constexpr int kFe = (MODE == EPI_FULL && FEAT >= 0) ? FEAT : 0;
        // (the GELU epilogues stay on the transposing path below: they are bound by their ~900 instructions per chunk, not by memory
        // latency, and measured 5-10 % slower here — two stores of 64-byte rows per chunk for fc1's result + saved derivative)
        constexpr bool kTmaCapable = EPW == 8 && (MODE != EPI_FULL || (FEAT >= 0 && !(FEAT & F_GELU)));
        bool tma_done = false;
        if constexpr (kTmaCapable) {
          if (use_tma) {                                                 // warp-uniform (kernel argument)
            tma_done = true;
            constexpr bool kGelu = kFe & F_GELU, kSave = kFe & F_SAVE_D, kMul = kFe & F_MULAUX, kDrop = kFe & F_DROP, kRes = kFe & F_RES,
                           kColsum = kFe & F_COLSUM;
            constexpr bool kLoad = kRes || kMul;
            constexpr uint32_t kBoxBytes = 32u * 32u * (uint32_t)sizeof(TC);
            constexpr int kChunkStep = (EPW / 4) * 32;
            const int ew = warp - 2;
            const uint32_t bufs = smem_u32(scratch) + (uint32_t)ew * (uint32_t)(nb * 4096);
            uint64_t* in_bar = epi_bar + ew * nb;
            const float k_out = kDrop ? inv_keep : 1.0f;                  // dropout keep-scale, folded into the constants below
            const float a_s = (kGelu ? 1.0f : k_out) * alpha;
            const float b_s = (kDrop && !kGelu) ? k_out : 1.0f;
            const GeluK gk(k_out);
            // this lane's row of a staged chunk: 128-byte rows under SWIZZLE_128B (fp32), 64-byte rows under SWIZZLE_64B (bf16)
            const uint32_t row_off = kSplit ? (uint32_t)lane * 128u : (uint32_t)lane * 64u;
            const uint32_t sx = kSplit ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
            // the chunk sequence of this warp, walked a second time ahead of the math by the loads of the input operand (lane 0)
            int ld_tile = tile0, ld_c0 = slot * 32;
            uint32_t ld_b = 0, cur_b = 0, cur_par = 0;               // staging buffer of the next load / of the current chunk (+ its barrier parity)
            auto ld_issue = [&]() {
                while (ld_tile < num_tiles) {
                    int mt2, nt2, ka, kb_;
                    decode(ld_tile, mt2, nt2, ka, kb_);
                    const int64_t c = (int64_t)nt2 * BN + ld_c0;
                    if (ld_c0 < BN && c < n) {
                        mbar_expect_tx(in_bar + ld_b, kBoxBytes);
                        tma_load_2d_u32(bufs + ld_b * 4096u, &tma_x, in_bar + ld_b, (int)c, mt2 * TM + (int)cta_rank * BM + quad * 32);
                        ld_c0 += kChunkStep;
                        if (++ld_b == (uint32_t)nb) ld_b = 0;
                        return;
                    }
                    ld_tile += tile_step;
                    ld_c0 = slot * 32;
                }
            };
            if constexpr (kLoad) {
                if (lane == 0) for (int q = 0; q < nb - 1; ++q) ld_issue();      // (host: nb >= 2 for the kernels that load)
            }
            // bias: lane L holds bias[col0 + L] of the chunk about to be processed, requested one chunk ahead (its latency would otherwise sit
            // between the accumulator read and the first FMA of every chunk) and handed to the row-owning lanes by shuffles
            float bias_cur = 0.f, bias_nxt = 0.f;
            int64_t bias_col = -1, bias_nxt_col = -1;
            auto bias_fetch = [&](int64_t c) { return bias ? __ldg(bias + min(c + lane, n - 1)) : 0.f; };
            for (int tile = tile0; tile < num_tiles; tile += tile_step) {
                int mt, nt, kb0, kb1;
                decode(tile, mt, nt, kb0, kb1);
                if (bias && bias_nxt_col != (int64_t)nt * BN + slot * 32) {      // first tile (or a skipped ragged one): nothing was requested ahead
                    bias_nxt_col = (int64_t)nt * BN + slot * 32;
                    bias_nxt = bias_fetch(bias_nxt_col);
                }
                mbar_wait(tmem_full + acc, acc_phase);
                tcgen05_fence_after();
                const int row0 = mt * TM + (int)cta_rank * BM + quad * 32;
                const int64_t row = (int64_t)row0 + lane;
                const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
                for (int c0 = slot * 32; c0 < BN; c0 += kChunkStep) {
                    const int64_t col0 = (int64_t)nt * BN + c0;
                    if (col0 >= n) break;                                     // warp-uniform
                    if (bias) {
                        bias_cur = bias_nxt; bias_col = bias_nxt_col;
                        if (bias_col != col0) { bias_col = col0; bias_cur = bias_fetch(col0); }     // (not reached on the regular chunk sequence)
                        // the chunk after this one: next in the tile, else the first of this warp's next tile
                        if (c0 + kChunkStep < BN && col0 + kChunkStep < n) bias_nxt_col = col0 + kChunkStep;
                        else {
                            int mt2 = 0, nt2 = 0, ka, kb_;
                            if (tile + tile_step < num_tiles) decode(tile + tile_step, mt2, nt2, ka, kb_);
                            bias_nxt_col = (int64_t)nt2 * BN + slot * 32;
                        }
                        bias_nxt = bias_fetch(bias_nxt_col);
                    }
                    const uint32_t b = cur_b, buf = bufs + b * 4096u;
                    const uint32_t my = buf + row_off;
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_row + (uint32_t)c0, r);
                    if constexpr (kLoad) {
                        mbar_wait(in_bar + b, cur_par);                      // the operand rows have landed in `buf`
                    } else {
                        if (lane == 0) {                                     // the store that last read `buf` is done with it
                            if (nb == 3) bulk_wait_read<2>(); else if (nb == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
                        }
                        __syncwarp();
                    }
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if (bias) {                                              // warp-uniform
                        const float bsc = bias_cur * b_s;
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], a_s, __shfl_sync(0xffffffffu, bsc, j));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= a_s;
                    }
                    float g[kSave ? 32 : 1];
                    if constexpr (kGelu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if constexpr (kSave) v[j] = gelu_and_grad_tanhfit(v[j], gk, g[j]);
                            else v[j] = gelu_tanhfit(v[j], gk);
                        }
                    }
                    if constexpr (kMul) {                                     // bf16: granule q = columns [8q, 8q + 8)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t w0, w1, w2, w3;
                            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(my + (((uint32_t)q ^ sx) << 4)) : "memory");
                            const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 a = unpack_bf16x2(w[j]);
                                v[8 * q + 2 * j] *= a.x; v[8 * q + 2 * j + 1] *= a.y;
                            }
                        }
                    }
                    if constexpr (kDrop) {
                        // Philox group = 8 consecutive elements of a row (n % 8 == 0, col0 % 32 == 0): four calls per lane and chunk
                        const uint64_t ctr = rng_offset + (uint64_t)((row * n + col0) >> 3);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 rr = philox4x32(ctr + (uint64_t)q, seed);
                            bool keep[8];
                            dropout_keep8(keep, rr, thr_hi);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                v[8 * q + j] = keep[j] ? v[8 * q + j] : 0.0f;
                                if constexpr (kSave) g[8 * q + j] = keep[j] ? g[8 * q + j] : 0.0f;
                            }
                        }
                    }
                    if constexpr (kSplit) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const uint32_t ga = my + (((uint32_t)q ^ sx) << 4);
                            if constexpr (kRes) {
                                float x0, x1, x2, x3;
                                lds128(ga, x0, x1, x2, x3);
                                v[4 * q] += x0; v[4 * q + 1] += x1; v[4 * q + 2] += x2; v[4 * q + 3] += x3;
                            }
                            sts128(ga, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t ga = my + (((uint32_t)q ^ sx) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ga), "r"(pack_bf16x2(v[8 * q], v[8 * q + 1])), "r"(pack_bf16x2(v[8 * q + 2], v[8 * q + 3])),
                                         "r"(pack_bf16x2(v[8 * q + 4], v[8 * q + 5])), "r"(pack_bf16x2(v[8 * q + 6], v[8 * q + 7])) : "memory");
                            if constexpr (kSave)
                                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ga + 2048u), "r"(pack_bf16x2(g[8 * q], g[8 * q + 1])), "r"(pack_bf16x2(g[8 * q + 2], g[8 * q + 3])),
                                             "r"(pack_bf16x2(g[8 * q + 4], g[8 * q + 5])), "r"(pack_bf16x2(g[8 * q + 6], g[8 * q + 7])) : "memory");
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (MODE == EPI_SPLITK) tma_reduce_add_2d_u32(&tma_c, buf, (int)col0, row0);
                        else tma_store_2d_u32(&tma_c, buf, (int)col0, row0);
                        if constexpr (kSave) tma_store_2d_u32(&tma_x, buf + 2048u, (int)col0, row0);
                        bulk_commit();
                        if constexpr (kLoad) {
                            bulk_wait_read<1>();                              // store(cnt - 1) has read its buffer: the next load may land there
                            ld_issue();
                        }
                    }
                    if constexpr (kColsum) {
                        // column sums over the chunk's 32 rows (rows beyond m excluded): transposing butterfly, lane L ends with column L
                        if (row >= m) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = 0.f;
                        }
#pragma unroll
                        for (int w = 16; w >= 1; w >>= 1) {
                            const bool up = lane & w;
#pragma unroll
                            for (int j = 0; j < w; ++j) {
                                const float give = up ? v[j] : v[j + w], keepv = up ? v[j + w] : v[j];
                                v[j] = keepv + __shfl_xor_sync(0xffffffffu, give, w);
                            }
                        }
                        if (col0 + lane < n) atomicAdd(colsum + col0 + lane, v[0]);
                    }
                    if (++cur_b == (uint32_t)nb) { cur_b = 0; cur_par ^= 1u; }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CTAS == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty + acc), 0));
                    else mbar_arrive(tmem_empty + acc);
                }
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
            if (lane == 0) bulk_wait_all();                               // shared memory stays valid until the last store has read it
          }
        }
        if (!tma_done)
        for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            int mt, nt, kb0, kb1;
            decode(tile, mt, nt, kb0, kb1);
            if constexpr (MODE == EPI_FULL && FEAT >= 0 && (FEAT & (F_RES | F_MULAUX)) != 0) {
                // Pull the NEXT tile's epilogue operand (residual rows / saved derivative) into L2 now: its loads are the
                // latency that bounds these K = 512 epilogues, and an L2 hit is ~3x closer than DRAM.  128 rows x BN columns.
                const int nxt = tile + tile_step;
                if (nxt < num_tiles) {
                    int mt2, nt2, ka, kb_;
                    decode(nxt, mt2, nt2, ka, kb_);
                    constexpr int kElt = (FEAT & F_RES) ? 4 : (int)sizeof(TC);
                    constexpr int kLinesPerRow = BN * kElt / 128;
                    const char* base = (FEAT & F_RES) ? reinterpret_cast<const char*>(residual) : reinterpret_cast<const char*>(aux_in);
                    const int64_t ld_b = ((FEAT & F_RES) ? ld_res : ld_aux) * kElt;
                    for (int idx = (warp - 2) * 32 + lane; idx < BM * kLinesPerRow; idx += EPW * 32) {
                        const int64_t r = (int64_t)mt2 * TM + (int64_t)cta_rank * BM + idx / kLinesPerRow;
                        const int64_t cbyte = ((int64_t)nt2 * BN) * kElt + (idx % kLinesPerRow) * 128;
                        if (r < m && cbyte < n * kElt) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + r * ld_b + cbyte));
                    }
                }
            }
            mbar_wait(tmem_full + acc, acc_phase);
            tcgen05_fence_after();
            const int64_t row_base = (int64_t)mt * TM + (int64_t)cta_rank * BM + quad * 32 + rr;          // + 8 * pass
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
            for (int c0 = slot * 32; c0 < BN; c0 += (EPW / 4) * 32) {
                const int64_t col0 = (int64_t)nt * BN + c0;
                if (col0 >= n) break;                                     // warp-uniform
                const int64_t col = col0 + (kSplit ? c8 * 4 : c8 * 8);
                const bool fast = vec_ok && (col0 + 32 <= n);             // warp-uniform: whole chunk on the vector path
                if constexpr (MODE == EPI_FULL && FEAT >= 0) {
                    if (fast) {                                           // warp-uniform
                        constexpr bool kBias = FEAT & F_BIAS, kGelu = FEAT & F_GELU, kSave = FEAT & F_SAVE_D, kMul = FEAT & F_MULAUX,
                                       kDrop = FEAT & F_DROP, kRes = FEAT & F_RES, kColsum = FEAT & F_COLSUM;
                        static_assert(!(kMul && kRes) && !(kSave && !kGelu), "unsupported specialised epilogue");
                        const float k_out = kDrop ? inv_keep : 1.0f;      // dropout keep-scale, folded into the constants below
                        const float a_s = (kGelu ? 1.0f : k_out) * alpha;
                        float b8[8];
                        if constexpr (kBias) {
                            load8<float, HI>(bias + col, b8);
                            if constexpr (kDrop && !kGelu) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) b8[j] *= k_out;
                            }
                        }
                        // operands of the four passes, requested before the accumulator is touched (rows beyond m: clamped, unused)
                        uint4 in_lo[4], in_hi[4];
                        if constexpr (kMul || kRes) {
#pragma unroll
                            for (int p = 0; p < 4; ++p) {
                                const int64_t row = min(row_base + 8 * p, m - 1);
                                if constexpr (kRes) {
                                    const float* rp = residual + row * ld_res + col;
                                    in_lo[p] = __ldg(reinterpret_cast<const uint4*>(rp)); in_hi[p] = __ldg(reinterpret_cast<const uint4*>(rp + HI));
                                } else {
                                    static_assert(!kMul || sizeof(TC) == 2, "MUL_AUX specialisation is bf16");
                                    in_lo[p] = __ldg(reinterpret_cast<const uint4*>(aux_in + row * ld_aux + col));
                                }
                            }
                        }
                        if constexpr (kRes) {                             // 32 residual registers are in flight: stage the chunk in two halves
#pragma unroll
                            for (int hlf = 0; hlf < 2; ++hlf) {
                                uint32_t r[16];
                                tmem_ld_32x32b_x16(t_row + (uint32_t)(c0 + 16 * hlf), r);
                                tmem_ld_wait();
#pragma unroll
                                for (int g = 0; g < 4; ++g)
                                    sts128(st_row + (((uint32_t)(4 * hlf + g) ^ st_x) << 4), __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                           __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                            }
                        } else {
                            uint32_t r[32];
                            tmem_ld_32x32b_x32(t_row + (uint32_t)c0, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                sts128(st_row + (((uint32_t)g ^ st_x) << 4), __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]),
                                       __uint_as_float(r[4 * g + 2]), __uint_as_float(r[4 * g + 3]));
                        }
                        __syncwarp();
                        const GeluK gk(k_out);
                        float cs[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) cs[j] = 0.f;
                        // Philox group (8 consecutive elements) counters; n % 8 == 0 (vec_ok): + n per pass of 8 rows.  bf16 C: the lane's 8
                        // columns ARE one group.  fp32 C: the lane holds half of two groups (g and g + 2); even lanes draw group g, odd lanes
                        // group g + 2, and the pair swaps the halves it does not need — still one Philox call per lane and pass.
                        const uint64_t ctr0 = rng_offset + (uint64_t)((row_base * n + col0) >> 3) +
                                              (uint64_t)(kSplit ? ((c8 & 1) ? 2 + c8 / 2 : c8 / 2) : c8);
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const int64_t row = row_base + 8 * p;
                            float v[8], g[8];
                            lds128(ld_lo + p * 1024, v[0], v[1], v[2], v[3]);
                            lds128(ld_hi + p * 1024, v[4], v[5], v[6], v[7]);
                            bool keep[8];
                            if constexpr (kDrop) {
                                uint4 r = philox4x32(ctr0 + (uint64_t)p * (uint64_t)n, seed);
                                if constexpr (kSplit) {
                                    const bool odd = c8 & 1;
                                    const uint32_t s0 = odd ? r.x : r.z, s1 = odd ? r.y : r.w;          // the half the partner needs
                                    const uint32_t q0 = __shfl_xor_sync(0xffffffffu, s0, 1), q1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                                    // columns [col, +4): fields of group g (even lane: own x,y; odd lane: partner's z,w);  [col+16, +4): group g + 2
                                    r = odd ? make_uint4(q0, q1, r.z, r.w) : make_uint4(r.x, r.y, q0, q1);
                                }
                                dropout_keep8(keep, r, thr_hi);
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                v[j] = kBias ? fmaf(v[j], a_s, b8[j]) : v[j] * a_s;
                                if constexpr (kGelu) v[j] = kSave ? gelu_and_grad_tanhfit(v[j], gk, g[j]) : gelu_tanhfit(v[j], gk);
                            }
                            if constexpr (kMul) {
                                const uint32_t w[4] = {in_lo[p].x, in_lo[p].y, in_lo[p].z, in_lo[p].w};
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 a = unpack_bf16x2(w[j]);
                                    v[2 * j] *= a.x; v[2 * j + 1] *= a.y;
                                }
                            }
                            if constexpr (kDrop) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    v[j] = keep[j] ? v[j] : 0.0f;
                                    if constexpr (kSave) g[j] = keep[j] ? g[j] : 0.0f;
                                }
                            }
                            if constexpr (kRes) {
                                const uint32_t w[8] = {in_lo[p].x, in_lo[p].y, in_lo[p].z, in_lo[p].w, in_hi[p].x, in_hi[p].y, in_hi[p].z, in_hi[p].w};
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] += __uint_as_float(w[j]);
                            }
                            if (row < m) {
                                store8<TC, HI>(C + row * ldc + col, v);
                                if constexpr (kSave) store8<TC, HI>(aux_out + row * ld_aux + col, g);
                                if constexpr (kColsum) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) cs[j] += v[j];
                                }
                            }
                        }
                        if constexpr (kColsum) colsum_flush<HI>(colsum, col, cs, lane);
                        __syncwarp();
                        continue;
                    }
                }
                float b8[8], cs[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b8[j] = cs[j] = 0.f;
                if (fast && bias) load8<float, HI>(bias + col, b8);
                {   // TMEM -> registers (row per lane) -> swizzled scratch
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_row + (uint32_t)c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        sts128(st_row + (((uint32_t)g ^ st_x) << 4), __uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                               __uint_as_float(r[4 * g + 3]));
                }
                __syncwarp();
                // four passes of 8 rows: each lane finishes 8 consecutive columns of one row per pass, every feature applied to
                // those 8 values before the next pass (small live state: 16 epilogue warps share the register file)
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int64_t row = row_base + 8 * p;
                    const bool row_ok = row < m;
                    float pre[8];                                         // residual row, else the saved GELU' / pre-activation
                    if constexpr (MODE == EPI_FULL) {
                        if (fast && row_ok) {
                            if (residual) load8<float, HI>(residual + row * ld_res + col, pre);
                            else if (f_dgelu || f_mul_aux) load8<TC, HI>(aux_in + row * ld_aux + col, pre);
                        }
                    }
                    float v[8];
                    lds128(ld_lo + p * 1024, v[0], v[1], v[2], v[3]);
                    lds128(ld_hi + p * 1024, v[4], v[5], v[6], v[7]);
                    if (!fast) {                                          // ragged tail / unaligned pitches
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                        {
                            const int64_t cj = col + (j < 4 ? j : HI + j - 4);
                            if (row_ok && cj < n) epilogue_scalar<TC>(epi, v[j], row, cj, n, C, ldc, split_k);
                        }
                        continue;
                    }
                    if (!row_ok) continue;
                    if constexpr (MODE == EPI_SPLITK) {
                        if constexpr (sizeof(TC) == 4) {
                            float* cp = reinterpret_cast<float*>(C) + row * ldc + col;
                            atomicAdd(reinterpret_cast<float4*>(cp), make_float4(v[0] * alpha, v[1] * alpha, v[2] * alpha, v[3] * alpha));
                            atomicAdd(reinterpret_cast<float4*>(cp + HI), make_float4(v[4] * alpha, v[5] * alpha, v[6] * alpha, v[7] * alpha));
                        }
                        continue;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], alpha, b8[j]);
                    if constexpr (MODE == EPI_FULL) {
                        float ds[8];
                        if (f_drop) {
                            if constexpr (kSplit) {               // two groups of four columns: the matching halves of two Philox calls
                                const float4 d0 = dropout_scale4((uint64_t)(row * n + col), drop_p, inv_keep, seed, rng_offset);
                                const float4 d1 = dropout_scale4((uint64_t)(row * n + col + HI), drop_p, inv_keep, seed, rng_offset);
                                ds[0] = d0.x; ds[1] = d0.y; ds[2] = d0.z; ds[3] = d0.w; ds[4] = d1.x; ds[5] = d1.y; ds[6] = d1.z; ds[7] = d1.w;
                            } else {
                                dropout_scale8(ds, (uint64_t)(row * n + col), drop_p, inv_keep, seed, rng_offset);
                            }
                        }
                        if (f_save_d) {                                   // y = GELU(v), aux_out = GELU'(v) [x dropout scale]: one CDF + one exponential for both
                            float d[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                v[j] = gelu_and_grad_fast(v[j], d[j]);
                                if (f_drop) d[j] *= ds[j];
                            }
                            store8<TC, HI>(aux_out + row * ld_aux + col, d);
                        } else {
                            if (aux_out) store8<TC, HI>(aux_out + row * ld_aux + col, v);
                            if (f_gelu) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = gelu_fast(v[j]);
                            }
                            if (f_relu) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                            }
                        }
                        if (f_dgelu || f_mul_aux) {
                            if (residual) load8<TC, HI>(aux_in + row * ld_aux + col, b8 /*scratch: bias already applied*/);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float a = residual ? b8[j] : pre[j];
                                v[j] *= f_mul_aux ? a : dgelu_fast(a);
                            }
                            if (residual && bias) load8<float, HI>(bias + col, b8);   // (never together on the path) restore for the next pass
                        }
                        if (f_drop) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] *= ds[j];
                        }
                        if (residual) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] += pre[j];
                        }
                        if constexpr (sizeof(TC) == 4) {
                            if (beta != 0.f) {
                                float o[8];
                                load8<float, HI>(reinterpret_cast<const float*>(C) + row * ldc + col, o);
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaf(beta, o[j], v[j]);
                            }
                        }
                    }
                    store8<TC, HI>(C + row * ldc + col, v);
                    if constexpr (MODE == EPI_FULL) {
                        if (c2) store8<__nv_bfloat16, HI>(c2 + row * ldc2 + col, v);
                        if (colsum) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) cs[j] += v[j];
                        }
                    }
                }
                if constexpr (MODE == EPI_FULL) {
                    if (colsum && fast) colsum_flush<HI>(colsum, col, cs, lane);      // (the scalar path adds element by element)
                }
                __syncwarp();                                             // scratch is rewritten by the next chunk
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {                                              // the MMA issuer (leader CTA) waits for every epilogue warp of the pair
                if (CTAS == 2 && cta_rank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty + acc), 0));
                else mbar_arrive(tmem_empty + acc);
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
    }
    tcgen05_fence_before();
    if constexpr (CTAS == 2) cluster_sync_all(); else __syncthreads();       // (pair: nobody leaves while the peer may still touch its smem / TMEM)
    if (warp == 1) {
        tcgen05_fence_after();
        if constexpr (CTAS == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::kTmemCols) : "memory");
    }
}

// ----------------------------------------------------------------------------- host: TMA descriptors
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}

struct MapKey {
    const void* ptr; uint64_t inner, outer, ld; uint32_t box_inner, box_outer; int swz, elt;
    bool operator==(const MapKey& o) const {
        return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner && box_outer == o.box_outer && swz == o.swz &&
               elt == o.elt;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = std::hash<const void*>()(k.ptr);
        auto mix = [&](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
        mix(k.inner); mix(k.outer); mix(k.ld); mix(((uint64_t)k.box_inner << 32) | k.box_outer); mix(((uint64_t)k.elt << 32) | (uint64_t)k.swz);
        return h;
    }
};

}  // namespace

// 2-D tensor [outer, inner] of bf16 (elt_bytes 2) or fp32 (4) with row stride ld (elements); box [box_outer, box_inner]; 128B / 64B swizzle; OOB -> 0.
int tc5::make_tma_map(CUtensorMap* out, const void* ptr, int elt_bytes, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                      int swizzle_bytes) {
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{ptr, inner, outer, ld, box_inner, box_outer, swizzle_bytes, elt_bytes};
    {
        std::lock_guard<std::mutex> g(mu);
        auto it = cache.find(key);
        if (it != cache.end()) { *out = it->second; return OVQA_OK; }
    }
    EncodeFn enc = get_encode();
    if (!enc) return OVQA_ERR_NO_DEVICE;
    const cuuint64_t dims[2] = {inner, outer};
    const cuuint64_t strides[1] = {ld * (uint64_t)elt_bytes};
    const cuuint32_t box[2] = {box_inner, box_outer};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(out, elt_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_last_cuda_error = (int)r; return OVQA_ERR_CUDA; }
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return OVQA_OK;
}
int tc5::make_tma_map_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                           int swizzle_bytes) {
    return make_tma_map(out, ptr, 2, inner, outer, ld, box_inner, box_outer, swizzle_bytes);
}

namespace {

inline int make_map(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer) {
    return tc5::make_tma_map_bf16(out, ptr, inner, outer, ld, box_inner, box_outer, 128);
}

template <int BN, bool A_MN, bool B_MN, int CTAS>
int launch(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
           int64_t ldc, const Epilogue& epi, int split_k, cudaStream_t st) {
    CUtensorMap ma, mb;
    int rc;
    // K-major: tensor [rows, k] box {64 k, ROWS};  MN-major: tensor [k, rows] box {64 rows, 64 k}.  A CTA of a pair loads half of B's rows.
    rc = A_MN ? make_map(&ma, A, (uint64_t)m, (uint64_t)k, (uint64_t)lda, 64, BK) : make_map(&ma, A, (uint64_t)k, (uint64_t)m, (uint64_t)lda, BK, BM);
    if (rc) return rc;
    rc = B_MN ? make_map(&mb, B, (uint64_t)n, (uint64_t)k, (uint64_t)ldb, 64, BK) : make_map(&mb, B, (uint64_t)k, (uint64_t)n, (uint64_t)ldb, BK, BN / CTAS);
    if (rc) return rc;
    const int num_k_blocks = (int)((k + BK - 1) / BK);
    if (split_k > num_k_blocks) split_k = num_k_blocks;
    // epilogue through the TMA unit: 32 x 32 boxes of C and of the specialised epilogue's one extra tensor (residual in, saved derivative in / out)
    CUtensorMap mc{}, mx{};
    int use_tma = 0;
    {
        static const bool enabled = [] { const char* e = getenv("OVQA_GEMM_TMA_EPILOGUE"); return !(e && e[0] == '0'); }();
        auto tma_ok = [](const void* p, int64_t ld, int eb) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * eb) % 16 == 0; };
        const int eb = epi.c_dtype == OVQA_F32 ? 4 : 2;
        const void* x = nullptr;
        int64_t ldx = 0;
        int ebx = eb;
        if (epi.residual) { x = epi.residual; ldx = epi.ld_res; ebx = 4; }
        else if (epi.aux_in) { x = epi.aux_in; ldx = epi.ld_aux; }
        else if (epi.aux_out) { x = epi.aux_out; ldx = epi.ld_aux; }
        if (enabled && m < (int64_t)1 << 30 && n < (int64_t)1 << 30 && tma_ok(C, ldc, eb) && (!x || tma_ok(x, ldx, ebx)) && (epi.dropout_p <= 0.f || n % 8 == 0) &&
            (!epi.colsum || (reinterpret_cast<uintptr_t>(epi.colsum) & 3) == 0)) {
            rc = tc5::make_tma_map(&mc, C, eb, (uint64_t)n, (uint64_t)m, (uint64_t)ldc, 32, 32, eb == 4 ? 128 : 64);
            if (!rc && x) rc = tc5::make_tma_map(&mx, x, ebx, (uint64_t)n, (uint64_t)m, (uint64_t)ldx, 32, 32, ebx == 4 ? 128 : 64);
            use_tma = rc == OVQA_OK;
        }
    }
    // k-blocks of one work unit: a short main loop (K <= 1024) leaves the epilogue exposed, a long one hides it and wants the deeper operand ring
    const bool short_k = (num_k_blocks + split_k - 1) / split_k <= 16;
    const int64_t tiles = ((m + BM * CTAS - 1) / (BM * CTAS)) * ((n + BN - 1) / BN) * split_k;
    const int64_t units = sm_count() / CTAS;
    const int grid = (int)(tiles < units ? tiles : units) * CTAS;
    const bool full = epi.flags || epi.aux_in || epi.aux_out || epi.residual || epi.c2 || epi.dropout_p > 0.f || epi.beta != 0.f || epi.colsum;
    const int mode = split_k > 1 ? EPI_SPLITK : full ? EPI_FULL : EPI_PLAIN;
    // compile-time feature set of this launch, if it is one of the specialised combinations (else F_RUNTIME)
    int feat = F_RUNTIME;
    if (mode == EPI_FULL && !epi.c2 && epi.beta == 0.f && !(epi.flags & (OVQA_EPI_DGELU | OVQA_EPI_RELU)) && !(epi.aux_out && !(epi.flags & OVQA_EPI_SAVE_DGELU))) {
        feat = (epi.bias ? F_BIAS : 0) | ((epi.flags & OVQA_EPI_GELU) ? F_GELU : 0) | ((epi.flags & OVQA_EPI_SAVE_DGELU) ? F_SAVE_D : 0) |
               ((epi.flags & OVQA_EPI_MUL_AUX) ? F_MULAUX : 0) | (epi.dropout_p > 0.f ? F_DROP : 0) | (epi.residual ? F_RES : 0) |
               (epi.colsum ? F_COLSUM : 0);
    }
#define OVQA_GO(TC_, MODE_, FEAT_)                                                                                              \
    do {                                                                                                                       \
        constexpr int kEpw = epw_for(MODE_, FEAT_), kThreads = 64 + kEpw * 32;                                                  \
        using cfg = Cfg<BN, CTAS, kEpw>;                                                                                        \
        auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, TC_, MODE_, FEAT_, CTAS, kEpw>;                                               \
        static OncePerDevice attr_set;                                                                                          \
        if (attr_set.pending()) { OVQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::kSmemLimit)); attr_set.mark(); } \
        /* TMA-fed epilogue: which instantiations have it (mirrors kTmaCapable in the kernel), how many staging buffers per warp */ \
        constexpr bool kCap = kEpw == 8 && (MODE_ != EPI_FULL || ((FEAT_) >= 0 && !((FEAT_) & F_GELU)));                          \
        constexpr bool kLoads = MODE_ == EPI_FULL && (FEAT_) >= 0 && ((FEAT_) & (F_RES | F_MULAUX));                              \
        const int tma = kCap ? use_tma : 0;                                                                                     \
        int nb = (!tma || MODE_ == EPI_SPLITK) ? 1 : short_k ? (kLoads ? 3 : 2) : (kLoads ? 2 : 1);                                                      \
        while (nb > 1 && cfg::stages_for(nb) < 3) --nb;                                                                         \
        if (kLoads && nb < 2) { nb = 1; }                                                                                       \
        const int tma_f = (kLoads && nb < 2) ? 0 : tma;                                                                         \
        const int stages = cfg::stages_for(nb), smem_bytes = cfg::smem_bytes(stages, nb);                                       \
        cudaLaunchConfig_t lc{};                                                                                               \
        lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(kThreads); lc.dynamicSmemBytes = smem_bytes; lc.stream = st;      \
        cudaLaunchAttribute at[2];                                                                                             \
        int na = 0;                                                                                                            \
        if (pdl_enabled()) { at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[na].val.programmaticStreamSerializationAllowed = 1; ++na; } \
        if constexpr (CTAS == 2) {                                                                                             \
            at[na].id = cudaLaunchAttributeClusterDimension;                                                                   \
            at[na].val.clusterDim.x = CTAS; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;                          \
            ++na;                                                                                                              \
        }                                                                                                                      \
        lc.attrs = at; lc.numAttrs = na;                                                                                       \
        if constexpr (CTAS == 2) {                                                                                             \
            /* persistent kernel: never more clusters than can be co-resident (GPCs with an odd SM count cannot pair them all) */ \
            static int max_clusters = -1;                                                                                      \
            if (max_clusters < 0) {                                                                                            \
                int nc = 0;                                                                                                    \
                lc.gridDim = dim3((unsigned)(sm_count() / CTAS * CTAS));                                                       \
                if (cudaOccupancyMaxActiveClusters(&nc, kern, &lc) != cudaSuccess || nc < 1) { (void)cudaGetLastError(); nc = sm_count() / CTAS; } \
                max_clusters = nc;                                                                                             \
                if (getenv("OVQA_DEBUG")) fprintf(stderr, "ovqa gemm pair kernel: %d co-resident clusters of %d\n", nc, CTAS);  \
            }                                                                                                                  \
            const int clusters = grid / CTAS < max_clusters ? grid / CTAS : max_clusters;                                      \
            lc.gridDim = dim3((unsigned)(clusters * CTAS));                                                                    \
        }                                                                                                                      \
        OVQA_CUDA_OK(cudaLaunchKernelEx(&lc, kern, ma, mb, mc, mx, (TC_*)C, m, n, k, ldc, epi, split_k, tma_f, nb, stages));   \
    } while (0)
    // the specialised sets exist for the operand forms that launch them: forward y = x W^T is (N,N), dgrad dx = dy W is (N,T)
    constexpr bool kFwd = !A_MN && !B_MN, kDgrad = !A_MN && B_MN;
    if (epi.c_dtype == OVQA_F32) {
        if (mode == EPI_SPLITK) OVQA_GO(float, EPI_SPLITK, F_RUNTIME);
        else if (mode == EPI_PLAIN) OVQA_GO(float, EPI_PLAIN, F_RUNTIME);
        else if (kFwd && feat == (F_BIAS | F_DROP | F_RES)) OVQA_GO(float, EPI_FULL, (kFwd ? F_BIAS | F_DROP | F_RES : F_RUNTIME));     // fc_o / fc2, training
        else if (kFwd && feat == (F_BIAS | F_RES)) OVQA_GO(float, EPI_FULL, (kFwd ? F_BIAS | F_RES : F_RUNTIME));                       // fc_o / fc2, eval
        else if (kDgrad && feat == F_RES) OVQA_GO(float, EPI_FULL, (kDgrad ? F_RES : F_RUNTIME));                                       // dx = dy W + other branch's gradient
        else if (kFwd && feat == (F_BIAS | F_GELU | F_SAVE_D | F_DROP)) OVQA_GO(float, EPI_FULL, (kFwd ? F_BIAS | F_GELU | F_SAVE_D | F_DROP : F_RUNTIME));   // FeatureEmbedding, training
        else if (kFwd && feat == (F_BIAS | F_GELU)) OVQA_GO(float, EPI_FULL, (kFwd ? F_BIAS | F_GELU : F_RUNTIME));                                           // FeatureEmbedding, eval
        else OVQA_GO(float, EPI_FULL, F_RUNTIME);
    } else {
        if (mode == EPI_PLAIN) OVQA_GO(__nv_bfloat16, EPI_PLAIN, F_RUNTIME);
        else if (kFwd && feat == (F_BIAS | F_GELU | F_SAVE_D | F_DROP)) OVQA_GO(__nv_bfloat16, EPI_FULL, (kFwd ? F_BIAS | F_GELU | F_SAVE_D | F_DROP : F_RUNTIME));   // fc1, training
        else if (kFwd && feat == (F_BIAS | F_GELU | F_SAVE_D)) OVQA_GO(__nv_bfloat16, EPI_FULL, (kFwd ? F_BIAS | F_GELU | F_SAVE_D : F_RUNTIME));                     // fc1, training, p = 0
        else if (kFwd && feat == (F_BIAS | F_GELU)) OVQA_GO(__nv_bfloat16, EPI_FULL, (kFwd ? F_BIAS | F_GELU : F_RUNTIME));                                           // fc1, eval
        else if (kDgrad && feat == F_MULAUX) OVQA_GO(__nv_bfloat16, EPI_FULL, (kDgrad ? F_MULAUX : F_RUNTIME));                                                       // dpre = (dy W2) * saved derivative
        else if (kDgrad && feat == (F_MULAUX | F_COLSUM)) OVQA_GO(__nv_bfloat16, EPI_FULL, (kDgrad ? F_MULAUX | F_COLSUM : F_RUNTIME));                               // ... + fc1 bias gradient
        else if (kDgrad && feat == F_COLSUM) OVQA_GO(__nv_bfloat16, EPI_FULL, (kDgrad ? F_COLSUM : F_RUNTIME));                                                       // dctx = dy Wo + value-bias gradient
        else OVQA_GO(__nv_bfloat16, EPI_FULL, F_RUNTIME);
    }
#undef OVQA_GO
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

}  // namespace

int gemm_tcgen05_bf16(const __nv_bfloat16* A, const __nv_bfloat16* B, void* C, int64_t m, int64_t n, int64_t k, int64_t lda,
                      int64_t ldb, int64_t ldc, int a_op, int b_op, const Epilogue& epi, int split_k, cudaStream_t st) {
    if (m == 0 || n == 0) return OVQA_OK;
    if (k <= 0) return OVQA_ERR_BAD_SHAPE;
    // TMA: base 16-byte aligned, row pitch multiple of 16 bytes
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda % 8) || (ldb % 8)) return OVQA_ERR_ALIGNMENT;
    const bool a_mn = a_op == OVQA_OP_T, b_mn = b_op == OVQA_OP_T;
    // CTA pairs (cta_group::2) for the 256-wide tile whenever a 256-row tile makes sense; OVQA_GEMM_2CTA=0 keeps single CTAs
    const char* env2 = getenv("OVQA_GEMM_2CTA");
    const bool allow_pair = !(env2 && env2[0] == '0') && m >= 256 && sm_count() % 2 == 0;
    bool wide;
    if (split_k <= 0) {
        // auto (long reductions, i.e. wgrad): 256-wide tiles halve the operand bytes pulled from L2 per flop (the
        // 128x128 tile needs 128 B/clk/SM of L2->smem at full MMA rate), then split K until ~2 waves of work units
        // exist, keeping at least 8 k-blocks per unit.  C must have been zero-filled by the caller.
        if (epi.c_dtype != OVQA_F32) return OVQA_ERR_BAD_ARG;
        wide = n >= 256;
        const bool pair = wide && allow_pair;
        const int64_t tm = pair ? 2 * BM : BM, units = pair ? sm_count() / 2 : sm_count();
        const int64_t tiles = ((m + tm - 1) / tm) * ((n + (wide ? 255 : 127)) / (wide ? 256 : 128));
        const int64_t kblocks = (k + BK - 1) / BK;
        static const int waves = [] { const char* e = getenv("OVQA_WGRAD_WAVES"); const int w = e ? atoi(e) : 2; return w < 1 ? 1 : w; }();
        int64_t want = (waves * units + tiles / 2) / tiles;
        if (want > kblocks / 8) want = kblocks / 8;
        split_k = (int)(want < 1 ? 1 : want);
    } else {
        if (split_k > 1 && epi.c_dtype != OVQA_F32) return OVQA_ERR_BAD_ARG;
        // wide-N tiles when there are enough of them to fill the machine; BN=128 otherwise
        const int64_t tiles256 = ((m + BM - 1) / BM) * ((n + 255) / 256) * split_k;
        wide = (n % 256 == 0 || n > 1024) && tiles256 >= 2 * sm_count();
    }
#define OVQA_DISPATCH(BN_, CTAS_)                                                                                          \
    do {                                                                                                                   \
        if (!a_mn && !b_mn) return launch<BN_, false, false, CTAS_>(A, B, C, m, n, k, lda, ldb, ldc, epi, split_k, st);    \
        if (!a_mn && b_mn) return launch<BN_, false, true, CTAS_>(A, B, C, m, n, k, lda, ldb, ldc, epi, split_k, st);      \
        if (a_mn && b_mn) return launch<BN_, true, true, CTAS_>(A, B, C, m, n, k, lda, ldb, ldc, epi, split_k, st);        \
        return launch<BN_, true, false, CTAS_>(A, B, C, m, n, k, lda, ldb, ldc, epi, split_k, st);                         \
    } while (0)
    if (wide && allow_pair) OVQA_DISPATCH(256, 2);
    if (wide) OVQA_DISPATCH(256, 1);
    OVQA_DISPATCH(128, 1);
#undef OVQA_DISPATCH
}

}  // namespace ovqa
