// Attention core on 5th-gen tensor cores for the MCAN shape class: head dim 64, Lq <= 128, Lk <= 128, padding (per-key) mask.
//
// One head of one sample is ONE 128-row UMMA tile: S = Q K^T is 4 tcgen05.mma instructions (128 x Lk16 x 16 each) into TMEM,
// the softmax runs with one THREAD per query row (tcgen05.ld gives lane = row, so max / sum need no shuffles), P goes back to
// shared memory as the K-major A operand of O = P V (Lk16/16 more instructions), and O is read from TMEM, normalised and stored.
// Against the mma.sync kernel (attention_mma.cu) this removes the ldmatrix / fragment traffic and ~3/4 of the issued
// instructions (profiles/r01e: the warp-tile kernel is issue-bound at ~27 % of the HBM roofline), and it makes the core a
// persistent, TMA-fed pipeline:
//   warp 0      TMA producer: Q, K, V tiles (SWIZZLE_128B), one stage per softmax group
//   warp 1      MMA issuer (one thread): S(i+1) is issued before waiting for P(i), so the tensor pipe never waits on softmax
//   warps 2-5   softmax group 0 (items 0, 2, 4, ...)      } two groups ping-pong on their own TMEM columns and P buffer: one
//   warps 6-9   softmax group 1 (items 1, 3, 5, ...)      } group's TMEM / MUFU latencies hide behind the other's arithmetic
// Two such CTAs are resident per SM (96 KB of shared memory, 256 TMEM columns, <= 100 registers each): four items in flight per
// SM hide each other's load, MMA-commit and TMEM round-trip latencies (profiles/r01g: one CTA per SM was latency-bound at 37 % of HBM).
// P overwrites the stage's Q | K tiles (dead once S is complete), O overwrites S in TMEM (dead once P is written).
// Work item = (sample, head); CTAs are persistent over items (grid = min(B*h, 2 #SM)).
// Rows of the K / V tiles beyond Lk belong to the next sample (or are TMA zero fill): their P columns are exactly 0 (the
// staged mask row is -inf there), so they contribute nothing.
#include "attention_simt.cuh"
#include "tcgen05.cuh"

namespace ovqa {

namespace {

using namespace tc5;

constexpr int kD = 64;                    // head dim (one SWIZZLE_128B row of bf16)
constexpr int kLoadStages = 2;            // one per softmax group: item i lives in stage i % 2 and is processed by group i % 2
constexpr int kTileBytes = 128 * 128;     // 128 rows x 128 B
constexpr int kStageBytes = 3 * kTileBytes;            // Q, K, V
constexpr int kPBytes = 2 * kTileBytes;                // P: two 64-key blocks of 128 rows x 128 B (forward: aliased onto the stage's dead Q | K tiles)
constexpr int kFwdThreads = 64 + 2 * 128;
constexpr int kFwdSmem = kLoadStages * kStageBytes + 2 * 128 * 4 /*mask rows*/ + 2 * 128 * 4 /*raw mask landing rows*/ + 256 /*barriers*/ + 1024 /*align*/;   // 2 CTAs per SM
constexpr int kFwdTmemCols = 256;         // group g: S in [128 g, +128), O aliased onto the first 64 of them (S is dead once P is written)
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

struct Tc5FwdArgs {
    SdpaArgs a;
    int nk16;               // Lk rounded up to 16 (N of the S MMA, K extent of the PV MMA, rows of the K / V TMA boxes)
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void lds128f(uint32_t addr, float& a, float& b, float& c, float& d) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(addr));
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(kFwdThreads, 2)
sdpa_fwd_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, Tc5FwdArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_load = smem;                                          // [kLoadStages][Q | K | V];  P of the stage's item overwrites Q | K
    float* s_mask = reinterpret_cast<float*>(smem + kLoadStages * kStageBytes);   // [2 groups][128]  (mask * log2 e; -inf beyond Lk)
    float* s_mraw = s_mask + 2 * 128;                                             // [2 groups][128]  next item's raw mask row (cp.async landing zone)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_mraw + 2 * 128);
    uint64_t* load_full = bars;                     // [kLoadStages]  TMA -> MMA
    uint64_t* load_empty = load_full + kLoadStages; // [kLoadStages]  MMA -> TMA
    uint64_t* s_full = load_empty + kLoadStages;    // [2]            MMA -> softmax group (S ready)
    uint64_t* p_full = s_full + 2;                  // [2]            softmax group -> MMA (P in smem), 128 arrivals
    uint64_t* o_full = p_full + 2;                  // [2]            MMA -> softmax group (O ready)
    uint64_t* acc_empty = o_full + 2;               // [2]            softmax group -> MMA (S/O columns and P buffer free), 128 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const SdpaArgs& a = args.a;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int Lq = (int)a.Lq, Lk = (int)a.Lk, H = (int)a.h, nk16 = args.nk16;
    const int n_items = (int)a.B * H;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_v)) : "memory");
        for (int i = 0; i < kLoadStages; ++i) { mbar_init(load_full + i, 1); mbar_init(load_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(s_full + i, 1); mbar_init(p_full + i, 128); mbar_init(o_full + i, 1); mbar_init(acc_empty + i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)kFwdTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();          // programmatic dependent launch: the set-up above overlaps the previous kernel's drain; no global access before this
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(Lq + 2 * nk16) * 128u;
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int ls = it % kLoadStages;
                const uint32_t par = (uint32_t)(it / kLoadStages) & 1u;
                const int b = item / H, hd = item % H;
                mbar_wait(load_empty + ls, par ^ 1u);
                mbar_expect_tx(load_full + ls, bytes);
                uint8_t* st = s_load + ls * kStageBytes;
                tma_load_2d(st, &map_q, load_full + ls, hd * kD, b * Lq);
                tma_load_2d(st + kTileBytes, &map_k, load_full + ls, hd * kD, b * Lk);
                tma_load_2d(st + 2 * kTileBytes, &map_v, load_full + ls, hd * kD, b * Lk);
                // the stage of the item after this one frees only when ITS predecessor's P V has retired: its tiles travel DRAM -> L2 meanwhile
                const int nx = item + 2 * (int)gridDim.x;
                if (nx < n_items) {
                    const int b2 = nx / H, h2 = nx % H;
                    tma_prefetch_2d(&map_q, h2 * kD, b2 * Lq);
                    tma_prefetch_2d(&map_k, h2 * kD, b2 * Lk);
                    tma_prefetch_2d(&map_v, h2 * kD, b2 * Lk);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc(128, nk16, false, false);
            const uint32_t idesc_o = make_idesc(128, kD, false, true);
            const int my_items = n_items > (int)blockIdx.x ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            auto issue_s = [&](int it) {
                const int ls = it % kLoadStages, g = it & 1;
                mbar_wait(load_full + ls, (uint32_t)(it / kLoadStages) & 1u);
                mbar_wait(acc_empty + g, ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t q_addr = smem_u32(s_load + ls * kStageBytes), k_addr = q_addr + kTileBytes;
                const uint32_t d_s = tmem_base + (uint32_t)(g * 128);
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    tcgen05_mma_bf16(d_s, make_smem_desc(q_addr + kk * 32, 16, 1024), make_smem_desc(k_addr + kk * 32, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
                tcgen05_commit(s_full + g);
            };
            auto issue_pv = [&](int it) {
                const int ls = it % kLoadStages, g = it & 1;
                mbar_wait(p_full + g, (uint32_t)(it >> 1) & 1u);
                tcgen05_fence_after();
                const uint32_t p_addr = smem_u32(s_load + ls * kStageBytes), v_addr = p_addr + 2 * kTileBytes;
                const uint32_t d_o = tmem_base + (uint32_t)(g * 128);
                for (int ks = 0; ks < nk16 / 16; ++ks)
                    tcgen05_mma_bf16(d_o, make_smem_desc(p_addr + (ks >> 2) * kTileBytes + (ks & 3) * 32, 16, 1024),
                                     make_smem_desc(v_addr + ks * 2048, 8192, 1024), idesc_o, ks > 0 ? 1u : 0u);
                tcgen05_commit(o_full + g);
                tcgen05_commit(load_empty + ls);     // Q, K, V of this stage are no longer read
            };
            if (my_items > 0) issue_s(0);
            for (int it = 0; it < my_items; ++it) {
                if (it + 1 < my_items) issue_s(it + 1);
                issue_pv(it);
            }
        }
    } else {
        const int g = (warp - 2) / 4;                 // softmax group
        const int quad = warp % 4;                    // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;             // query row of this thread
        const int tid_g = (warp - 2) % 4 * 32 + lane; // 0..127 inside the group
        float* mrow = s_mask + g * 128;
        const uint32_t mrow_u32 = smem_u32(mrow);
        const uint32_t p_row = smem_u32(s_load + g * kStageBytes) + (uint32_t)row * 128u, p_x = (uint32_t)(row & 7);   // stage g == group g
        const uint32_t t_s = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 128);
        const uint32_t t_o = t_s;
        const float sc2 = a.scale * kLog2e;
        const int nch = nk16 / 16;
        int u = 0;
        // mask value of key `tid_g` for an item (raw; -inf beyond Lk): fetched one item ahead so its DRAM latency is off the critical path —
        // by cp.async into the thread's own slot of a landing row, not into a register (the register was spilled across the item and the
        // spill store waited for the load: 5 % of the stall samples, profiles/r02_kernels_ncu.md)
        float* mraw = s_mraw + g * 128 + tid_g;
        auto mask_fetch = [&](int item) {
            if (tid_g < Lk && item < n_items && a.mask.ptr)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(mraw)),
                             "l"(a.mask.ptr + (int64_t)(item / H) * a.mask.sb + (int64_t)(item % H) * a.mask.sh + tid_g) : "memory");
            else *mraw = (tid_g >= Lk || item >= n_items) ? -INFINITY : 0.0f;
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        mask_fetch(blockIdx.x + g * gridDim.x);
        for (int item = blockIdx.x + g * gridDim.x; item < n_items; item += 2 * gridDim.x, ++u) {
            const int b = item / H, hd = item % H;
            const uint32_t par = (uint32_t)u & 1u;
            // per-key mask row of this (sample, head), pre-multiplied by log2 e; -inf beyond Lk
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            const float mv_cur = *mraw;
            named_bar_sync(1 + g, 128);               // the previous item's readers are done
            mrow[tid_g] = mv_cur * kLog2e;
            mask_fetch(item + 2 * gridDim.x);
            named_bar_sync(1 + g, 128);
            mbar_wait(s_full + g, par);
            tcgen05_fence_after();
            // Two passes over the S row in TMEM (row max, then exponentials), 16 columns at a time, the next chunk's tcgen05.ld in
            // flight while the current one is processed: ~60 live registers (two resident CTAs leave 100 per thread).
            auto scaled = [&](const uint32_t (&r)[16], int ch, float (&v)[16]) {   // v = S * scale * log2e + mask
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    float m0, m1, m2, m3;
                    lds128f(mrow_u32 + (uint32_t)(16 * ch + 4 * q4) * 4u, m0, m1, m2, m3);
                    v[4 * q4] = fmaf(__uint_as_float(r[4 * q4]), sc2, m0);
                    v[4 * q4 + 1] = fmaf(__uint_as_float(r[4 * q4 + 1]), sc2, m1);
                    v[4 * q4 + 2] = fmaf(__uint_as_float(r[4 * q4 + 2]), sc2, m2);
                    v[4 * q4 + 3] = fmaf(__uint_as_float(r[4 * q4 + 3]), sc2, m3);
                }
            };
            uint32_t ra[16], rb[16];
            float mx = -INFINITY;
            tmem_ld_32x32b_x16(t_s, ra);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c < nch) {
                    tmem_ld_wait();
                    if (c + 1 < nch) tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * (c + 1)), (c & 1) ? ra : rb);
                    float v[16];
                    scaled((c & 1) ? rb : ra, c, v);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) mx = fmaxf(fmaxf(mx, fmaxf(v[j], v[j + 1])), fmaxf(v[j + 2], v[j + 3]));
                }
            }
            float l = 0.f;
            const bool live = row < Lq;
            tmem_ld_32x32b_x16(t_s, ra);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c < nch) {
                    tmem_ld_wait();
                    if (c + 1 < nch) tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * (c + 1)), (c & 1) ? ra : rb);
                    float v[16];
                    scaled((c & 1) ? rb : ra, c, v);
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float p0 = ex2_approx(v[2 * j] - mx), p1 = ex2_approx(v[2 * j + 1] - mx);
                        l += p0 + p1;
                        pk[j] = live ? pack_bf16x2(p0, p1) : 0u;
                    }
                    // keys [16 c, +16) = 16-byte chunks 2c, 2c+1 of the row; chunk i of key block i/8 sits at (i%8) ^ (row%8).
                    // (this overwrites the Q | K tiles of the stage: S = Q K^T is complete, nobody reads them any more)
                    const uint32_t blk = p_row + (uint32_t)(c >> 2) * kTileBytes;
                    sts128u(blk + ((((uint32_t)(2 * c) & 7u) ^ p_x) << 4), pk[0], pk[1], pk[2], pk[3]);
                    sts128u(blk + ((((uint32_t)(2 * c + 1) & 7u) ^ p_x) << 4), pk[4], pk[5], pk[6], pk[7]);
                }
            }
            fence_proxy_async_smem();                 // generic-proxy smem writes -> visible to the tensor core (async proxy)
            mbar_arrive(p_full + g);
            mbar_wait(o_full + g, par);
            tcgen05_fence_after();
            // O in two halves of 32 columns (64 accumulator registers at once were the kernel's register peak: it spilled under the 96 that two
            // resident CTAs leave)
            const float inv = 1.0f / l;
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + ((int64_t)b * Lq + row) * a.ldo + (int64_t)hd * kD;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t o[32];
                tmem_ld_32x32b_x32(t_o + (uint32_t)(32 * hf), o);
                tmem_ld_wait();
                if (hf == 1) {
                    tcgen05_fence_before();
                    mbar_arrive(acc_empty + g);       // S / O columns and the P buffer may be overwritten by item u+1 of this group
                }
                if (live) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 w;
                        w.x = pack_bf16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
                        w.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
                        w.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
                        w.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
                        *reinterpret_cast<uint4*>(op + 32 * hf + 8 * c) = w;
                    }
                }
            }
            if (live && a.lse) a.lse[((int64_t)b * H + hd) * Lq + row] = fmaf(mx, kLn2, __logf(l));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kFwdTmemCols) : "memory");
    }
}


// ------------------------------------------------------------------------------------------------ forward, wide heads / long rows
// Head dim 64 + TW (TW = 32: M4C's 96-wide heads, TW = 64: 128), Lq of any length as 128-row tiles (work item = sample, head, row tile),
// Lk <= 256.  Same pipeline as above with ONE CTA per SM and two load stages (a stage is up to 96 KB here): producer warp, MMA warp, two
// softmax groups.  A head is two column blocks: 64 columns as SWIZZLE_128B rows and a TW-column tail block as SWIZZLE_64B rows (32
// columns) or a second SWIZZLE_128B block (64), so nothing is padded in shared memory:  S = Q K^T is 4 + TW/16 MMAs,  O = P V is one MMA per
// 16 keys into columns [0,64) plus one into [64, 64+TW) (the tail block of V is its own MN-major operand).
struct Tc5WideArgs {
    SdpaArgs a;
    int nk16, q_rows, n_mt;      // keys rounded to 16; rows of the Q box; row tiles per (sample, head)
};
// Softmax groups have TWO threads per query row (8 warps per group: warps h*4 .. h*4+3 own the 16-key chunks c with c % 2 == h).  With one
// thread per row the element-wise chain of an item (~3 900 instructions for 192 keys) ran at an IPC of 0.16 per warp with two warps per
// scheduler: 92 us per launch of the M4C shape against a 22 us HBM bound (profiles/r02_kernels_ncu.md).
constexpr int kWideThreads = 64 + 2 * 256;

template <int TW>
__global__ void __launch_bounds__(kWideThreads, 1)
sdpa_fwd_tc5w_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                     const __grid_constant__ CUtensorMap map_k0, const __grid_constant__ CUtensorMap map_k1,
                     const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_v1, Tc5WideArgs args) {
    constexpr int D = 64 + TW;
    constexpr int kRowB = TW == 64 ? 128 : 64;           // bytes of a tail-block row
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SdpaArgs& a = args.a;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int Lq = (int)a.Lq, Lk = (int)a.Lk, H = (int)a.h, nk16 = args.nk16, n_mt = args.n_mt;
    // stage: Q0 [128 x 128 B] | Q1 [128 x kRowB] | K0 [nk16 x 128] | K1 [nk16 x kRowB] | V0 | V1;  P (nk16/64 blocks of 128 x 128 B) overwrites Q | K
    const uint32_t q1_off = kTileBytes, k0_off = q1_off + 128 * kRowB, k1_off = k0_off + (uint32_t)nk16 * 128u,
                   v0_off = k1_off + (uint32_t)nk16 * kRowB, v1_off = v0_off + (uint32_t)nk16 * 128u, stage_bytes = v1_off + (uint32_t)nk16 * kRowB;
    uint8_t* s_load = smem;
    float* s_mask = reinterpret_cast<float*>(smem + 2 * stage_bytes);     // [2 groups][256]
    float* s_red = s_mask + 2 * 256;                                      // [2 groups][max | sum][2 halves][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_red + 2 * 2 * 2 * 128);
    uint64_t* load_full = bars;          // [2]
    uint64_t* load_empty = bars + 2;     // [2]
    uint64_t* s_full = bars + 4;         // [2]
    uint64_t* p_full = bars + 6;         // [2]  128 arrivals
    uint64_t* o_full = bars + 8;         // [2]
    uint64_t* acc_empty = bars + 10;     // [2]  128 arrivals
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
    const int n_items = (int)a.B * H * n_mt;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(load_full + i, 1); mbar_init(load_empty + i, 1); mbar_init(s_full + i, 1); mbar_init(p_full + i, 256);
            mbar_init(o_full + i, 1); mbar_init(acc_empty + i, 256);
        }
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
    pdl_wait();          // programmatic dependent launch: the set-up above overlaps the previous kernel's drain; no global access before this
    pdl_trigger();           // group g: S in columns [256 g, +nk16), O aliased onto [256 g, +D)

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(args.q_rows + 2 * nk16) * (128u + kRowB);
            int it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int ls = it & 1;
                const int mt = item % n_mt, bh = item / n_mt, b = bh / H, hd = bh % H;
                mbar_wait(load_empty + ls, (((uint32_t)(it >> 1)) & 1u) ^ 1u);
                mbar_expect_tx(load_full + ls, bytes);
                uint8_t* st = s_load + ls * stage_bytes;
                const int qrow = b * Lq + mt * 128, krow = b * Lk, c0 = hd * D, c1 = hd * D + 64;
                tma_load_2d(st, &map_q0, load_full + ls, c0, qrow);
                tma_load_2d(st + q1_off, &map_q1, load_full + ls, c1, qrow);
                tma_load_2d(st + k0_off, &map_k0, load_full + ls, c0, krow);
                tma_load_2d(st + k1_off, &map_k1, load_full + ls, c1, krow);
                tma_load_2d(st + v0_off, &map_v0, load_full + ls, c0, krow);
                tma_load_2d(st + v1_off, &map_v1, load_full + ls, c1, krow);
                const int nx = item + 2 * (int)gridDim.x;            // (see sdpa_fwd_tc5_kernel: the item two ahead waits for a stage; pull it into L2)
                if (nx < n_items) {
                    const int mt2 = nx % n_mt, bh2 = nx / n_mt, b2 = bh2 / H, h2 = bh2 % H;
                    const int qrow2 = b2 * Lq + mt2 * 128, krow2 = b2 * Lk, d0 = h2 * D, d1 = h2 * D + 64;
                    tma_prefetch_2d(&map_q0, d0, qrow2); tma_prefetch_2d(&map_q1, d1, qrow2);
                    tma_prefetch_2d(&map_k0, d0, krow2); tma_prefetch_2d(&map_k1, d1, krow2);
                    tma_prefetch_2d(&map_v0, d0, krow2); tma_prefetch_2d(&map_v1, d1, krow2);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc(128, nk16, false, false);
            const uint32_t idesc_o0 = make_idesc(128, 64, false, true), idesc_o1 = make_idesc(128, TW, false, true);
            const int my_items = n_items > (int)blockIdx.x ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
            auto tail_desc_k = [&](uint32_t addr, int t) {           // K-major tail block, k-step t (16 columns)
                return TW == 64 ? make_smem_desc(addr + t * 32, 16, 1024) : make_smem_desc_sw64(addr + t * 32, 16, 512);
            };
            auto issue_s = [&](int it) {
                const int ls = it & 1, g = it & 1;
                mbar_wait(load_full + ls, (uint32_t)(it >> 1) & 1u);
                mbar_wait(acc_empty + g, ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t st = smem_u32(s_load + ls * stage_bytes);
                const uint32_t d_s = tmem_base + (uint32_t)(g * 256);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    tcgen05_mma_bf16(d_s, make_smem_desc(st + kk * 32, 16, 1024), make_smem_desc(st + k0_off + kk * 32, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int t = 0; t < TW / 16; ++t)
                    tcgen05_mma_bf16(d_s, tail_desc_k(st + q1_off, t), tail_desc_k(st + k1_off, t), idesc_s, 1u);
                tcgen05_commit(s_full + g);
            };
            auto issue_pv = [&](int it) {
                const int ls = it & 1, g = it & 1;
                mbar_wait(p_full + g, (uint32_t)(it >> 1) & 1u);
                tcgen05_fence_after();
                const uint32_t st = smem_u32(s_load + ls * stage_bytes);
                const uint32_t d_o = tmem_base + (uint32_t)(g * 256);
                for (int ks = 0; ks < nk16 / 16; ++ks) {
                    const uint64_t p_desc = make_smem_desc(st + (ks >> 2) * kTileBytes + (ks & 3) * 32, 16, 1024);
                    tcgen05_mma_bf16(d_o, p_desc, make_smem_desc(st + v0_off + ks * 2048, 8192, 1024), idesc_o0, ks > 0 ? 1u : 0u);
                    // tail block of V: MN-major rows of kRowB bytes (16 keys = 16 rows)
                    const uint64_t v1 = TW == 64 ? make_smem_desc(st + v1_off + ks * 2048, 8192, 1024) : make_smem_desc_sw64(st + v1_off + ks * 1024, 4096, 512);
                    tcgen05_mma_bf16(d_o + 64u, p_desc, v1, idesc_o1, ks > 0 ? 1u : 0u);
                }
                tcgen05_commit(o_full + g);
                tcgen05_commit(load_empty + ls);
            };
            if (my_items > 0) issue_s(0);
            for (int it = 0; it < my_items; ++it) {
                if (it + 1 < my_items) issue_s(it + 1);
                issue_pv(it);
            }
        }
    } else {
        const int g = (warp - 2) / 8, half = ((warp - 2) / 4) % 2, quad = warp % 4, row = quad * 32 + lane, tid_g = (warp - 2) % 8 * 32 + lane;
        float* mrow = s_mask + g * 256;
        float* red_max = s_red + g * 512;             // [2 halves][128]
        float* red_sum = red_max + 256;               // [2 halves][128]
        const uint32_t mrow_u32 = smem_u32(mrow);
        const uint32_t p_row = smem_u32(s_load + g * stage_bytes) + (uint32_t)row * 128u, p_x = (uint32_t)(row & 7);   // stage g == group g
        const uint32_t t_s = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(g * 256);
        const float sc2 = a.scale * kLog2e;
        const int nch = nk16 / 16;
        const int ctail = (int)a.mask.causal_tail;
        const float cval2 = a.mask.causal_value * kLog2e;
        auto mask_of = [&](int item, int key) -> float {
            if (key >= Lk || item >= n_items) return -INFINITY;
            const int bh = item / n_mt;
            return a.mask.ptr ? __ldg(a.mask.ptr + (int64_t)(bh / H) * a.mask.sb + (int64_t)(bh % H) * a.mask.sh + key) : 0.0f;
        };
        const int first = blockIdx.x + g * gridDim.x;
        float mv = mask_of(first, tid_g);
        int u = 0;
        for (int item = first; item < n_items; item += 2 * gridDim.x, ++u) {
            const int mt = item % n_mt, bh = item / n_mt, b = bh / H, hd = bh % H;
            const uint32_t par = (uint32_t)u & 1u;
            named_bar_sync(1 + g, 256);               // the previous item's readers of the mask row / reduction scratch are done
            mrow[tid_g] = mv * kLog2e;
            mv = mask_of(item + 2 * gridDim.x, tid_g);
            named_bar_sync(1 + g, 256);
            mbar_wait(s_full + g, par);
            tcgen05_fence_after();
            // prefix-LM tail: keys beyond `kmax` are in this query's causal future (kmax = INT_MAX for rows outside the tail)
            const int qi_ = mt * 128 + row;
            const int kmax = (ctail > 0 && qi_ >= Lq - ctail) ? (Lk - ctail) + (qi_ - (Lq - ctail)) : 0x7fffffff;
            auto scaled = [&](const uint32_t (&r)[16], int ch, float (&v)[16]) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    float m0, m1, m2, m3;
                    lds128f(mrow_u32 + (uint32_t)(16 * ch + 4 * q4) * 4u, m0, m1, m2, m3);
                    v[4 * q4] = fmaf(__uint_as_float(r[4 * q4]), sc2, m0);
                    v[4 * q4 + 1] = fmaf(__uint_as_float(r[4 * q4 + 1]), sc2, m1);
                    v[4 * q4 + 2] = fmaf(__uint_as_float(r[4 * q4 + 2]), sc2, m2);
                    v[4 * q4 + 3] = fmaf(__uint_as_float(r[4 * q4 + 3]), sc2, m3);
                }
                if (16 * ch + 15 >= Lk - ctail && ctail > 0) {          // (warp-uniform) the chunk reaches into the causal block
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += (16 * ch + j > kmax) ? cval2 : 0.0f;
                }
            };
            // this thread's chunks: half, half + 2, ...  (nmine of them; chunk index of the i-th: 2 i + half)
            const int nmine = (nch - half + 1) / 2;
            uint32_t ra[16], rb[16];
            float mx = -INFINITY;
            // (two chunks per trip of a rolled loop: fully unrolled, the two passes were ~5 600 SASS lines and a quarter of the stall samples
            // were instruction-cache misses — profiles/r02_kernels_ncu.md)
            if (nmine > 0) tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * half), ra);
            auto max_of = [&](const uint32_t (&r)[16], int c) {
                float v[16];
                scaled(r, c, v);
#pragma unroll
                for (int j = 0; j < 16; j += 4) mx = fmaxf(fmaxf(mx, fmaxf(v[j], v[j + 1])), fmaxf(v[j + 2], v[j + 3]));
            };
#pragma unroll 1
            for (int i = 0; i < nmine; i += 2) {
                const int c = 2 * i + half;
                tmem_ld_wait();
                if (i + 1 < nmine) tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * (c + 2)), rb);
                max_of(ra, c);
                if (i + 1 < nmine) {
                    tmem_ld_wait();
                    if (i + 2 < nmine) tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * (c + 4)), ra);
                    max_of(rb, c + 2);
                }
            }
            red_max[half * 128 + row] = mx;
            named_bar_sync(1 + g, 256);
            mx = fmaxf(mx, red_max[(half ^ 1) * 128 + row]);
            float l = 0.f;
            const int qi = mt * 128 + row;
            const bool live = qi < Lq;
            if (nmine > 0) tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * half), ra);
            auto probs_of = [&](const uint32_t (&r)[16], int c) {
                float v[16];
                scaled(r, c, v);
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float p0 = ex2_approx(v[2 * j] - mx), p1 = ex2_approx(v[2 * j + 1] - mx);
                    l += p0 + p1;
                    pk[j] = live ? pack_bf16x2(p0, p1) : 0u;
                }
                const uint32_t blk = p_row + (uint32_t)(c >> 2) * kTileBytes;
                sts128u(blk + ((((uint32_t)(2 * c) & 7u) ^ p_x) << 4), pk[0], pk[1], pk[2], pk[3]);
                sts128u(blk + ((((uint32_t)(2 * c + 1) & 7u) ^ p_x) << 4), pk[4], pk[5], pk[6], pk[7]);
            };
#pragma unroll 1
            for (int i = 0; i < nmine; i += 2) {
                const int c = 2 * i + half;
                tmem_ld_wait();
                if (i + 1 < nmine) tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * (c + 2)), rb);
                probs_of(ra, c);
                if (i + 1 < nmine) {
                    tmem_ld_wait();
                    if (i + 2 < nmine) tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * (c + 4)), ra);
                    probs_of(rb, c + 2);
                }
            }
            red_sum[half * 128 + row] = l;            // (visible to the partner after the p_full -> o_full barrier chain: release / acquire)
            fence_proxy_async_smem();
            mbar_arrive(p_full + g);
            mbar_wait(o_full + g, par);
            tcgen05_fence_after();
            l += red_sum[(half ^ 1) * 128 + row];
            const float inv = 1.0f / l;
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(a.out) + ((int64_t)b * Lq + qi) * a.ldo + (int64_t)hd * D;
            // O drain: 16-column chunks, alternating between the two threads of a row
            constexpr int kOch = D / 16;
#pragma unroll
            for (int c16 = 0; c16 < kOch; ++c16) {
                if ((c16 & 1) != half) continue;
                uint32_t o[16];
                tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * c16), o);
                tmem_ld_wait();
                if (c16 + 2 >= kOch) {               // this thread's last read of the accumulator
                    tcgen05_fence_before();
                    mbar_arrive(acc_empty + g);
                }
                if (live) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        uint4 w;
                        w.x = pack_bf16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
                        w.y = pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
                        w.z = pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
                        w.w = pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
                        *reinterpret_cast<uint4*>(op + 16 * c16 + 8 * c) = w;
                    }
                }
            }
            if (live && half == 0 && a.lse) a.lse[((int64_t)b * H + hd) * Lq + qi] = fmaf(mx, kLn2, __logf(l));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ backward
// Per item (sample, head):   S = Q K^T,  dP = dO V^T                          (tcgen05, accumulators in TMEM)
//   phase A, two threads per query row (alternate 16-key chunks; the element-wise chain is the critical path of an item):
//                                       P = exp(S*scale + mask - lse),  delta = sum_j P dP  (== dO . O, so O is never read; the two
//                                       partial sums meet in shared memory),  dS = P (dP - delta);  P and dS go to shared memory as
//                                       bf16 [query][key] tiles
//   dV = P^T dO,  dK = dS^T Q,  dQ = dS K    — the SAME two tiles serve as MN-major A operands (transposed use) and as a
//                                       K-major A operand: no transposition pass anywhere
//   phase B, one thread per output row: TMEM -> bf16 -> global (dK, dQ scaled by 1/sqrt(d) here, in fp32)
// warp 0 = TMA producer (Q, dO, K, V; 2 stages), warp 1 = MMA issuer, warps 2-9 = phase A, warps 10-13 = phase B, so the
// element-wise work of item i+1 overlaps the output drain of item i and the loads of item i+2.  S(i+1) / dP(i+1) are issued
// BEFORE dV/dK/dQ(i): phase A never waits for the three long products.
constexpr int kBwdStages = 2;
constexpr int kBwdStageBytes = 4 * kTileBytes;         // Q, dO, K, V
constexpr int kBwdAThreads = 256;                      // phase A: TWO threads per query row, each owning every other 16-key chunk
constexpr int kBwdThreads = 64 + kBwdAThreads + 128;
constexpr int kBwdSmem = kBwdStages * kBwdStageBytes + 2 * kPBytes /*P, dS*/ + 128 * 4 /*mask row*/ + 2 * 2 * 128 * 4 /*delta partials*/ +
                         256 /*barriers*/ + 1024 /*align*/;

struct Tc5BwdArgs {
    SdpaArgs a;
    void *dq, *dk, *dv;
    int64_t lddq, lddk, lddv;
    int nk16, nq16;
};

__device__ __forceinline__ void store_row64(__nv_bfloat16* dst, const uint32_t (&r)[kD], float mul) {
#pragma unroll
    for (int c = 0; c < kD / 8; ++c) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(r[8 * c]) * mul, __uint_as_float(r[8 * c + 1]) * mul);
        w.y = pack_bf16x2(__uint_as_float(r[8 * c + 2]) * mul, __uint_as_float(r[8 * c + 3]) * mul);
        w.z = pack_bf16x2(__uint_as_float(r[8 * c + 4]) * mul, __uint_as_float(r[8 * c + 5]) * mul);
        w.w = pack_bf16x2(__uint_as_float(r[8 * c + 6]) * mul, __uint_as_float(r[8 * c + 7]) * mul);
        *reinterpret_cast<uint4*>(dst + 8 * c) = w;
    }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
sdpa_bwd_tc5_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v, Tc5BwdArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* s_load = smem;                                          // [kBwdStages][Q | dO | K | V]
    uint8_t* s_p = smem + kBwdStages * kBwdStageBytes;               // P  [2 key blocks][128 queries][128 B]
    uint8_t* s_ds = s_p + kPBytes;                                   // dS, same layout
    float* s_mask = reinterpret_cast<float*>(s_ds + kPBytes);        // [128]
    float* s_delta = s_mask + 128;                                   // [2 item parities][2 halves][128]  partial sums of P dP
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_delta + 2 * 2 * 128);
    uint64_t* load_full = bars;                     // [2]  TMA -> MMA (and phase A)
    uint64_t* load_empty = load_full + kBwdStages;  // [2]  MMA -> TMA
    // Tensor memory is two REGIONS of 256 columns; item `it` lives in region it & 1: S | dP while phase A works on it, then dV | dK | dQ on
    // top of them (dead once P and dS are in shared memory).  So S / dP of item it + 1 are issued — and usually complete — while phase A is
    // still on item it: phase A never waits for the tensor pipe between items (it did for ~15 % of an item with one S | dP buffer).
    uint64_t* sdp_full = load_empty + kBwdStages;   // [2] MMA -> phase A: S and dP of the region's item ready
    uint64_t* pds_full = sdp_full + 2;              // phase A -> MMA: P and dS in smem, S / dP columns free   (256 arrivals)
    uint64_t* out_full = pds_full + 1;              // [2] MMA -> phase B (dV, dK, dQ ready) and -> phase A (P / dS tiles no longer read)
    uint64_t* out_free = out_full + 2;              // [2] phase B -> MMA: the region's dV / dK / dQ columns drained   (128 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_free + 2);

    const SdpaArgs& a = args.a;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int Lq = (int)a.Lq, Lk = (int)a.Lk, H = (int)a.h, nk16 = args.nk16, nq16 = args.nq16;
    const int n_items = (int)a.B * H;
    const int my_items = n_items > (int)blockIdx.x ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_do)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_k)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_v)) : "memory");
        for (int i = 0; i < kBwdStages; ++i) { mbar_init(load_full + i, 1); mbar_init(load_empty + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(sdp_full + i, 1); mbar_init(out_full + i, 1); mbar_init(out_free + i, 128); }
        mbar_init(pds_full, kBwdAThreads);
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
    pdl_wait();          // programmatic dependent launch: the set-up above overlaps the previous kernel's drain; no global access before this
    pdl_trigger();
    constexpr uint32_t kRegion = 256, kColS = 0, kColDP = 128, kColDV = 0, kColDK = 64, kColDQ = 128;      // column offsets inside a region

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(2 * nq16 + 2 * nk16) * 128u;
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int ls = it % kBwdStages;
                const int b = item / H, hd = item % H;
                mbar_wait(load_empty + ls, ((uint32_t)(it / kBwdStages) & 1u) ^ 1u);
                mbar_expect_tx(load_full + ls, bytes);
                uint8_t* st = s_load + ls * kBwdStageBytes;
                tma_load_2d(st, &map_q, load_full + ls, hd * kD, b * Lq);
                tma_load_2d(st + kTileBytes, &map_do, load_full + ls, hd * kD, b * Lq);
                tma_load_2d(st + 2 * kTileBytes, &map_k, load_full + ls, hd * kD, b * Lk);
                tma_load_2d(st + 3 * kTileBytes, &map_v, load_full + ls, hd * kD, b * Lk);
                const int nx = item + kBwdStages * (int)gridDim.x;   // the item that will wait for this stage: DRAM -> L2 while it waits
                if (nx < n_items) {
                    const int b2 = nx / H, h2 = nx % H;
                    tma_prefetch_2d(&map_q, h2 * kD, b2 * Lq); tma_prefetch_2d(&map_do, h2 * kD, b2 * Lq);
                    tma_prefetch_2d(&map_k, h2 * kD, b2 * Lk); tma_prefetch_2d(&map_v, h2 * kD, b2 * Lk);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc(128, nk16, false, false);      // S, dP: A, B K-major
            const uint32_t idesc_t = make_idesc(128, kD, true, true);          // dV, dK: A = tile^T (MN-major), B MN-major
            const uint32_t idesc_q = make_idesc(128, kD, false, true);         // dQ: A K-major, B MN-major
            auto issue_sdp = [&](int it) {
                const int ls = it % kBwdStages;
                const uint32_t reg = tmem_base + (uint32_t)(it & 1) * kRegion;
                mbar_wait(load_full + ls, (uint32_t)(it / kBwdStages) & 1u);
                tcgen05_fence_after();
                const uint32_t q_addr = smem_u32(s_load + ls * kBwdStageBytes), do_addr = q_addr + kTileBytes, k_addr = q_addr + 2 * kTileBytes,
                               v_addr = q_addr + 3 * kTileBytes;
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    tcgen05_mma_bf16(reg + kColS, make_smem_desc(q_addr + kk * 32, 16, 1024), make_smem_desc(k_addr + kk * 32, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
                for (int kk = 0; kk < kD / 16; ++kk)
                    tcgen05_mma_bf16(reg + kColDP, make_smem_desc(do_addr + kk * 32, 16, 1024), make_smem_desc(v_addr + kk * 32, 16, 1024), idesc_s, kk > 0 ? 1u : 0u);
                tcgen05_commit(sdp_full + (it & 1));
            };
            if (my_items > 0) issue_sdp(0);
            if (my_items > 1) issue_sdp(1);
            for (int it = 0; it < my_items; ++it) {
                const int ls = it % kBwdStages;
                const uint32_t reg = tmem_base + (uint32_t)(it & 1) * kRegion;
                mbar_wait(pds_full, (uint32_t)it & 1u);   // P(it), dS(it) written; S / dP columns of the region consumed
                tcgen05_fence_after();
                const uint32_t q_addr = smem_u32(s_load + ls * kBwdStageBytes), do_addr = q_addr + kTileBytes, k_addr = q_addr + 2 * kTileBytes;
                const uint32_t p_addr = smem_u32(s_p), ds_addr = smem_u32(s_ds);
                // dV[key, d] = sum_q P[q, key] dO[q, d]: A = P^T (MN-major: 64-key blocks kTileBytes apart, 16 queries = 2048 B), B = dO (MN-major)
                for (int ks = 0; ks < nq16 / 16; ++ks)
                    tcgen05_mma_bf16(reg + kColDV, make_smem_desc(p_addr + ks * 2048, kTileBytes, 1024), make_smem_desc(do_addr + ks * 2048, 8192, 1024), idesc_t, ks > 0 ? 1u : 0u);
                for (int ks = 0; ks < nq16 / 16; ++ks)
                    tcgen05_mma_bf16(reg + kColDK, make_smem_desc(ds_addr + ks * 2048, kTileBytes, 1024), make_smem_desc(q_addr + ks * 2048, 8192, 1024), idesc_t, ks > 0 ? 1u : 0u);
                // dQ[q, d] = sum_key dS[q, key] K[key, d]: A = dS (K-major), B = K (MN-major)
                for (int ks = 0; ks < nk16 / 16; ++ks)
                    tcgen05_mma_bf16(reg + kColDQ, make_smem_desc(ds_addr + (ks >> 2) * kTileBytes + (ks & 3) * 32, 16, 1024),
                                     make_smem_desc(k_addr + ks * 2048, 8192, 1024), idesc_q, ks > 0 ? 1u : 0u);
                tcgen05_commit(out_full + (it & 1));
                tcgen05_commit(load_empty + ls);
                if (it + 2 < my_items) {
                    mbar_wait(out_free + (it & 1), (uint32_t)(it >> 1) & 1u);     // dV / dK / dQ of item `it` drained: the region takes S | dP of item it + 2
                    issue_sdp(it + 2);
                }
            }
        }
    } else if (warp < 2 + kBwdAThreads / 32) {
        // ---------------------------------------------------------------- phase A: two threads per query row
        const int quad = warp % 4, row = quad * 32 + lane, half = (warp - 2) / 4, tid_g = (warp - 2) * 32 + lane;
        const uint32_t mrow_u32 = smem_u32(s_mask);
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const uint32_t p_row = smem_u32(s_p) + (uint32_t)row * 128u, ds_row = smem_u32(s_ds) + (uint32_t)row * 128u, p_x = (uint32_t)(row & 7);
        const float sc2 = a.scale * kLog2e;
        const int nch = nk16 / 16;
        const int n_mine = (nch - half + 1) / 2;         // this thread's chunks: half, half + 2, ...
        const bool live = row < Lq;
        auto mask_of = [&](int item) -> float {          // fetched one item ahead (as lse): off the critical path
            if (tid_g >= Lk || item >= n_items) return -INFINITY;
            return a.mask.ptr ? __ldg(a.mask.ptr + (int64_t)(item / H) * a.mask.sb + (int64_t)(item % H) * a.mask.sh + tid_g) : 0.0f;
        };
        auto lse_of = [&](int item) -> float { return (live && item < n_items) ? __ldg(a.lse + (int64_t)item * Lq + row) : 0.0f; };
        float mv_next = mask_of(blockIdx.x), lse_next = lse_of(blockIdx.x);
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            named_bar_sync(1, kBwdAThreads);
            if (tid_g < 128) s_mask[tid_g] = mv_next * kLog2e;
            const float lse2 = lse_next * kLog2e;
            mv_next = mask_of(item + gridDim.x);
            lse_next = lse_of(item + gridDim.x);
            named_bar_sync(1, kBwdAThreads);
            const uint32_t t_s = t_lane + (uint32_t)(it & 1) * kRegion + kColS, t_dp = t_lane + (uint32_t)(it & 1) * kRegion + kColDP;
            mbar_wait(sdp_full + (it & 1), (uint32_t)(it >> 1) & 1u);
            tcgen05_fence_after();
            auto probs = [&](const uint32_t (&rs)[16], int ch, float (&pv)[16]) {       // P for keys [16 ch, +16)
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    float m0, m1, m2, m3;
                    lds128f(mrow_u32 + (uint32_t)(16 * ch + 4 * q4) * 4u, m0, m1, m2, m3);
                    pv[4 * q4] = ex2_approx(fmaf(__uint_as_float(rs[4 * q4]), sc2, m0) - lse2);
                    pv[4 * q4 + 1] = ex2_approx(fmaf(__uint_as_float(rs[4 * q4 + 1]), sc2, m1) - lse2);
                    pv[4 * q4 + 2] = ex2_approx(fmaf(__uint_as_float(rs[4 * q4 + 2]), sc2, m2) - lse2);
                    pv[4 * q4 + 3] = ex2_approx(fmaf(__uint_as_float(rs[4 * q4 + 3]), sc2, m3) - lse2);
                }
            };
            auto load_pair = [&](uint32_t (&rs)[16], uint32_t (&rd)[16], int ch) {
                tmem_ld_32x32b_x16(t_s + (uint32_t)(16 * ch), rs);
                tmem_ld_32x32b_x16(t_dp + (uint32_t)(16 * ch), rd);
            };
            uint32_t sa[16], da[16], sb[16], db[16];
            // pass 1: delta = sum_j P_j dP_j  (this thread's chunks; the other half's partial sum arrives through shared memory)
            float delta = 0.f;
            if (n_mine > 0) load_pair(sa, da, half);
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {
                if (kq < n_mine) {
                    const int ch = half + 2 * kq;
                    tmem_ld_wait();
                    if (kq + 1 < n_mine) load_pair((kq & 1) ? sa : sb, (kq & 1) ? da : db, ch + 2);
                    float pv[16];
                    probs((kq & 1) ? sb : sa, ch, pv);
#pragma unroll
                    for (int j = 0; j < 16; ++j) delta = fmaf(pv[j], __uint_as_float(((kq & 1) ? db : da)[j]), delta);
                }
            }
            float* dpart = s_delta + (it & 1) * 256;
            dpart[half * 128 + row] = delta;
            // start re-reading while the partial sums cross
            if (n_mine > 0) load_pair(sa, da, half);
            named_bar_sync(2, kBwdAThreads);
            delta += dpart[(half ^ 1) * 128 + row];
            // the P / dS tiles are still being read by dV/dK/dQ of the previous item until out_full(it-1) fires
            if (it > 0) mbar_wait(out_full + ((it - 1) & 1), (uint32_t)((it - 1) >> 1) & 1u);
            // pass 2: P and dS = P (dP - delta) -> shared memory (bf16, [query][key], SWIZZLE_128B rows)
#pragma unroll
            for (int kq = 0; kq < 4; ++kq) {
                if (kq < n_mine) {
                    const int ch = half + 2 * kq;
                    tmem_ld_wait();
                    if (kq + 1 < n_mine) load_pair((kq & 1) ? sa : sb, (kq & 1) ? da : db, ch + 2);
                    float pv[16];
                    probs((kq & 1) ? sb : sa, ch, pv);
                    uint32_t pp[8], pd[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float d0 = pv[2 * j] * (__uint_as_float(((kq & 1) ? db : da)[2 * j]) - delta);
                        const float d1 = pv[2 * j + 1] * (__uint_as_float(((kq & 1) ? db : da)[2 * j + 1]) - delta);
                        pp[j] = live ? pack_bf16x2(pv[2 * j], pv[2 * j + 1]) : 0u;
                        pd[j] = live ? pack_bf16x2(d0, d1) : 0u;
                    }
                    const uint32_t off = (uint32_t)(ch >> 2) * kTileBytes;
                    const uint32_t c0 = ((((uint32_t)(2 * ch)) & 7u) ^ p_x) << 4, c1 = ((((uint32_t)(2 * ch + 1)) & 7u) ^ p_x) << 4;
                    sts128u(p_row + off + c0, pp[0], pp[1], pp[2], pp[3]);
                    sts128u(p_row + off + c1, pp[4], pp[5], pp[6], pp[7]);
                    sts128u(ds_row + off + c0, pd[0], pd[1], pd[2], pd[3]);
                    sts128u(ds_row + off + c1, pd[4], pd[5], pd[6], pd[7]);
                }
            }
            fence_proxy_async_smem();
            tcgen05_fence_before();
            mbar_arrive(pds_full);
        }
    } else {
        // ---------------------------------------------------------------- phase B: one thread per output row (key for dV, dK; query for dQ)
        const int quad = warp % 4, row = quad * 32 + lane;
        const uint32_t t_row0 = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int b = item / H, hd = item % H;
            mbar_wait(out_full + (it & 1), (uint32_t)(it >> 1) & 1u);
            tcgen05_fence_after();
            const uint32_t t_row = t_row0 + (uint32_t)(it & 1) * kRegion;
            uint32_t r0[kD];
            tmem_ld_32x32b_x32(t_row + kColDV, r0); tmem_ld_32x32b_x32(t_row + kColDV + 32u, r0 + 32);
            tmem_ld_wait();
            if (row < Lk) store_row64(reinterpret_cast<__nv_bfloat16*>(args.dv) + ((int64_t)b * Lk + row) * args.lddv + (int64_t)hd * kD, r0, 1.0f);
            tmem_ld_32x32b_x32(t_row + kColDK, r0); tmem_ld_32x32b_x32(t_row + kColDK + 32u, r0 + 32);
            tmem_ld_wait();
            if (row < Lk) store_row64(reinterpret_cast<__nv_bfloat16*>(args.dk) + ((int64_t)b * Lk + row) * args.lddk + (int64_t)hd * kD, r0, a.scale);
            tmem_ld_32x32b_x32(t_row + kColDQ, r0); tmem_ld_32x32b_x32(t_row + kColDQ + 32u, r0 + 32);
            tmem_ld_wait();
            tcgen05_fence_before();
            mbar_arrive(out_free + (it & 1));
            if (row < Lq) store_row64(reinterpret_cast<__nv_bfloat16*>(args.dq) + ((int64_t)b * Lq + row) * args.lddq + (int64_t)hd * kD, r0, a.scale);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace

// OVQA_SDPA_TC5 / OVQA_SDPA_TC5_BWD: "0" = never, "1" = whenever the shape is supported, unset = where it is the faster core
// (profiles/r01f: a 128-row tile that is 3/4 padding loses to the 16-row warp tiles of attention_mma.cu).  Read per call.
static int tc5_mode(const char* name) {
    const char* e = getenv(name);
    return !e ? -1 : e[0] == '0' ? 0 : 1;
}

static bool tc5_shape_ok(const SdpaArgs& a) {
    if (a.geo.boxes || a.att || a.mask.causal_tail > 0) return false;
    if (a.dk != kD || a.dv != kD || a.Lq > 128 || a.Lk > 128 || a.Lq < 1 || a.Lk < 1) return false;
    if (a.mask.ptr && a.mask.sq != 0) return false;                           // per-key (padding) masks only
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(a.out) || (a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (a.ldo % 8)) return false;
    return a.B * a.h >= 1;
}

bool sdpa_fwd_tc5_supported(const SdpaArgs& a) {
    const int mode = tc5_mode("OVQA_SDPA_TC5");
    if (mode == 0 || !tc5_shape_ok(a)) return false;
    return mode == 1 || a.Lq > 48;
}

int sdpa_fwd_tc5_bf16(const SdpaArgs& a, cudaStream_t st) {
    Tc5FwdArgs args{a, (int)((a.Lk + 15) / 16 * 16)};
    CUtensorMap mq, mk, mv;
    int rc;
    const uint64_t cols = (uint64_t)a.h * kD;
    if ((rc = make_tma_map_bf16(&mq, a.q, cols, (uint64_t)(a.B * a.Lq), (uint64_t)a.ldq, kD, (uint32_t)a.Lq))) return rc;
    if ((rc = make_tma_map_bf16(&mk, a.k, cols, (uint64_t)(a.B * a.Lk), (uint64_t)a.ldk, kD, (uint32_t)args.nk16))) return rc;
    if ((rc = make_tma_map_bf16(&mv, a.v, cols, (uint64_t)(a.B * a.Lk), (uint64_t)a.ldv, kD, (uint32_t)args.nk16))) return rc;
    static OncePerDevice attr_set;
    if (attr_set.pending()) { OVQA_CUDA_OK(cudaFuncSetAttribute(sdpa_fwd_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem)); attr_set.mark(); }
    const int64_t items = a.B * a.h;
    const int64_t slots = 2 * (int64_t)sm_count();            // two CTAs per SM
    const int grid = (int)(items < slots ? items : slots);
    OVQA_CUDA_OK(launch_pdl(sdpa_fwd_tc5_kernel, dim3(grid), dim3(kFwdThreads), (size_t)(kFwdSmem), st, mq, mk, mv, args));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

// wide heads (d = 96 / 128) and rows longer than one tile: sdpa_fwd_tc5w_kernel
static size_t tc5w_stage_bytes(int tw, int nk16) { const int rowb = tw == 64 ? 128 : 64; return (size_t)(128 + 2 * nk16) * (128 + rowb); }
static size_t tc5w_smem_bytes(int tw, int nk16) { return 2 * tc5w_stage_bytes(tw, nk16) + 2 * 256 * 4 /*mask rows*/ + 2 * 512 * 4 /*max / sum exchange*/ + 256 + 1024; }

bool sdpa_fwd_tc5w_supported(const SdpaArgs& a) {
    const int mode = tc5_mode("OVQA_SDPA_TC5");
    if (mode == 0 || a.geo.boxes || a.att) return false;
    if (a.dk != a.dv || !(a.dk == 96 || a.dk == 128) || a.Lq < 1 || a.Lk < 1 || a.Lk > 256) return false;
    if (a.mask.ptr && a.mask.sq != 0) return false;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(a.out) || (a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (a.ldo % 8)) return false;
    const int tw = (int)a.dk - 64, nk16 = (int)((a.Lk + 15) / 16 * 16), rowb = tw == 64 ? 128 : 64;
    if (tc5w_smem_bytes(tw, nk16) > 227 * 1024) return false;
    // P (nk16/64 blocks of 128 x 128 B) must fit into the stage's Q | K tiles it overwrites
    if ((size_t)((nk16 + 63) / 64) * kTileBytes > (size_t)(128 + nk16) * (128 + rowb)) return false;
    return mode == 1 || a.Lq > 48;
}

template <int TW>
static int launch_tc5w(const SdpaArgs& a, cudaStream_t st) {
    const int D = 64 + TW;
    Tc5WideArgs args{a, (int)((a.Lk + 15) / 16 * 16), (int)(a.Lq < 128 ? a.Lq : 128), (int)((a.Lq + 127) / 128)};
    CUtensorMap m[6];
    const void* ptr[3] = {a.q, a.k, a.v};
    const int64_t ld[3] = {a.ldq, a.ldk, a.ldv}, rows[3] = {a.B * a.Lq, a.B * a.Lk, a.B * a.Lk};
    const uint32_t box_rows[3] = {(uint32_t)args.q_rows, (uint32_t)args.nk16, (uint32_t)args.nk16};
    const uint64_t cols = (uint64_t)a.h * D;
    for (int i = 0; i < 3; ++i) {
        int rc;
        if ((rc = make_tma_map_bf16(&m[2 * i], ptr[i], cols, (uint64_t)rows[i], (uint64_t)ld[i], 64, box_rows[i], 128))) return rc;
        if ((rc = make_tma_map_bf16(&m[2 * i + 1], ptr[i], cols, (uint64_t)rows[i], (uint64_t)ld[i], TW, box_rows[i], TW == 64 ? 128 : 64))) return rc;
    }
    const int smem = (int)tc5w_smem_bytes(TW, args.nk16);
    OVQA_CUDA_OK(cudaFuncSetAttribute(sdpa_fwd_tc5w_kernel<TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t items = a.B * a.h * args.n_mt;
    const int grid = (int)(items < sm_count() ? items : sm_count());
    OVQA_CUDA_OK(launch_pdl(sdpa_fwd_tc5w_kernel<TW>, dim3(grid), dim3(kWideThreads), (size_t)(smem), st, m[0], m[1], m[2], m[3], m[4], m[5], args));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

int sdpa_fwd_tc5w_bf16(const SdpaArgs& a, cudaStream_t st) { return a.dk == 96 ? launch_tc5w<32>(a, st) : launch_tc5w<64>(a, st); }

bool sdpa_bwd_tc5_supported(const SdpaArgs& a, const void* dout, const void* dq, const void* dk, const void* dv, int64_t lddo, int64_t lddq,
                            int64_t lddk, int64_t lddv) {
    const int mode = tc5_mode("OVQA_SDPA_TC5_BWD");
    if (mode == 0 || !tc5_shape_ok(a) || !a.lse) return false;
    if (mode != 1 && a.Lq <= 48 && a.Lk <= 48) return false;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return al16(dout) && al16(dq) && al16(dk) && al16(dv) && lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0;
}

int sdpa_bwd_tc5_bf16(const SdpaArgs& a, const void* dout, void* dq, void* dk, void* dv, int64_t lddo, int64_t lddq, int64_t lddk,
                      int64_t lddv, cudaStream_t st) {
    Tc5BwdArgs args{a, dq, dk, dv, lddq, lddk, lddv, (int)((a.Lk + 15) / 16 * 16), (int)((a.Lq + 15) / 16 * 16)};
    CUtensorMap mq, mdo, mk, mv;
    int rc;
    const uint64_t cols = (uint64_t)a.h * kD;
    if ((rc = make_tma_map_bf16(&mq, a.q, cols, (uint64_t)(a.B * a.Lq), (uint64_t)a.ldq, kD, (uint32_t)args.nq16))) return rc;
    if ((rc = make_tma_map_bf16(&mdo, dout, cols, (uint64_t)(a.B * a.Lq), (uint64_t)lddo, kD, (uint32_t)args.nq16))) return rc;
    if ((rc = make_tma_map_bf16(&mk, a.k, cols, (uint64_t)(a.B * a.Lk), (uint64_t)a.ldk, kD, (uint32_t)args.nk16))) return rc;
    if ((rc = make_tma_map_bf16(&mv, a.v, cols, (uint64_t)(a.B * a.Lk), (uint64_t)a.ldv, kD, (uint32_t)args.nk16))) return rc;
    static OncePerDevice attr_set;
    if (attr_set.pending()) { OVQA_CUDA_OK(cudaFuncSetAttribute(sdpa_bwd_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem)); attr_set.mark(); }
    const int64_t items = a.B * a.h;
    const int grid = (int)(items < sm_count() ? items : sm_count());
    OVQA_CUDA_OK(launch_pdl(sdpa_bwd_tc5_kernel, dim3(grid), dim3(kBwdThreads), (size_t)(kBwdSmem), st, mq, mdo, mk, mv, args));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

}  // namespace ovqa
