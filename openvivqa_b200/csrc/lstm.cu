// Recurrent part of the single-layer LSTM of LSTMTextEmbedding (models/modules/text_embeddings.py:234,244).
//
// cuDNN runs the 30 time steps as 30 x (GEMM + cell kernel) forward and 30 x (cell kernel + GEMM) backward: ~180 dependent launches, each
// re-reading W_hh (4H x H) from L2 (profiles/r01l_launches.md, tools/step_audit.py: ~0.8 ms of an 11 ms MCAN step).  Here the whole
// recurrence is ONE kernel per direction:
//   * a thread-block CLUSTER of H/32 CTAs (16 for H = 512; non-portable size) owns 32 samples for all L steps; independent clusters take
//     the other samples, so nothing is grid-wide and there is no co-residency requirement between clusters;
//   * CTA r of the cluster owns hidden units [32 r, 32 r + 32): the 4 x 32 rows of W_hh that produce their gates live in REGISTERS as
//     mma.sync A fragments for the whole kernel (8 warps x 16 rows x H columns of bf16 = H/4 registers per thread) — the weight is read
//     from global memory exactly once per cluster;
//   * per step: gates[128 own rows, 32 samples] = W_slice h_{t-1}^T by mma.sync.m16n8k16 (bf16 in, fp32 accumulate), h_{t-1} (32 x H bf16)
//     read by ldmatrix from the CTA's shared memory; + the pre-computed input projection xg (one tensor-core GEMM over all steps);
//     sigmoid / tanh / cell update in fp32 registers (c_t never leaves them); the new 32 x 32 slice of h_t is PUSHED into the shared
//     memory of all CTAs of the cluster (st.shared::cluster, 16-byte vectors) and a cluster barrier (release / acquire) ends the step.
//     Two h buffers alternate, so one barrier per step suffices.
// Backward (same partition): per step the CTA forms the pre-activation gradients of its own 128 gate rows from dh_t, multiplies them by
// its register-resident slice  W_hh[own rows, :]^T  (a K-split of dh_{t-1} = dgates_t W_hh), and pushes the fp32 partial sums to the CTAs
// that own the units (a reduce-scatter through distributed shared memory); the owners add the 16 partials at the start of the next step.
// The weight / input gradients are ordinary GEMMs over all steps afterwards (dgates is written out in bf16).
//
// h_0 = c_0 = 0 (the reference calls nn.LSTM without an initial state).  Gate order i, f, g, o (PyTorch).
#include <cstdio>

#include "common.cuh"

namespace ovqa {
namespace {

constexpr int kUnits = 32;       // hidden units per CTA
constexpr int kLstmThreads = 256;
constexpr int kStRow = 36;       // fp32 words per row of the 32 x 32 staging tiles (16-byte aligned rows, conflict-free cell writes)
constexpr int kGRow = 136;       // bf16 elements per row of the 32 x 128 gate tiles (272 B: conflict-free ldmatrix)

__device__ __forceinline__ uint32_t lstm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lstm_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t lstm_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// asynchronous 16-byte store into a peer CTA's shared memory that completes (by 16 transaction bytes) on an mbarrier of the SAME peer
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d), "r"(mbar) : "memory");
}
__device__ __forceinline__ void lstm_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lstm_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void lstm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lstm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool lstm_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(lstm_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void lstm_mbar_wait(uint64_t* bar, uint32_t parity) {
    if (lstm_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!lstm_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) { printf("ovqa lstm: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void lstm_cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lstm_smem_u32(dst)), "l"(src) : "memory");
}
// activations on the SFU: tanh.approx (max relative error 2^-11, below the bf16 rounding h and the saved gates go through anyway)
__device__ __forceinline__ float tanh_f(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_f(float x) { return fmaf(0.5f, tanh_f(0.5f * x), 0.5f); }

// ------------------------------------------------------------------------------------------------ forward
// shared memory: h buffers [2][NB samples][H + 8] bf16 | xg tiles [2][NB][148] fp32 | h stage [NB][36] fp32 | c stage [NB][36] fp32 |
// gate stage [NB][136] bf16 | 2 mbarriers
constexpr int kXRow = 148;       // fp32 words per sample row of the staged xg tile: 4 gates x 36 (+4): conflict-free cell reads, 16-byte rows
constexpr int kXGate = 36;
template <int H, int NB> constexpr size_t lstm_fwd_smem() {
    return (size_t)2 * NB * (H + 8) * 2 + (size_t)2 * NB * kXRow * 4 + 2 * (size_t)NB * kStRow * 4 + (size_t)NB * kGRow * 2 + 64;
}

template <int H, int NB>
__global__ void __launch_bounds__(kLstmThreads, 1)
lstm_fwd_kernel(const float* __restrict__ xg, const __nv_bfloat16* __restrict__ whh, const float* __restrict__ bhh, float* __restrict__ y,
                __nv_bfloat16* __restrict__ y16, __nv_bfloat16* __restrict__ hprev16, __nv_bfloat16* __restrict__ gates16, float* __restrict__ c_out,
                int B, int L) {
    constexpr int CS = H / kUnits;            // CTAs per cluster
    constexpr int KS = H / 16;                // k-steps of the recurrent product
    constexpr int HROW = H + 8;               // bf16 elements per staged h row
    constexpr int NT = NB / 8;                // 8-sample n-tiles
    static_assert(NB % 16 == 0, "two n-tiles per ldmatrix.x4");
    extern __shared__ __align__(16) uint8_t lstm_smem[];
    __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(lstm_smem);
    float* xbuf = reinterpret_cast<float*>(lstm_smem + (size_t)2 * NB * HROW * 2);
    float* hst = xbuf + 2 * NB * kXRow;
    float* cst = hst + NB * kStRow;
    __nv_bfloat16* gst = reinterpret_cast<__nv_bfloat16*>(cst + NB * kStRow);
    uint64_t* hbar = reinterpret_cast<uint64_t*>(gst + NB * kGRow);          // [2]: all of h_t has landed in h buffer (t + 1) & 1

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int r = (int)lstm_cluster_rank();
    const int b0 = ((int)blockIdx.x / CS) * NB;
    const int ra = lane / 4, q = lane % 4;
    const int ul = ra % 4, typeB = ra / 4;                    // this thread's unit inside the warp's four; rows: (i, g) for type A, (f, o) for type B
    const int uc = 4 * warp + ul, ug = kUnits * r + uc;       // unit inside the CTA / of the layer
    const int g_lo = typeB, g_hi = 2 + typeB;                 // gates of accumulator rows ra and ra + 8

    // W_hh slice -> A fragments (rows of the warp's tile: gate-major, 4 units each)
    uint32_t afr[KS][4];
    {
        const __nv_bfloat16* w_lo = whh + (size_t)(g_lo * H + ug) * H;
        const __nv_bfloat16* w_hi = whh + (size_t)(g_hi * H + ug) * H;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int k0 = 16 * ks + 2 * q;
            afr[ks][0] = *reinterpret_cast<const uint32_t*>(w_lo + k0);
            afr[ks][1] = *reinterpret_cast<const uint32_t*>(w_hi + k0);
            afr[ks][2] = *reinterpret_cast<const uint32_t*>(w_lo + k0 + 8);
            afr[ks][3] = *reinterpret_cast<const uint32_t*>(w_hi + k0 + 8);
        }
    }
    const float b_lo = bhh ? __ldg(bhh + g_lo * H + ug) : 0.f, b_hi = bhh ? __ldg(bhh + g_hi * H + ug) : 0.f;
    // the CTA's slice of xg for one step -> shared memory (cp.async, 16 bytes = 4 units of one gate of one sample); samples beyond B: zeros
    auto load_xg = [&](int t) {
        float* xb = xbuf + (size_t)(t & 1) * NB * kXRow;
        for (int idx = tid; idx < NB * 32; idx += kLstmThreads) {
            const int s = idx / 32, gate = (idx % 32) / 8, c4 = (idx % 8) * 4;
            float* dst = xb + s * kXRow + gate * kXGate + c4;
            if (b0 + s < B) lstm_cp_async16(dst, xg + ((size_t)(b0 + s) * L + t) * (4 * H) + gate * H + kUnits * r + c4);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_xg(0);
    // h_{-1} = 0
    for (int i = tid; i < NB * HROW / 2; i += kLstmThreads) reinterpret_cast<uint32_t*>(hbuf)[i] = 0u;
    if (tid == 0) {
        lstm_mbar_init(hbar, 1);
        lstm_mbar_init(hbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_arrive();
    cluster_wait();                                            // every CTA of the cluster runs, has cleared its buffer and initialised its barriers

    float cst_reg[NT];                                         // c_{t-1} of the thread's cells (one per n-tile)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) cst_reg[nt] = 0.f;
    // ldmatrix row address of this lane: matrix m = lane / 8 -> samples (m / 2) * 8 + lane % 8 (+16 per further instruction), k offset (m % 2) * 8
    const uint32_t ld_off = (uint32_t)(((lane / 16) * 8 + lane % 8) * HROW + ((lane / 8) % 2) * 8) * 2u;
    const uint32_t hbuf_u32 = lstm_smem_u32(hbuf), hbar_u32 = lstm_smem_u32(hbar);
    uint32_t ph[2] = {0u, 0u};

    for (int t = 0; t < L; ++t) {
        if (t > 0) { lstm_mbar_wait(hbar + (t & 1), ph[t & 1]); ph[t & 1] ^= 1u; }
        const uint32_t hb = hbuf_u32 + (uint32_t)(t & 1) * (uint32_t)(NB * HROW * 2) + ld_off;
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[nt][c] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) {
                uint32_t b0r, b1r, b2r, b3r;
                ldmatrix_x4(hb + (uint32_t)(16 * np * HROW * 2) + (uint32_t)ks * 32u, b0r, b1r, b2r, b3r);
                mma_bf16_16816(acc[2 * np], afr[ks], b0r, b1r);
                mma_bf16_16816(acc[2 * np + 1], afr[ks], b2r, b3r);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // this step's xg tile (issued a step ago)
        __syncthreads();                                       // ... from every thread; and the previous step's copy-out has finished reading the staging tiles
        const float* xb = xbuf + (size_t)(t & 1) * NB * kXRow + g_lo * kXGate + uc;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const float* xs = xb + (8 * nt + 2 * q) * kXRow;
            const float lo0 = sigmoid_f(acc[nt][0] + xs[0] + b_lo), lo1 = sigmoid_f(acc[nt][1] + xs[kXRow] + b_lo);
            const float p2 = acc[nt][2] + xs[2 * kXGate] + b_hi, p3 = acc[nt][3] + xs[kXRow + 2 * kXGate] + b_hi;
            const float hi0 = typeB ? sigmoid_f(p2) : tanh_f(p2), hi1 = typeB ? sigmoid_f(p3) : tanh_f(p3);
            // type A keeps the even sample and needs (f, o) of it; type B keeps the odd one and needs (i, g)
            const float r_lo = __shfl_xor_sync(0xffffffffu, typeB ? lo0 : lo1, 16);
            const float r_hi = __shfl_xor_sync(0xffffffffu, typeB ? hi0 : hi1, 16);
            const float gi = typeB ? r_lo : lo0, gf = typeB ? lo1 : r_lo, gg = typeB ? r_hi : hi0, go = typeB ? hi1 : r_hi;
            const float cn = fmaf(gf, cst_reg[nt], gi * gg);
            cst_reg[nt] = cn;
            const float hn = go * tanh_f(cn);
            const int s = 8 * nt + 2 * q + typeB;
            hst[s * kStRow + uc] = hn;
            cst[s * kStRow + uc] = cn;
            __nv_bfloat16* gp = gst + s * kGRow + uc;
            gp[0] = __float2bfloat16_rn(gi); gp[kUnits] = __float2bfloat16_rn(gf); gp[2 * kUnits] = __float2bfloat16_rn(gg); gp[3 * kUnits] = __float2bfloat16_rn(go);
        }
        __syncthreads();
        if (t + 1 < L) {
            // push the CTA's NB x 32 slice of h_t (bf16) into every CTA's next h buffer: 16-byte asynchronous stores that complete on the
            // DESTINATION's mbarrier (no fence, no cluster barrier: the barrier's transaction count is the hand-over)
            if (tid == 0) lstm_mbar_expect_tx(hbar + ((t + 1) & 1), (uint32_t)(NB * H * 2));
            const uint32_t nxt = hbuf_u32 + (uint32_t)((t + 1) & 1) * (uint32_t)(NB * HROW * 2) + (uint32_t)(kUnits * r) * 2u;
            const uint32_t bar = hbar_u32 + (uint32_t)((t + 1) & 1) * 8u;
            // thread = one 16-byte chunk (sample, 8 units) of the slice, read and packed once, stored to its share of the destinations
            constexpr int CH = NB * 4, GROUPS = kLstmThreads / CH;      // NB = 32: two groups of 128 threads, NB = 48: one group of 192
            static_assert(GROUPS >= 1, "one thread per chunk");
            const int ch = tid % CH, grp = tid / CH;
            if (grp < GROUPS) {
                const int s = ch / 4, u8 = (ch % 4) * 8;
                const float4 v0 = *reinterpret_cast<const float4*>(hst + s * kStRow + u8), v1 = *reinterpret_cast<const float4*>(hst + s * kStRow + u8 + 4);
                const uint32_t p0 = pack_bf16x2(v0.x, v0.y), p1 = pack_bf16x2(v0.z, v0.w), p2 = pack_bf16x2(v1.x, v1.y), p3 = pack_bf16x2(v1.z, v1.w);
                const uint32_t dst = nxt + (uint32_t)(s * HROW + u8) * 2u;
#pragma unroll
                for (int d = grp; d < CS; d += GROUPS) st_async_v4(lstm_mapa(dst, (uint32_t)d), p0, p1, p2, p3, lstm_mapa(bar, (uint32_t)d));
            }
            load_xg(t + 1);
        }
        // copy-out
        for (int idx = tid; idx < NB * 8; idx += kLstmThreads) {
            const int s = idx / 8, u4 = (idx % 8) * 4;
            if (b0 + s < B) {
                const size_t row = (size_t)(b0 + s) * L + t;
                const float4 hv = *reinterpret_cast<const float4*>(hst + s * kStRow + u4);
                *reinterpret_cast<float4*>(y + row * H + kUnits * r + u4) = hv;
                const uint2 h16 = make_uint2(pack_bf16x2(hv.x, hv.y), pack_bf16x2(hv.z, hv.w));
                if (y16) *reinterpret_cast<uint2*>(y16 + row * H + kUnits * r + u4) = h16;
                if (hprev16 && t + 1 < L) *reinterpret_cast<uint2*>(hprev16 + (row + 1) * H + kUnits * r + u4) = h16;
                if (hprev16 && t == 0) *reinterpret_cast<uint2*>(hprev16 + row * H + kUnits * r + u4) = make_uint2(0u, 0u);
                if (c_out) *reinterpret_cast<float4*>(c_out + row * H + kUnits * r + u4) = *reinterpret_cast<const float4*>(cst + s * kStRow + u4);
            }
        }
        if (gates16) {
            for (int ch = tid; ch < NB * 16; ch += kLstmThreads) {
                const int gs = ch / 16, part = ch % 16, gate = part / 4, u8 = (part % 4) * 8;
                if (b0 + gs < B)
                    *reinterpret_cast<uint4*>(gates16 + ((size_t)(b0 + gs) * L + t) * (4 * H) + gate * H + kUnits * r + u8) =
                        *reinterpret_cast<const uint4*>(gst + gs * kGRow + gate * kUnits + u8);
            }
        }
    }
    cluster_arrive();                                          // nobody leaves while a peer's stores into it may be in flight
    cluster_wait();
}

// ------------------------------------------------------------------------------------------------ backward
// shared memory: partial sums [2][CS sources][32 units][NB samples] fp32 | own pre-activation gradients [NB samples][136] bf16 | 2 mbarriers
template <int H, int NB> constexpr size_t lstm_bwd_smem() { return (size_t)2 * (H / kUnits) * kUnits * NB * 4 + (size_t)NB * kGRow * 2 + 64; }

template <int H, int NB>
__global__ void __launch_bounds__(kLstmThreads, 1)
lstm_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ whh, const __nv_bfloat16* __restrict__ gates16,
                const float* __restrict__ c_all, __nv_bfloat16* __restrict__ dgates16, int B, int L) {
    constexpr int CS = H / kUnits;
    constexpr int MT = H / 128;               // 16-unit m-tiles per warp (the warp owns H / 8 units of the K-split product)
    constexpr int KS = 4 * kUnits / 16;       // k-steps over the CTA's 128 gate rows
    constexpr int NP = (NB + 31) / 32;        // element-wise passes: a thread owns (unit, 4 consecutive samples) per pass of 32 samples
    extern __shared__ __align__(16) uint8_t lstm_smem[];
    float* part = reinterpret_cast<float*>(lstm_smem);                          // [2][CS][32 u][NB s]
    __nv_bfloat16* dgb = reinterpret_cast<__nv_bfloat16*>(part + (size_t)2 * CS * kUnits * NB);   // [NB s][136]
    uint64_t* pbar = reinterpret_cast<uint64_t*>(dgb + NB * kGRow);             // [2]: every CTA's partial sums for step t have landed in part[t & 1]

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int r = (int)lstm_cluster_rank();
    const int b0 = ((int)blockIdx.x / CS) * NB;
    const int ra = lane / 4, q = lane % 4;

    // A[m][k] = W_hh[row(k)][m]:  m = hidden unit of the layer (the warp's H/8), k = the CTA's gate rows (gate-major, 32 units each)
    uint32_t afr[MT][KS][4];
    {
        auto wrow = [&](int k) { return (size_t)((k / kUnits) * H + kUnits * r + k % kUnits) * H; };
        auto pair = [&](int k, int m) {
            const uint32_t lo = *reinterpret_cast<const unsigned short*>(whh + wrow(k) + m), hi = *reinterpret_cast<const unsigned short*>(whh + wrow(k + 1) + m);
            return lo | (hi << 16);
        };
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const int m = (H / 8) * warp + 16 * mt + ra, k0 = 16 * ks + 2 * q;
                afr[mt][ks][0] = pair(k0, m);
                afr[mt][ks][1] = pair(k0, m + 8);
                afr[mt][ks][2] = pair(k0 + 8, m);
                afr[mt][ks][3] = pair(k0 + 8, m + 8);
            }
    }
    if (tid == 0) {
        lstm_mbar_init(pbar, 1);
        lstm_mbar_init(pbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_arrive();
    cluster_wait();                                            // every CTA of the cluster runs and has initialised its barriers

    // element-wise phase: thread = (unit, 4 consecutive samples) per pass
    const int eu = tid / 8, es = (tid % 8) * 4, ug = kUnits * r + eu;
    float dc_next[NP][4];
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) dc_next[p][j] = 0.f;
    const uint32_t ld_off = (uint32_t)(((lane / 16) * 8 + lane % 8) * kGRow + ((lane / 8) % 2) * 8) * 2u;
    const uint32_t dgb_u32 = lstm_smem_u32(dgb), part_u32 = lstm_smem_u32(part), pbar_u32 = lstm_smem_u32(pbar);
    uint32_t ph[2] = {0u, 0u};

    for (int t = L - 1; t >= 0; --t) {
        // ---- dh_t = dy_t + sum over the cluster of the partial products pushed during step t + 1
        if (t + 1 < L) { lstm_mbar_wait(pbar + (t & 1), ph[t & 1]); ph[t & 1] ^= 1u; }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int sl = 32 * p + es;                          // first of this thread's 4 samples (local)
            if (sl < NB) {
                float rec[4] = {0.f, 0.f, 0.f, 0.f};
                if (t + 1 < L) {
                    const float* pb = part + (size_t)(t & 1) * CS * kUnits * NB + eu * NB + sl;
#pragma unroll
                    for (int src = 0; src < CS; ++src) {
                        const float4 v = *reinterpret_cast<const float4*>(pb + (size_t)src * kUnits * NB);
                        rec[0] += v.x; rec[1] += v.y; rec[2] += v.z; rec[3] += v.w;
                    }
                }
                // all of the four cells' operands are requested before the first is used (samples beyond B: clamped, result discarded)
                float dyv[4], ctv[4], cpv[4];
                unsigned short gr[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int s = min(b0 + sl + j, B - 1);
                    const size_t row = (size_t)s * L + t;
                    dyv[j] = __ldg(dy + row * H + ug);
                    ctv[j] = __ldg(c_all + row * H + ug);
                    cpv[j] = t > 0 ? __ldg(c_all + (row - 1) * H + ug) : 0.f;
                    const unsigned short* gp = reinterpret_cast<const unsigned short*>(gates16) + row * (4 * H) + ug;
#pragma unroll
                    for (int g = 0; g < 4; ++g) gr[j][g] = __ldg(gp + g * H);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int s = b0 + sl + j;
                    float dai = 0.f, daf = 0.f, dag = 0.f, dao = 0.f;
                    if (s < B) {
                        const size_t row = (size_t)s * L + t;
                        const float dh = dyv[j] + rec[j];
                        const float gi = __uint_as_float((uint32_t)gr[j][0] << 16), gf = __uint_as_float((uint32_t)gr[j][1] << 16),
                                    gg = __uint_as_float((uint32_t)gr[j][2] << 16), go = __uint_as_float((uint32_t)gr[j][3] << 16);
                        const float ct = ctv[j], cp = cpv[j];
                        const float tc = tanh_f(ct);
                        const float dc = fmaf(dh * go, 1.0f - tc * tc, dc_next[p][j]);
                        dc_next[p][j] = dc * gf;
                        dai = dc * gg * gi * (1.0f - gi);
                        daf = dc * cp * gf * (1.0f - gf);
                        dag = dc * gi * (1.0f - gg * gg);
                        dao = dh * tc * go * (1.0f - go);
                        __nv_bfloat16* op = dgates16 + row * (4 * H) + ug;
                        op[0] = __float2bfloat16_rn(dai); op[H] = __float2bfloat16_rn(daf); op[2 * H] = __float2bfloat16_rn(dag); op[3 * H] = __float2bfloat16_rn(dao);
                    }
                    __nv_bfloat16* bp = dgb + (sl + j) * kGRow + eu;
                    bp[0] = __float2bfloat16_rn(dai); bp[kUnits] = __float2bfloat16_rn(daf); bp[2 * kUnits] = __float2bfloat16_rn(dag); bp[3 * kUnits] = __float2bfloat16_rn(dao);
                }
            }
        }
        if (t == 0) break;
        __syncthreads();
        // ---- partial dh_{t-1}[all units, NB samples] over the CTA's 128 gate rows, pushed to the owners of the units (16-byte asynchronous
        // stores completing on the owner's mbarrier: the transaction count is the hand-over, no fence and no cluster barrier)
        if (tid == 0) lstm_mbar_expect_tx(pbar + ((t - 1) & 1), (uint32_t)(CS * kUnits * NB * 4));
        const uint32_t dst_buf = part_u32 + (uint32_t)((t - 1) & 1) * (uint32_t)(CS * kUnits * NB * 4) + (uint32_t)r * (uint32_t)(kUnits * NB * 4);
        const uint32_t bar = pbar_u32 + (uint32_t)((t - 1) & 1) * 8u;
        const bool odd = q & 1;
#pragma unroll 1
        for (int nh = 0; nh < NB / 16; ++nh) {
            float acc[MT][2][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[mt][nt][c] = 0.f;
            uint32_t bfr[2][4];
            const uint32_t bsrc = dgb_u32 + ld_off + (uint32_t)(16 * nh * kGRow * 2);
            ldmatrix_x4(bsrc, bfr[0][0], bfr[0][1], bfr[0][2], bfr[0][3]);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                if (ks + 1 < KS) ldmatrix_x4(bsrc + (uint32_t)(ks + 1) * 32u, bfr[(ks + 1) & 1][0], bfr[(ks + 1) & 1][1], bfr[(ks + 1) & 1][2], bfr[(ks + 1) & 1][3]);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma_bf16_16816(acc[mt][0], afr[mt][ks], bfr[ks & 1][0], bfr[ks & 1][1]);
                    mma_bf16_16816(acc[mt][1], afr[mt][ks], bfr[ks & 1][2], bfr[ks & 1][3]);
                }
            }
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int m = (H / 8) * warp + 16 * mt + ra;              // unit of accumulator rows ra (m + 8: rows ra + 8); both live in CTA m / 32
                const uint32_t dcta = (uint32_t)(m / kUnits);
                const uint32_t base = lstm_mapa(dst_buf, dcta), rbar = lstm_mapa(bar, dcta);
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    // lanes q, q ^ 1 hold adjacent sample pairs of the same two units: the even lane takes four samples of unit m, the odd lane of unit m + 8
                    const float x0 = __shfl_xor_sync(0xffffffffu, odd ? acc[mt][nt][0] : acc[mt][nt][2], 1);
                    const float x1 = __shfl_xor_sync(0xffffffffu, odd ? acc[mt][nt][1] : acc[mt][nt][3], 1);
                    const int s4 = 16 * nh + 8 * nt + 2 * (q & 2);       // first of four consecutive samples (q even: 2q, q odd: 2(q - 1))
                    const int unit = (odd ? m + 8 : m) % kUnits;
                    const float v0 = odd ? x0 : acc[mt][nt][0], v1 = odd ? x1 : acc[mt][nt][1], v2 = odd ? acc[mt][nt][2] : x0, v3 = odd ? acc[mt][nt][3] : x1;
                    st_async_v4(base + (uint32_t)((unit * NB + s4) * 4), __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3), rbar);
                }
            }
        }
        // pull the next step's operands towards L2 while the partial sums travel
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int sl = 32 * p + es;
            if (sl < NB) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (b0 + sl + j < B) {
                        const size_t row = (size_t)(b0 + sl + j) * L + (t - 1);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(dy + row * H + ug));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(c_all + row * H + ug));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(gates16 + row * (4 * H) + ug));
                    }
                }
            }
        }
    }
    cluster_arrive();                                          // nobody leaves while a peer's stores into it may be in flight
    cluster_wait();
}

template <typename K>
int lstm_prepare(K kern, OncePerDevice& latch, size_t smem, int cs, int grid, cudaLaunchConfig_t& lc, cudaLaunchAttribute* at, cudaStream_t st) {
    lc = cudaLaunchConfig_t{};
    lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(kLstmThreads); lc.dynamicSmemBytes = smem; lc.stream = st;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    // once per device and kernel: opt into the shared memory and (for 16 CTAs) the non-portable cluster size, and check that one cluster fits
    if (latch.pending()) {
        OVQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (cs > 8) OVQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, kern, &lc) != cudaSuccess || nc < 1) { (void)cudaGetLastError(); return OVQA_ERR_BAD_SHAPE; }
        if (getenv("OVQA_DEBUG")) fprintf(stderr, "ovqa lstm kernel: %d co-resident clusters of %d CTAs\n", nc, cs);
        latch.mark();
    }
    return OVQA_OK;
}

// Samples per cluster: 32, or 48 when that saves a wave (B200: 7 clusters of 16 CTAs are co-resident, so 256 samples are 8 clusters = two
// waves of 32-sample clusters but one wave of six 48-sample clusters; a 48-sample step costs ~1.35x a 32-sample one).
inline int lstm_pick_nb(int B, int cs) {
    const int co = cs >= 16 ? 7 : cs >= 8 ? 16 : cs >= 4 ? 32 : 64;             // co-resident clusters (measured for 16; generous for the smaller sizes)
    auto waves = [&](int nb) { return (((B + nb - 1) / nb) + co - 1) / co; };
    return 135 * waves(48) < 100 * waves(32) ? 48 : 32;
}

template <int H, int NB>
int lstm_fwd_launch(const float* xg, const void* whh, const float* bhh, float* y, void* y16, void* hprev16, void* gates16, float* c_out, int B, int L,
                    cudaStream_t st) {
    constexpr int CS = H / kUnits;
    const int grid = (B + NB - 1) / NB * CS;
    cudaLaunchConfig_t lc;
    cudaLaunchAttribute at[1];
    static OncePerDevice latch;
    const int rc = lstm_prepare(lstm_fwd_kernel<H, NB>, latch, lstm_fwd_smem<H, NB>(), CS, grid, lc, at, st);
    if (rc) return rc;
    OVQA_CUDA_OK(cudaLaunchKernelEx(&lc, lstm_fwd_kernel<H, NB>, xg, (const __nv_bfloat16*)whh, bhh, y, (__nv_bfloat16*)y16, (__nv_bfloat16*)hprev16,
                                    (__nv_bfloat16*)gates16, c_out, B, L));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

template <int H, int NB>
int lstm_bwd_launch(const float* dy, const void* whh, const void* gates16, const float* c_all, void* dgates16, int B, int L, cudaStream_t st) {
    constexpr int CS = H / kUnits;
    const int grid = (B + NB - 1) / NB * CS;
    cudaLaunchConfig_t lc;
    cudaLaunchAttribute at[1];
    static OncePerDevice latch;
    const int rc = lstm_prepare(lstm_bwd_kernel<H, NB>, latch, lstm_bwd_smem<H, NB>(), CS, grid, lc, at, st);
    if (rc) return rc;
    OVQA_CUDA_OK(cudaLaunchKernelEx(&lc, lstm_bwd_kernel<H, NB>, dy, (const __nv_bfloat16*)whh, (const __nv_bfloat16*)gates16, c_all, (__nv_bfloat16*)dgates16, B, L));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

inline bool lstm_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace ovqa

using namespace ovqa;

extern "C" int ovqa_lstm_supported(int64_t H) { return H == 128 || H == 256 || H == 512; }

extern "C" int ovqa_lstm_fwd(const float* xg, const void* w_hh_bf16, const float* b_hh, float* y, void* y_bf16, void* hprev_bf16, void* gates_bf16,
                             float* c_out, int64_t B, int64_t L, int64_t H, void* stream) {
    if (!xg || !w_hh_bf16 || !y || B < 0 || L < 0) return OVQA_ERR_BAD_ARG;
    if (!ovqa_lstm_supported(H) || B > (1 << 24) || L > (1 << 20)) return OVQA_ERR_BAD_SHAPE;
    if (!lstm_al16(xg) || !lstm_al16(w_hh_bf16) || !lstm_al16(y) || !lstm_al16(y_bf16) || !lstm_al16(hprev_bf16) || !lstm_al16(gates_bf16) || !lstm_al16(c_out))
        return OVQA_ERR_ALIGNMENT;
    if (B == 0 || L == 0) return OVQA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide = lstm_pick_nb((int)B, (int)H / kUnits) == 48;
#define OVQA_LSTM_GO(H_) (wide ? lstm_fwd_launch<H_, 48>(xg, w_hh_bf16, b_hh, y, y_bf16, hprev_bf16, gates_bf16, c_out, (int)B, (int)L, st) \
                               : lstm_fwd_launch<H_, 32>(xg, w_hh_bf16, b_hh, y, y_bf16, hprev_bf16, gates_bf16, c_out, (int)B, (int)L, st))
    switch ((int)H) {
        case 128: return OVQA_LSTM_GO(128);
        case 256: return OVQA_LSTM_GO(256);
        default: return OVQA_LSTM_GO(512);
    }
#undef OVQA_LSTM_GO
}

extern "C" int ovqa_lstm_bwd(const float* dy, const void* w_hh_bf16, const void* gates_bf16, const float* c_all, void* dgates_bf16, int64_t B, int64_t L,
                             int64_t H, void* stream) {
    if (!dy || !w_hh_bf16 || !gates_bf16 || !c_all || !dgates_bf16 || B < 0 || L < 0) return OVQA_ERR_BAD_ARG;
    if (!ovqa_lstm_supported(H) || B > (1 << 24) || L > (1 << 20)) return OVQA_ERR_BAD_SHAPE;
    if (B == 0 || L == 0) return OVQA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide = lstm_pick_nb((int)B, (int)H / kUnits) == 48;
#define OVQA_LSTM_GO(H_) (wide ? lstm_bwd_launch<H_, 48>(dy, w_hh_bf16, gates_bf16, c_all, dgates_bf16, (int)B, (int)L, st) \
                               : lstm_bwd_launch<H_, 32>(dy, w_hh_bf16, gates_bf16, c_all, dgates_bf16, (int)B, (int)L, st))
    switch ((int)H) {
        case 128: return OVQA_LSTM_GO(128);
        case 256: return OVQA_LSTM_GO(256);
        default: return OVQA_LSTM_GO(512);
    }
#undef OVQA_LSTM_GO
}
