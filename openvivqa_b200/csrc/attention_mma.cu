// Attention core on bf16 tensor-core tiles for the short-sequence regime (Lq ~ 30-182, Lk ~ 30-600):
//   forward  : out = softmax(q k^T * scale + mask) v, flash-style over 64-key chunks staged in shared memory
//   backward : dq kernel (CTA = query tile, streams keys) + dk/dv kernel (CTA = key tile, streams queries),
//              probabilities recomputed from the saved row log-sum-exp.
// One warp owns a 16-row MMA tile (mma.sync.m16n8k16 bf16 -> fp32); the softmax lives in registers with quad
// shuffles; masks are added in fp32 before the running max (reference order).  The core is HBM-bound
// (~23 FLOP/B at Lq=30, Lk=100): a 128-row tcgen05 tile would be 3/4 padding here, so the dense projections get
// tcgen05 (gemm_tcgen05.cu) and the core gets warp-level tiles, as BASELINE.json's north_star prescribes.
#include "attention_simt.cuh"

namespace ovqa {

namespace {

constexpr int KCH = 64;    // keys (or queries, in the dk/dv kernel) per streamed chunk
constexpr int PAD = 8;     // bf16 elements of row padding: row pitch (D+8)*2 B -> conflict-free ldmatrix

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// Stage `rows` x D bf16 (global row pitch ld elements) into smem with pitch D+PAD; rows >= valid are zero-filled.
template <int D>
__device__ __forceinline__ void stage_rows(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int rows, int valid, int tid, int nthreads) {
    constexpr int CPR = D / 8;  // 16-byte chunks per row
    for (int i = tid; i < rows * CPR; i += nthreads) {
        const int r = i / CPR, c = i % CPR;
        __nv_bfloat16* d = dst + r * (D + PAD) + c * 8;
        if (r < valid) cp_async16(d, src + (int64_t)r * ld + c * 8);
        else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
    }
}

// A-operand fragments of a 16 x D row-major smem tile (pitch D+PAD), rows [row0, row0+16)
template <int D>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[D / 16][4], const __nv_bfloat16* tile, int row0, int lane) {
    const int r = row0 + (lane % 8) + 8 * ((lane / 8) % 2), c = 8 * (lane / 16);
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) ldsm_x4(smem_u32(tile + r * (D + PAD) + ks * 16 + c), a[ks][0], a[ks][1], a[ks][2], a[ks][3]);
}

// acc[nt] (16 x 8 each, nt over CH "column" rows of `rows_tile`) += A(16 x D) . rows_tile^T      (S = Q K^T form)
template <int D, int CH = KCH>
__device__ __forceinline__ void mma_a_rowsT(float (&acc)[CH / 8][4], const uint32_t (&a)[D / 16][4], const __nv_bfloat16* rows_tile, int lane) {
    const int rr = (lane % 8) + 8 * (lane / 16), cc = 8 * ((lane / 8) % 2);
#pragma unroll
    for (int np = 0; np < CH / 16; ++np) {
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(smem_u32(rows_tile + (np * 16 + rr) * (D + PAD) + ks * 16 + cc), b0, b1, b2, b3);
            mma_bf16(acc[2 * np], a[ks], b0, b1);
            mma_bf16(acc[2 * np + 1], a[ks], b2, b3);
        }
    }
}

// out[nt] (16 x 8 each, nt over D columns) += P(16 x CH, from accumulator-layout registers) . rows_tile(CH x D)
template <int D, int CH = KCH>
__device__ __forceinline__ void mma_p_rows(float (&out)[D / 8][4], const float (&p)[CH / 8][4], const __nv_bfloat16* rows_tile, int lane) {
    const int rr = (lane % 8) + 8 * ((lane / 8) % 2), cc = 8 * (lane / 16);
#pragma unroll
    for (int ks = 0; ks < CH / 16; ++ks) {
        uint32_t a[4];
        a[0] = pack_bf16x2(p[2 * ks][0], p[2 * ks][1]);
        a[1] = pack_bf16x2(p[2 * ks][2], p[2 * ks][3]);
        a[2] = pack_bf16x2(p[2 * ks + 1][0], p[2 * ks + 1][1]);
        a[3] = pack_bf16x2(p[2 * ks + 1][2], p[2 * ks + 1][3]);
#pragma unroll
        for (int np = 0; np < D / 16; ++np) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(smem_u32(rows_tile + (ks * 16 + rr) * (D + PAD) + np * 16 + cc), b0, b1, b2, b3);
            mma_bf16(out[2 * np], a, b0, b1);
            mma_bf16(out[2 * np + 1], a, b2, b3);
        }
    }
}

__device__ __forceinline__ float quad_max(float v) { v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1)); return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2)); }
__device__ __forceinline__ float quad_sum(float v) { v += __shfl_xor_sync(0xffffffffu, v, 1); return v + __shfl_xor_sync(0xffffffffu, v, 2); }

// ------------------------------------------------------------------------------------------------ forward
// FULL = the head's whole K and V (Lk rows padded to 64) fit next to the Q tile: everything is requested at once with
// cp.async, ONE wait + barrier, and the chunk loop then runs with no further synchronisation (one DRAM round trip
// per CTA instead of one per chunk — these CTAs are latency-bound, not bandwidth-bound).  Otherwise K/V stream
// through one 64-row chunk buffer.  When the mask is a per-key row (stride over queries == 0: padding masks) or
// absent it is staged as one float row with -inf beyond Lk, so the key-range predicate disappears as well.
template <int D, int NWARPS, bool FULL>
__global__ void __launch_bounds__(NWARPS * 32) sdpa_fwd_mma_kernel(SdpaArgs a) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int P = D + PAD;
    const int Lq = (int)a.Lq, Lk = (int)a.Lk;
    const int Lkp = (Lk + KCH - 1) / KCH * KCH;
    const int kv_rows = FULL ? Lkp : KCH;
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [NWARPS*16][P]
    __nv_bfloat16* sK = sQ + NWARPS * 16 * P;                               // [kv_rows][P]
    __nv_bfloat16* sV = sK + kv_rows * P;                                   // [kv_rows][P]
    float* sMask = reinterpret_cast<float*>(sV + kv_rows * P);              // [Lkp]
    const int b = blockIdx.z, hd = blockIdx.y, q0 = blockIdx.x * NWARPS * 16;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(a.q) + ((int64_t)b * Lq + q0) * a.ldq + (int64_t)hd * D;
    const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(a.k) + (int64_t)b * Lk * a.ldk + (int64_t)hd * D;
    const __nv_bfloat16* V = reinterpret_cast<const __nv_bfloat16*>(a.v) + (int64_t)b * Lk * a.ldv + (int64_t)hd * D;

    stage_rows<D>(sQ, Q, a.ldq, NWARPS * 16, min(NWARPS * 16, Lq - q0), threadIdx.x, NWARPS * 32);
    if constexpr (FULL) {
        stage_rows<D>(sK, K, a.ldk, Lkp, Lk, threadIdx.x, NWARPS * 32);
        stage_rows<D>(sV, V, a.ldv, Lkp, Lk, threadIdx.x, NWARPS * 32);
    }
    const float* mbase = a.mask.ptr ? a.mask.ptr + (int64_t)b * a.mask.sb + (int64_t)hd * a.mask.sh : nullptr;
    const bool row_mask = mbase == nullptr || a.mask.sq == 0;               // one mask row serves every query
    if (row_mask)
        for (int j = threadIdx.x; j < Lkp; j += NWARPS * 32) sMask[j] = j < Lk ? (mbase ? __ldg(mbase + j) : 0.f) : -INFINITY;
    cp_async_wait_all();
    __syncthreads();
    uint32_t qa[D / 16][4];
    load_a_frags<D>(qa, sQ, warp * 16, lane);

    const int row_lo = q0 + warp * 16 + g, row_hi = row_lo + 8;      // the two query rows this thread holds
    const float* mrow_lo = nullptr;
    const float* mrow_hi = nullptr;
    if (!row_mask) {
        mrow_lo = mbase + (int64_t)min(row_lo, Lq - 1) * a.mask.sq;
        mrow_hi = mbase + (int64_t)min(row_hi, Lq - 1) * a.mask.sq;
    }
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;

    for (int k0 = 0; k0 < Lk; k0 += KCH) {
        const __nv_bfloat16* kc = sK;
        const __nv_bfloat16* vc = sV;
        if constexpr (FULL) {
            kc = sK + k0 * P;
            vc = sV + k0 * P;
        } else {
            const int nk = min(KCH, Lk - k0);
            __syncthreads();                               // previous chunk fully consumed
            stage_rows<D>(sK, K + (int64_t)k0 * a.ldk, a.ldk, KCH, nk, threadIdx.x, NWARPS * 32);
            stage_rows<D>(sV, V + (int64_t)k0 * a.ldv, a.ldv, KCH, nk, threadIdx.x, NWARPS * 32);
            cp_async_wait_all();
            __syncthreads();
        }
        float s[KCH / 8][4];
#pragma unroll
        for (int i = 0; i < KCH / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        mma_a_rowsT<D>(s, qa, kc, lane);
        // scale, additive mask (fp32, before the max), key-range predicate
        float cmax_lo = -INFINITY, cmax_hi = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < KCH / 8; ++nt) {
            const int j0 = k0 + nt * 8 + 2 * t;
            float mk_lo[2], mk_hi[2];
            if (row_mask) {
                const float2 m2 = *reinterpret_cast<const float2*>(sMask + j0);
                mk_lo[0] = mk_hi[0] = m2.x; mk_lo[1] = mk_hi[1] = m2.y;
            } else {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const bool in = j0 + e < Lk;
                    mk_lo[e] = in ? __ldg(mrow_lo + j0 + e) : -INFINITY;
                    mk_hi[e] = in ? __ldg(mrow_hi + j0 + e) : -INFINITY;
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float v_lo = fmaf(s[nt][e], a.scale, mk_lo[e]), v_hi = fmaf(s[nt][2 + e], a.scale, mk_hi[e]);
                s[nt][e] = v_lo; s[nt][2 + e] = v_hi;
                cmax_lo = fmaxf(cmax_lo, v_lo); cmax_hi = fmaxf(cmax_hi, v_hi);
            }
        }
        const float mn_lo = fmaxf(m_lo, quad_max(cmax_lo)), mn_hi = fmaxf(m_hi, quad_max(cmax_hi));
        const float corr_lo = __expf(m_lo - mn_lo), corr_hi = __expf(m_hi - mn_hi);   // exp(-inf) = 0 on the first chunk
        m_lo = mn_lo; m_hi = mn_hi;
        float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
        for (int nt = 0; nt < KCH / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float p_lo = __expf(s[nt][e] - m_lo), p_hi = __expf(s[nt][2 + e] - m_hi);
                s[nt][e] = p_lo; s[nt][2 + e] = p_hi;
                sum_lo += p_lo; sum_hi += p_hi;
            }
        }
        l_lo = l_lo * corr_lo + quad_sum(sum_lo);
        l_hi = l_hi * corr_hi + quad_sum(sum_hi);
#pragma unroll
        for (int i = 0; i < D / 8; ++i) { o[i][0] *= corr_lo; o[i][1] *= corr_lo; o[i][2] *= corr_hi; o[i][3] *= corr_hi; }
        mma_p_rows<D>(o, s, vc, lane);
    }
    const float inv_lo = 1.f / l_lo, inv_hi = 1.f / l_hi;
    __nv_bfloat16* O = reinterpret_cast<__nv_bfloat16*>(a.out) + (int64_t)b * Lq * a.ldo + (int64_t)hd * D;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        const int c = nt * 8 + 2 * t;
        if (row_lo < Lq) *reinterpret_cast<uint32_t*>(O + (int64_t)row_lo * a.ldo + c) = pack_bf16x2(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
        if (row_hi < Lq) *reinterpret_cast<uint32_t*>(O + (int64_t)row_hi * a.ldo + c) = pack_bf16x2(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
    }
    if (a.lse && t == 0) {
        float* lse = a.lse + ((int64_t)b * a.h + hd) * Lq;
        if (row_lo < Lq) lse[row_lo] = m_lo + logf(l_lo);
        if (row_hi < Lq) lse[row_hi] = m_hi + logf(l_hi);
    }
}

// ------------------------------------------------------------------------------------------------ backward
// One CTA per (batch, head): Q, dO (Lq rows) and K, V (Lk rows) of the head are staged ONCE in shared memory
// (zero-padded to multiples of BCH rows), delta_i = dO_i . O_i and lse_i go to shared memory too, and then the
// warps pull independent work items with no further block synchronisation:
//   query-tile item (16 query rows):  S = Q K^T, P = exp(S*scale + mask - lse), dP = dO V^T, dS = P (dP - delta),
//                                     dQ = scale * dS K                       (registers -> global)
//   key-tile item   (16 key rows)  :  S^T = K Q^T, P^T, dP^T = V dO^T, dS^T,  dV = P^T dO,  dK = scale * dS^T Q
// Computing S^T directly (instead of transposing P) keeps every product in the "A from registers, B rows from
// smem" form of the forward, so the same two MMA helpers serve all seven contractions.
constexpr int BCH = 32;    // rows of the streamed operand per inner step

struct SdpaBwdMmaArgs {
    SdpaArgs f;
    const void* dout;
    void *dq, *dk, *dv;
    int64_t lddo, lddq, lddk, lddv;
    float* dbias;        // [B,h,Lq,Lk] fp32 or NULL: dS, the gradient of an additive (b,h,q,k) bias such as the box-geometry term
};

// Register budget: the register file is per SM sub-partition (16K), so 12 resident warps (3 per sub-partition) may use
// up to 170 registers each: 6-warp CTAs x 2 or 4-warp CTAs x 3 (MINB) for d <= 64.  A-operand fragments are re-read
// from shared memory inside the chunk loops (ldmatrix is cheap) instead of living in registers across them.
template <int D, int NWARPS, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB) sdpa_bwd_mma_kernel(SdpaBwdMmaArgs a) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ __align__(16) uint8_t smem_raw[];
    constexpr int P = D + PAD;
    const SdpaArgs& f = a.f;
    const int Lq = (int)f.Lq, Lk = (int)f.Lk;
    const int Lqp = (Lq + BCH - 1) / BCH * BCH, Lkp = (Lk + BCH - 1) / BCH * BCH;
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);      // [Lqp][P]
    __nv_bfloat16* sdO = sQ + Lqp * P;                                    // [Lqp][P]
    __nv_bfloat16* sK = sdO + Lqp * P;                                    // [Lkp][P]
    __nv_bfloat16* sV = sK + Lkp * P;                                     // [Lkp][P]
    __nv_bfloat16* sO = sV + Lkp * P;                                     // [Lqp][P]  (only until delta is known)
    float* sLse = reinterpret_cast<float*>(sO + Lqp * P);                 // [Lqp]
    float* sDelta = sLse + Lqp;                                           // [Lqp]
    float* sMask = sDelta + Lqp;                                          // [Lkp]    per-key mask row (when broadcast over queries)
    const int b = blockIdx.y, hd = blockIdx.x;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, g = lane / 4, t = lane % 4;
    const __nv_bfloat16* Q = reinterpret_cast<const __nv_bfloat16*>(f.q) + (int64_t)b * Lq * f.ldq + (int64_t)hd * D;
    const __nv_bfloat16* K = reinterpret_cast<const __nv_bfloat16*>(f.k) + (int64_t)b * Lk * f.ldk + (int64_t)hd * D;
    const __nv_bfloat16* V = reinterpret_cast<const __nv_bfloat16*>(f.v) + (int64_t)b * Lk * f.ldv + (int64_t)hd * D;
    const __nv_bfloat16* O = reinterpret_cast<const __nv_bfloat16*>(f.out) + (int64_t)b * Lq * f.ldo + (int64_t)hd * D;
    const __nv_bfloat16* dO = reinterpret_cast<const __nv_bfloat16*>(a.dout) + (int64_t)b * Lq * a.lddo + (int64_t)hd * D;

    stage_rows<D>(sQ, Q, f.ldq, Lqp, Lq, threadIdx.x, NWARPS * 32);
    stage_rows<D>(sdO, dO, a.lddo, Lqp, Lq, threadIdx.x, NWARPS * 32);
    stage_rows<D>(sK, K, f.ldk, Lkp, Lk, threadIdx.x, NWARPS * 32);
    stage_rows<D>(sV, V, f.ldv, Lkp, Lk, threadIdx.x, NWARPS * 32);
    stage_rows<D>(sO, O, f.ldo, Lqp, Lq, threadIdx.x, NWARPS * 32);
    // lse (+inf on padded rows: their probabilities vanish) and the mask row while the copies fly
    const float* lse_g = f.lse + ((int64_t)b * f.h + hd) * Lq;
    for (int i = threadIdx.x; i < Lqp; i += NWARPS * 32) sLse[i] = i < Lq ? __ldg(lse_g + i) : INFINITY;
    const float* mbase = f.mask.ptr ? f.mask.ptr + (int64_t)b * f.mask.sb + (int64_t)hd * f.mask.sh : nullptr;
    const bool row_bcast = mbase == nullptr || f.mask.sq == 0;
    if (row_bcast)
        for (int j = threadIdx.x; j < Lkp; j += NWARPS * 32) sMask[j] = (mbase && j < Lk) ? __ldg(mbase + j) : 0.f;
    cp_async_wait_all();
    __syncthreads();
    // delta_i = sum_d dO[i,d] * O[i,d]: 8 lanes per row, each one 16-byte piece at a time, all operands already in smem
    for (int i0 = 0; i0 < Lqp; i0 += NWARPS * 4) {                       // block-uniform trip count (full-warp shuffles below)
        const int i = i0 + threadIdx.x / 8;
        float s = 0.f;
        for (int c = threadIdx.x % 8; i < Lqp && c < D / 8; c += 8) {
            const uint4 o4 = *reinterpret_cast<const uint4*>(sO + i * P + c * 8), d4 = *reinterpret_cast<const uint4*>(sdO + i * P + c * 8);
            const uint32_t ow[4] = {o4.x, o4.y, o4.z, o4.w}, dw[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 o2 = unpack_bf16x2(ow[e]), d2 = unpack_bf16x2(dw[e]);
                s = fmaf(o2.x, d2.x, fmaf(o2.y, d2.y, s));
            }
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (threadIdx.x % 8 == 0 && i < Lqp) sDelta[i] = s;
    }
    __syncthreads();

    const int nqt = Lqp / 16, nkt = Lkp / 16;
    // heavier key-tile items first (4 contractions each), then the query tiles
    for (int item = warp; item < nqt + nkt; item += NWARPS) {
        if (item >= nkt) {
            // ------------------------------------------------ query-tile item
            const int r0 = (item - nkt) * 16;
            if (r0 >= Lq) continue;
            const int i_lo = r0 + g, i_hi = i_lo + 8;
            const float lse_lo = sLse[i_lo], lse_hi = sLse[i_hi], dl_lo = sDelta[i_lo], dl_hi = sDelta[i_hi];
            const float* mrow_lo = !row_bcast ? mbase + (int64_t)min(i_lo, Lq - 1) * f.mask.sq : nullptr;
            const float* mrow_hi = !row_bcast ? mbase + (int64_t)min(i_hi, Lq - 1) * f.mask.sq : nullptr;
            float acc[D / 8][4];
#pragma unroll
            for (int i = 0; i < D / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
            for (int k0 = 0; k0 < Lkp; k0 += BCH) {
                float s[BCH / 8][4], dp[BCH / 8][4];
#pragma unroll
                for (int i = 0; i < BCH / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
                {
                    uint32_t fa[D / 16][4];
                    load_a_frags<D>(fa, sQ, r0, lane);
                    mma_a_rowsT<D, BCH>(s, fa, sK + k0 * P, lane);
                    load_a_frags<D>(fa, sdO, r0, lane);
                    mma_a_rowsT<D, BCH>(dp, fa, sV + k0 * P, lane);
                }
#pragma unroll
                for (int nt = 0; nt < BCH / 8; ++nt) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = k0 + nt * 8 + 2 * t + e;
                        float p_lo = 0.f, p_hi = 0.f;
                        if (j < Lk) {
                            float v_lo = s[nt][e] * f.scale, v_hi = s[nt][2 + e] * f.scale;
                            if (mrow_lo) { v_lo += __ldg(mrow_lo + j); v_hi += __ldg(mrow_hi + j); }
                            else { const float mk = sMask[j]; v_lo += mk; v_hi += mk; }
                            p_lo = __expf(v_lo - lse_lo); p_hi = __expf(v_hi - lse_hi);
                        }
                        s[nt][e] = p_lo * (dp[nt][e] - dl_lo);
                        s[nt][2 + e] = p_hi * (dp[nt][2 + e] - dl_hi);
                        if (a.dbias && j < Lk) {
                            float* db = a.dbias + (((int64_t)b * f.h + hd) * Lq) * Lk + j;
                            if (i_lo < Lq) db[(int64_t)i_lo * Lk] = s[nt][e];
                            if (i_hi < Lq) db[(int64_t)i_hi * Lk] = s[nt][2 + e];
                        }
                    }
                }
                mma_p_rows<D, BCH>(acc, s, sK + k0 * P, lane);
            }
            __nv_bfloat16* dQ = reinterpret_cast<__nv_bfloat16*>(a.dq) + (int64_t)b * Lq * a.lddq + (int64_t)hd * D;
#pragma unroll
            for (int nt = 0; nt < D / 8; ++nt) {
                const int c = nt * 8 + 2 * t;
                if (i_lo < Lq) *reinterpret_cast<uint32_t*>(dQ + (int64_t)i_lo * a.lddq + c) = pack_bf16x2(acc[nt][0] * f.scale, acc[nt][1] * f.scale);
                if (i_hi < Lq) *reinterpret_cast<uint32_t*>(dQ + (int64_t)i_hi * a.lddq + c) = pack_bf16x2(acc[nt][2] * f.scale, acc[nt][3] * f.scale);
            }
        } else {
            // ------------------------------------------------ key-tile item (everything transposed: rows = keys)
            const int j0 = item * 16;
            if (j0 >= Lk) continue;
            const int j_lo = j0 + g, j_hi = j_lo + 8;
            const bool ok_lo = j_lo < Lk, ok_hi = j_hi < Lk;
            const float* mcol_lo = mbase ? mbase + min(j_lo, Lk - 1) : nullptr;   // + i * sq
            const float* mcol_hi = mbase ? mbase + min(j_hi, Lk - 1) : nullptr;
            const float m0_lo = row_bcast ? sMask[j_lo] : 0.f, m0_hi = row_bcast ? sMask[j_hi] : 0.f;
            float acck[D / 8][4], accv[D / 8][4];
#pragma unroll
            for (int i = 0; i < D / 8; ++i) { acck[i][0] = acck[i][1] = acck[i][2] = acck[i][3] = 0.f; accv[i][0] = accv[i][1] = accv[i][2] = accv[i][3] = 0.f; }
            for (int i0 = 0; i0 < Lqp; i0 += BCH) {
                float st[BCH / 8][4], dpt[BCH / 8][4];
#pragma unroll
                for (int i = 0; i < BCH / 8; ++i) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
                {
                    uint32_t fa[D / 16][4];
                    load_a_frags<D>(fa, sK, j0, lane);
                    mma_a_rowsT<D, BCH>(st, fa, sQ + i0 * P, lane);
                    load_a_frags<D>(fa, sV, j0, lane);
                    mma_a_rowsT<D, BCH>(dpt, fa, sdO + i0 * P, lane);
                }
#pragma unroll
                for (int nt = 0; nt < BCH / 8; ++nt) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int i = i0 + nt * 8 + 2 * t + e;           // query index (column of the transposed tile)
                        const float lse_i = sLse[i], dl_i = sDelta[i];   // +inf on padded queries -> p = 0
                        float v_lo = st[nt][e] * f.scale, v_hi = st[nt][2 + e] * f.scale;
                        if (row_bcast) { v_lo += m0_lo; v_hi += m0_hi; }
                        else {
                            { const int64_t off = (int64_t)min(i, Lq - 1) * f.mask.sq; v_lo += __ldg(mcol_lo + off); v_hi += __ldg(mcol_hi + off); }
                        }
                        const float p_lo = ok_lo ? __expf(v_lo - lse_i) : 0.f, p_hi = ok_hi ? __expf(v_hi - lse_i) : 0.f;
                        st[nt][e] = p_lo; st[nt][2 + e] = p_hi;
                        dpt[nt][e] = p_lo * (dpt[nt][e] - dl_i);
                        dpt[nt][2 + e] = p_hi * (dpt[nt][2 + e] - dl_i);
                    }
                }
                mma_p_rows<D, BCH>(accv, st, sdO + i0 * P, lane);
                mma_p_rows<D, BCH>(acck, dpt, sQ + i0 * P, lane);
            }
            __nv_bfloat16* dK = reinterpret_cast<__nv_bfloat16*>(a.dk) + (int64_t)b * Lk * a.lddk + (int64_t)hd * D;
            __nv_bfloat16* dV = reinterpret_cast<__nv_bfloat16*>(a.dv) + (int64_t)b * Lk * a.lddv + (int64_t)hd * D;
#pragma unroll
            for (int nt = 0; nt < D / 8; ++nt) {
                const int c = nt * 8 + 2 * t;
                if (ok_lo) {
                    *reinterpret_cast<uint32_t*>(dK + (int64_t)j_lo * a.lddk + c) = pack_bf16x2(acck[nt][0] * f.scale, acck[nt][1] * f.scale);
                    *reinterpret_cast<uint32_t*>(dV + (int64_t)j_lo * a.lddv + c) = pack_bf16x2(accv[nt][0], accv[nt][1]);
                }
                if (ok_hi) {
                    *reinterpret_cast<uint32_t*>(dK + (int64_t)j_hi * a.lddk + c) = pack_bf16x2(acck[nt][2] * f.scale, acck[nt][3] * f.scale);
                    *reinterpret_cast<uint32_t*>(dV + (int64_t)j_hi * a.lddv + c) = pack_bf16x2(accv[nt][2], accv[nt][3]);
                }
            }
        }
    }
}

size_t bwd_smem_bytes(int D, int64_t Lq, int64_t Lk) {
    const int64_t Lqp = (Lq + BCH - 1) / BCH * BCH, Lkp = (Lk + BCH - 1) / BCH * BCH;
    return (size_t)((3 * Lqp + 2 * Lkp) * (D + PAD)) * sizeof(__nv_bfloat16) + (size_t)(2 * Lqp + Lkp) * sizeof(float);
}

template <int D>
int launch_bwd(const SdpaBwdMmaArgs& a, cudaStream_t st) {
    const size_t smem = bwd_smem_bytes(D, a.f.Lq, a.f.Lk);
    const int items = (int)((a.f.Lq + 15) / 16 + (a.f.Lk + 15) / 16);
    auto go = [&](auto kern, int nw) -> int {
        if (smem > 48 * 1024) OVQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)a.f.h, (unsigned)a.f.B);
        OVQA_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(nw * 32), (size_t)(smem), st, a));
        OVQA_LAUNCH_OK();
        return OVQA_OK;
    };
    if constexpr (D <= 64) {
        if (items <= 4) return go(sdpa_bwd_mma_kernel<D, 4, 3>, 4);
        return go(sdpa_bwd_mma_kernel<D, 6, 2>, 6);
    } else {
        if (items <= 4) return go(sdpa_bwd_mma_kernel<D, 4, 1>, 4);
        return go(sdpa_bwd_mma_kernel<D, 8, 1>, 8);
    }
}

template <int D>
int launch_fwd(const SdpaArgs& a, cudaStream_t st) {
    const int Lq = (int)a.Lq;
    const int64_t Lkp = (a.Lk + KCH - 1) / KCH * KCH;
    // warps per CTA: cover Lq with as few CTAs as make sense (16 rows per warp), at most 8 warps
    const int tiles = (Lq + 15) / 16;
    auto go = [&](auto kern_full, auto kern_chunk, int nw) -> int {
        const size_t smem_full = sizeof(__nv_bfloat16) * (size_t)(nw * 16 + 2 * Lkp) * (D + PAD) + sizeof(float) * Lkp;
        const bool full = smem_full <= 100 * 1024;                      // keeps >= 2 CTAs per SM resident
        const size_t smem = full ? smem_full : sizeof(__nv_bfloat16) * (size_t)(nw * 16 + 2 * KCH) * (D + PAD) + sizeof(float) * Lkp;
        dim3 grid((unsigned)((Lq + nw * 16 - 1) / (nw * 16)), (unsigned)a.h, (unsigned)a.B);
        if (full) {
            if (smem > 48 * 1024) OVQA_CUDA_OK(cudaFuncSetAttribute(kern_full, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            OVQA_CUDA_OK(launch_pdl(kern_full, dim3(grid), dim3(nw * 32), (size_t)(smem), st, a));
        } else {
            if (smem > 48 * 1024) OVQA_CUDA_OK(cudaFuncSetAttribute(kern_chunk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            OVQA_CUDA_OK(launch_pdl(kern_chunk, dim3(grid), dim3(nw * 32), (size_t)(smem), st, a));
        }
        OVQA_LAUNCH_OK();
        return OVQA_OK;
    };
#define OVQA_FWD(NW) go(sdpa_fwd_mma_kernel<D, NW, true>, sdpa_fwd_mma_kernel<D, NW, false>, NW)
    if (tiles <= 2) return OVQA_FWD(2);
    if (tiles <= 4) return OVQA_FWD(4);
    if (tiles == 7 || tiles == 13 || tiles == 14) return OVQA_FWD(7);   // 100 or 200-224 rows: no empty warps
    if (tiles % 6 == 0 || tiles == 11 || tiles == 12) return OVQA_FWD(6);      // 182 rows -> 12 tiles
    return OVQA_FWD(8);
#undef OVQA_FWD
}

}  // namespace

bool sdpa_fwd_mma_supported(const SdpaArgs& a) {
    if (a.geo.boxes || a.att) return false;                     // geometry / materialised probabilities: fp32-math kernel
    if (a.dk != a.dv || !(a.dk == 32 || a.dk == 64 || a.dk == 96 || a.dk == 128)) return false;
    // 16-byte cp.async rows and 4-byte output stores
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!al16(a.q) || !al16(a.k) || !al16(a.v) || (a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (a.ldo % 2)) return false;
    if (reinterpret_cast<uintptr_t>(a.out) & 3) return false;
    return true;
}

int sdpa_fwd_mma_bf16(const SdpaArgs& a, cudaStream_t st) {
    switch ((int)a.dk) {
        case 32: return launch_fwd<32>(a, st);
        case 64: return launch_fwd<64>(a, st);
        case 96: return launch_fwd<96>(a, st);
        case 128: return launch_fwd<128>(a, st);
    }
    return OVQA_ERR_BAD_SHAPE;
}

bool sdpa_bwd_mma_supported(const SdpaArgs& a, const void* dout, const void* dq, const void* dk, const void* dv, int64_t lddo,
                            int64_t lddq, int64_t lddk, int64_t lddv) {
    if (a.geo.boxes) return false;                              // geometry: fp32-math kernels (also produce the fc_g gradients)
    if (a.dk != a.dv || !(a.dk == 32 || a.dk == 64 || a.dk == 96)) return false;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    auto al4 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3) == 0; };
    if (!al16(a.q) || !al16(a.k) || !al16(a.v) || !al16(dout) || (a.ldq % 8) || (a.ldk % 8) || (a.ldv % 8) || (lddo % 8)) return false;
    if (!al4(a.out) || !al4(dq) || !al4(dk) || !al4(dv) || (a.ldo % 2) || (lddq % 2) || (lddk % 2) || (lddv % 2)) return false;
    return bwd_smem_bytes((int)a.dk, a.Lq, a.Lk) <= 200 * 1024;
}

int sdpa_bwd_mma_bf16(const SdpaArgs& f, const void* dout, void* dq, void* dk, void* dv, int64_t lddo, int64_t lddq, int64_t lddk,
                      int64_t lddv, float* dbias, cudaStream_t st) {
    SdpaBwdMmaArgs a{f, dout, dq, dk, dv, lddo, lddq, lddk, lddv, dbias};
    switch ((int)f.dk) {
        case 32: return launch_bwd<32>(a, st);
        case 64: return launch_bwd<64>(a, st);
        case 96: return launch_bwd<96>(a, st);
    }
    return OVQA_ERR_BAD_SHAPE;
}

}  // namespace ovqa
