// Attention core, fp32-math path (any storage type): forward, probabilities, backward.
//   out = softmax(q k^T * scale + mask [+ log-geometry]) v            (attentions.py:53-57 / :121-134)
// One CTA owns a (batch, head, 16-query tile); scores for the tile live in shared memory (16 x Lk fp32), K and
// V stream through a 32-key smem chunk.  The additive mask is applied in fp32 BEFORE the row max, exactly as
// the reference does (`att += mask; softmax`), so fully-masked rows reproduce softmax(s - 1e5).
// This is the 1e-4 parity path and the reference for the bf16 tensor-core kernel in attention_mma.cu.
#include "attention_simt.cuh"

namespace ovqa {

namespace {

constexpr int QT = 16;       // query rows per CTA
constexpr int KC = 32;       // keys per streamed chunk (== warp size: lane <-> key)
constexpr int NW = 4;        // warps per CTA
constexpr int RPW = QT / NW; // rows per warp
constexpr int MAXD = 128;    // max head dim

template <typename T>
__global__ void __launch_bounds__(NW * 32) sdpa_fwd_simt(SdpaArgs a) {
    extern __shared__ float sm[];
    const int dk = (int)a.dk, dv = (int)a.dv, Lk = (int)a.Lk;
    const int dmax = dk > dv ? dk : dv;
    float* sQ = sm;                         // [QT][dk]
    float* sKV = sQ + QT * dk;              // [KC][dmax+1]
    float* sS = sKV + KC * (dmax + 1);      // [QT][Lk]
    BoxGeom* sBox = reinterpret_cast<BoxGeom*>(sS + QT * Lk);  // [Lk] (geometry only)

    const int b = blockIdx.z, hd = blockIdx.y, q0 = blockIdx.x * QT;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int nq = min(QT, (int)a.Lq - q0);
    const T* Q = reinterpret_cast<const T*>(a.q) + ((int64_t)b * a.Lq + q0) * a.ldq + (int64_t)hd * dk;
    const T* K = reinterpret_cast<const T*>(a.k) + (int64_t)b * a.Lk * a.ldk + (int64_t)hd * dk;
    const T* V = reinterpret_cast<const T*>(a.v) + (int64_t)b * a.Lk * a.ldv + (int64_t)hd * dv;
    const bool geo = a.geo.boxes != nullptr;

    for (int i = threadIdx.x; i < QT * dk; i += blockDim.x) {
        const int r = i / dk, d = i % dk;
        sQ[i] = r < nq ? to_f32<T>(Q[(int64_t)r * a.ldq + d]) : 0.0f;
    }
    if (geo)
        for (int j = threadIdx.x; j < Lk; j += blockDim.x) sBox[j] = box_geom(a.geo.boxes + ((int64_t)b * a.Lk + j) * 4);
    __syncthreads();

    // ---- phase 1: scores
    for (int k0 = 0; k0 < Lk; k0 += KC) {
        const int nk = min(KC, Lk - k0);
        for (int i = threadIdx.x; i < KC * dk; i += blockDim.x) {
            const int r = i / dk, d = i % dk;
            sKV[r * (dmax + 1) + d] = r < nk ? to_f32<T>(K[(int64_t)(k0 + r) * a.ldk + d]) : 0.0f;
        }
        __syncthreads();
        if (lane < nk) {
            const float* kr = sKV + lane * (dmax + 1);
            const int j = k0 + lane;
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const int r = warp * RPW + rr;
                if (r >= nq) break;
                const float* qr = sQ + r * dk;
                float s = 0.0f;
                for (int d = 0; d < dk; ++d) s = fmaf(qr[d], kr[d], s);
                s *= a.scale;
                if (a.mask.ptr) s += __ldg(a.mask.ptr + (int64_t)b * a.mask.sb + (int64_t)hd * a.mask.sh + (int64_t)(q0 + r) * a.mask.sq + j);
                if (geo) s = geo_log_bias(geo_raw(a.geo, hd, sBox[q0 + r], sBox[j])) + s;
                sS[r * Lk + j] = s;
            }
        }
        __syncthreads();
    }
    // ---- phase 2: row softmax (warp per row)
    for (int rr = 0; rr < RPW; ++rr) {
        const int r = warp * RPW + rr;
        if (r >= nq) break;
        float* row = sS + r * Lk;
        float mx = -INFINITY;
        for (int j = lane; j < Lk; j += 32) mx = fmaxf(mx, row[j]);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < Lk; j += 32) { const float e = expf(row[j] - mx); row[j] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float* att = a.att ? a.att + (((int64_t)b * a.h + hd) * a.Lq + q0 + r) * a.Lk : nullptr;
        const uint64_t row_idx = (uint64_t)((((int64_t)b * a.h + hd) * a.Lq + q0 + r) * a.Lk);
        for (int j = lane; j < Lk; j += 32) {
            float p = row[j] * inv;
            if (a.pdrop > 0.f) p *= dropout_scale(row_idx + j, a.pdrop, 1.0f / (1.0f - a.pdrop), a.pdrop_seed, effective_offset(a.pdrop_offset, a.pdrop_offset_base));
            row[j] = p;
            if (att) att[j] = p;
        }
        if (a.lse && lane == 0) a.lse[((int64_t)b * a.h + hd) * a.Lq + q0 + r] = mx + logf(sum);
    }
    __syncthreads();
    // ---- phase 3: out = P V   (lanes over the value dim)
    float acc[RPW][MAXD / 32];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
        for (int t = 0; t < MAXD / 32; ++t) acc[rr][t] = 0.0f;
    for (int k0 = 0; k0 < Lk; k0 += KC) {
        const int nk = min(KC, Lk - k0);
        for (int i = threadIdx.x; i < KC * dv; i += blockDim.x) {
            const int r = i / dv, d = i % dv;
            sKV[r * (dmax + 1) + d] = r < nk ? to_f32<T>(V[(int64_t)(k0 + r) * a.ldv + d]) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr;
            if (r < nq) {
                const float* prow = sS + r * Lk + k0;
                for (int j = 0; j < nk; ++j) {
                    const float p = prow[j];
                    const float* vr = sKV + j * (dmax + 1);
#pragma unroll
                    for (int t = 0; t < MAXD / 32; ++t)
                        if (lane + 32 * t < dv) acc[rr][t] = fmaf(p, vr[lane + 32 * t], acc[rr][t]);
                }
            }
        }
        __syncthreads();
    }
    T* O = reinterpret_cast<T*>(a.out) + ((int64_t)b * a.Lq + q0) * a.ldo + (int64_t)hd * dv;
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const int r = warp * RPW + rr;
        if (r >= nq) continue;
#pragma unroll
        for (int t = 0; t < MAXD / 32; ++t)
            if (lane + 32 * t < dv) O[(int64_t)r * a.ldo + lane + 32 * t] = from_f32<T>(acc[rr][t]);
    }
}

// ------------------------------------------------------------------------------------------------ backward
struct SdpaBwdArgs {
    SdpaArgs f;          // forward tensors (out = forward output)
    const void* dout;
    void *dq, *dk, *dv;
    int64_t lddo, lddq, lddk, lddv;
    float* delta;        // [B,h,Lq]
    float* geo_dw;       // [h,d_g] accumulated
    float* geo_db;       // [h]
    const float* att;    // [B,h,Lq,Lk] forward probabilities, with datt: the caller differentiates THROUGH the returned probabilities
    const float* datt;   // [B,h,Lq,Lk] their upstream gradient (mmf_lorra.py:127-146 builds features from `att`), or NULL
};

// delta[b,h,i] = sum_d dout[b,i,h,d] * out[b,i,h,d]  (+ sum_j att[b,h,i,j] * datt[b,h,i,j])     (warp per row)
template <typename T>
__global__ void sdpa_delta_kernel(SdpaBwdArgs a) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;  // over B*Lq*h
    const int lane = threadIdx.x % 32;
    if (row >= a.f.B * a.f.Lq * a.f.h) return;
    const int hd = (int)(row % a.f.h);
    const int64_t bi = row / a.f.h;  // b*Lq + i
    const T* o = reinterpret_cast<const T*>(a.f.out) + bi * a.f.ldo + (int64_t)hd * a.f.dv;
    const T* d_o = reinterpret_cast<const T*>(a.dout) + bi * a.lddo + (int64_t)hd * a.f.dv;
    float s = 0.0f;
    for (int d = lane; d < a.f.dv; d += 32) s = fmaf(to_f32<T>(o[d]), to_f32<T>(d_o[d]), s);
    const int64_t stat = ((bi / a.f.Lq) * a.f.h + hd) * a.f.Lq + bi % a.f.Lq;
    if (a.datt) {
        const float* pr = a.att + stat * a.f.Lk;
        const float* dr = a.datt + stat * a.f.Lk;
        for (int j = lane; j < a.f.Lk; j += 32) s = fmaf(__ldg(pr + j), __ldg(dr + j), s);
    }
    s = warp_sum(s);
    if (lane == 0) a.delta[stat] = s;
}

// score recomputation for one (row, key) pair; identical arithmetic to the forward
template <typename T>
__device__ __forceinline__ float recompute_score(const SdpaArgs& a, const float* qr, const float* kr, int b, int hd, int i, int j,
                                                 const BoxGeom* sBox, float* raw_out) {
    float s = 0.0f;
    for (int d = 0; d < (int)a.dk; ++d) s = fmaf(qr[d], kr[d], s);
    s *= a.scale;
    if (a.mask.ptr) s += __ldg(a.mask.ptr + (int64_t)b * a.mask.sb + (int64_t)hd * a.mask.sh + (int64_t)i * a.mask.sq + j);
    if (a.geo.boxes) {
        const float raw = geo_raw(a.geo, hd, sBox[i], sBox[j]);
        if (raw_out) *raw_out = raw;
        s = geo_log_bias(raw) + s;
    }
    return s;
}

// dq kernel: CTA = (b, head, 16-query tile); streams K,V chunks.  Also produces the geometry weight gradient.
template <typename T>
__global__ void __launch_bounds__(NW * 32) sdpa_bwd_dq_simt(SdpaBwdArgs a) {
    extern __shared__ float sm[];
    const SdpaArgs& f = a.f;
    const int dk = (int)f.dk, dv = (int)f.dv, Lk = (int)f.Lk;
    float* sQ = sm;                      // [QT][dk]
    float* sdO = sQ + QT * dk;           // [QT][dv]
    float* sK = sdO + QT * dv;           // [KC][dk+1]
    float* sV = sK + KC * (dk + 1);      // [KC][dv+1]
    float* sdS = sV + KC * (dv + 1);     // [QT][KC]
    BoxGeom* sBox = reinterpret_cast<BoxGeom*>(sdS + QT * KC);
    const int b = blockIdx.z, hd = blockIdx.y, q0 = blockIdx.x * QT;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int nq = min(QT, (int)f.Lq - q0);
    const T* Q = reinterpret_cast<const T*>(f.q) + ((int64_t)b * f.Lq + q0) * f.ldq + (int64_t)hd * dk;
    const T* dO = reinterpret_cast<const T*>(a.dout) + ((int64_t)b * f.Lq + q0) * a.lddo + (int64_t)hd * dv;
    const T* K = reinterpret_cast<const T*>(f.k) + (int64_t)b * f.Lk * f.ldk + (int64_t)hd * dk;
    const T* V = reinterpret_cast<const T*>(f.v) + (int64_t)b * f.Lk * f.ldv + (int64_t)hd * dv;
    const bool geo = f.geo.boxes != nullptr;
    for (int i = threadIdx.x; i < QT * dk; i += blockDim.x) { const int r = i / dk, d = i % dk; sQ[i] = r < nq ? to_f32<T>(Q[(int64_t)r * f.ldq + d]) : 0.f; }
    for (int i = threadIdx.x; i < QT * dv; i += blockDim.x) { const int r = i / dv, d = i % dv; sdO[i] = r < nq ? to_f32<T>(dO[(int64_t)r * a.lddo + d]) : 0.f; }
    if (geo) for (int j = threadIdx.x; j < Lk; j += blockDim.x) sBox[j] = box_geom(f.geo.boxes + ((int64_t)b * f.Lk + j) * 4);
    float acc[RPW][MAXD / 32];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
        for (int t = 0; t < MAXD / 32; ++t) acc[rr][t] = 0.f;
    const int64_t stat0 = ((int64_t)b * f.h + hd) * f.Lq + q0;
    __syncthreads();
    for (int k0 = 0; k0 < Lk; k0 += KC) {
        const int nk = min(KC, Lk - k0);
        for (int i = threadIdx.x; i < KC * dk; i += blockDim.x) { const int r = i / dk, d = i % dk; sK[r * (dk + 1) + d] = r < nk ? to_f32<T>(K[(int64_t)(k0 + r) * f.ldk + d]) : 0.f; }
        for (int i = threadIdx.x; i < KC * dv; i += blockDim.x) { const int r = i / dv, d = i % dv; sV[r * (dv + 1) + d] = r < nk ? to_f32<T>(V[(int64_t)(k0 + r) * f.ldv + d]) : 0.f; }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr;
            float ds = 0.f;
            if (r < nq && lane < nk) {
                float raw = 1.f;
                const float s = recompute_score<T>(f, sQ + r * dk, sK + lane * (dk + 1), b, hd, q0 + r, k0 + lane, sBox, &raw);
                const float p = expf(s - f.lse[stat0 + r]);
                float dp = 0.f;
                const float* dor = sdO + r * dv;
                const float* vr = sV + lane * (dv + 1);
                for (int d = 0; d < dv; ++d) dp = fmaf(dor[d], vr[d], dp);
                if (a.datt) dp += __ldg(a.datt + (stat0 + r) * Lk + k0 + lane);
                if (f.pdrop > 0.f) dp *= dropout_scale((uint64_t)((stat0 + r) * Lk + k0 + lane), f.pdrop, 1.0f / (1.0f - f.pdrop), f.pdrop_seed, effective_offset(f.pdrop_offset, f.pdrop_offset_base));
                ds = p * (dp - a.delta[stat0 + r]);
                if (geo && a.geo_dw && raw > 1e-6f) {
                    // d log(clamp(relu(raw))) / d raw = 1/raw when raw > 1e-6; chain into w_h, b_h (attentions.py:116-132)
                    const float gcoef = ds / raw;
                    float d4[4];
                    box_deltas(sBox[q0 + r], sBox[k0 + lane], d4);
                    if (f.geo.trig) {
                        const int nf = f.geo.d_g / 8;
                        for (int t = 0; t < 4; ++t)
                            for (int kf = 0; kf < nf; ++kf) {
                                float sn, cs;
                                sincosf(100.0f * d4[t] * geo_freq(f.geo, kf, nf), &sn, &cs);
                                atomicAdd(a.geo_dw + (int64_t)hd * f.geo.d_g + t * nf + kf, gcoef * sn);
                                atomicAdd(a.geo_dw + (int64_t)hd * f.geo.d_g + 4 * nf + t * nf + kf, gcoef * cs);
                            }
                    } else {
                        for (int t = 0; t < 4; ++t) atomicAdd(a.geo_dw + (int64_t)hd * f.geo.d_g + t, gcoef * d4[t]);
                    }
                    atomicAdd(a.geo_db + hd, gcoef);
                }
            }
            sdS[r * KC + lane] = ds;
        }
        __syncwarp();
        // dq_r += scale * sum_j dS_rj k_j   (lanes over head dim)
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int r = warp * RPW + rr;
            if (r < nq)
                for (int j = 0; j < nk; ++j) {
                    const float ds = sdS[r * KC + j];
                    const float* kr = sK + j * (dk + 1);
#pragma unroll
                    for (int t = 0; t < MAXD / 32; ++t)
                        if (lane + 32 * t < dk) acc[rr][t] = fmaf(ds, kr[lane + 32 * t], acc[rr][t]);
                }
        }
        __syncthreads();
    }
    T* dQ = reinterpret_cast<T*>(a.dq) + ((int64_t)b * f.Lq + q0) * a.lddq + (int64_t)hd * dk;
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const int r = warp * RPW + rr;
        if (r >= nq) continue;
#pragma unroll
        for (int t = 0; t < MAXD / 32; ++t)
            if (lane + 32 * t < dk) dQ[(int64_t)r * a.lddq + lane + 32 * t] = from_f32<T>(acc[rr][t] * f.scale);
    }
}

// dk/dv kernel: CTA = (b, head, 16-key tile); streams Q,dO chunks of 32 query rows (lane <-> query row).
template <typename T>
__global__ void __launch_bounds__(NW * 32) sdpa_bwd_dkv_simt(SdpaBwdArgs a) {
    extern __shared__ float sm[];
    const SdpaArgs& f = a.f;
    const int dk = (int)f.dk, dv = (int)f.dv, Lq = (int)f.Lq, Lk = (int)f.Lk;
    float* sK = sm;                      // [QT][dk]      (QT keys here)
    float* sV = sK + QT * dk;            // [QT][dv]
    float* sQ = sV + QT * dv;            // [KC][dk+1]    (KC query rows)
    float* sdO = sQ + KC * (dk + 1);     // [KC][dv+1]
    float* sP = sdO + KC * (dv + 1);     // [QT][KC]
    float* sdS = sP + QT * KC;           // [QT][KC]
    BoxGeom* sBox = reinterpret_cast<BoxGeom*>(sdS + QT * KC);
    const int b = blockIdx.z, hd = blockIdx.y, j0 = blockIdx.x * QT;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int nkeys = min(QT, Lk - j0);
    const T* Kp = reinterpret_cast<const T*>(f.k) + ((int64_t)b * f.Lk + j0) * f.ldk + (int64_t)hd * dk;
    const T* Vp = reinterpret_cast<const T*>(f.v) + ((int64_t)b * f.Lk + j0) * f.ldv + (int64_t)hd * dv;
    const T* Q = reinterpret_cast<const T*>(f.q) + (int64_t)b * f.Lq * f.ldq + (int64_t)hd * dk;
    const T* dO = reinterpret_cast<const T*>(a.dout) + (int64_t)b * f.Lq * a.lddo + (int64_t)hd * dv;
    const bool geo = f.geo.boxes != nullptr;
    for (int i = threadIdx.x; i < QT * dk; i += blockDim.x) { const int r = i / dk, d = i % dk; sK[i] = r < nkeys ? to_f32<T>(Kp[(int64_t)r * f.ldk + d]) : 0.f; }
    for (int i = threadIdx.x; i < QT * dv; i += blockDim.x) { const int r = i / dv, d = i % dv; sV[i] = r < nkeys ? to_f32<T>(Vp[(int64_t)r * f.ldv + d]) : 0.f; }
    if (geo) for (int j = threadIdx.x; j < Lk; j += blockDim.x) sBox[j] = box_geom(f.geo.boxes + ((int64_t)b * f.Lk + j) * 4);
    float acck[RPW][MAXD / 32], accv[RPW][MAXD / 32];
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
        for (int t = 0; t < MAXD / 32; ++t) { acck[rr][t] = 0.f; accv[rr][t] = 0.f; }
    const int64_t stat0 = ((int64_t)b * f.h + hd) * f.Lq;
    __syncthreads();
    for (int i0 = 0; i0 < Lq; i0 += KC) {
        const int nq = min(KC, Lq - i0);
        for (int i = threadIdx.x; i < KC * dk; i += blockDim.x) { const int r = i / dk, d = i % dk; sQ[r * (dk + 1) + d] = r < nq ? to_f32<T>(Q[(int64_t)(i0 + r) * f.ldq + d]) : 0.f; }
        for (int i = threadIdx.x; i < KC * dv; i += blockDim.x) { const int r = i / dv, d = i % dv; sdO[r * (dv + 1) + d] = r < nq ? to_f32<T>(dO[(int64_t)(i0 + r) * a.lddo + d]) : 0.f; }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int jr = warp * RPW + rr;  // key row in tile
            float p = 0.f, ds = 0.f;
            if (jr < nkeys && lane < nq) {
                const float s = recompute_score<T>(f, sQ + lane * (dk + 1), sK + jr * dk, b, hd, i0 + lane, j0 + jr, sBox, nullptr);
                p = expf(s - f.lse[stat0 + i0 + lane]);
                float dp = 0.f;
                const float* dor = sdO + lane * (dv + 1);
                const float* vr = sV + jr * dv;
                for (int d = 0; d < dv; ++d) dp = fmaf(dor[d], vr[d], dp);
                if (a.datt) dp += __ldg(a.datt + (stat0 + i0 + lane) * Lk + j0 + jr);
                float keep = 1.f;
                if (f.pdrop > 0.f) keep = dropout_scale((uint64_t)((stat0 + i0 + lane) * Lk + j0 + jr), f.pdrop, 1.0f / (1.0f - f.pdrop), f.pdrop_seed, effective_offset(f.pdrop_offset, f.pdrop_offset_base));
                dp *= keep;
                ds = p * (dp - a.delta[stat0 + i0 + lane]);
                p *= keep;                                  // dV = (dropped probabilities)^T dO
            }
            sP[jr * KC + lane] = p;
            sdS[jr * KC + lane] = ds;
        }
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const int jr = warp * RPW + rr;
            if (jr < nkeys)
                for (int i = 0; i < nq; ++i) {
                    const float p = sP[jr * KC + i], ds = sdS[jr * KC + i];
                    const float* qr = sQ + i * (dk + 1);
                    const float* dor = sdO + i * (dv + 1);
#pragma unroll
                    for (int t = 0; t < MAXD / 32; ++t) {
                        if (lane + 32 * t < dk) acck[rr][t] = fmaf(ds, qr[lane + 32 * t], acck[rr][t]);
                        if (lane + 32 * t < dv) accv[rr][t] = fmaf(p, dor[lane + 32 * t], accv[rr][t]);
                    }
                }
        }
        __syncthreads();
    }
    T* dK = reinterpret_cast<T*>(a.dk) + ((int64_t)b * f.Lk + j0) * a.lddk + (int64_t)hd * dk;
    T* dV = reinterpret_cast<T*>(a.dv) + ((int64_t)b * f.Lk + j0) * a.lddv + (int64_t)hd * dv;
#pragma unroll
    for (int rr = 0; rr < RPW; ++rr) {
        const int jr = warp * RPW + rr;
        if (jr >= nkeys) continue;
#pragma unroll
        for (int t = 0; t < MAXD / 32; ++t) {
            if (lane + 32 * t < dk) dK[(int64_t)jr * a.lddk + lane + 32 * t] = from_f32<T>(acck[rr][t] * f.scale);
            if (lane + 32 * t < dv) dV[(int64_t)jr * a.lddv + lane + 32 * t] = from_f32<T>(accv[rr][t]);
        }
    }
}

template <typename K>
int set_smem(K kern, size_t bytes) {
    if (bytes > 48 * 1024) OVQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return OVQA_OK;
}

}  // namespace

template <typename T>
int sdpa_fwd_simt_launch(const SdpaArgs& a, cudaStream_t st) {
    const int dmax = (int)(a.dk > a.dv ? a.dk : a.dv);
    size_t smem = sizeof(float) * (QT * a.dk + KC * (dmax + 1) + QT * a.Lk) + (a.geo.boxes ? sizeof(BoxGeom) * a.Lk : 0);
    if (smem > 227 * 1024) return OVQA_ERR_BAD_SHAPE;
    int rc = set_smem(sdpa_fwd_simt<T>, smem);
    if (rc) return rc;
    dim3 grid((unsigned)((a.Lq + QT - 1) / QT), (unsigned)a.h, (unsigned)a.B);
    sdpa_fwd_simt<T><<<grid, NW * 32, smem, st>>>(a);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
template int sdpa_fwd_simt_launch<float>(const SdpaArgs&, cudaStream_t);
template int sdpa_fwd_simt_launch<__nv_bfloat16>(const SdpaArgs&, cudaStream_t);

template <typename T>
int sdpa_bwd_simt_launch(const SdpaArgs& f, const void* dout, void* dq, void* dk, void* dv, int64_t lddo, int64_t lddq,
                         int64_t lddk, int64_t lddv, float* delta, float* geo_dw, float* geo_db, const float* att, const float* datt,
                         cudaStream_t st) {
    SdpaBwdArgs a{f, dout, dq, dk, dv, lddo, lddq, lddk, lddv, delta, geo_dw, geo_db, att, datt};
    const int64_t rows = f.B * f.Lq * f.h;
    sdpa_delta_kernel<T><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(a);
    OVQA_LAUNCH_OK();
    const size_t geo_b = f.geo.boxes ? sizeof(BoxGeom) * f.Lk : 0;
    size_t smem1 = sizeof(float) * (QT * f.dk + QT * f.dv + KC * (f.dk + 1) + KC * (f.dv + 1) + QT * KC) + geo_b;
    size_t smem2 = sizeof(float) * (QT * f.dk + QT * f.dv + KC * (f.dk + 1) + KC * (f.dv + 1) + 2 * QT * KC) + geo_b;
    int rc = set_smem(sdpa_bwd_dq_simt<T>, smem1);
    if (rc) return rc;
    rc = set_smem(sdpa_bwd_dkv_simt<T>, smem2);
    if (rc) return rc;
    dim3 g1((unsigned)((f.Lq + QT - 1) / QT), (unsigned)f.h, (unsigned)f.B);
    sdpa_bwd_dq_simt<T><<<g1, NW * 32, smem1, st>>>(a);
    OVQA_LAUNCH_OK();
    dim3 g2((unsigned)((f.Lk + QT - 1) / QT), (unsigned)f.h, (unsigned)f.B);
    sdpa_bwd_dkv_simt<T><<<g2, NW * 32, smem2, st>>>(a);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
template int sdpa_bwd_simt_launch<float>(const SdpaArgs&, const void*, void*, void*, void*, int64_t, int64_t, int64_t, int64_t, float*, float*, float*, const float*, const float*, cudaStream_t);
template int sdpa_bwd_simt_launch<__nv_bfloat16>(const SdpaArgs&, const void*, void*, void*, void*, int64_t, int64_t, int64_t, int64_t, float*, float*, float*, const float*, const float*, cudaStream_t);

// ------------------------------------------------------------------------------------------------ materialised geometry bias
__global__ void box_geometry_bias_kernel(GeoArg g, float* out, int64_t B, int64_t N, int64_t h) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over B*N*N
    if (idx >= B * N * N) return;
    const int j = (int)(idx % N), i = (int)((idx / N) % N);
    const int64_t b = idx / (N * N);
    const BoxGeom bi = box_geom(g.boxes + (b * N + i) * 4), bj = box_geom(g.boxes + (b * N + j) * 4);
    for (int hd = 0; hd < (int)h; ++hd) out[((b * h + hd) * N + i) * N + j] = geo_log_bias(geo_raw(g, hd, bi, bj));
}
int box_geometry_bias_launch(const GeoArg& g, float* out, int64_t B, int64_t N, int64_t h, cudaStream_t st) {
    const int64_t total = B * N * N;
    if (total == 0) return OVQA_OK;
    box_geometry_bias_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(g, out, B, N, h);
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

}  // namespace ovqa
