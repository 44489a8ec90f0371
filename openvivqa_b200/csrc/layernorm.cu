// Residual + LayerNorm (forward/backward) and the small element-wise helpers of the path.
// HBM-bound: one warp per row, the row lives in registers (128-bit loads/stores), statistics by warp shuffle;
// two-pass mean/variance in fp32 like ATen's LayerNorm.  Algorithmic bytes: fwd 2-3 x 4*D per row.
#include "common.cuh"

namespace ovqa {

namespace {

constexpr int kMaxVec = 8;  // float4 per lane -> D <= 1024
constexpr int kRowsPerCta = 4;
constexpr int kBwdWarps = 8;  // backward: warps per CTA.  Every CTA ends with one atomic per column vector into dgamma/dbeta/dbias, and
                              // same-address atomics serialise in L2: with 4-warp CTAs (1036 of them) and scalar atomics that tail cost
                              // ~40 us regardless of the row count (profiles/r01e: 47 us for 7680 rows, 62 us for 25600)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// NV = float4 vectors per lane (D <= 128*NV): sized at compile time so the row really lives in registers
template <int NV>
__global__ void __launch_bounds__(kRowsPerCta * 32)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ gamma,
                     const float* __restrict__ beta, const float* __restrict__ post_add, int64_t post_rows,
                     float* __restrict__ y, __nv_bfloat16* __restrict__ y16, float* __restrict__ sum_out,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int D, float eps) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int64_t row = (int64_t)blockIdx.x * kRowsPerCta + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= rows) return;
    const int nvec = D / 4;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
        const int c = lane + 32 * t;
        if (c < nvec) {
            v[t] = ld4(x + row * D + 4 * c);
            if (residual) v[t] = add4(v[t], ld4(residual + row * D + 4 * c));
            if (sum_out) *reinterpret_cast<float4*>(sum_out + row * D + 4 * c) = v[t];
            s += (v[t].x + v[t].y) + (v[t].z + v[t].w);
        }
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < NV; ++t) {
        const int c = lane + 32 * t;
        if (c < nvec) {
            const float a = v[t].x - mean, b = v[t].y - mean, cc = v[t].z - mean, d = v[t].w - mean;
            q += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + eps);
    if (lane == 0) { if (mean_out) mean_out[row] = mean; if (rstd_out) rstd_out[row] = rstd; }
#pragma unroll
    for (int t = 0; t < NV; ++t) {
        const int c = lane + 32 * t;
        if (c < nvec) {
            const float4 g = ld4(gamma + 4 * c), bb = ld4(beta + 4 * c);
            float4 o = make_float4((v[t].x - mean) * rstd * g.x + bb.x, (v[t].y - mean) * rstd * g.y + bb.y,
                                   (v[t].z - mean) * rstd * g.z + bb.z, (v[t].w - mean) * rstd * g.w + bb.w);
            if (post_add) o = add4(o, ld4(post_add + (row % post_rows) * D + 4 * c));
            if (y) *reinterpret_cast<float4*>(y + row * D + 4 * c) = o;
            if (y16) *reinterpret_cast<uint2*>(y16 + row * D + 4 * c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
}

// Backward.  One warp per row, persistent over rows.  The row's x and dy are brought in by cp.async into a per-warp
// double buffer in shared memory, the NEXT row's copies being issued before the current row is reduced: two rows of loads
// in flight per warp without holding them in registers (the kernel is DRAM-latency bound: ~150 registers allowed only 12
// warps per SM and one row in flight each; profiles/r01d: 36 % of HBM peak, 80 % of stall samples on long_scoreboard).
__device__ __forceinline__ void ln_cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <int NV>
__global__ void __launch_bounds__(kBwdWarps * 32, 2)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ dx_add,
                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ dbias, int64_t rows, int D, float p, float inv_keep, uint64_t seed, uint64_t offset_,
                     const uint64_t* offset_base) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    extern __shared__ __align__(16) float ln_smem[];       // [kBwdWarps][2 stages][2 (x, dy)][NV*128]; reused for the final reduction
    constexpr int W = NV * 128;
    const uint64_t offset = p > 0.f ? effective_offset(offset_, offset_base) : 0;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int nvec = D / 4;
    float* wbuf = ln_smem + (size_t)warp * 4 * W;
    float4 ag[NV], ab[NV], ad[NV];      // column sums for dgamma, dbeta and (masked dx) the producing GEMM's bias gradient
#pragma unroll
    for (int t = 0; t < NV; ++t) ag[t] = ab[t] = ad[t] = make_float4(0, 0, 0, 0);
    const int64_t stride = (int64_t)gridDim.x * kBwdWarps;
    auto prefetch = [&](int64_t row, int stage) {
        float* bx = wbuf + stage * 2 * W;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
            const int c = lane + 32 * t;
            if (c < nvec) {
                ln_cp_async16(bx + 4 * c, x + row * D + 4 * c);
                ln_cp_async16(bx + W + 4 * c, dy + row * D + 4 * c);
            }
        }
    };
    int64_t row = (int64_t)blockIdx.x * kBwdWarps + warp;
    if (row < rows) prefetch(row, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    // the row's statistics are requested one row ahead as well: loaded where they are used they were the kernel's top stall
    // (27 % of the samples on the first subtraction of the mean, gpurun_out/ln_full.ncu-rep)
    float mu_n = row < rows ? __ldg(mean + row) : 0.f, rs_n = row < rows ? __ldg(rstd + row) : 0.f;
    for (int it = 0; row < rows; row += stride, ++it) {
        const int stage = it & 1;
        const int64_t next = row + stride;
        if (next < rows) prefetch(next, stage ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        const float mu = mu_n, rs = rs_n;
        if (next < rows) { mu_n = __ldg(mean + next); rs_n = __ldg(rstd + next); }
        asm volatile("cp.async.wait_group 1;" ::: "memory");          // everything but the newest group: this row has landed
        __syncwarp();
        const float* bx = wbuf + stage * 2 * W;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
            const int c = lane + 32 * t;
            if (c < nvec) {
                const float4 xv = ld4(bx + 4 * c), d = ld4(bx + W + 4 * c), g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
                const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                ab[t] = add4(ab[t], d);
                ag[t] = add4(ag[t], make_float4(d.x * xh.x, d.y * xh.y, d.z * xh.z, d.w * xh.w));
                const float4 dg = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);  // dy * gamma
                s1 += (dg.x + dg.y) + (dg.z + dg.w);
                s2 += (dg.x * xh.x + dg.y * xh.y) + (dg.z * xh.z + dg.w * xh.w);
            }
        }
        s1 = warp_sum(s1) / (float)D;
        s2 = warp_sum(s2) / (float)D;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
            const int c = lane + 32 * t;
            if (c < nvec) {
                const float4 xv = ld4(bx + 4 * c), d = ld4(bx + W + 4 * c), g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
                const float4 xh = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
                float4 o = make_float4(rs * (d.x * g.x - s1 - xh.x * s2), rs * (d.y * g.y - s1 - xh.y * s2),
                                       rs * (d.z * g.z - s1 - xh.z * s2), rs * (d.w * g.w - s1 - xh.w * s2));
                if (dx_add) o = add4(o, ld4(dx_add + row * D + 4 * c));
                if (dx) *reinterpret_cast<float4*>(dx + row * D + 4 * c) = o;
                if (dx16 || dbias) {
                    if (p > 0.f) {
                        const float4 s4 = dropout_scale4((uint64_t)(row * D + 4 * c), p, inv_keep, seed, offset);
                        o = make_float4(o.x * s4.x, o.y * s4.y, o.z * s4.z, o.w * s4.w);
                    }
                    if (dx16) *reinterpret_cast<uint2*>(dx16 + row * D + 4 * c) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
                    if (dbias) ad[t] = add4(ad[t], o);
                }
            }
        }
        __syncwarp();                                                 // this stage is overwritten by the prefetch of iteration it+1
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    // dgamma / dbeta / dbias: reduce the warps through smem (the staging buffers are free now), then ONE 16-byte vector
    // atomic per 4 columns per CTA
    if (!dgamma && !dbeta && !dbias) return;
    float4* red = reinterpret_cast<float4*>(ln_smem);                 // [kBwdWarps][NV*32] float4
    for (int pass = 0; pass < 3; ++pass) {
        float* dst = pass == 0 ? dgamma : pass == 1 ? dbeta : dbias;
        if (!dst) continue;
        __syncthreads();                                              // staging buffers / previous pass fully consumed
#pragma unroll
        for (int t = 0; t < NV; ++t) red[warp * (NV * 32) + lane + 32 * t] = pass == 0 ? ag[t] : pass == 1 ? ab[t] : ad[t];
        __syncthreads();
        const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
        for (int c = threadIdx.x; c < nvec; c += kBwdWarps * 32) {
            float4 sum = red[c];
#pragma unroll
            for (int w = 1; w < kBwdWarps; ++w) sum = add4(sum, red[w * (NV * 32) + c]);
            if (vec) {
                atomicAdd(reinterpret_cast<float4*>(dst) + c, sum);
            } else {
                atomicAdd(dst + 4 * c, sum.x); atomicAdd(dst + 4 * c + 1, sum.y); atomicAdd(dst + 4 * c + 2, sum.z); atomicAdd(dst + 4 * c + 3, sum.w);
            }
        }
    }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        if (i + 8 <= n) {
            const float4 a = ld4(x + i), b = ld4(x + i + 4);
            *reinterpret_cast<uint4*>(y + i) = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
        } else {
            for (int64_t j = i; j < n; ++j) y[j] = __float2bfloat16_rn(x[j]);
        }
    }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int64_t n) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __bfloat162float(x[i]);
}
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) *reinterpret_cast<float4*>(y + i) = add4(ld4(a + i), ld4(b + i));
        else for (int64_t j = i; j < n; ++j) y[j] = a[j] + b[j];
    }
}

// out[c] += sum over a slab of rows; block (32 cols, 8 row lanes)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t m, int64_t n, int64_t ldx, int64_t rows_per_block) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    __shared__ float red[8][33];
    const int64_t col = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
    float s = 0.f;
    if (col < n)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += to_f32<T>(x[r * ldx + col]);
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < n) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        atomicAdd(out + col, t);
    }
}

// Vector form (n % 8 == 0, 16-byte aligned rows): a lane owns 8 consecutive columns (one 16-byte load of bf16, two of fp32), a
// warp 256 columns of a row (512 contiguous bytes), the 8 warps of a CTA take rows r, r+8, ... of the CTA's slab four at a time
// (four independent loads in flight per lane); warps are reduced through shared memory, then one 16-byte atomic per 4 columns.
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t m, int64_t n, int64_t ldx,
                                                         int64_t rows_per_block) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    __shared__ float4 red[8][64];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t col = (int64_t)blockIdx.x * 256 + lane * 8;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(m, r0 + rows_per_block);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    auto load8v = [&](int64_t r, float (&v)[8]) {
        if constexpr (sizeof(T) == 2) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + r * ldx + col));
            const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y), p2 = unpack_bf16x2(u.z), p3 = unpack_bf16x2(u.w);
            v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y; v[4] = p2.x; v[5] = p2.y; v[6] = p3.x; v[7] = p3.y;
        } else {
            const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ldx + col)), b = __ldg(reinterpret_cast<const float4*>(x + r * ldx + col) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        }
    };
    if (col < n) {
        int64_t r = r0 + warp;
        for (; r + 24 < r1; r += 32) {
            float v0[8], v1[8], v2[8], v3[8];
            load8v(r, v0); load8v(r + 8, v1); load8v(r + 16, v2); load8v(r + 24, v3);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += (v0[j] + v1[j]) + (v2[j] + v3[j]);
        }
        for (; r < r1; r += 8) {
            float v0[8];
            load8v(r, v0);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v0[j];
        }
    }
    red[warp][2 * lane] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    red[warp][2 * lane + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    __syncthreads();
    if (threadIdx.x < 64) {
        const int64_t c4 = (int64_t)blockIdx.x * 256 + threadIdx.x * 4;
        if (c4 < n) {
            float4 sum = red[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < 8; ++w) sum = add4(sum, red[w][threadIdx.x]);
            atomicAdd(reinterpret_cast<float4*>(out + c4), sum);
        }
    }
}

template <typename TO>
__global__ void dropout_apply_kernel(const float* __restrict__ x, TO* __restrict__ y, int64_t n, float p, float inv_keep, uint64_t seed, uint64_t offset_,
                                     const uint64_t* offset_base) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const uint64_t offset = p > 0.f ? effective_offset(offset_, offset_base) : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p > 0.f) s4 = dropout_scale4((uint64_t)i, p, inv_keep, seed, offset);
        const float sc[4] = {s4.x, s4.y, s4.z, s4.w};
        for (int j = 0; j < 4 && i + j < n; ++j) y[i + j] = from_f32<TO>(x[i + j] * sc[j]);
    }
}

__global__ void dropout_mask_kernel(float* out, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_base) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = dropout_scale((uint64_t)i, p, 1.0f, seed, effective_offset(offset, offset_base));
}
__global__ void advance_u64_kernel(uint64_t* word, uint64_t delta) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger(); *word += delta; }

// FeatureEmbedding front-end (vision_embeddings.py:19-25 + models/utils.py:44-57): ONE pass over the fp32 region features writes the
// bf16 GEMM operand AND the padding mask (a row whose features sum to 0 is padding: additive mask_value there, 0 elsewhere).  Warp per row,
// 16-byte loads; the reference reads the features twice (torch.sum for the mask, then the Linear) plus a cast under autocast.
__global__ void __launch_bounds__(256) cast_rows_padmask_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, float* __restrict__ mask,
                                                                int64_t rows, int64_t K, int64_t ldx, float mask_value) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= rows) return;
    const float* xr = x + row * ldx;
    __nv_bfloat16* yr = y + row * K;
    float s = 0.f;
    const int64_t nv = K / 8;                                    // 8 elements per lane and iteration: two 16-byte loads, one 16-byte store
    for (int64_t c = lane; c < nv; c += 32) {
        const float4 a = ld4(xr + 8 * c), b = ld4(xr + 8 * c + 4);
        s += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
        *reinterpret_cast<uint4*>(yr + 8 * c) = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    for (int64_t c = nv * 8 + lane; c < K; c += 32) { const float v = xr[c]; s += v; yr[c] = __float2bfloat16_rn(v); }
    s = warp_sum(s);
    if (lane == 0 && mask) mask[row] = s == 0.0f ? mask_value : 0.0f;
}

// the padding mask alone, for features that already arrive as bf16 (pinned bf16 staging of the input feed): warp per row
__global__ void __launch_bounds__(256) rows_padmask_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ mask, int64_t rows, int64_t K,
                                                                int64_t ldx, float mask_value) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (row >= rows) return;
    const __nv_bfloat16* xr = x + row * ldx;
    float s = 0.f;
    const int64_t nv = K / 8;
    for (int64_t c = lane; c < nv; c += 32) {
        const uint4 u = *reinterpret_cast<const uint4*>(xr + 8 * c);
        const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c2 = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        s += ((a.x + a.y) + (b.x + b.y)) + ((c2.x + c2.y) + (d.x + d.y));
    }
    for (int64_t c = nv * 8 + lane; c < K; c += 32) s += __bfloat162float(xr[c]);
    s = warp_sum(s);
    if (lane == 0) mask[row] = s == 0.0f ? mask_value : 0.0f;
}

// out_bf16 = a (fp32) * b (fp32 | bf16): the gradient of a fused activation from its saved derivative
template <typename TB>
__global__ void mul_to_bf16_kernel(const float* __restrict__ a, const TB* __restrict__ b, __nv_bfloat16* __restrict__ y, int64_t n) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        if (i + 4 <= n) {
            const float4 av = ld4(a + i);
            float b0, b1, b2, b3;
            if constexpr (sizeof(TB) == 4) { const float4 bv = ld4(reinterpret_cast<const float*>(b) + i); b0 = bv.x; b1 = bv.y; b2 = bv.z; b3 = bv.w; }
            else { const uint2 u = *reinterpret_cast<const uint2*>(b + i); const float2 p0 = unpack_bf16x2(u.x), p1 = unpack_bf16x2(u.y); b0 = p0.x; b1 = p0.y; b2 = p1.x; b3 = p1.y; }
            *reinterpret_cast<uint2*>(y + i) = make_uint2(pack_bf16x2(av.x * b0, av.y * b1), pack_bf16x2(av.z * b2, av.w * b3));
        } else {
            for (int64_t j = i; j < n; ++j) y[j] = __float2bfloat16_rn(a[j] * to_f32<TB>(b[j]));
        }
    }
}

// many (src fp32 -> dst fp32|bf16) contiguous copies in one launch: blockIdx.y = descriptor, blockIdx.x strides over it
__global__ void stage_many_kernel(const ovqa_stage_desc* __restrict__ descs) {
    pdl_wait();          // programmatic dependent launch (common.cuh): nothing above touches global memory
    pdl_trigger();
    const ovqa_stage_desc d = descs[blockIdx.y];
    const float* __restrict__ x = d.src;
    const int64_t n = d.n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 8;
    const bool vec = ((reinterpret_cast<uintptr_t>(d.src) | reinterpret_cast<uintptr_t>(d.dst)) & 15) == 0;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
        if (vec && i + 8 <= n) {
            const float4 a = ld4(x + i), b = ld4(x + i + 4);
            if (d.dst_dtype == OVQA_BF16) {
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.dst) + i) =
                    make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
            } else {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.dst) + i) = a;
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.dst) + i + 4) = b;
            }
        } else {
            for (int64_t j = i; j < n && j < i + 8; ++j) {
                if (d.dst_dtype == OVQA_BF16) reinterpret_cast<__nv_bfloat16*>(d.dst)[j] = __float2bfloat16_rn(x[j]);
                else reinterpret_cast<float*>(d.dst)[j] = x[j];
            }
        }
    }
}

int grid_for(int64_t work_items, int threads) {
    int64_t g = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

}  // namespace ovqa

using namespace ovqa;

// kernel instantiation whose per-lane register row (NV float4) just covers D
#define OVQA_LN_DISPATCH(K) (nv <= 1 ? K<1> : nv <= 2 ? K<2> : nv <= 3 ? K<3> : nv <= 4 ? K<4> : nv <= 6 ? K<6> : K<8>)

extern "C" int ovqa_layernorm_fwd(const float* x, const float* residual, const float* gamma, const float* beta,
                                  const float* post_add, int64_t post_add_rows, float* y_f32, void* y_bf16, float* sum_out,
                                  float* mean, float* rstd, int64_t rows, int64_t D, float eps, void* stream) {
    if (!x || !gamma || !beta || rows < 0 || D <= 0) return OVQA_ERR_BAD_ARG;
    if (D % 4 != 0 || D > kMaxVec * 128) return OVQA_ERR_BAD_SHAPE;
    if (rows == 0) return OVQA_OK;
    if (post_add && post_add_rows <= 0) return OVQA_ERR_BAD_ARG;
    const int nv = (int)((D + 127) / 128);
    OVQA_CUDA_OK(launch_pdl(OVQA_LN_DISPATCH(layernorm_fwd_kernel), dim3((unsigned)((rows + kRowsPerCta - 1) / kRowsPerCta)), dim3(kRowsPerCta * 32), (size_t)(0), (cudaStream_t)stream, 
        x, residual, gamma, beta, post_add, post_add ? post_add_rows : 1, y_f32, (__nv_bfloat16*)y_bf16, sum_out, mean, rstd, rows, (int)D, eps));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

extern "C" int ovqa_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                  const float* dx_add, float* dx, void* dx_bf16, float* dgamma, float* dbeta, float* dbias,
                                  int64_t rows, int64_t D, float dropout_p, uint64_t seed, uint64_t offset, const uint64_t* offset_base,
                                  void* stream) {
    if (!dy || !x || !mean || !rstd || !gamma || rows < 0 || D <= 0) return OVQA_ERR_BAD_ARG;
    if (D % 4 != 0 || D > kMaxVec * 128) return OVQA_ERR_BAD_SHAPE;
    if (dropout_p < 0.f || dropout_p >= 1.f) return OVQA_ERR_BAD_ARG;
    if (rows == 0) return OVQA_OK;
    // persistent: as many CTAs as stay resident (two rows of cp.async per warp in flight), never more than the rows need
    const int nv = (int)((D + 127) / 128);
    const size_t smem = (size_t)kBwdWarps * 4 * nv * 128 * sizeof(float);        // >= the [kBwdWarps][nv*32] float4 reduction scratch
    auto kern = OVQA_LN_DISPATCH(layernorm_bwd_kernel);
    if (smem > 48 * 1024) OVQA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    OVQA_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBwdWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t g = (rows + kBwdWarps - 1) / kBwdWarps;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (g > cap) g = cap;
    OVQA_CUDA_OK(launch_pdl(kern, dim3((unsigned)g), dim3(kBwdWarps * 32), (size_t)(smem), (cudaStream_t)stream, 
        dy, x, mean, rstd, gamma, dx_add, dx, (__nv_bfloat16*)dx_bf16, dgamma, dbeta, dbias, rows, (int)D, dropout_p,
        dropout_p > 0.f ? 1.0f / (1.0f - dropout_p) : 1.0f, seed, offset, offset_base));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

extern "C" int ovqa_cast_f32_to_bf16(const float* x, void* y, int64_t n, void* stream) {
    if (!x || !y || n < 0) return OVQA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return OVQA_ERR_ALIGNMENT;
    if (n == 0) return OVQA_OK;
    OVQA_CUDA_OK(launch_pdl(cast_f32_bf16_kernel, dim3(grid_for((n + 7) / 8, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, x, (__nv_bfloat16*)y, n));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_cast_bf16_to_f32(const void* x, float* y, int64_t n, void* stream) {
    if (!x || !y || n < 0) return OVQA_ERR_BAD_ARG;
    if (n == 0) return OVQA_OK;
    OVQA_CUDA_OK(launch_pdl(cast_bf16_f32_kernel, dim3(grid_for(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)x, y, n));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_add_f32(const float* a, const float* b, float* y, int64_t n, void* stream) {
    if (!a || !b || !y || n < 0) return OVQA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return OVQA_ERR_ALIGNMENT;
    if (n == 0) return OVQA_OK;
    OVQA_CUDA_OK(launch_pdl(add_f32_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, a, b, y, n));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_colsum(const void* x, float* out, int64_t m, int64_t n, int64_t ldx, int dtype, int accumulate, void* stream) {
    if (!x || !out || m < 0 || n <= 0) return OVQA_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (!accumulate) OVQA_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float) * n, st));
    if (m == 0) return OVQA_OK;
    const int esz = dtype == OVQA_F32 ? 4 : 2;
    if ((dtype == OVQA_F32 || dtype == OVQA_BF16) && n % 8 == 0 && (ldx * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const int64_t cb = (n + 255) / 256;
        int64_t rb = ((int64_t)sm_count() * 6 + cb - 1) / cb;              // ~6 CTAs of 8 warps per SM
        if (rb > (m + 63) / 64) rb = (m + 63) / 64;
        if (rb < 1) rb = 1;
        const int64_t rpb_v = ((m + rb - 1) / rb + 7) / 8 * 8;
        dim3 grid_v((unsigned)cb, (unsigned)((m + rpb_v - 1) / rpb_v));
        if (dtype == OVQA_F32) OVQA_CUDA_OK(launch_pdl(colsum_vec_kernel<float>, dim3(grid_v), dim3(256), (size_t)(0), st, (const float*)x, out, m, n, ldx, rpb_v));
        else OVQA_CUDA_OK(launch_pdl(colsum_vec_kernel<__nv_bfloat16>, dim3(grid_v), dim3(256), (size_t)(0), st, (const __nv_bfloat16*)x, out, m, n, ldx, rpb_v));
        OVQA_LAUNCH_OK();
        return OVQA_OK;
    }
    const int64_t col_blocks = (n + 31) / 32;
    int64_t row_blocks = (int64_t)sm_count() * 4 / col_blocks;
    if (row_blocks < 1) row_blocks = 1;
    if (row_blocks > (m + 63) / 64) row_blocks = (m + 63) / 64;
    const int64_t rpb = (m + row_blocks - 1) / row_blocks;
    dim3 grid((unsigned)col_blocks, (unsigned)row_blocks), block(32, 8);
    if (dtype == OVQA_F32) OVQA_CUDA_OK(launch_pdl(colsum_kernel<float>, dim3(grid), dim3(block), (size_t)(0), st, (const float*)x, out, m, n, ldx, rpb));
    else if (dtype == OVQA_BF16) OVQA_CUDA_OK(launch_pdl(colsum_kernel<__nv_bfloat16>, dim3(grid), dim3(block), (size_t)(0), st, (const __nv_bfloat16*)x, out, m, n, ldx, rpb));
    else return OVQA_ERR_BAD_ARG;
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_dropout_apply(const float* x, void* y, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_base,
                                  int out_dtype, void* stream) {
    if (!x || !y || n < 0 || p < 0.f || p >= 1.f) return OVQA_ERR_BAD_ARG;
    if (n == 0) return OVQA_OK;
    const float inv = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
    if (out_dtype == OVQA_F32) OVQA_CUDA_OK(launch_pdl(dropout_apply_kernel<float>, dim3(grid_for((n + 3) / 4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, x, (float*)y, n, p, inv, seed, offset, offset_base));
    else if (out_dtype == OVQA_BF16) OVQA_CUDA_OK(launch_pdl(dropout_apply_kernel<__nv_bfloat16>, dim3(grid_for((n + 3) / 4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, x, (__nv_bfloat16*)y, n, p, inv, seed, offset, offset_base));
    else return OVQA_ERR_BAD_ARG;
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_advance_u64(uint64_t* word, uint64_t delta, void* stream) {
    if (!word) return OVQA_ERR_BAD_ARG;
    OVQA_CUDA_OK(launch_pdl(advance_u64_kernel, dim3(1), dim3(1), (size_t)(0), (cudaStream_t)stream, word, delta));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_stage_many(const ovqa_stage_desc* descs, int64_t n_desc, void* stream) {
    if (n_desc < 0 || (n_desc > 0 && !descs)) return OVQA_ERR_BAD_ARG;
    if (n_desc == 0) return OVQA_OK;
    if (n_desc > 65535) return OVQA_ERR_BAD_SHAPE;
    OVQA_CUDA_OK(launch_pdl(stage_many_kernel, dim3(dim3(32, (unsigned)n_desc)), dim3(256), (size_t)(0), (cudaStream_t)stream, descs));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_dropout_mask(float* out, int64_t n, float p, uint64_t seed, uint64_t offset, const uint64_t* offset_base, void* stream) {
    if (!out || n < 0) return OVQA_ERR_BAD_ARG;
    if (n == 0) return OVQA_OK;
    OVQA_CUDA_OK(launch_pdl(dropout_mask_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, out, n, p, seed, offset, offset_base));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}

extern "C" int ovqa_cast_rows_padmask(const float* x, void* y_bf16, float* mask, int64_t rows, int64_t K, int64_t ldx, float mask_value, void* stream) {
    if (!x || !y_bf16 || rows < 0 || K <= 0 || ldx < K) return OVQA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y_bf16) & 15) || (ldx % 4) || (K % 8)) return OVQA_ERR_ALIGNMENT;
    if (rows == 0) return OVQA_OK;
    OVQA_CUDA_OK(launch_pdl(cast_rows_padmask_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)(0), (cudaStream_t)stream, x, (__nv_bfloat16*)y_bf16, mask, rows, K, ldx, mask_value));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_rows_padmask_bf16(const void* x_bf16, float* mask, int64_t rows, int64_t K, int64_t ldx, float mask_value, void* stream) {
    if (!x_bf16 || !mask || rows < 0 || K <= 0 || ldx < K) return OVQA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x_bf16) & 15) || (ldx % 8)) return OVQA_ERR_ALIGNMENT;
    if (rows == 0) return OVQA_OK;
    OVQA_CUDA_OK(launch_pdl(rows_padmask_bf16_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), (size_t)(0), (cudaStream_t)stream, (const __nv_bfloat16*)x_bf16, mask, rows, K, ldx, mask_value));
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
extern "C" int ovqa_mul_to_bf16(const float* a, const void* b, int b_dtype, void* y_bf16, int64_t n, void* stream) {
    if (!a || !b || !y_bf16 || n < 0) return OVQA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15) || (reinterpret_cast<uintptr_t>(y_bf16) & 7)) return OVQA_ERR_ALIGNMENT;
    if (n == 0) return OVQA_OK;
    if (b_dtype == OVQA_F32) OVQA_CUDA_OK(launch_pdl(mul_to_bf16_kernel<float>, dim3(grid_for((n + 3) / 4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, a, (const float*)b, (__nv_bfloat16*)y_bf16, n));
    else if (b_dtype == OVQA_BF16) OVQA_CUDA_OK(launch_pdl(mul_to_bf16_kernel<__nv_bfloat16>, dim3(grid_for((n + 3) / 4, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, a, (const __nv_bfloat16*)b, (__nv_bfloat16*)y_bf16, n));
    else return OVQA_ERR_BAD_ARG;
    OVQA_LAUNCH_OK();
    return OVQA_OK;
}
