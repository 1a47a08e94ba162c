// pm_norm.cu — the two normalisations of the reference's networks at the shapes PyTorch's kernels handle worst
// (the PPO update around the PyTorch networks), as sm_100a kernels:
//
//  pm_add_layernorm32   y = LayerNorm_32(a + b) * gamma + beta — the post-norm residual of the critic's
//                       nn.TransformerEncoderLayer (d_model 32, pacman_mappo_resnet.py:131-136): norm1(x + attn(x)),
//                       norm2(x + ff(x)) over [tokens x batch] = 630 784 rows of 32 floats per minibatch.  PyTorch's
//                       layer-norm kernel gives every 128-byte row a CTA (120 us per call, 18 % of a PPO update);
//                       here 8 lanes share a row (one float4 each, three shuffles per moment), a warp covers 512
//                       contiguous bytes, the residual add is fused in, and the backward pass needs no saved x-hat.
//  pm_groupnorm_4x8     GroupNorm(4, 32) of the actor's ResidualBlocks (`:49-68`) on [B][32][H][W] (NCHW): a group is
//                       8 channels x H*W contiguous floats; one warp per (sample, group) keeps the group in registers
//                       between the moment pass and the write, with the block's GELU and skip connection fused in:
//                       `act(gn1(conv1(x)))` and `act(gn2(conv2(.)) + x)` are one kernel each.  Boards above 192 cells
//                       (bloxCapture: 400) use one CTA of 256 threads per (sample, group) with the same arithmetic.
// Both are memory-bound elementwise-plus-row-reduction ops: 2 x 4 B per element forward, 3-4 x backward.  fp32, biased
// variance, eps inside the square root, like torch.  Per-column (gamma / beta) gradients are accumulated per block and
// summed by the caller from a [blocks][2*C] partial buffer (deterministic).
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/pacman_b200.h"

#define LN_D 32
#define LN_THREADS 256

__device__ __forceinline__ float sum8(float v) {          // sum over the 8 lanes that share a row
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}

__global__ void __launch_bounds__(LN_THREADS) pm_add_layernorm32_fwd_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b,
                                                                            const float4 *__restrict__ gamma, const float4 *__restrict__ beta,
                                                                            float4 *__restrict__ y, float2 *__restrict__ stats, int64_t rows,
                                                                            float eps) {
    const int sub = threadIdx.x & 7;                        // which float4 of the row
    const float4 g = __ldg(gamma + sub), be = __ldg(beta + sub);
    const int64_t stride = (int64_t)gridDim.x * (LN_THREADS / 8);
    // the trip count is the same for all lanes of a warp (its four rows are consecutive): the shuffles below are
    // executed by the whole warp, rows past the end just carry zeros and are not written
    for (int64_t r0 = (int64_t)blockIdx.x * (LN_THREADS / 8) + ((threadIdx.x >> 5) << 2); r0 < rows; r0 += stride) {
        const int64_t r = r0 + ((threadIdx.x >> 3) & 3);
        const bool live = r < rows;
        float4 v = live ? __ldg(a + r * 8 + sub) : make_float4(0, 0, 0, 0);
        if (b && live) { const float4 w = __ldg(b + r * 8 + sub); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        const float mean = sum8((v.x + v.y) + (v.z + v.w)) * (1.0f / LN_D);
        const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
        const float var = sum8(fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, dw * dw)))) * (1.0f / LN_D);
        const float rstd = rsqrtf(var + eps);
        float4 o;
        o.x = fmaf(dx * rstd, g.x, be.x); o.y = fmaf(dy * rstd, g.y, be.y); o.z = fmaf(dz * rstd, g.z, be.z); o.w = fmaf(dw * rstd, g.w, be.w);
        if (live) y[r * 8 + sub] = o;
        if (stats && live && sub == 0) stats[r] = make_float2(mean, rstd);
    }
}

// dx = rstd * (dxh - mean(dxh) - xh * mean(dxh * xh)) with dxh = dy * gamma; x-hat is recomputed from the forward call's
// inputs `a`, `b` and the saved (mean, rstd), so nothing but 8 bytes per row is kept for the backward pass.  dx is the
// gradient of both `a` and `b`.  partial[block][0..31] = this block's share of dgamma, [32..63] of dbeta.
__global__ void __launch_bounds__(LN_THREADS) pm_add_layernorm32_bwd_kernel(const float4 *__restrict__ a, const float4 *__restrict__ b,
                                                                            const float4 *__restrict__ gamma, const float2 *__restrict__ stats,
                                                                            const float4 *__restrict__ dy, float4 *__restrict__ dx,
                                                                            float *__restrict__ partial, int64_t rows) {
    __shared__ float red[LN_THREADS / 32][2 * LN_D];
    const int sub = threadIdx.x & 7;
    const float4 g = __ldg(gamma + sub);
    float4 dg = make_float4(0, 0, 0, 0), db = dg;
    const int64_t stride = (int64_t)gridDim.x * (LN_THREADS / 8);
    for (int64_t r0 = (int64_t)blockIdx.x * (LN_THREADS / 8) + ((threadIdx.x >> 5) << 2); r0 < rows; r0 += stride) {
        const int64_t r = r0 + ((threadIdx.x >> 3) & 3);
        const bool live = r < rows;                         // (warp-uniform trip count, see the forward kernel)
        float4 v = live ? __ldg(a + r * 8 + sub) : make_float4(0, 0, 0, 0);
        if (b && live) { const float4 w = __ldg(b + r * 8 + sub); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        const float2 st = live ? __ldg(stats + r) : make_float2(0.f, 0.f);
        const float4 d = live ? __ldg(dy + r * 8 + sub) : make_float4(0, 0, 0, 0);
        const float hx = (v.x - st.x) * st.y, hy = (v.y - st.x) * st.y, hz = (v.z - st.x) * st.y, hw = (v.w - st.x) * st.y;
        const float ex = d.x * g.x, ey = d.y * g.y, ez = d.z * g.z, ew = d.w * g.w;
        const float m1 = sum8((ex + ey) + (ez + ew)) * (1.0f / LN_D);
        const float m2 = sum8(fmaf(ex, hx, fmaf(ey, hy, fmaf(ez, hz, ew * hw)))) * (1.0f / LN_D);
        float4 o;
        o.x = st.y * (ex - m1 - hx * m2); o.y = st.y * (ey - m1 - hy * m2); o.z = st.y * (ez - m1 - hz * m2); o.w = st.y * (ew - m1 - hw * m2);
        if (live) dx[r * 8 + sub] = o;
        dg.x = fmaf(d.x, hx, dg.x); dg.y = fmaf(d.y, hy, dg.y); dg.z = fmaf(d.z, hz, dg.z); dg.w = fmaf(d.w, hw, dg.w);
        db.x += d.x; db.y += d.y; db.z += d.z; db.w += d.w;
    }
    // lanes with the same `sub` hold the same columns: fold the four row-groups of a warp, then the warps of the block
    float acc[8] = {dg.x, dg.y, dg.z, dg.w, db.x, db.y, db.z, db.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) { acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 8); acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 16); }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < 8) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { red[warp][4 * lane + k] = acc[k]; red[warp][LN_D + 4 * lane + k] = acc[4 + k]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * LN_D) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_THREADS / 32; ++w) s += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 2 * LN_D + threadIdx.x] = s;
    }
}

static unsigned ln_grid(int64_t rows) {
    const int64_t want = (rows + LN_THREADS / 8 - 1) / (LN_THREADS / 8);
    return (unsigned)(want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8);
}

extern "C" int pm_layernorm32_blocks(int64_t rows) { return (int)ln_grid(rows); }

extern "C" int pm_add_layernorm32_fwd(const float *a, const float *b, const float *gamma, const float *beta, float *y, float *stats,
                                      int64_t rows, double eps, pm_stream_t stream) {
    if (!a || !gamma || !beta || !y || rows <= 0 || (((size_t)a | (size_t)b | (size_t)gamma | (size_t)beta | (size_t)y) & 15) ||
        ((size_t)stats & 7))
        return PM_EINVAL;
    pm_add_layernorm32_fwd_kernel<<<ln_grid(rows), LN_THREADS, 0, (cudaStream_t)stream>>>(
        (const float4 *)a, (const float4 *)b, (const float4 *)gamma, (const float4 *)beta, (float4 *)y, (float2 *)stats, rows, (float)eps);
    return (int)cudaGetLastError();
}

extern "C" int pm_add_layernorm32_bwd(const float *a, const float *b, const float *gamma, const float *stats, const float *dy, float *dx,
                                      float *partial, int64_t rows, pm_stream_t stream) {
    if (!a || !gamma || !stats || !dy || !dx || !partial || rows <= 0 ||
        (((size_t)a | (size_t)b | (size_t)gamma | (size_t)dy | (size_t)dx) & 15) || ((size_t)stats & 7))
        return PM_EINVAL;
    pm_add_layernorm32_bwd_kernel<<<ln_grid(rows), LN_THREADS, 0, (cudaStream_t)stream>>>(
        (const float4 *)a, (const float4 *)b, (const float4 *)gamma, (const float2 *)stats, (const float4 *)dy, (float4 *)dx, partial, rows);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm(4, 32) on NCHW: group = 8 channels x HW contiguous floats, one warp per (sample, group)
// ---------------------------------------------------------------------------------------------------------
#define GN_C 32
#define GN_G 4
#define GN_CPG 8
#define GN_MAXV 12                        // float4s per lane: group sizes up to 32 * 12 * 4 = 1536 floats (HW <= 192) stay in registers
#define GN_THREADS 128

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

// channel (within the group) of element j of float4 number i: a float4 may straddle two channels when HW % 4 != 0
__device__ __forceinline__ void gn_channels(int i, int HW, int &c_lo, int &split) {
    c_lo = (4 * i) / HW;
    split = (c_lo + 1) * HW - 4 * i;                      // elements j >= split belong to channel c_lo + 1 (split >= 4: none)
}

// y = [gelu](groupnorm(x) [+ res]); 8 * HW / 4 <= 32 * GN_MAXV for the register path (host checks)
template <int GELU>
__global__ void __launch_bounds__(GN_THREADS) pm_groupnorm_fwd_kernel(const float *__restrict__ x, const float *__restrict__ res,
                                                                      const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                      float *__restrict__ y, float2 *__restrict__ stats, int64_t groups, int HW,
                                                                      float eps) {
    const int lane = threadIdx.x & 31;
    const int n4 = GN_CPG * HW / 4;                       // float4s per group
    const float inv = 1.0f / (float)(GN_CPG * HW);
    for (int64_t gi = (int64_t)blockIdx.x * (GN_THREADS / 32) + (threadIdx.x >> 5); gi < groups; gi += (int64_t)gridDim.x * (GN_THREADS / 32)) {
        const float4 *src = reinterpret_cast<const float4 *>(x) + gi * n4;
        float4 v[GN_MAXV];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < GN_MAXV; ++k) {
            const int i = lane + 32 * k;
            v[k] = i < n4 ? __ldg(src + i) : make_float4(0, 0, 0, 0);
            s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
        const float mean = warp_sum(s) * inv;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < GN_MAXV; ++k) {
            if (lane + 32 * k < n4) {
                const float a0 = v[k].x - mean, a1 = v[k].y - mean, a2 = v[k].z - mean, a3 = v[k].w - mean;
                q += fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, a3 * a3)));
            }
        }
        const float rstd = rsqrtf(warp_sum(q) * inv + eps);
        const int c0 = (int)(gi % GN_G) * GN_CPG;
        float4 *dst = reinterpret_cast<float4 *>(y) + gi * n4;
        const float4 *rs = res ? reinterpret_cast<const float4 *>(res) + gi * n4 : nullptr;
#pragma unroll
        for (int k = 0; k < GN_MAXV; ++k) {
            const int i = lane + 32 * k;
            if (i < n4) {
                int cl, split;
                gn_channels(i, HW, cl, split);
                const float g0 = __ldg(gamma + c0 + cl) * rstd, b0 = __ldg(beta + c0 + cl) - mean * g0;
                float g1 = g0, b1 = b0;
                if (split < 4) { g1 = __ldg(gamma + c0 + cl + 1) * rstd; b1 = __ldg(beta + c0 + cl + 1) - mean * g1; }
                float4 o = make_float4(fmaf(v[k].x, split > 0 ? g0 : g1, split > 0 ? b0 : b1), fmaf(v[k].y, split > 1 ? g0 : g1, split > 1 ? b0 : b1),
                                       fmaf(v[k].z, split > 2 ? g0 : g1, split > 2 ? b0 : b1), fmaf(v[k].w, split > 3 ? g0 : g1, split > 3 ? b0 : b1));
                if (rs) { const float4 r = __ldg(rs + i); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
                if (GELU) { o.x = gelu_f(o.x); o.y = gelu_f(o.y); o.z = gelu_f(o.z); o.w = gelu_f(o.w); }
                dst[i] = o;
            }
        }
        if (stats && lane == 0) stats[gi] = make_float2(mean, rstd);
    }
}

// Backward of y = [gelu](z), z = xh * gamma + beta [+ res], xh = (x - mean) * rstd:
//   dz = dy * gelu'(z) (z recomputed), dres = dz, dxh = dz * gamma, dx = rstd * (dxh - mean_g(dxh) - xh * mean_g(dxh * xh)),
//   dgamma[c] += sum dz * xh, dbeta[c] += sum dz   (per-block partials [blocks][2*32])
template <int GELU>
__global__ void __launch_bounds__(GN_THREADS) pm_groupnorm_bwd_kernel(const float *__restrict__ x, const float *__restrict__ res,
                                                                      const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                      const float2 *__restrict__ stats, const float *__restrict__ dy,
                                                                      float *__restrict__ dx, float *__restrict__ dres, float *__restrict__ partial,
                                                                      int64_t groups, int HW) {
    __shared__ float red[2 * GN_C];
    for (int i = threadIdx.x; i < 2 * GN_C; i += GN_THREADS) red[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n4 = GN_CPG * HW / 4;
    const float inv = 1.0f / (float)(GN_CPG * HW);
    for (int64_t gi = (int64_t)blockIdx.x * (GN_THREADS / 32) + (threadIdx.x >> 5); gi < groups; gi += (int64_t)gridDim.x * (GN_THREADS / 32)) {
        const float4 *sx = reinterpret_cast<const float4 *>(x) + gi * n4, *sd = reinterpret_cast<const float4 *>(dy) + gi * n4;
        const float4 *rs = res ? reinterpret_cast<const float4 *>(res) + gi * n4 : nullptr;
        float4 *dr = dres ? reinterpret_cast<float4 *>(dres) + gi * n4 : nullptr;
        const float2 st = __ldg(stats + gi);
        const int c0 = (int)(gi % GN_G) * GN_CPG;
        float4 xh[GN_MAXV], e[GN_MAXV];                   // x-hat and dxh = dz * gamma
        float s1 = 0.f, s2 = 0.f;
        float dgc[GN_CPG], dbc[GN_CPG];
#pragma unroll
        for (int c = 0; c < GN_CPG; ++c) { dgc[c] = 0.f; dbc[c] = 0.f; }
#pragma unroll
        for (int k = 0; k < GN_MAXV; ++k) {
            const int i = lane + 32 * k;
            xh[k] = make_float4(0, 0, 0, 0); e[k] = xh[k];
            if (i < n4) {
                const float4 v = __ldg(sx + i);
                const float4 d4 = __ldg(sd + i);
                int cl, split;
                gn_channels(i, HW, cl, split);
                const float g0 = __ldg(gamma + c0 + cl), g1 = split < 4 ? __ldg(gamma + c0 + cl + 1) : g0;
                float h[4] = {(v.x - st.x) * st.y, (v.y - st.x) * st.y, (v.z - st.x) * st.y, (v.w - st.x) * st.y};
                float d[4] = {d4.x, d4.y, d4.z, d4.w};
                float ga[4] = {split > 0 ? g0 : g1, split > 1 ? g0 : g1, split > 2 ? g0 : g1, split > 3 ? g0 : g1};
                if (GELU) {
                    const float b0 = __ldg(beta + c0 + cl), b1 = split < 4 ? __ldg(beta + c0 + cl + 1) : b0;
                    float r[4] = {0.f, 0.f, 0.f, 0.f};
                    if (rs) { const float4 r4 = __ldg(rs + i); r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w; }
#pragma unroll
                    for (int j = 0; j < 4; ++j) d[j] *= gelu_grad_f(fmaf(h[j], ga[j], (split > j ? b0 : b1)) + r[j]);
                }
                if (dr) dr[i] = make_float4(d[0], d[1], d[2], d[3]);
                float g_lo = 0.f, b_lo = 0.f, g_hi = 0.f, b_hi = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (split > j) { g_lo = fmaf(d[j], h[j], g_lo); b_lo += d[j]; } else { g_hi = fmaf(d[j], h[j], g_hi); b_hi += d[j]; }
                }
#pragma unroll
                for (int c = 0; c < GN_CPG; ++c) {
                    if (c == cl) { dgc[c] += g_lo; dbc[c] += b_lo; }
                    if (c == cl + 1) { dgc[c] += g_hi; dbc[c] += b_hi; }
                }
                xh[k] = make_float4(h[0], h[1], h[2], h[3]);
                e[k] = make_float4(d[0] * ga[0], d[1] * ga[1], d[2] * ga[2], d[3] * ga[3]);
                s1 += (e[k].x + e[k].y) + (e[k].z + e[k].w);
                s2 += fmaf(e[k].x, xh[k].x, fmaf(e[k].y, xh[k].y, fmaf(e[k].z, xh[k].z, e[k].w * xh[k].w)));
            }
        }
        const float m1 = warp_sum(s1) * inv, m2 = warp_sum(s2) * inv;
        float4 *dst = reinterpret_cast<float4 *>(dx) + gi * n4;
#pragma unroll
        for (int k = 0; k < GN_MAXV; ++k) {
            const int i = lane + 32 * k;
            if (i < n4)
                dst[i] = make_float4(st.y * (e[k].x - m1 - xh[k].x * m2), st.y * (e[k].y - m1 - xh[k].y * m2),
                                     st.y * (e[k].z - m1 - xh[k].z * m2), st.y * (e[k].w - m1 - xh[k].w * m2));
        }
#pragma unroll
        for (int c = 0; c < GN_CPG; ++c) {
            const float a = warp_sum(dgc[c]), bsum = warp_sum(dbc[c]);
            if (lane == 0) { atomicAdd(&red[c0 + c], a); atomicAdd(&red[GN_C + c0 + c], bsum); }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * GN_C; i += GN_THREADS) partial[(size_t)blockIdx.x * 2 * GN_C + i] = red[i];
}

// ---- larger boards (H*W > 192, e.g. bloxCapture's 400 cells): one CTA of 256 threads per (sample, group) ----------
#define GNC_THREADS 256
#define GNC_MAXV 8                        // float4s per thread: groups up to 256 * 8 * 4 = 8192 floats (H*W <= 1024)

__device__ __forceinline__ float2 block_sum2(float a, float b, float2 *red) {
    a = warp_sum(a); b = warp_sum(b);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                                      // red[] may still be read from the previous use
    if (lane == 0) red[warp] = make_float2(a, b);
    __syncthreads();
    float2 r = make_float2(0.f, 0.f);
#pragma unroll
    for (int w = 0; w < GNC_THREADS / 32; ++w) { r.x += red[w].x; r.y += red[w].y; }
    return r;
}

template <int GELU>
__global__ void __launch_bounds__(GNC_THREADS) pm_groupnorm_cta_fwd_kernel(const float *__restrict__ x, const float *__restrict__ res,
                                                                           const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                           float *__restrict__ y, float2 *__restrict__ stats, int64_t groups,
                                                                           int HW, float eps) {
    __shared__ float2 red[GNC_THREADS / 32];
    const int n4 = GN_CPG * HW / 4;
    const float inv = 1.0f / (float)(GN_CPG * HW);
    for (int64_t gi = blockIdx.x; gi < groups; gi += gridDim.x) {
        const float4 *src = reinterpret_cast<const float4 *>(x) + gi * n4;
        float4 v[GNC_MAXV];
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < GNC_MAXV; ++k) {
            const int i = threadIdx.x + GNC_THREADS * k;
            v[k] = i < n4 ? __ldg(src + i) : make_float4(0, 0, 0, 0);
            s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
        }
        const float mean = block_sum2(s, 0.f, red).x * inv;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < GNC_MAXV; ++k) {
            if (threadIdx.x + GNC_THREADS * k < n4) {
                const float a0 = v[k].x - mean, a1 = v[k].y - mean, a2 = v[k].z - mean, a3 = v[k].w - mean;
                q += fmaf(a0, a0, fmaf(a1, a1, fmaf(a2, a2, a3 * a3)));
            }
        }
        const float rstd = rsqrtf(block_sum2(q, 0.f, red).x * inv + eps);
        const int c0 = (int)(gi % GN_G) * GN_CPG;
        float4 *dst = reinterpret_cast<float4 *>(y) + gi * n4;
        const float4 *rs = res ? reinterpret_cast<const float4 *>(res) + gi * n4 : nullptr;
#pragma unroll
        for (int k = 0; k < GNC_MAXV; ++k) {
            const int i = threadIdx.x + GNC_THREADS * k;
            if (i < n4) {
                int cl, split;
                gn_channels(i, HW, cl, split);
                const float g0 = __ldg(gamma + c0 + cl) * rstd, b0 = __ldg(beta + c0 + cl) - mean * g0;
                float g1 = g0, b1 = b0;
                if (split < 4) { g1 = __ldg(gamma + c0 + cl + 1) * rstd; b1 = __ldg(beta + c0 + cl + 1) - mean * g1; }
                float4 o = make_float4(fmaf(v[k].x, split > 0 ? g0 : g1, split > 0 ? b0 : b1), fmaf(v[k].y, split > 1 ? g0 : g1, split > 1 ? b0 : b1),
                                       fmaf(v[k].z, split > 2 ? g0 : g1, split > 2 ? b0 : b1), fmaf(v[k].w, split > 3 ? g0 : g1, split > 3 ? b0 : b1));
                if (rs) { const float4 r = __ldg(rs + i); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
                if (GELU) { o.x = gelu_f(o.x); o.y = gelu_f(o.y); o.z = gelu_f(o.z); o.w = gelu_f(o.w); }
                dst[i] = o;
            }
        }
        if (stats && threadIdx.x == 0) stats[gi] = make_float2(mean, rstd);
    }
}

template <int GELU>
__global__ void __launch_bounds__(GNC_THREADS) pm_groupnorm_cta_bwd_kernel(const float *__restrict__ x, const float *__restrict__ res,
                                                                           const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                           const float2 *__restrict__ stats, const float *__restrict__ dy,
                                                                           float *__restrict__ dx, float *__restrict__ dres,
                                                                           float *__restrict__ partial, int64_t groups, int HW) {
    __shared__ float2 red[GNC_THREADS / 32];
    __shared__ float acc[2 * GN_C];
    for (int i = threadIdx.x; i < 2 * GN_C; i += GNC_THREADS) acc[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n4 = GN_CPG * HW / 4;
    const float inv = 1.0f / (float)(GN_CPG * HW);
    for (int64_t gi = blockIdx.x; gi < groups; gi += gridDim.x) {
        const float4 *sx = reinterpret_cast<const float4 *>(x) + gi * n4, *sd = reinterpret_cast<const float4 *>(dy) + gi * n4;
        const float4 *rs = res ? reinterpret_cast<const float4 *>(res) + gi * n4 : nullptr;
        float4 *dr = dres ? reinterpret_cast<float4 *>(dres) + gi * n4 : nullptr;
        const float2 st = __ldg(stats + gi);
        const int c0 = (int)(gi % GN_G) * GN_CPG;
        float4 xh[GNC_MAXV], e[GNC_MAXV];
        float s1 = 0.f, s2 = 0.f;
        float dgc[GN_CPG], dbc[GN_CPG];
#pragma unroll
        for (int c = 0; c < GN_CPG; ++c) { dgc[c] = 0.f; dbc[c] = 0.f; }
#pragma unroll
        for (int k = 0; k < GNC_MAXV; ++k) {
            const int i = threadIdx.x + GNC_THREADS * k;
            xh[k] = make_float4(0, 0, 0, 0); e[k] = xh[k];
            if (i < n4) {
                const float4 v = __ldg(sx + i);
                const float4 d4 = __ldg(sd + i);
                int cl, split;
                gn_channels(i, HW, cl, split);
                const float g0 = __ldg(gamma + c0 + cl), g1 = split < 4 ? __ldg(gamma + c0 + cl + 1) : g0;
                float h[4] = {(v.x - st.x) * st.y, (v.y - st.x) * st.y, (v.z - st.x) * st.y, (v.w - st.x) * st.y};
                float d[4] = {d4.x, d4.y, d4.z, d4.w};
                float ga[4] = {split > 0 ? g0 : g1, split > 1 ? g0 : g1, split > 2 ? g0 : g1, split > 3 ? g0 : g1};
                if (GELU) {
                    const float b0 = __ldg(beta + c0 + cl), b1 = split < 4 ? __ldg(beta + c0 + cl + 1) : b0;
                    float r[4] = {0.f, 0.f, 0.f, 0.f};
                    if (rs) { const float4 r4 = __ldg(rs + i); r[0] = r4.x; r[1] = r4.y; r[2] = r4.z; r[3] = r4.w; }
#pragma unroll
                    for (int j = 0; j < 4; ++j) d[j] *= gelu_grad_f(fmaf(h[j], ga[j], (split > j ? b0 : b1)) + r[j]);
                }
                if (dr) dr[i] = make_float4(d[0], d[1], d[2], d[3]);
                float g_lo = 0.f, b_lo = 0.f, g_hi = 0.f, b_hi = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (split > j) { g_lo = fmaf(d[j], h[j], g_lo); b_lo += d[j]; } else { g_hi = fmaf(d[j], h[j], g_hi); b_hi += d[j]; }
                }
#pragma unroll
                for (int c = 0; c < GN_CPG; ++c) {
                    if (c == cl) { dgc[c] += g_lo; dbc[c] += b_lo; }
                    if (c == cl + 1) { dgc[c] += g_hi; dbc[c] += b_hi; }
                }
                xh[k] = make_float4(h[0], h[1], h[2], h[3]);
                e[k] = make_float4(d[0] * ga[0], d[1] * ga[1], d[2] * ga[2], d[3] * ga[3]);
                s1 += (e[k].x + e[k].y) + (e[k].z + e[k].w);
                s2 += fmaf(e[k].x, xh[k].x, fmaf(e[k].y, xh[k].y, fmaf(e[k].z, xh[k].z, e[k].w * xh[k].w)));
            }
        }
        const float2 m = block_sum2(s1, s2, red);
        const float m1 = m.x * inv, m2 = m.y * inv;
        float4 *dst = reinterpret_cast<float4 *>(dx) + gi * n4;
#pragma unroll
        for (int k = 0; k < GNC_MAXV; ++k) {
            const int i = threadIdx.x + GNC_THREADS * k;
            if (i < n4)
                dst[i] = make_float4(st.y * (e[k].x - m1 - xh[k].x * m2), st.y * (e[k].y - m1 - xh[k].y * m2),
                                     st.y * (e[k].z - m1 - xh[k].z * m2), st.y * (e[k].w - m1 - xh[k].w * m2));
        }
#pragma unroll
        for (int c = 0; c < GN_CPG; ++c) {
            const float a = warp_sum(dgc[c]), bsum = warp_sum(dbc[c]);
            if (lane == 0) { atomicAdd(&acc[c0 + c], a); atomicAdd(&acc[GN_C + c0 + c], bsum); }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * GN_C; i += GNC_THREADS) partial[(size_t)blockIdx.x * 2 * GN_C + i] = acc[i];
}

static unsigned gn_grid(int64_t groups) {
    const int64_t want = (groups + GN_THREADS / 32 - 1) / (GN_THREADS / 32);
    return (unsigned)(want < 148 * 16 ? (want > 0 ? want : 1) : 148 * 16);
}
extern "C" int pm_groupnorm_blocks(int64_t n_samples) { return (int)gn_grid(n_samples * GN_G); }

static int gn_ok(const void *x, const void *y, int64_t B, int HW) {
    return x && y && B > 0 && HW > 0 && (HW & 1) == 0 && GN_CPG * HW / 4 <= GNC_THREADS * GNC_MAXV && (((size_t)x | (size_t)y) & 15) == 0;
}
static bool gn_warp_path(int HW) { return GN_CPG * HW / 4 <= 32 * GN_MAXV; }      // the group fits one warp's registers
extern "C" int pm_groupnorm_4x8_max_hw(void) { return GNC_THREADS * GNC_MAXV * 4 / GN_CPG; }

extern "C" int pm_groupnorm_4x8_fwd(const float *x, const float *res, const float *gamma, const float *beta, float *y, float *stats,
                                    int64_t n_samples, int HW, double eps, int gelu, pm_stream_t stream) {
    if (!gn_ok(x, y, n_samples, HW) || !gamma || !beta || ((size_t)stats & 7) || ((size_t)res & 15)) return PM_EINVAL;
    const int64_t groups = n_samples * GN_G;
    if (!gn_warp_path(HW)) {
        if (gelu)
            pm_groupnorm_cta_fwd_kernel<1><<<gn_grid(groups), GNC_THREADS, 0, (cudaStream_t)stream>>>(x, res, gamma, beta, y, (float2 *)stats, groups, HW, (float)eps);
        else
            pm_groupnorm_cta_fwd_kernel<0><<<gn_grid(groups), GNC_THREADS, 0, (cudaStream_t)stream>>>(x, res, gamma, beta, y, (float2 *)stats, groups, HW, (float)eps);
        return (int)cudaGetLastError();
    }
    if (gelu)
        pm_groupnorm_fwd_kernel<1><<<gn_grid(groups), GN_THREADS, 0, (cudaStream_t)stream>>>(x, res, gamma, beta, y, (float2 *)stats, groups, HW, (float)eps);
    else
        pm_groupnorm_fwd_kernel<0><<<gn_grid(groups), GN_THREADS, 0, (cudaStream_t)stream>>>(x, res, gamma, beta, y, (float2 *)stats, groups, HW, (float)eps);
    return (int)cudaGetLastError();
}

extern "C" int pm_groupnorm_4x8_bwd(const float *x, const float *res, const float *gamma, const float *beta, const float *stats,
                                    const float *dy, float *dx, float *dres, float *partial, int64_t n_samples, int HW, int gelu,
                                    pm_stream_t stream) {
    if (!gn_ok(x, dx, n_samples, HW) || !gamma || !beta || !stats || !dy || !partial || (((size_t)dy | (size_t)res | (size_t)dres) & 15) ||
        ((size_t)stats & 7))
        return PM_EINVAL;
    const int64_t groups = n_samples * GN_G;
    if (!gn_warp_path(HW)) {
        if (gelu)
            pm_groupnorm_cta_bwd_kernel<1><<<gn_grid(groups), GNC_THREADS, 0, (cudaStream_t)stream>>>(x, res, gamma, beta, (const float2 *)stats, dy, dx, dres, partial, groups, HW);
        else
            pm_groupnorm_cta_bwd_kernel<0><<<gn_grid(groups), GNC_THREADS, 0, (cudaStream_t)stream>>>(x, res, gamma, beta, (const float2 *)stats, dy, dx, dres, partial, groups, HW);
        return (int)cudaGetLastError();
    }
    if (gelu)
        pm_groupnorm_bwd_kernel<1><<<gn_grid(groups), GN_THREADS, 0, (cudaStream_t)stream>>>(x, res, gamma, beta, (const float2 *)stats, dy, dx, dres, partial, groups, HW);
    else
        pm_groupnorm_bwd_kernel<0><<<gn_grid(groups), GN_THREADS, 0, (cudaStream_t)stream>>>(x, res, gamma, beta, (const float2 *)stats, dy, dx, dres, partial, groups, HW);
    return (int)cudaGetLastError();
}
