// pm_attn.cu — the critic's self-attention core (the PPO update around the PyTorch networks) as
// sm_100a kernels: softmax(Q K^T / sqrt(8)) V and its gradient for heads of dimension 8 over the H*W board tokens.
//
// The reference's centralised critic (pacman_mappo_resnet.py:131-136, 168-176) is nn.TransformerEncoder(d_model 32,
// nhead 4) over one token per board cell: 140-154 tokens (tiny / smallCapture), head_dim 8.  At that shape PyTorch's
// fused fp32 attention kernels (mem-efficient fmha, 64x64 tiles built for head_dim >= 32) spend 5 + 17 ms per layer
// (forward + backward) on a 4 096-sample minibatch, and the unfused bmm / softmax / bmm sequence materialises a
// [B*4, S, S] probability tensor (1.5 GB) - together most of a PPO update.  The arithmetic is tiny (2.5 GFLOP per
// layer): one CTA per (sample, head) keeps K and V (forward) or Q, K, V, dO (backward) of its head in shared memory,
// one thread per token row streams over them with broadcast 128-bit shared-memory loads, and nothing but the
// log-sum-exp per row is kept for the backward pass.  Plain fp32 FMAs (no TF32), explicit fmaf: the library is
// compiled with -fmad=false for the reward / GAE arithmetic.
//
// Layouts are nn.MultiheadAttention's own (batch_first = False):
//   qkv  [S][B][3E]  = in_proj(x), E = 8 * n_heads: q of head h at channel 8h, k at E + 8h, v at 2E + 8h
//   out  [S][B][E]   the concatenated heads, input of out_proj
//   lse  [B][n_heads][S]   log-sum-exp of the scaled scores of every row (saved for the backward pass)
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/pacman_b200.h"

#define PM_ATTN_HD 8
#define PM_ATTN_MAX_S 1024

__device__ __forceinline__ float dot8(const float4 a0, const float4 a1, const float4 b0, const float4 b1) {
    float s = a0.x * b0.x;
    s = fmaf(a0.y, b0.y, s); s = fmaf(a0.z, b0.z, s); s = fmaf(a0.w, b0.w, s);
    s = fmaf(a1.x, b1.x, s); s = fmaf(a1.y, b1.y, s); s = fmaf(a1.z, b1.z, s); s = fmaf(a1.w, b1.w, s);
    return s;
}
__device__ __forceinline__ void axpy8(float4 &y0, float4 &y1, const float a, const float4 x0, const float4 x1) {
    y0.x = fmaf(a, x0.x, y0.x); y0.y = fmaf(a, x0.y, y0.y); y0.z = fmaf(a, x0.z, y0.z); y0.w = fmaf(a, x0.w, y0.w);
    y1.x = fmaf(a, x1.x, y1.x); y1.y = fmaf(a, x1.y, y1.y); y1.z = fmaf(a, x1.z, y1.z); y1.w = fmaf(a, x1.w, y1.w);
}
__device__ __forceinline__ float4 scale4(const float4 v, const float a) { return make_float4(v.x * a, v.y * a, v.z * a, v.w * a); }

// grid = B * n_heads (head fastest: the four heads of a sample read neighbouring 32-byte pieces of the same lines),
// block = S rounded up to a warp; thread t = token row t.
__global__ void __launch_bounds__(PM_ATTN_MAX_S) pm_attn_fwd_kernel(const float *__restrict__ qkv, float *__restrict__ out, float *__restrict__ lse, int S, int B,
                                   int nh, float scale) {
    extern __shared__ float4 sm4[];
    float4 *Ks = sm4, *Vs = sm4 + 2 * S;
    const int b = blockIdx.x / nh, h = blockIdx.x - b * nh, E = nh * PM_ATTN_HD, t = threadIdx.x;
    float4 q0 = make_float4(0, 0, 0, 0), q1 = q0;
    if (t < S) {
        const float *p = qkv + ((size_t)t * B + b) * 3 * E + h * PM_ATTN_HD;
        q0 = scale4(__ldg(reinterpret_cast<const float4 *>(p)), scale);          // q / sqrt(head_dim), like F.multi_head_attention
        q1 = scale4(__ldg(reinterpret_cast<const float4 *>(p) + 1), scale);
        Ks[2 * t] = __ldg(reinterpret_cast<const float4 *>(p + E)); Ks[2 * t + 1] = __ldg(reinterpret_cast<const float4 *>(p + E) + 1);
        Vs[2 * t] = __ldg(reinterpret_cast<const float4 *>(p + 2 * E)); Vs[2 * t + 1] = __ldg(reinterpret_cast<const float4 *>(p + 2 * E) + 1);
    }
    __syncthreads();
    if (t >= S) return;
    float m = -INFINITY;
#pragma unroll 4
    for (int j = 0; j < S; ++j) m = fmaxf(m, dot8(q0, q1, Ks[2 * j], Ks[2 * j + 1]));
    float l = 0.f;
    float4 o0 = make_float4(0, 0, 0, 0), o1 = o0;
#pragma unroll 2
    for (int j = 0; j < S; ++j) {
        const float p = expf(dot8(q0, q1, Ks[2 * j], Ks[2 * j + 1]) - m);
        l += p;
        axpy8(o0, o1, p, Vs[2 * j], Vs[2 * j + 1]);
    }
    const float inv = 1.f / l;
    float *po = out + ((size_t)t * B + b) * E + h * PM_ATTN_HD;
    reinterpret_cast<float4 *>(po)[0] = scale4(o0, inv);
    reinterpret_cast<float4 *>(po)[1] = scale4(o1, inv);
    if (lse) lse[(size_t)blockIdx.x * S + t] = m + logf(l);
}

// Backward.  With P = softmax(scale * Q K^T), O = P V, D_i = dO_i . O_i:
//   dV_j = sum_i P_ij dO_i      dS_ij = P_ij (dO_i . V_j - D_i)      dQ_i = scale * sum_j dS_ij K_j
//   dK_j = scale * sum_i dS_ij Q_i
// Pass 1 has thread i own a query row (dQ_i), pass 2 has thread j own a key row (dK_j, dV_j); both recompute P from
// the saved log-sum-exp, so no S x S tensor exists anywhere and no atomics are needed.
__global__ void __launch_bounds__(PM_ATTN_MAX_S) pm_attn_bwd_kernel(const float *__restrict__ qkv, const float *__restrict__ out, const float *__restrict__ dout,
                                   const float *__restrict__ lse, float *__restrict__ dqkv, int S, int B, int nh, float scale) {
    extern __shared__ float4 sm4[];
    float4 *Qs = sm4, *Ks = Qs + 2 * S, *Vs = Ks + 2 * S, *Gs = Vs + 2 * S;      // Gs = dO
    float *Ls = reinterpret_cast<float *>(Gs + 2 * S), *Ds = Ls + S;
    const int b = blockIdx.x / nh, h = blockIdx.x - b * nh, E = nh * PM_ATTN_HD, t = threadIdx.x;
    float4 q0, q1, k0, k1, v0, v1, g0, g1;
    float li = 0.f, di = 0.f;
    if (t < S) {
        const float *p = qkv + ((size_t)t * B + b) * 3 * E + h * PM_ATTN_HD;
        q0 = scale4(__ldg(reinterpret_cast<const float4 *>(p)), scale); q1 = scale4(__ldg(reinterpret_cast<const float4 *>(p) + 1), scale);
        k0 = __ldg(reinterpret_cast<const float4 *>(p + E)); k1 = __ldg(reinterpret_cast<const float4 *>(p + E) + 1);
        v0 = __ldg(reinterpret_cast<const float4 *>(p + 2 * E)); v1 = __ldg(reinterpret_cast<const float4 *>(p + 2 * E) + 1);
        const size_t oi = ((size_t)t * B + b) * E + h * PM_ATTN_HD;
        g0 = __ldg(reinterpret_cast<const float4 *>(dout + oi)); g1 = __ldg(reinterpret_cast<const float4 *>(dout + oi) + 1);
        const float4 o0 = __ldg(reinterpret_cast<const float4 *>(out + oi)), o1 = __ldg(reinterpret_cast<const float4 *>(out + oi) + 1);
        li = lse[(size_t)blockIdx.x * S + t];
        di = dot8(g0, g1, o0, o1);
        Qs[2 * t] = q0; Qs[2 * t + 1] = q1; Ks[2 * t] = k0; Ks[2 * t + 1] = k1;
        Vs[2 * t] = v0; Vs[2 * t + 1] = v1; Gs[2 * t] = g0; Gs[2 * t + 1] = g1;
        Ls[t] = li; Ds[t] = di;
    }
    __syncthreads();
    if (t >= S) return;
    // pass 1: this thread's query row
    float4 dq0 = make_float4(0, 0, 0, 0), dq1 = dq0;
#pragma unroll 2
    for (int j = 0; j < S; ++j) {
        const float4 kj0 = Ks[2 * j], kj1 = Ks[2 * j + 1];
        const float p = expf(dot8(q0, q1, kj0, kj1) - li);
        const float ds = p * (dot8(g0, g1, Vs[2 * j], Vs[2 * j + 1]) - di);
        axpy8(dq0, dq1, ds, kj0, kj1);
    }
    // pass 2: this thread's key row
    float4 dk0 = make_float4(0, 0, 0, 0), dk1 = dk0, dv0 = dk0, dv1 = dk0;
#pragma unroll 2
    for (int i = 0; i < S; ++i) {
        const float4 qi0 = Qs[2 * i], qi1 = Qs[2 * i + 1], gi0 = Gs[2 * i], gi1 = Gs[2 * i + 1];
        const float p = expf(dot8(qi0, qi1, k0, k1) - Ls[i]);
        axpy8(dv0, dv1, p, gi0, gi1);
        const float ds = p * (dot8(gi0, gi1, v0, v1) - Ds[i]);
        axpy8(dk0, dk1, ds, qi0, qi1);                    // Qs already carries the 1/sqrt(head_dim)
    }
    float *pd = dqkv + ((size_t)t * B + b) * 3 * E + h * PM_ATTN_HD;
    reinterpret_cast<float4 *>(pd)[0] = scale4(dq0, scale); reinterpret_cast<float4 *>(pd)[1] = scale4(dq1, scale);
    reinterpret_cast<float4 *>(pd + E)[0] = dk0; reinterpret_cast<float4 *>(pd + E)[1] = dk1;
    reinterpret_cast<float4 *>(pd + 2 * E)[0] = dv0; reinterpret_cast<float4 *>(pd + 2 * E)[1] = dv1;
}

static int attn_args_ok(const void *a, const void *b, int S, int64_t B, int nh) {
    return a && b && S >= 1 && S <= PM_ATTN_MAX_S && B >= 1 && nh >= 1 && B * nh <= 0x7fffffff && (((size_t)a | (size_t)b) & 15) == 0;
}

extern "C" int pm_attn_fwd(const float *qkv, float *out, float *lse, int S, int64_t B, int n_heads, pm_stream_t stream) {
    if (!attn_args_ok(qkv, out, S, B, n_heads)) return PM_EINVAL;
    const int threads = (S + 31) & ~31;
    const size_t smem = (size_t)4 * S * sizeof(float4);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(pm_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    pm_attn_fwd_kernel<<<(unsigned)(B * n_heads), threads, smem, (cudaStream_t)stream>>>(qkv, out, lse, S, (int)B, n_heads,
                                                                                          1.0f / sqrtf((float)PM_ATTN_HD));
    return (int)cudaGetLastError();
}

extern "C" int pm_attn_bwd(const float *qkv, const float *out, const float *dout, const float *lse, float *dqkv, int S, int64_t B,
                           int n_heads, pm_stream_t stream) {
    if (!attn_args_ok(qkv, out, S, B, n_heads) || !dout || !lse || !dqkv || (((size_t)dout | (size_t)dqkv) & 15)) return PM_EINVAL;
    const int threads = (S + 31) & ~31;
    const size_t smem = (size_t)8 * S * sizeof(float4) + (size_t)2 * S * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(pm_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    pm_attn_bwd_kernel<<<(unsigned)(B * n_heads), threads, smem, (cudaStream_t)stream>>>(qkv, out, dout, lse, dqkv, S, (int)B, n_heads,
                                                                                          1.0f / sqrtf((float)PM_ATTN_HD));
    return (int)cudaGetLastError();
}
