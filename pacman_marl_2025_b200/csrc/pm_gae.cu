// pm_gae.cu — rollout post-processing kernels (sm_100a).
//
//  pm_gae      compute_gae (pacman_mappo_resnet.py:277-290) for M independent
//              columns of a [T][M] rollout buffer: one thread per column (per four
//              adjacent columns for wide buffers) walks t from T-1 down to 0, so every
//              load/store of a warp is a coalesced 128 B / 512 B row segment.  20 algorithmic bytes per (t, column): read
//              reward, value, done; write advantage, return.  The arithmetic
//              reproduces torch's fp32 op order (Python float scalars are rounded
//              to fp32 before meeting a tensor; float*float stays double), with
//              explicit _rn intrinsics so nothing contracts into FMA.
//  pm_adv_norm per-minibatch (adv - mean) / (std_unbiased + 1e-8)
//              (pacman_mappo_resnet.py:577): one CTA per minibatch, gather by
//              index, warp-shuffle + shared-memory block reductions.
#include <cuda_runtime.h>

#include "../../include/pacman_b200.h"

#define GAE_UNROLL 8

// V columns per thread (1 or 4): with V = 4 every load / store is 128-bit and a thread keeps 3 x GAE_UNROLL of them in
// flight, which is what it takes to hide HBM latency with M/4 threads; the arithmetic per column is the same sequence
// of _rn operations either way, so results do not depend on V.
template <int V> struct GaeVec;
template <> struct GaeVec<1> { typedef float T; };
template <> struct GaeVec<4> { typedef float4 T; };
__device__ __forceinline__ float gae_get(const float &v, int) { return v; }
__device__ __forceinline__ float gae_get(const float4 &v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ void gae_set(float &v, int, float x) { v = x; }
__device__ __forceinline__ void gae_set(float4 &v, int i, float x) { if (i == 0) v.x = x; else if (i == 1) v.y = x; else if (i == 2) v.z = x; else v.w = x; }

template <int V>
__global__ void __launch_bounds__(128) pm_gae_kernel(const float *__restrict__ rew, const float *__restrict__ val,
                                                     const float *__restrict__ done, const float *__restrict__ last_value,
                                                     float gamma_f, float gl_f, double gamma_d, float *__restrict__ adv,
                                                     float *__restrict__ ret, int T, int64_t M) {
    typedef typename GaeVec<V>::T vec;
    const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;     // first of this thread's V columns
    if (m >= M) return;
    // t = T-1 bootstraps from last_value: gamma * next_val is Python float * float (double), then meets a tensor
    float gv_next[V], lastgaelam[V], v_next[V];
#pragma unroll
    for (int c = 0; c < V; ++c) { gv_next[c] = (float)(gamma_d * (double)last_value[m + c]); lastgaelam[c] = 0.0f; v_next[c] = 0.0f; }
    bool first = true;
    for (int t0 = T - 1; t0 >= 0; t0 -= GAE_UNROLL) {
        vec r[GAE_UNROLL], v[GAE_UNROLL], d[GAE_UNROLL];
#pragma unroll
        for (int u = 0; u < GAE_UNROLL; ++u) {           // independent loads first: GAE_UNROLL rows in flight per thread
            const int t = t0 - u;
            if (t >= 0) {
                const size_t o = (size_t)t * M + m;
                r[u] = __ldcs(reinterpret_cast<const vec *>(rew + o)); v[u] = __ldcs(reinterpret_cast<const vec *>(val + o));
                d[u] = __ldcs(reinterpret_cast<const vec *>(done + o));
            }
        }
#pragma unroll
        for (int u = 0; u < GAE_UNROLL; ++u) {
            const int t = t0 - u;
            if (t >= 0) {
                vec a_out, r_out;
#pragma unroll
                for (int c = 0; c < V; ++c) {
                    const float vc = gae_get(v[u], c);
                    const float nonterm = __fsub_rn(1.0f, gae_get(d[u], c));
                    const float gv = first ? gv_next[c] : __fmul_rn(gamma_f, v_next[c]);
                    const float delta = __fsub_rn(__fadd_rn(gae_get(r[u], c), __fmul_rn(gv, nonterm)), vc);
                    lastgaelam[c] = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl_f, nonterm), lastgaelam[c]));
                    gae_set(a_out, c, lastgaelam[c]);
                    gae_set(r_out, c, __fadd_rn(lastgaelam[c], vc));
                    v_next[c] = vc;
                }
                first = false;
                const size_t o = (size_t)t * M + m;
                __stcs(reinterpret_cast<vec *>(adv + o), a_out);
                __stcs(reinterpret_cast<vec *>(ret + o), r_out);
            }
        }
    }
}

extern "C" int pm_gae(const float *rew, const float *val, const float *done, const float *last_value, double gamma,
                      double lam, float *adv, float *ret, int T, int64_t M, pm_stream_t stream) {
    if (!rew || !val || !done || !last_value || !adv || !ret || T <= 0 || M <= 0) return PM_EINVAL;
    const bool aligned = ((((size_t)rew | (size_t)val | (size_t)done | (size_t)adv | (size_t)ret) & 15) == 0) && (M & 3) == 0;
    if (aligned && M >= 4 * 128 * 64) {                  // enough columns to give every SM work at four per thread
        const unsigned grid = (unsigned)((M / 4 + 127) / 128);
        pm_gae_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>(rew, val, done, last_value, (float)gamma, (float)(gamma * lam),
                                                                 gamma, adv, ret, T, M);
    } else {
        const unsigned grid = (unsigned)((M + 127) / 128);
        pm_gae_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(rew, val, done, last_value, (float)gamma, (float)(gamma * lam),
                                                                 gamma, adv, ret, T, M);
    }
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float x, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();                                    // red[] may still be read from the previous call
    if (lane == 0) red[warp] = x;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w];           // same order in every thread -> identical result
    return s;
}

// CACHE = 1: the minibatch's gathered advantages (mb floats) are kept in shared memory after the first pass, so the
// random 4-byte gathers - each costs a 32-byte sector of HBM traffic - happen once instead of three times.
template <int CACHE>
__global__ void __launch_bounds__(1024) pm_adv_norm_kernel(const float *__restrict__ adv, const int64_t *__restrict__ inds,
                                                          float *__restrict__ out, int64_t n, int mb) {
    extern __shared__ float vals[];
    __shared__ float red[32];
    const int64_t s = (int64_t)blockIdx.x * mb;
    const int k = (int)min((int64_t)mb, n - s);
    float part = 0.f;
    // the gather is latency-bound (index, then a random 4-byte read): eight independent pairs in flight per thread
    for (int i0 = threadIdx.x; i0 < k; i0 += 8 * blockDim.x) {
        int64_t ix[8];
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; ix[u] = i < k ? __ldg(inds + s + i) : -1; }
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = ix[u] >= 0 ? __ldg(adv + ix[u]) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < k) { if (CACHE) vals[i] = a[u]; part += a[u]; }
        }
    }
    const float mean = block_sum(part, red) / (float)k;
    part = 0.f;
    for (int i = threadIdx.x; i < k; i += blockDim.x) { const float dlt = (CACHE ? vals[i] : adv[inds[s + i]]) - mean; part += dlt * dlt; }
    // unbiased variance; a trailing group of ONE sample has none (torch.std gives NaN there): treat it as 0, so the
    // lone sample normalises to (a - a) / 1e-8 = 0 instead of poisoning the loss
    const float ssq = block_sum(part, red);
    const float var = k > 1 ? ssq / (float)(k - 1) : 0.0f;
    const float denom = sqrtf(var) + 1e-8f;
    for (int i = threadIdx.x; i < k; i += blockDim.x) out[s + i] = ((CACHE ? vals[i] : adv[inds[s + i]]) - mean) / denom;
}

extern "C" int pm_adv_norm(const float *adv, const int64_t *inds, float *out, int64_t n, int mb, pm_stream_t stream) {
    if (!adv || !inds || !out || n <= 0 || mb <= 1) return PM_EINVAL;
    const unsigned grid = (unsigned)((n + mb - 1) / mb);
    const size_t smem = (size_t)mb * sizeof(float);
    // about eight samples per thread (B200, 33.5 M shuffled samples: 0.37-0.40 ms for minibatches of 512 ... 16 384;
    // 256 threads on 8 192 samples took 0.93 ms)
    const int threads = mb >= 8192 ? 1024 : mb >= 4096 ? 512 : 256;
    if (smem <= 160 * 1024) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(pm_adv_norm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        pm_adv_norm_kernel<1><<<grid, threads, smem, (cudaStream_t)stream>>>(adv, inds, out, n, mb);
    } else {
        pm_adv_norm_kernel<0><<<grid, threads, 0, (cudaStream_t)stream>>>(adv, inds, out, n, mb);
    }
    return (int)cudaGetLastError();
}
