// pm_ppo.cu — the PPO minibatch pipeline around the (PyTorch) networks, SURVEY §8 row f3
// (pacman_mappo_resnet.py:555-600), as sm_100a kernels:
//
//  pm_gather_rows   b_obs[mb].to(device) / b_merged[mb] (`:569-573`): rows of the uint8 rollout buffer picked by
//                   a shuffled index list and widened to the networks' input type (fp32 or bf16) in one pass -
//                   the `.float()` is fused into the gather.  HBM-bound: 1 B read + 4 (2) B written per element.
//  pm_ppo_loss      everything between the networks' outputs and `loss.backward()` (`:575-585, 597-600`):
//                   log-softmax, log-prob of the stored action, entropy, ratio, clipped surrogate, value loss,
//                   the scalar loss AND its gradient w.r.t. logits / values, plus the four logged scalars
//                   (pg, vl, entropy mean, clip fraction).  One launch; deterministic two-stage reduction (the
//                   last block to finish sums the per-block partials in a fixed order).
//  pm_adam_ema      clip_grad_norm_ + Adam step + EMA of the policy (`:588-595`) over FLAT fp32 parameter /
//                   gradient / moment / EMA buffers: one sum-of-squares reduction and one elementwise pass
//                   (36 B per parameter) instead of a multi-tensor Adam, a norm, a scale and two foreach ops.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/pacman_b200.h"

// ---------------------------------------------------------------------------------------------------------
// gather + widen
// ---------------------------------------------------------------------------------------------------------
template <typename OUT>
__device__ __forceinline__ void store8(OUT *dst, const uint2 v);
template <>
__device__ __forceinline__ void store8<float>(float *dst, const uint2 v) {
    float4 a = make_float4((float)(v.x & 0xff), (float)((v.x >> 8) & 0xff), (float)((v.x >> 16) & 0xff), (float)(v.x >> 24));
    float4 b = make_float4((float)(v.y & 0xff), (float)((v.y >> 8) & 0xff), (float)((v.y >> 16) & 0xff), (float)(v.y >> 24));
    reinterpret_cast<float4 *>(dst)[0] = a;
    reinterpret_cast<float4 *>(dst)[1] = b;
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16 *dst, const uint2 v) {
    __nv_bfloat162 p[4];
    p[0] = __floats2bfloat162_rn((float)(v.x & 0xff), (float)((v.x >> 8) & 0xff));
    p[1] = __floats2bfloat162_rn((float)((v.x >> 16) & 0xff), (float)(v.x >> 24));
    p[2] = __floats2bfloat162_rn((float)(v.y & 0xff), (float)((v.y >> 8) & 0xff));
    p[3] = __floats2bfloat162_rn((float)((v.y >> 16) & 0xff), (float)(v.y >> 24));
    *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(p);
}

// one warp per output row at a time; consecutive lanes = consecutive 8-byte chunks of the source row
template <typename OUT>
__global__ void __launch_bounds__(256) pm_gather_rows_kernel(const uint8_t *__restrict__ src, const int64_t *__restrict__ idx,
                                                             int64_t n_rows, int chunks, OUT *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_rows; r += nwarps) {
        const uint2 *s = reinterpret_cast<const uint2 *>(src + (size_t)idx[r] * chunks * 8);
        OUT *d = dst + (size_t)r * chunks * 8;
        for (int c = lane; c < chunks; c += 32) store8<OUT>(d + 8 * c, __ldg(s + c));
    }
}

extern "C" int pm_gather_rows(const uint8_t *src, const int64_t *idx, int64_t n_rows, int row_bytes, void *dst, int dst_bf16,
                              pm_stream_t stream) {
    if (!src || !idx || !dst || n_rows <= 0 || row_bytes <= 0 || (row_bytes & 7) || ((size_t)src & 7) || ((size_t)dst & 15))
        return PM_EINVAL;
    const int64_t warps = n_rows < 148 * 64 ? n_rows : 148 * 64;
    const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
    if (dst_bf16)
        pm_gather_rows_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(src, idx, n_rows, row_bytes / 8, (__nv_bfloat16 *)dst);
    else
        pm_gather_rows_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(src, idx, n_rows, row_bytes / 8, (float *)dst);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// fused PPO loss (+ gradient)
// ---------------------------------------------------------------------------------------------------------
#define PPO_THREADS 256
#define PPO_NSUM 4      // pg, value, entropy, clipped

__device__ __forceinline__ void block_sums(float (&x)[PPO_NSUM], float (*red)[PPO_NSUM]) {
#pragma unroll
    for (int k = 0; k < PPO_NSUM; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x[k] += __shfl_xor_sync(0xffffffffu, x[k], o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < PPO_NSUM; ++k) red[warp][k] = x[k];
    __syncthreads();
    if (threadIdx.x == 0)
        for (int w = 1; w < PPO_THREADS / 32; ++w)
#pragma unroll
            for (int k = 0; k < PPO_NSUM; ++k) red[0][k] += red[w][k];
    __syncthreads();
}

// logits [B][A] (A <= 8), values [nv] (sample i reads values[i % nv]: nv == B, or B/2 when both learners of an
// env-step share one critic value), actions [B] int64, old_logp / adv / ret [B].
// dlogits [B][A], dvalues [B] (per SAMPLE; the caller folds pairs), stats[8] = loss, pg, vl, entropy, clip_frac.
__global__ void __launch_bounds__(PPO_THREADS) pm_ppo_loss_kernel(const float *__restrict__ logits, const float *__restrict__ values,
                                                                  const int64_t *__restrict__ actions, const float *__restrict__ old_logp,
                                                                  const float *__restrict__ adv, const float *__restrict__ ret, int64_t B,
                                                                  int64_t nv, int A, float clip_eps, float vf_coef, float ent_coef,
                                                                  float *__restrict__ dlogits, float *__restrict__ dvalues,
                                                                  float *__restrict__ partial, unsigned *__restrict__ counter,
                                                                  float *__restrict__ stats) {
    __shared__ float red[PPO_THREADS / 32][PPO_NSUM];
    __shared__ bool last;
    float acc[PPO_NSUM] = {0.f, 0.f, 0.f, 0.f};
    const float invB = 1.0f / (float)B;
    for (int64_t i = (int64_t)blockIdx.x * PPO_THREADS + threadIdx.x; i < B; i += (int64_t)gridDim.x * PPO_THREADS) {
        float z[8], mx = -INFINITY;
        for (int k = 0; k < A; ++k) { z[k] = logits[i * A + k]; mx = fmaxf(mx, z[k]); }
        float se = 0.f;
        for (int k = 0; k < A; ++k) se += expf(z[k] - mx);
        const float lse = mx + logf(se);
        const int a = (int)actions[i];
        float ent = 0.f, lp = 0.f, p[8], lg[8];
        for (int k = 0; k < A; ++k) {
            lg[k] = z[k] - lse; p[k] = expf(lg[k]);
            ent -= p[k] * lg[k];
            if (k == a) lp = lg[k];
        }
        const float ad = adv[i];
        const float ratio = expf(lp - old_logp[i]);
        const float s1 = ad * ratio, s2 = ad * fminf(fmaxf(ratio, 1.0f - clip_eps), 1.0f + clip_eps);
        const float pg = -fminf(s1, s2);
        const float v = values[i % nv], dv = v - ret[i];
        acc[0] += pg; acc[1] += dv * dv; acc[2] += ent;
        acc[3] += (fabsf(ratio - 1.0f) > clip_eps) ? 1.f : 0.f;
        // d loss: torch.min passes the gradient to s1 when s1 <= s2 (a tie means the ratio is inside the clip range,
        // where the clamp branch has the same derivative, or adv == 0)
        const float g_lp = (s1 <= s2 ? -ad * ratio : 0.f) * invB;
        const float g_ent = -ent_coef * invB;
        for (int k = 0; k < A; ++k)
            dlogits[i * A + k] = g_lp * ((k == a ? 1.f : 0.f) - p[k]) - g_ent * p[k] * (lg[k] + ent);
        dvalues[i] = vf_coef * dv * invB;
    }
    block_sums(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < PPO_NSUM; ++k) partial[(size_t)blockIdx.x * PPO_NSUM + k] = red[0][k];
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {                  // fixed summation order over blocks: deterministic
        __threadfence();
        float t[PPO_NSUM] = {0.f, 0.f, 0.f, 0.f};
        for (unsigned b = 0; b < gridDim.x; ++b)
#pragma unroll
            for (int k = 0; k < PPO_NSUM; ++k) t[k] += partial[(size_t)b * PPO_NSUM + k];
        const float pg = t[0] * invB, vl = 0.5f * t[1] * invB, en = t[2] * invB;
        stats[0] = pg + vf_coef * vl - ent_coef * en;
        stats[1] = pg; stats[2] = vl; stats[3] = en; stats[4] = t[3] * invB;
        *counter = 0;                                // ready for the next launch
    }
}

#define PPO_MAX_BLOCKS 1024

extern "C" size_t pm_ppo_scratch_bytes(void) { return (size_t)PPO_MAX_BLOCKS * PPO_NSUM * 4 + 16; }

extern "C" int pm_ppo_loss(const float *logits, const float *values, const int64_t *actions, const float *old_logp,
                           const float *adv, const float *ret, int64_t B, int64_t n_values, int n_actions, double clip_eps,
                           double vf_coef, double ent_coef, float *dlogits, float *dvalues, float *stats, void *scratch,
                           pm_stream_t stream) {
    if (!logits || !values || !actions || !old_logp || !adv || !ret || !dlogits || !dvalues || !stats || !scratch) return PM_EINVAL;
    if (B <= 0 || n_values <= 0 || B % n_values != 0 || n_actions < 1 || n_actions > 8) return PM_EINVAL;
    int64_t blocks = (B + PPO_THREADS - 1) / PPO_THREADS;
    if (blocks > PPO_MAX_BLOCKS) blocks = PPO_MAX_BLOCKS;
    unsigned *counter = reinterpret_cast<unsigned *>(scratch);          // zero-initialised by the caller once
    float *partial = reinterpret_cast<float *>(scratch) + 4;
    pm_ppo_loss_kernel<<<(unsigned)blocks, PPO_THREADS, 0, (cudaStream_t)stream>>>(
        logits, values, actions, old_logp, adv, ret, B, n_values, n_actions, (float)clip_eps, (float)vf_coef, (float)ent_coef,
        dlogits, dvalues, partial, counter, stats);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam + EMA over flat buffers
// ---------------------------------------------------------------------------------------------------------
#define ADAM_THREADS 256
#define ADAM_MAX_BLOCKS 1184     // 148 SMs x 8

__global__ void __launch_bounds__(ADAM_THREADS) pm_sumsq_kernel(const float *__restrict__ g, int64_t n, float *__restrict__ partial,
                                                                unsigned *__restrict__ counter, float *__restrict__ norm_out,
                                                                float grad_scale) {
    __shared__ float red[ADAM_THREADS / 32];
    __shared__ bool last;
    float s = 0.f;
    const int64_t n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (int64_t i = (int64_t)blockIdx.x * ADAM_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * ADAM_THREADS) {
        const float4 v = g4[i];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += ADAM_THREADS) s += g[i] * g[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < ADAM_THREADS / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += (double)partial[b];
        *norm_out = (float)sqrt(t) * grad_scale;       // norm of the (averaged) gradient
        *counter = 0;
    }
}

// torch.optim.Adam (no weight decay, no amsgrad) in its own operation order:
//   m.lerp_(g, 1-b1); v = v*b2 + (1-b2)*g*g; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// with g = grad * grad_scale * min(1, max_norm / (norm + 1e-6))  (clip_grad_norm_), then
//   ema = ema * decay + p * (1 - decay)
__global__ void __launch_bounds__(ADAM_THREADS) pm_adam_ema_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                                                   float *__restrict__ v, float *__restrict__ ema, int64_t n,
                                                                   const float *__restrict__ norm, float grad_scale, float max_norm,
                                                                   float step_size, float b1, float b2, float inv_sqrt_bc2, float eps,
                                                                   float ema_decay) {
    const float coef = grad_scale * fminf(1.0f, max_norm / (*norm + 1e-6f));
    for (int64_t i = (int64_t)blockIdx.x * ADAM_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * ADAM_THREADS) {
        const float gi = g[i] * coef;
        float mi = m[i], vi = v[i];
        mi = mi + (gi - mi) * (1.0f - b1);
        vi = vi * b2 + (1.0f - b2) * gi * gi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        const float pi = p[i] - step_size * (mi / denom);
        m[i] = mi; v[i] = vi; p[i] = pi;
        if (ema) ema[i] = ema[i] * ema_decay + pi * (1.0f - ema_decay);
    }
}

extern "C" size_t pm_adam_scratch_bytes(void) { return (size_t)ADAM_MAX_BLOCKS * 4 + 32; }

extern "C" int pm_adam_ema(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, float *ema, int64_t n, int64_t step,
                           double lr, double beta1, double beta2, double eps, double max_grad_norm, double grad_scale,
                           double ema_decay, float *grad_norm_out, void *scratch, pm_stream_t stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !grad_norm_out || !scratch || n <= 0 || step <= 0) return PM_EINVAL;
    if (((size_t)grads & 15)) return PM_EINVAL;
    int64_t blocks = (n / 4 + ADAM_THREADS - 1) / ADAM_THREADS;
    if (blocks < 1) blocks = 1;
    if (blocks > ADAM_MAX_BLOCKS) blocks = ADAM_MAX_BLOCKS;
    unsigned *counter = reinterpret_cast<unsigned *>(scratch);
    float *partial = reinterpret_cast<float *>(scratch) + 8;
    pm_sumsq_kernel<<<(unsigned)blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(grads, n, partial, counter, grad_norm_out, (float)grad_scale);
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    int64_t b2 = (n + ADAM_THREADS - 1) / ADAM_THREADS;
    if (b2 > ADAM_MAX_BLOCKS) b2 = ADAM_MAX_BLOCKS;
    pm_adam_ema_kernel<<<(unsigned)b2, ADAM_THREADS, 0, (cudaStream_t)stream>>>(
        params, grads, exp_avg, exp_avg_sq, ema, n, grad_norm_out, (float)grad_scale, (float)max_grad_norm, (float)(lr / bc1),
        (float)beta1, (float)beta2, (float)(1.0 / sqrt(bc2)), (float)eps, (float)ema_decay);
    return (int)cudaGetLastError();
}
