// Dev tool: what store bandwidth can a write-only kernel reach on this B200?
// The step+obs kernel is ~99 % stores, while MEASURED_PEAKS.json's hbm_gbs is a
// read+write copy; this measures the write-only ceilings it should be compared with:
//   (1) grid-stride float4 st.global                 (plain / .cs streaming hint)
//   (2) per-warp 512 B x k contiguous chunks like phase B (one observation per warp)
//   (3) cp.async.bulk shared::cta -> global (TMA bulk store) of one observation per warp
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_peak tools/write_peak.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int CS>
__global__ void k_stream(float4 *dst, size_t n4) {
    const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        if (CS) __stcs(dst + i, v); else dst[i] = v;
    }
}

// data dependence: the same grid-stride float4 store with different VALUES (the observations are ~90 % zeros)
template <int PAT>
__global__ void k_pattern(float4 *dst, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v;
        if (PAT == 0) v = make_float4(0.f, 0.f, 0.f, 0.f);
        else if (PAT == 1) v = make_float4(1.f, 1.f, 1.f, 1.f);
        else if (PAT == 2) v = make_float4(1.f, 0.f, 1.f, 0.f);
        else {
            unsigned h = (unsigned)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
            v = make_float4(__uint_as_float(h | 0x3f000000u), __uint_as_float(h * 3u), __uint_as_float(h * 7u), __uint_as_float(h * 11u));
        }
        dst[i] = v;
    }
}

// one warp writes whole "observations" of obs_f4 float4s, observation index strided by total warps
template <int CS>
__global__ void k_obs(float4 *dst, size_t nobs, int obs_f4) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    for (size_t o = warp; o < nobs; o += nwarps) {
        float4 *p = dst + o * obs_f4;
        for (int f = lane; f < obs_f4; f += 32) { if (CS) __stcs(p + f, v); else p[f] = v; }
    }
}

// same, with 256-bit stores (st.global.v8.f32 -> STG.E.ENL2.256, sm_100+): a warp store covers 1024 B
__global__ void k_obs256(float *dst, size_t nobs, int obs_f8) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t o = warp; o < nobs; o += nwarps) {
        float *p = dst + o * obs_f8 * 8;
        for (int f = lane; f < obs_f8; f += 32)
            asm volatile("st.global.v8.f32 [%0], {%1,%1,%2,%2,%1,%1,%2,%2};" ::"l"(p + 8 * (size_t)f), "f"(1.f), "f"(0.f) : "memory");
    }
}

// TMA bulk store: each warp owns NBUF smem buffers of obs_bytes; lanes fill the buffer (st.shared.v4),
// lane 0 issues cp.async.bulk.global.shared::cta; buffers recycle with wait_group.read.
template <int NBUF, int FILL>
__global__ void k_bulk(float4 *dst, size_t nobs, int obs_f4) {
    extern __shared__ __align__(128) float4 smem4[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    float4 *mine = smem4 + (size_t)wib * NBUF * obs_f4;
    const float4 v = make_float4(1.f, 0.f, 1.f, 0.f);
    if (!FILL) {
        for (int f = lane; f < NBUF * obs_f4; f += 32) mine[f] = v;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
    int b = 0;
    for (size_t o = warp; o < nobs; o += nwarps) {
        float4 *buf = mine + (size_t)b * obs_f4;
        if (FILL) {
            // wait until the bulk store that last read this buffer has finished reading
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
            __syncwarp();
            for (int f = lane; f < obs_f4; f += 32) buf[f] = v;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
        }
        if (lane == 0) {
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + o * obs_f4), "r"(s),
                         "r"(obs_f4 * 16)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        b = (b + 1 == NBUF) ? 0 : b + 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static double time_ms(F launch, int iters = 20) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / iters;
}

int main() {
    const size_t bytes = (size_t)65536 * 4 * 4480;            // one tinyCapture step of observations: 1.17 GB
    float4 *dst;
    CK(cudaMalloc(&dst, bytes + 4096));
    CK(cudaMemset(dst, 0, bytes));
    const size_t n4 = bytes / 16;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs; %zu MB per launch\n", prop.name, prop.multiProcessorCount, bytes >> 20);
    const int sms = prop.multiProcessorCount;
    {
        double ms = time_ms([&] { CK(cudaMemsetAsync(dst, 1, bytes)); });
        printf("cudaMemsetAsync                      %8.1f us  %7.1f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    }
    for (int tpb : {128, 256, 512})
        for (int cps : {4, 8, 16}) {
            if (tpb * cps > 2048) continue;
            double a = time_ms([&] { k_stream<0><<<sms * cps, tpb>>>(dst, n4); });
            double b = time_ms([&] { k_stream<1><<<sms * cps, tpb>>>(dst, n4); });
            printf("grid-stride float4  tpb=%3d ctas/SM=%2d  plain %7.1f GB/s   .cs %7.1f GB/s\n", tpb, cps, bytes / a / 1e6,
                   bytes / b / 1e6);
        }
    {
        double z = time_ms([&] { k_pattern<0><<<sms * 8, 256>>>(dst, n4); });
        double o = time_ms([&] { k_pattern<1><<<sms * 8, 256>>>(dst, n4); });
        double m = time_ms([&] { k_pattern<2><<<sms * 8, 256>>>(dst, n4); });
        double r = time_ms([&] { k_pattern<3><<<sms * 8, 256>>>(dst, n4); });
        printf("value patterns (grid-stride float4, 256x8/SM): zeros %7.1f  ones %7.1f  1,0,1,0 %7.1f  hashed %7.1f GB/s\n",
               bytes / z / 1e6, bytes / o / 1e6, bytes / m / 1e6, bytes / r / 1e6);
    }
    for (int obs_bytes : {4480, 4928, 12800}) {
        const int f4 = obs_bytes / 16;
        const size_t nobs = bytes / obs_bytes;
        for (int cps : {4, 8, 16}) {
            double a = time_ms([&] { k_obs<0><<<sms * cps, 128>>>(dst, nobs, f4); });
            double b = time_ms([&] { k_obs<1><<<sms * cps, 128>>>(dst, nobs, f4); });
            printf("warp-per-obs %5d B  ctas/SM=%2d          plain %7.1f GB/s   .cs %7.1f GB/s\n", obs_bytes, cps,
                   nobs * obs_bytes / a / 1e6, nobs * obs_bytes / b / 1e6);
        }
        for (int cps : {4, 8, 16}) {
            double a = time_ms([&] { k_obs256<<<sms * cps, 128>>>((float *)dst, nobs, obs_bytes / 32); });
            printf("warp-per-obs %5d B  ctas/SM=%2d  256-bit stores  %7.1f GB/s\n", obs_bytes, cps, nobs * obs_bytes / a / 1e6);
        }
        for (int cps : {2, 4, 8}) {
            const size_t smem = (size_t)4 * 2 * obs_bytes;
            if (smem * cps > 220 * 1024) continue;
            CK(cudaFuncSetAttribute(k_bulk<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(k_bulk<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            double a = time_ms([&] { k_bulk<2, 0><<<sms * cps, 128, smem>>>(dst, nobs, f4); });
            double b = time_ms([&] { k_bulk<2, 1><<<sms * cps, 128, smem>>>(dst, nobs, f4); });
            printf("bulk S2G     %5d B  ctas/SM=%2d          nofill %7.1f GB/s  fill %7.1f GB/s\n", obs_bytes, cps,
                   nobs * obs_bytes / a / 1e6, nobs * obs_bytes / b / 1e6);
        }
    }
    // one-shot grid like the env kernel: 2048 CTAs x 128 threads, each CTA writes 128 observations
    {
        const int f4 = 4480 / 16;
        const size_t nobs = bytes / 4480;
        double a = time_ms([&] { k_obs<0><<<2048, 128>>>(dst, nobs, f4); });
        printf("warp-per-obs 4480 B, grid 2048x128 (env-kernel geometry)  %7.1f GB/s\n", nobs * 4480 / a / 1e6);
    }
    CK(cudaFree(dst));
    return 0;
}
