// pm_env.cu — sm_100a kernels + C-ABI entry points for reset / step / observe
// of the batched capture environment.  The per-thread logic lives in
// pm_env_core.cuh; this file owns the CTA choreography, launch geometry and the
// ABI glue (include/pacman_b200.h).
//
// CTA choreography (one CTA = PM_E = 32 consecutive envs, 4 warps; B200 sweeps in DESIGN.md §3.1 / §3.2):
//   the rules threads issue their state loads (pm_phaseA_fetch) -> stage geometry masks (+ the layout record
//   when the batch has one layout) into shared memory -> phase A (lane el of warp 0 steps env n0+el and leaves
//   four snapshots in smem) -> phase B (every warp builds the observation bit
//   string of one (env, agent) at a time in registers and streams it to HBM as
//   fp32 planes with 128-bit stores).  Several CTAs are resident per SM (5 x 96 registers for the plain step,
//   8 x 64 for the views / bots variants), so the latency-bound phase A of one CTA runs underneath the
//   HBM-bound phase B of the others; pm_step_chain extends that overlap ACROSS launches (programmatic dependent
//   launch + one release/acquire counter per env tile).  (Persistent producer/consumer CTAs, a per-sub-move
//   pipeline with named barriers, early stores of the static planes, staggered first waves and
//   higher-occupancy encoders were all measured on B200 and did not beat this
//   plain schedule — DESIGN.md §3.1.)
// HBM traffic per env-step: ~2 x (64 + 4*NW) B of state, 4 B actions, ~70 B of
// small outputs, and 4 * 32*H*W B of observations (98+ % of all bytes).
#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <new>

#include "../../include/pacman_b200.h"
#include "pm_env_core.cuh"

#ifndef PM_THREADS
#define PM_THREADS 128
#endif
#ifndef PM_MIN_CTAS
#define PM_MIN_CTAS 5                             // 96 registers/thread, no spills in the rules phase, 5 CTAs x 4 warps per SM:
#endif                                            // the step with fp32 observations (with or without views), reset, observe -
                                                  // round-2 B200 sweep, DESIGN.md 3.2 (round 1: 8 CTAs x 64 registers)
#ifndef PM_MIN_CTAS_VIEWS_ONLY
#define PM_MIN_CTAS_VIEWS_ONLY 8                  // the views-only step (64 envs per CTA) is bound by its rules phase and by
#endif                                            // instruction issue, not by stores: more resident warps win (86 vs 90 us)
#ifndef PM_MIN_CTAS_BOTS
#define PM_MIN_CTAS_BOTS 8                         // step with the baselineTeam look-ahead (B200 sweep: 4 -> 448 us, 6 -> 400, 8 -> 350 on tiny)
#endif
#ifndef PM_MIN_CTAS_OBS
#define PM_MIN_CTAS_OBS PM_MIN_CTAS                // encoder-only launches (pm_observe); tools/gpu_sweep_obs.sh
#endif
#ifndef PM_E
#define PM_E 32                                   // envs per CTA (power of two <= PM_THREADS)
#endif
#ifndef PM_E_RULES
#define PM_E_RULES 64                             // envs per CTA of the launches bound by the rules phase (see launch_env)
#endif
#ifndef PM_E_CHAIN
#define PM_E_CHAIN PM_E_RULES                     // envs per CTA of the plain step inside a pm_step_chain run (large batches)
#endif
#ifndef PM_E_BOTS
#define PM_E_BOTS PM_E_RULES                      // envs per CTA of the launches with the baselineTeam look-ahead
#endif
#ifndef PM_E_SMALL
#define PM_E_SMALL 8                              // envs per CTA for batches too small to give every SM work at PM_E
#endif

// BOTS = 1: the variant with the baselineTeam look-ahead compiled in (one more bitset slot of shared
// memory as scratch); the plain variant pays nothing for it.
// VIEWS = 1: also writes the trainer's learner views (uint8, canonicalised + merged); two more bitset
// slots hold the mirrored snapshots of red learners.
#ifdef PM_TRACE
// dev instrumentation (tools/trace_step.py): [grid][4] = smid, t_start, t_rules_done, t_end (globaltimer ns)
__device__ unsigned long long *pm_trace_buf;
__device__ __forceinline__ unsigned long long pm_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
extern "C" int pm_trace_set(unsigned long long *buf) { return (int)cudaMemcpyToSymbol(pm_trace_buf, &buf, sizeof(buf)); }
#endif

// E = envs per CTA: PM_E (32) for big batches; PM_E_SMALL (8) when the batch would otherwise leave most SMs
// without a CTA (a trainer-sized batch of 1 024 envs is 32 CTAs of 32 envs but 128 CTAs of 8).
template <int MODE, int BOTS, int VIEWS, int E>
__global__ void __launch_bounds__(PM_THREADS, BOTS ? PM_MIN_CTAS_BOTS
                                                    : MODE == PM_MODE_OBSERVE ? PM_MIN_CTAS_OBS
                                                    : (MODE == PM_MODE_STEP && VIEWS && E == PM_E_RULES) ? PM_MIN_CTAS_VIEWS_ONLY : PM_MIN_CTAS)
pm_env_kernel(const PmParams P) {
    extern __shared__ __align__(16) uint32_t sm[];
    const PmSmem M = pm_smem_map(P.HW, P.LS, E, P.has_caps, P.L, 4 + BOTS + 2 * VIEWS, BOTS ? P.dist_stage_words : 0, VIEWS);
    const int tid = threadIdx.x;
    const int64_t n0 = P.n_begin + (int64_t)blockIdx.x * E;
    // pm_step_chain: the next launch of the run may take the SM slots this launch no longer needs as soon as every CTA
    // of this one has started (programmatic dependent launch); what orders the two is the per-tile counter below
    const bool chained = MODE == PM_MODE_STEP && P.chain_flags != nullptr;
    if (chained) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (chained && P.chain_k > 0) {
        // Launch k of a chain started before launch k-1 finished.  This CTA's envs, observation rows and output records
        // were last touched by the CTA with the same index of launch k-1: wait until that CTA has published k
        // (release / acquire at GPU scope; the barrier extends the ordering to the whole CTA).  Launch k-1's CTAs
        // are all resident or finished before any CTA of launch k is scheduled, so the wait always ends.
        if (tid == 0) {
            const uint32_t *f = P.chain_flags + blockIdx.x;
            uint32_t v = 0;
            for (unsigned spin = 0;; ++spin) {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if (v == P.chain_k) break;
                if (spin > (1u << 24)) __trap();        // (seconds: the chain was mis-used; fail instead of hanging the GPU)
                __nanosleep(100);
            }
            __threadfence();
        }
        __syncthreads();
    }
    // the rules threads' state loads go out first, so that they are in flight under the table staging below
    PmFetch F;
    if (MODE == PM_MODE_STEP && tid < E) pm_phaseA_fetch(P, sm, M, tid, n0, F);
#ifdef PM_TRACE
    if (MODE == PM_MODE_STEP && tid == 0 && pm_trace_buf) {
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        pm_trace_buf[4 * blockIdx.x + 0] = smid; pm_trace_buf[4 * blockIdx.x + 1] = pm_gtime();
    }
#endif

    // ---- stage per-batch tables ----
    if (tid < 16) {
        float4 v = make_float4((tid & 1) ? 1.f : 0.f, (tid & 2) ? 1.f : 0.f, (tid & 4) ? 1.f : 0.f, (tid & 8) ? 1.f : 0.f);
        reinterpret_cast<float4 *>(sm + M.lut)[tid] = v;
    }
    for (int i = tid; i < PM_G_NMASKS * P.NCAT; i += PM_THREADS) sm[M.masks + i] = __ldg(P.geom + PM_G_HDR + i);
    if (P.L == 1)
        for (int i = tid; i < P.LS; i += PM_THREADS) sm[M.lay + i] = __ldg(P.layouts + i);
    if (BOTS && M.dist) {
        // one board for the whole batch and a table that fits: the baselineTeam look-ahead reads its maze distances
        // (Distancer.getDistance, distanceCalculator.py:38-62) from shared memory instead of L2
        const int dw = (P.dist_nmax * P.dist_nmax * 2 + 3) >> 2, nw = (P.HW * 2 + 3) >> 2;
        const uint32_t *gd = reinterpret_cast<const uint32_t *>(P.dist), *gn = reinterpret_cast<const uint32_t *>(P.node_of_cell);
        for (int i = tid; i < dw; i += PM_THREADS) sm[M.dist + i] = __ldg(gd + i);
        for (int i = tid; i < nw; i += PM_THREADS) sm[M.dist + dw + i] = __ldg(gn + i);
    }
    __syncthreads();

    // ---- phase A: rules, one thread per env ----
    if (tid < E) {
        if (MODE == PM_MODE_STEP) pm_phaseA_step<BOTS, VIEWS>(P, sm, sm, M, tid, n0, F);
        else pm_phaseA_static<MODE, VIEWS>(P, sm, sm, M, tid, n0);
    }
    __syncthreads();
#ifdef PM_TRACE
    if (MODE == PM_MODE_STEP && tid == 0 && pm_trace_buf) pm_trace_buf[4 * blockIdx.x + 2] = pm_gtime();
#endif
    if (MODE == PM_MODE_STEP && P.packed && P.a1 == 4)
        for (int t = tid; t < 3 * E; t += PM_THREADS) pm_records_out(P, sm, M, t, n0);
    // ---- phase B: observations, one warp per (env, agent) ----
    if (P.obs != nullptr) pm_phaseB<MODE>(P, sm, sm, M, tid >> 5, PM_THREADS / 32, tid & 31, n0);
    // ---- phase B': learner views, one (view, word) item per lane ----
    if (VIEWS && P.view_actor != nullptr) pm_phaseB_views<MODE>(P, sm, sm, M, tid >> 5, PM_THREADS / 32, tid & 31, n0, 4 + BOTS);
    // ---- phase B'': packed observations for a host-side consumer ----
    if (P.obs_bits != nullptr) pm_phaseB_bits<MODE>(P, sm, sm, M, tid >> 5, PM_THREADS / 32, tid & 31, n0);
#ifdef PM_TRACE
    __syncthreads();
    if (MODE == PM_MODE_STEP && tid == 0 && pm_trace_buf) pm_trace_buf[4 * blockIdx.x + 3] = pm_gtime();
#endif
    if (chained) {
        // every store of this CTA (state, observations, records) happens-before the counter the next launch polls
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(P.chain_flags + blockIdx.x), "r"(P.chain_k + 1u) : "memory");
        }
    }
}

// ---- state export / import (GameState accessors; tests, checkpoints) ---------
__global__ void pm_get_state_kernel(const PmParams P, int32_t *agents, uint8_t *food, uint8_t *caps, int32_t *misc) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.N) return;
    const int64_t NP = pm_np(P.N);
    const pm_u4 A = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agA(P.N))[n];
    const pm_u4 B = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agB(P.N))[n];
    const pm_u4 Mi = reinterpret_cast<const pm_u4 *>(P.state + pm_st_misc(P.N))[n];
    const uint32_t a[4] = {A.x, A.y, A.z, A.w}, b[4] = {B.x, B.y, B.z, B.w};
    if (agents)
        for (int i = 0; i < 4; ++i) {
            int32_t *o = agents + (n * 4 + i) * 7;
            o[0] = a[i] & 0xff; o[1] = (a[i] >> 8) & 0xff; o[2] = (a[i] >> 16) & 7; o[3] = (a[i] >> 19) & 1;
            o[4] = (a[i] >> 20) & 0xff; o[5] = b[i] & 0xffff; o[6] = b[i] >> 16;
        }
    if (misc) { misc[3 * n] = (int32_t)Mi.x; misc[3 * n + 1] = (int32_t)Mi.y; misc[3 * n + 2] = (int32_t)Mi.w; }
    const uint32_t *gfood = reinterpret_cast<const uint32_t *>(P.state + pm_st_food(P.N));
    const uint32_t *gcaps = reinterpret_cast<const uint32_t *>(P.state + pm_st_caps(P.N, P.HW));
    for (int c = 0; c < P.HW; ++c) {
        if (food) food[n * P.HW + c] = (gfood[(size_t)(c >> 5) * NP + n] >> (c & 31)) & 1;
        if (caps) caps[n * P.HW + c] = (gcaps[(size_t)(c >> 5) * NP + n] >> (c & 31)) & 1;
    }
}

__global__ void pm_set_state_kernel(const PmParams P, const int32_t *agents, const uint8_t *food, const uint8_t *caps,
                                    const int32_t *misc) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= P.N) return;
    const int64_t NP = pm_np(P.N);
    uint32_t a[4], b[4], s[4];
    for (int i = 0; i < 4; ++i) {
        const int32_t *o = agents + (n * 4 + i) * 7;
        a[i] = pm_agA_pack(o[0], o[1], o[2], o[3], o[4]);
        b[i] = (uint32_t)o[5] | (uint32_t)o[6] << 16;
        s[i] = (uint32_t)o[0] | (uint32_t)o[1] << 8 | (uint32_t)o[5] << 16;
    }
    pm_u4 A, B, S, Mi;
    A.x = a[0]; A.y = a[1]; A.z = a[2]; A.w = a[3];
    B.x = b[0]; B.y = b[1]; B.z = b[2]; B.w = b[3];
    S.x = s[0]; S.y = s[1]; S.z = s[2]; S.w = s[3];
    uint32_t *gfood = reinterpret_cast<uint32_t *>(P.state + pm_st_food(P.N));
    uint32_t *gcaps = reinterpret_cast<uint32_t *>(P.state + pm_st_caps(P.N, P.HW));
    int rf = 0, bf = 0;
    for (int w = 0; w < P.NW; ++w) {
        uint32_t fw = 0, cw = 0;
        for (int bit = 0; bit < 32; ++bit) {
            const int c = w * 32 + bit;
            if (c >= P.HW) break;
            if (food[n * P.HW + c]) { fw |= 1u << bit; if (c % P.W < (P.W >> 1)) ++rf; else ++bf; }
            if (caps[n * P.HW + c]) cw |= 1u << bit;
        }
        gfood[(size_t)w * NP + n] = fw;
        gcaps[(size_t)w * NP + n] = cw;
    }
    Mi.x = (uint32_t)misc[3 * n]; Mi.y = (uint32_t)misc[3 * n + 1]; Mi.z = (uint32_t)rf | (uint32_t)bf << 16;
    Mi.w = (uint32_t)misc[3 * n + 2];                  // episode counter: keys the opponent draws and the layout pool
    reinterpret_cast<pm_u4 *>(P.state + pm_st_agA(P.N))[n] = A;
    reinterpret_cast<pm_u4 *>(P.state + pm_st_agB(P.N))[n] = B;
    reinterpret_cast<pm_u4 *>(P.state + pm_st_snap(P.N))[n] = S;
    reinterpret_cast<pm_u4 *>(P.state + pm_st_misc(P.N))[n] = Mi;
}

// ---------------------------------------------------------------------------
// host glue
// ---------------------------------------------------------------------------
#define PM_MAX_DEVICES 64

// SM count of the current device (cached per device, mutex-guarded); < 0 = PM_E* error
static int pm_sm_count() {
    static int count[PM_MAX_DEVICES];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PM_ENOGPU;
    if (dev < 0 || dev >= PM_MAX_DEVICES) return PM_EINVAL;
    std::lock_guard<std::mutex> lock(mu);
    if (!count[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        count[dev] = v;
    }
    return count[dev];
}

static int fill_params(const pm_env_desc *d, PmParams &P) { return pm_fill_params(d, P, PM_E); }

static void set_envs_per_cta(PmParams &P, int E) {
    P.E = E; P.lgE = 0;
    while ((1 << P.lgE) < E) ++P.lgE;
}

template <int MODE, int BOTS, int VIEWS, int E>
static int launch_env_e(const PmParams &P0, cudaStream_t st) {
    PmParams P = P0;
    set_envs_per_cta(P, E);
    P.dist_stage_words = 0;
    if (BOTS && P.L == 1 && P.dist && (((size_t)P.dist | (size_t)P.node_of_cell) & 3) == 0 && ((P.dist_nmax * P.dist_nmax) & 1) == 0 &&
        pm_dist_stage_words(P.HW, P.dist_nmax) <= PM_DIST_STAGE_MAX_WORDS)
        P.dist_stage_words = pm_dist_stage_words(P.HW, P.dist_nmax);
    const PmSmem M = pm_smem_map(P.HW, P.LS, E, P.has_caps, P.L, 4 + BOTS + 2 * VIEWS, BOTS ? P.dist_stage_words : 0, VIEWS);
    const size_t smem = (size_t)M.total * 4;
    // Per-device cache of the dynamic shared-memory size this kernel instance is configured for.  The launch stays
    // under the lock whenever the attribute may still change, so another host thread can neither launch with a size
    // that is not configured yet nor shrink it in between; the common path (already large enough) takes the lock
    // only to read.
    static size_t configured[PM_MAX_DEVICES];
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PM_ENOGPU;
    if (dev < 0 || dev >= PM_MAX_DEVICES) return PM_EINVAL;
    const unsigned grid = (unsigned)((P.n_end - P.n_begin + E - 1) / E);
    std::lock_guard<std::mutex> lock(mu);
    if (smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(pm_env_kernel<MODE, BOTS, VIEWS, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        // the kernel streams its output and keeps its working set in shared memory: prefer smem over L1
        e = cudaFuncSetAttribute(pm_env_kernel<MODE, BOTS, VIEWS, E>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        configured[dev] = smem;
    }
    if (MODE == PM_MODE_STEP && P.chain_flags && P.chain_k > 0) {
        // launch k > 0 of a chain: allowed to start while the previous launch of the stream (launch k-1) is still running
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return (int)cudaLaunchKernelEx(&cfg, pm_env_kernel<MODE, BOTS, VIEWS, E>, P);
    }
    pm_env_kernel<MODE, BOTS, VIEWS, E><<<grid, PM_THREADS, smem, st>>>(P);
    return (int)cudaGetLastError();
}

template <int MODE, int BOTS, int VIEWS>
static int launch_env(const PmParams &P, cudaStream_t st) {
    // fewer than ~2 CTAs of PM_E envs per SM: use small CTAs so that the whole GPU takes part
    // (pm_env_desc.envs_per_cta overrides the choice)
    bool small = P.epc == PM_E_SMALL;                   // (epc == 32 forces the large-batch choice)
    if (P.epc == 0) {
        int sms = pm_sm_count();
        if (sms < 0) return sms;
        small = (P.n_end - P.n_begin) < (int64_t)2 * PM_E * sms;
    }
    if (small) return launch_env_e<MODE, BOTS, VIEWS, PM_E_SMALL>(P, st);
    // Launches whose time is set by the rules phase rather than by the fp32 stores - the baselineTeam look-ahead, and
    // steps that write only the uint8 learner views - run PM_E_RULES envs per CTA: twice the rules warps per CTA and
    // all CTAs resident in one wave (B200, tinyCapture 65,536 envs: bots 292 -> 274 us, bots + views 178 -> 143 us,
    // views only 94 -> 86 us; the fp32 step itself is 3 % slower that way and stays at PM_E).
    if constexpr (MODE == PM_MODE_STEP && BOTS) {
        return launch_env_e<MODE, BOTS, VIEWS, PM_E_BOTS>(P, st);
    } else {
        if constexpr (MODE == PM_MODE_STEP && VIEWS)
            if (P.obs == nullptr) return launch_env_e<MODE, BOTS, VIEWS, PM_E_RULES>(P, st);
        // the launches of a chain overlap, so the wave structure of a single launch does not matter to them and the
        // larger tiles win (tinyCapture 65,536 envs: 175.3 -> 171.7 us per chained step; alone: 194 -> 219 us)
        if constexpr (MODE == PM_MODE_STEP && !VIEWS)
            if (P.chain_flags != nullptr && (P.n_end - P.n_begin) >= (int64_t)4 * PM_E_CHAIN * pm_sm_count())
                return launch_env_e<MODE, BOTS, VIEWS, PM_E_CHAIN>(P, st);
        return launch_env_e<MODE, BOTS, VIEWS, PM_E>(P, st);
    }
}

static int set_views(PmParams &P, uint8_t *actor, int64_t stride, uint8_t *critic, int red) {
    if (!actor && !critic) return 0;
    if (!actor || !critic || P.W > 32 || ((size_t)actor & 7) || ((size_t)critic & 7) || (stride & 7)) return PM_EINVAL;
    P.view_actor = actor; P.view_actor_stride = stride; P.view_critic = critic; P.view_red = red != 0;
    return 0;
}

static int set_bits(PmParams &P, uint32_t *bits, int64_t stride) {
    if (!bits) return 0;
    if (((size_t)bits & 15) || (stride & 3) || stride < (int64_t)P.NB) return PM_EINVAL;
    P.obs_bits = bits; P.obs_bits_stride = stride;
    return 0;
}

static bool has_lookahead_bots(const PmParams &P) {
    for (int i = 0; i < 4; ++i) if (((P.bot_kind >> (4 * i)) & 15) >= 2) return true;
    return false;
}

extern "C" int pm_reset(const pm_env_desc *d, const uint8_t *mask, float *obs, int64_t obs_agent_stride, uint8_t *legal,
                        pm_stream_t stream) {
    PmParams P;
    if (int rc = fill_params(d, P)) return rc;
    if (obs && ((obs_agent_stride & 3) || ((size_t)obs & 15))) return PM_EINVAL;      // 128-bit stores
    P.mask = mask; P.obs = obs; P.obs_agent_stride = obs_agent_stride; P.legal = legal;
    return launch_env<PM_MODE_RESET, 0, 0>(P, (cudaStream_t)stream);
}

extern "C" int pm_observe(const pm_env_desc *d, float *obs, int64_t obs_agent_stride, pm_stream_t stream) {
    PmParams P;
    if (int rc = fill_params(d, P)) return rc;
    if (!obs || (obs_agent_stride & 3) || ((size_t)obs & 15)) return PM_EINVAL;
    P.obs = obs; P.obs_agent_stride = obs_agent_stride;
    return launch_env<PM_MODE_OBSERVE, 0, 0>(P, (cudaStream_t)stream);
}

extern "C" int pm_observe_bits(const pm_env_desc *d, uint32_t *obs_bits, int64_t obs_bits_agent_stride, pm_stream_t stream) {
    PmParams P;
    if (int rc = fill_params(d, P)) return rc;
    if (!obs_bits) return PM_EINVAL;
    if (int rc = set_bits(P, obs_bits, obs_bits_agent_stride)) return rc;
    return launch_env<PM_MODE_OBSERVE, 0, 0>(P, (cudaStream_t)stream);
}

extern "C" int pm_observe_views(const pm_env_desc *d, int view_red, uint8_t *view_actor, int64_t view_actor_stride,
                                uint8_t *view_critic, pm_stream_t stream) {
    PmParams P;
    if (int rc = fill_params(d, P)) return rc;
    if (!view_actor || !view_critic) return PM_EINVAL;
    if (int rc = set_views(P, view_actor, view_actor_stride, view_critic, view_red)) return rc;
    return launch_env<PM_MODE_OBSERVE, 0, 1>(P, (cudaStream_t)stream);
}

static int step_impl(const pm_env_desc *d, const void *actions, int action_itemsize, const pm_step_out *out,
                     pm_stream_t stream, uint32_t *chain_flags, uint32_t chain_k) {
    PmParams P;
    if (int rc = fill_params(d, P)) return rc;
    P.chain_flags = chain_flags; P.chain_k = chain_k;
    if (!actions || !out || (action_itemsize != 1 && action_itemsize != 8)) return PM_EINVAL;
    if (out->obs && ((out->obs_agent_stride & 3) || ((size_t)out->obs & 15))) return PM_EINVAL;
    P.actions = actions; P.act_size = action_itemsize;
    P.obs = out->obs; P.obs_agent_stride = out->obs_agent_stride;
    P.obs_elide = out->obs_static_valid != 0;
    P.reward = out->reward; P.done = out->done; P.legal = out->legal; P.score_change = out->score_change;
    P.shaping = out->shaping; P.shaped = out->shaped_reward; P.packed = out->packed; P.act_used = out->actions_used;
    if (int rc = set_views(P, out->view_actor, out->view_actor_stride, out->view_critic, out->view_red)) return rc;
    if (P.view_actor && (P.a0 != 0 || P.a1 != 4)) return PM_EINVAL;                    // views need both learners' snapshots
    if (int rc = set_bits(P, out->obs_bits, out->obs_bits_agent_stride)) return rc;
    const bool bots = has_lookahead_bots(P), views = P.view_actor != nullptr;
    // steps with the baselineTeam look-ahead are bound by their rules phase and run as ONE wave of 64-env CTAs: there
    // is no store-only tail for the next launch to hide under, and sharing the SMs between two steps costs more than
    // it returns (B200, tinyCapture 65,536 envs: 142 us per step alone, 198 us overlapped) - a chain of such steps is
    // issued as ordinary launches
    if (bots) P.chain_flags = nullptr;
    if (bots) return views ? launch_env<PM_MODE_STEP, 1, 1>(P, (cudaStream_t)stream) : launch_env<PM_MODE_STEP, 1, 0>(P, (cudaStream_t)stream);
    return views ? launch_env<PM_MODE_STEP, 0, 1>(P, (cudaStream_t)stream) : launch_env<PM_MODE_STEP, 0, 0>(P, (cudaStream_t)stream);
}

extern "C" int pm_step(const pm_env_desc *d, const void *actions, int action_itemsize, const pm_step_out *out,
                       pm_stream_t stream) {
    return step_impl(d, actions, action_itemsize, out, stream, nullptr, 0);
}

// ---- a run of steps whose actions are already resident ------------------------------
extern "C" size_t pm_chain_scratch_bytes(int64_t n_envs) {
    return n_envs > 0 ? (size_t)((n_envs + PM_E_SMALL - 1) / PM_E_SMALL) * sizeof(uint32_t) : 0;   // one counter per CTA
}

extern "C" int pm_step_chain(const pm_env_desc *d, const void *actions, int action_itemsize, int64_t actions_step_stride,
                             int n_steps, const pm_step_out *outs, int n_outs, void *scratch, pm_stream_t stream) {
    if (!d || !actions || !outs || n_steps < 1 || n_outs < 1 || !scratch || ((size_t)scratch & 3) || actions_step_stride < 0)
        return PM_EINVAL;
    if (action_itemsize != 1 && action_itemsize != 8) return PM_EINVAL;
    if (d->agent_begin || d->agent_end || d->env_begin || d->env_count) return PM_EINVAL;     // whole steps of the whole batch
    // the counters start at 0; launch 0 is an ordinary (fully stream-ordered) launch, so everything queued before the
    // chain - this memset included - is complete before any launch of the chain reads or writes
    cudaError_t e = cudaMemsetAsync(scratch, 0, pm_chain_scratch_bytes(d->n_envs), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    const char *a = (const char *)actions;
    for (int k = 0; k < n_steps; ++k) {
        const int rc = step_impl(d, a + (size_t)k * (size_t)actions_step_stride * (size_t)action_itemsize, action_itemsize,
                                 outs + (k % n_outs), stream, (uint32_t *)scratch, (uint32_t)k);
        if (rc) return rc;
    }
    return 0;
}

// ---- host-buffer step ------------------------------------------------------
extern "C" int pm_step_host(const pm_env_desc *d, const uint8_t *actions_host, uint8_t *actions_dev,
                            const pm_step_out *od, const pm_step_out *oh, pm_stream_t stream) {
    if (!d || !actions_host || !actions_dev || !od || !oh) return PM_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t N = (size_t)d->n_envs, HW = (size_t)d->W * d->H;
    cudaError_t e = cudaMemcpyAsync(actions_dev, actions_host, N * 4, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    if (int rc = pm_step(d, actions_dev, 1, od, stream)) return rc;
#define PM_D2H(member, bytes)                                                                  \
    if (oh->member) {                                                                          \
        if (!od->member) return PM_EINVAL;                                                     \
        e = cudaMemcpyAsync(oh->member, od->member, bytes, cudaMemcpyDeviceToHost, st);        \
        if (e != cudaSuccess) return (int)e;                                                   \
    }
    PM_D2H(reward, N * 16)
    PM_D2H(done, N)
    PM_D2H(legal, N * 4)
    PM_D2H(score_change, N * 4)
    PM_D2H(shaping, N * 32)
    PM_D2H(shaped_reward, N * 16)
#undef PM_D2H
    if (oh->obs) {
        if (!od->obs) return PM_EINVAL;
        for (int a = 0; a < 4; ++a) {
            e = cudaMemcpyAsync(oh->obs + a * oh->obs_agent_stride, od->obs + a * od->obs_agent_stride, N * 8 * HW * 4,
                                cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    // ABI 7: the compact per-step products a host-side consumer may want instead of the fp32 observations
    if (oh->view_actor || oh->view_critic) {
        if (!od->view_actor || !od->view_critic) return PM_EINVAL;
        for (int l = 0; l < 2 && oh->view_actor; ++l) {
            e = cudaMemcpyAsync(oh->view_actor + l * oh->view_actor_stride, od->view_actor + l * od->view_actor_stride,
                                N * 8 * HW, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return (int)e;
        }
        if (oh->view_critic) {
            e = cudaMemcpyAsync(oh->view_critic, od->view_critic, N * 8 * HW, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    if (oh->obs_bits) {
        if (!od->obs_bits) return PM_EINVAL;
        const size_t nb = (size_t)pm_obs_bits_words_((int)HW);
        for (int a = 0; a < 4; ++a) {
            e = cudaMemcpyAsync(oh->obs_bits + a * oh->obs_bits_agent_stride, od->obs_bits + a * od->obs_bits_agent_stride,
                                N * nb * 4, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return (int)e;
        }
    }
    return (int)cudaStreamSynchronize(st);
}

// ---- pipelined host-buffer step ------------------------------------------------
struct pm_pipe {
    int n_chunks, device;
    cudaStream_t s[2];
    cudaEvent_t start, done[2];
};

extern "C" int pm_pipe_create(pm_pipe **pipe, int n_chunks) {
    if (!pipe || n_chunks < 1 || n_chunks > 64) return PM_EINVAL;
    pm_pipe *p = new pm_pipe();
    p->n_chunks = n_chunks;
    if (cudaGetDevice(&p->device) != cudaSuccess) { delete p; return PM_ENOGPU; }
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&p->s[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->start, cudaEventDisableTiming);
    if (e != cudaSuccess) { pm_pipe_destroy(p); return (int)e; }       // releases whatever was created so far
    *pipe = p;
    return 0;
}

extern "C" int pm_pipe_destroy(pm_pipe *p) {
    if (!p) return PM_EINVAL;
    for (int i = 0; i < 2; ++i) {
        if (p->s[i]) cudaStreamDestroy(p->s[i]);
        if (p->done[i]) cudaEventDestroy(p->done[i]);
    }
    if (p->start) cudaEventDestroy(p->start);
    delete p;
    return 0;
}

extern "C" int pm_step_host_pipelined(const pm_env_desc *d, pm_pipe *p, const uint8_t *actions_host, uint8_t *actions_dev,
                                      const pm_step_out *od, pm_host_record *records_host, pm_stream_t stream) {
    if (!d || !p || !actions_host || !actions_dev || !od || !od->packed || !records_host) return PM_EINVAL;
    if (d->env_begin != 0 || d->env_count != 0) return PM_EINVAL;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != p->device) return PM_EINVAL;      // a pipe belongs to one device
    cudaStream_t caller = (cudaStream_t)stream;
    const int64_t N = d->n_envs;
    const int C = p->n_chunks;
    int64_t per = ((N + C - 1) / C + 31) & ~(int64_t)31;           // chunk size, multiple of 32 envs
    cudaError_t e = cudaEventRecord(p->start, caller);
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < 2; ++i)
        if ((e = cudaStreamWaitEvent(p->s[i], p->start, 0)) != cudaSuccess) return (int)e;
    pm_env_desc dc = *d;
    int used = 0, rc = 0;
    for (int64_t b = 0; b < N && rc == 0; b += per, ++used) {
        const int64_t cnt = (N - b < per) ? N - b : per;
        cudaStream_t st = p->s[used & 1];
        e = cudaMemcpyAsync(actions_dev + 4 * b, actions_host + 4 * b, (size_t)cnt * 4, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) { rc = (int)e; break; }
        dc.env_begin = b; dc.env_count = cnt;
        if ((rc = pm_step(&dc, actions_dev, 1, od, (pm_stream_t)st)) != 0) break;
        e = cudaMemcpyAsync(records_host + b, od->packed + b, (size_t)cnt * sizeof(pm_host_record), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = (int)e;
    }
    // Always join both internal streams back into the caller's stream - also after a failure, so that later work
    // on the caller's stream cannot overtake the ranges that were already queued.
    for (int i = 0; i < 2; ++i) {
        if ((e = cudaEventRecord(p->done[i], p->s[i])) == cudaSuccess) e = cudaStreamWaitEvent(caller, p->done[i], 0);
        if (e != cudaSuccess && rc == 0) rc = (int)e;
    }
    for (int i = 0; i < 2; ++i)
        if ((e = cudaStreamSynchronize(p->s[i])) != cudaSuccess && rc == 0) rc = (int)e;
    return rc;
}

// ---- zero-copy host step ----------------------------------------------------------
extern "C" int pm_step_host_mapped_async(const pm_env_desc *d, const uint8_t *actions_pinned, const pm_step_out *od,
                                         pm_host_record *records_pinned, pm_stream_t stream) {
    if (!d || !actions_pinned || !od || !records_pinned) return PM_EINVAL;
    void *a_dev = nullptr, *r_dev = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&a_dev, (void *)actions_pinned, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return PM_EINVAL; }        // not pinned / not mapped
    e = cudaHostGetDevicePointer(&r_dev, (void *)records_pinned, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return PM_EINVAL; }
    pm_step_out o = *od;
    o.packed = (pm_host_record *)r_dev;
    if (o.obs_bits) {                                   // may itself be mapped pinned host memory (zero-copy bits)
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, o.obs_bits) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            o.obs_bits = (uint32_t *)at.devicePointer;
        else
            cudaGetLastError();
    }
    return pm_step(d, a_dev, 1, &o, stream);
}

// ---- a run of steps fed from / delivered to pinned host memory ----------------------------------------
extern "C" int pm_step_chain_host_mapped(const pm_env_desc *d, const uint8_t *actions_pinned, int n_steps, const pm_step_out *od,
                                         pm_host_record *records_pinned, void *scratch, pm_stream_t stream) {
    if (!d || !actions_pinned || !od || !records_pinned || n_steps < 1 || n_steps > 65536) return PM_EINVAL;
    void *a_dev = nullptr, *r_dev = nullptr;
    cudaError_t e = cudaHostGetDevicePointer(&a_dev, (void *)actions_pinned, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return PM_EINVAL; }        // not pinned / not mapped
    e = cudaHostGetDevicePointer(&r_dev, (void *)records_pinned, 0);
    if (e != cudaSuccess) { cudaGetLastError(); return PM_EINVAL; }
    pm_step_out *outs = new (std::nothrow) pm_step_out[n_steps];
    if (!outs) return PM_EINVAL;
    for (int t = 0; t < n_steps; ++t) {
        outs[t] = *od;
        outs[t].packed = (pm_host_record *)r_dev + (size_t)t * (size_t)d->n_envs;
    }
    const int rc = pm_step_chain(d, a_dev, 1, 4 * d->n_envs, n_steps, outs, n_steps, scratch, stream);
    delete[] outs;
    return rc;
}

extern "C" int pm_step_host_mapped(const pm_env_desc *d, const uint8_t *actions_pinned, const pm_step_out *od,
                                   pm_host_record *records_pinned, pm_stream_t stream) {
    if (int rc = pm_step_host_mapped_async(d, actions_pinned, od, records_pinned, stream)) return rc;
    return (int)cudaStreamSynchronize((cudaStream_t)stream);
}

__global__ void pm_make_observation_kernel(const PmParams P, const uint32_t *draws, uint8_t *visible, int32_t *noisy) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < P.N) pm_make_observation_env(P, n, draws, visible, noisy);
}

extern "C" int pm_make_observation(const pm_env_desc *d, const uint32_t *draws, uint8_t *visible, int32_t *noisy_distances,
                                   pm_stream_t stream) {
    PmParams P;
    if (int rc = fill_params(d, P)) return rc;
    if (!visible || !noisy_distances) return PM_EINVAL;
    const unsigned grid = (unsigned)((P.N + 127) / 128);
    pm_make_observation_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, draws, visible, noisy_distances);
    return (int)cudaGetLastError();
}

extern "C" int pm_get_state(const pm_env_desc *d, int32_t *agents, uint8_t *food, uint8_t *caps, int32_t *misc,
                            pm_stream_t stream) {
    PmParams P;
    if (int rc = fill_params(d, P)) return rc;
    const unsigned grid = (unsigned)((P.N + 127) / 128);
    pm_get_state_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, agents, food, caps, misc);
    return (int)cudaGetLastError();
}

extern "C" int pm_set_state(const pm_env_desc *d, const int32_t *agents, const uint8_t *food, const uint8_t *caps,
                            const int32_t *misc, pm_stream_t stream) {
    PmParams P;
    if (int rc = fill_params(d, P)) return rc;
    if (!agents || !food || !caps || !misc) return PM_EINVAL;
    const unsigned grid = (unsigned)((P.N + 127) / 128);
    pm_set_state_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(P, agents, food, caps, misc);
    return (int)cudaGetLastError();
}
