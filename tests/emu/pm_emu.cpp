// pm_emu.cpp — TEST-ONLY host emulation of the CUDA kernels' CTA choreography.
//
// Compiles the very same per-thread phase functions the sm_100a kernels run
// (pacman_marl_2025_b200/csrc/pm_env_core.cuh) with g++ and executes them
// "one CTA at a time, one thread at a time" on host memory, so the kernel logic
// can be replayed against the golden fixtures in the GPU-less CI container.
// It is never loaded by the product package: the product path has no CPU
// fallback and fails loudly without the CUDA library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/pacman_b200.h"
#include "../../pacman_marl_2025_b200/csrc/pm_env_core.cuh"

static const int EMU_THREADS = 256;

static int fill(const pm_env_desc *d, PmParams &P, int E, int /*KB*/) { return pm_fill_params(d, P, E); }

template <int MODE>
static void run(const PmParams &P) {
    bool bots = false;
    for (int i = 0; i < 4; ++i) bots |= MODE == PM_MODE_STEP && ((P.bot_kind >> (4 * i)) & 15) >= 2;   // like the C ABI glue
    const bool views = P.view_actor != nullptr;
    // like launch_env_e: one board + a small table -> the bots' distance table is staged in (emulated) shared memory
    int dsw = 0;
    if (bots && P.L == 1 && P.dist && ((P.dist_nmax * P.dist_nmax) & 1) == 0 && pm_dist_stage_words(P.HW, P.dist_nmax) <= PM_DIST_STAGE_MAX_WORDS)
        dsw = pm_dist_stage_words(P.HW, P.dist_nmax);
    const PmSmem M = pm_smem_map(P.HW, P.LS, P.E, P.has_caps, P.L, 4 + (bots ? 1 : 0) + (views ? 2 : 0), dsw, views ? 1 : 0);
    std::vector<uint32_t> smv((size_t)M.total + 4);
    uint32_t *sm = smv.data();
    const int64_t nblk = (P.n_end - P.n_begin + P.E - 1) / P.E;
    for (int64_t b = 0; b < nblk; ++b) {
        std::fill(smv.begin(), smv.end(), 0xdeadbeefu);            // catch reads of unwritten smem
        const int64_t n0 = P.n_begin + b * P.E;
        for (int t = 0; t < 16; ++t) {
            float *v = reinterpret_cast<float *>(sm + M.lut) + 4 * t;
            v[0] = (t & 1) ? 1.f : 0.f; v[1] = (t & 2) ? 1.f : 0.f; v[2] = (t & 4) ? 1.f : 0.f; v[3] = (t & 8) ? 1.f : 0.f;
        }
        for (int i = 0; i < PM_G_NMASKS * P.NCAT; ++i) sm[M.masks + i] = P.geom[PM_G_HDR + i];
        if (P.L == 1) for (int i = 0; i < P.LS; ++i) sm[M.lay + i] = P.layouts[i];
        if (M.dist) {
            const int dw = (P.dist_nmax * P.dist_nmax * 2 + 3) >> 2;
            memcpy(sm + M.dist, P.dist, (size_t)P.dist_nmax * P.dist_nmax * 2);
            memcpy(sm + M.dist + dw, P.node_of_cell, (size_t)P.HW * 2);
        }
        for (int el = 0; el < P.E; ++el) {
            if (MODE == PM_MODE_STEP) {
                PmFetch F;
                pm_phaseA_fetch(P, sm, M, el, n0, F);
                if (bots && views) pm_phaseA_step<1, 1>(P, sm, sm, M, el, n0, F);
                else if (bots) pm_phaseA_step<1, 0>(P, sm, sm, M, el, n0, F);
                else if (views) pm_phaseA_step<0, 1>(P, sm, sm, M, el, n0, F);
                else pm_phaseA_step<0, 0>(P, sm, sm, M, el, n0, F);
            } else if (views) pm_phaseA_static<MODE, 1>(P, sm, sm, M, el, n0);
            else pm_phaseA_static<MODE, 0>(P, sm, sm, M, el, n0);
        }
        if (MODE == PM_MODE_STEP && P.packed && P.a1 == 4)
            for (int t = 0; t < 3 * P.E; ++t) pm_records_out(P, sm, M, t, n0);
        if (P.obs)
            for (int w = 0; w < EMU_THREADS / 32; ++w)
                for (int lane = 0; lane < 32; ++lane) pm_phaseB<MODE>(P, sm, sm, M, w, EMU_THREADS / 32, lane, n0);
        if (views)
            for (int w = 0; w < EMU_THREADS / 32; ++w)
                for (int lane = 0; lane < 32; ++lane)
                    pm_phaseB_views<MODE>(P, sm, sm, M, w, EMU_THREADS / 32, lane, n0, 4 + (bots ? 1 : 0));
        if (P.obs_bits)
            for (int w = 0; w < EMU_THREADS / 32; ++w)
                for (int lane = 0; lane < 32; ++lane) pm_phaseB_bits<MODE>(P, sm, sm, M, w, EMU_THREADS / 32, lane, n0);
    }
}

extern "C" int emu_reset(const pm_env_desc *d, const uint8_t *mask, float *obs, int64_t stride, uint8_t *legal, int E, int KB) {
    PmParams P; if (int rc = fill(d, P, E, KB)) return rc;
    P.mask = mask; P.obs = obs; P.obs_agent_stride = stride; P.legal = legal;
    run<PM_MODE_RESET>(P);
    return 0;
}
extern "C" int emu_observe(const pm_env_desc *d, float *obs, int64_t stride, int E, int KB) {
    PmParams P; if (int rc = fill(d, P, E, KB)) return rc;
    P.obs = obs; P.obs_agent_stride = stride;
    run<PM_MODE_OBSERVE>(P);
    return 0;
}
extern "C" int emu_observe_views(const pm_env_desc *d, int red, uint8_t *actor, int64_t stride, uint8_t *critic, int E, int KB) {
    PmParams P; if (int rc = fill(d, P, E, KB)) return rc;
    P.view_actor = actor; P.view_actor_stride = stride; P.view_critic = critic; P.view_red = red;
    run<PM_MODE_OBSERVE>(P);
    return 0;
}
extern "C" int emu_step(const pm_env_desc *d, const void *actions, int itemsize, const pm_step_out *out, int E, int KB) {
    PmParams P; if (int rc = fill(d, P, E, KB)) return rc;
    P.actions = actions; P.act_size = itemsize;
    P.obs = out->obs; P.obs_agent_stride = out->obs_agent_stride;
    P.reward = out->reward; P.done = out->done; P.legal = out->legal; P.score_change = out->score_change;
    P.shaping = out->shaping; P.shaped = out->shaped_reward; P.packed = out->packed; P.act_used = out->actions_used;
    P.view_actor = out->view_actor; P.view_actor_stride = out->view_actor_stride; P.view_critic = out->view_critic;
    P.view_red = out->view_red;
    P.obs_bits = out->obs_bits; P.obs_bits_stride = out->obs_bits_agent_stride;
    P.obs_elide = out->obs_static_valid != 0;
    run<PM_MODE_STEP>(P);
    return 0;
}
extern "C" int emu_observe_bits(const pm_env_desc *d, uint32_t *bits, int64_t stride, int E, int KB) {
    PmParams P; if (int rc = fill(d, P, E, KB)) return rc;
    P.obs_bits = bits; P.obs_bits_stride = stride;
    run<PM_MODE_OBSERVE>(P);
    return 0;
}
extern "C" int emu_make_observation(const pm_env_desc *d, const uint32_t *draws, uint8_t *visible, int32_t *noisy) {
    PmParams P; if (int rc = fill(d, P, 64, 4)) return rc;
    for (int64_t n = 0; n < P.N; ++n) pm_make_observation_env(P, n, draws, visible, noisy);
    return 0;
}
// host mirror of pm_get_state_kernel
extern "C" int emu_get_state(const pm_env_desc *d, int32_t *agents, uint8_t *food, uint8_t *caps, int32_t *misc) {
    PmParams P; if (int rc = fill(d, P, 64, 4)) return rc;
    const int64_t NP = pm_np(P.N);
    for (int64_t n = 0; n < P.N; ++n) {
        const pm_u4 A = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agA(P.N))[n];
        const pm_u4 B = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agB(P.N))[n];
        const pm_u4 Mi = reinterpret_cast<const pm_u4 *>(P.state + pm_st_misc(P.N))[n];
        const uint32_t a[4] = {A.x, A.y, A.z, A.w}, b[4] = {B.x, B.y, B.z, B.w};
        for (int i = 0; i < 4; ++i) {
            int32_t *o = agents + (n * 4 + i) * 7;
            o[0] = a[i] & 0xff; o[1] = (a[i] >> 8) & 0xff; o[2] = (a[i] >> 16) & 7; o[3] = (a[i] >> 19) & 1;
            o[4] = (a[i] >> 20) & 0xff; o[5] = b[i] & 0xffff; o[6] = b[i] >> 16;
        }
        misc[3 * n] = (int32_t)Mi.x; misc[3 * n + 1] = (int32_t)Mi.y; misc[3 * n + 2] = (int32_t)Mi.w;
        const uint32_t *gfood = reinterpret_cast<const uint32_t *>(P.state + pm_st_food(P.N));
        const uint32_t *gcaps = reinterpret_cast<const uint32_t *>(P.state + pm_st_caps(P.N, P.HW));
        for (int c = 0; c < P.HW; ++c) {
            food[n * P.HW + c] = (gfood[(size_t)(c >> 5) * NP + n] >> (c & 31)) & 1;
            caps[n * P.HW + c] = (gcaps[(size_t)(c >> 5) * NP + n] >> (c & 31)) & 1;
        }
    }
    return 0;
}
