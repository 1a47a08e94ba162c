// pm_env_core.cuh — the rules + observation logic of the batched capture env,
// written as three per-thread phases that the sm_100a kernel (pm_env.cu) runs
// back to back inside one CTA:
//
//   phase A  (one thread per env)   the four ordered sub-moves of
//            gymPacMan.step (gymPacMan.py:143-193): legal actions, move, pacman
//            flag, food return (+ the agentState rebinding of capture.py:505),
//            eating, capsules, collisions / respawns / food dump, scared timer,
//            score, rewards, termination, shaping — and after each sub-move a
//            compact SNAPSHOT (board bitsets + four cell positions) in shared
//            memory, because the reference observes agent i right after ITS
//            sub-move (gymPacMan.py:166).
//   phase B  (one warp per observation) snapshot -> observation bit string
//            (see pm_format.h) in registers -> fp32 planes with coalesced
//            128-bit stores; this is the HBM-bound part.
//
// The functions are PM_HD (host + device) on purpose: tests/emu compiles the
// very same code with g++ and replays the golden fixtures through it without a
// GPU.  That emulation is test-only; the product never runs it.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "pm_format.h"

#if defined(__CUDACC__)
typedef uint4 pm_u4;
typedef float4 pm_f4;
#define PM_NOINLINE __host__ __device__ __noinline__
#else
struct alignas(16) pm_u4 { uint32_t x, y, z, w; };
struct alignas(16) pm_f4 { float x, y, z, w; };
#define PM_NOINLINE inline
#endif

#define PM_SCARED_TIME 40   // capture.py:75
#define PM_MODE_STEP 0
#define PM_MODE_RESET 1
#define PM_MODE_OBSERVE 2

// ---- exact fp64 helpers: the reference accumulates rewards as Python floats,
// one rounding per operation; never let the compiler contract into FMA --------
PM_HD double pm_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
PM_HD double pm_dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
PM_HD uint64_t pm_dbits(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
PM_HD double pm_dfrom(uint32_t lo, uint32_t hi) {
    const uint64_t u = (uint64_t)lo | (uint64_t)hi << 32;
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}
PM_HD uint32_t pm_fbits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PM_HD uint32_t pm_funnel_r(uint32_t lo, uint32_t hi, int r) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, r);
#else
    return r ? (lo >> r) | (hi << (32 - r)) : lo;
#endif
}

struct PmParams {
    int W, H, HW, NW, NCAT, ND, LS;
    int L, has_caps;                    // layouts in table, capsules anywhere
    int E, lgE;                         // envs per CTA (power of two), log2(E)
    int64_t N;
    const uint32_t *geom, *layouts, *layout_id;
    uint8_t *state;
    int length, reward_legal, defence;
    double shaping_scale;
    int bot_kind; unsigned long long bot_seed;   // per-agent nibble: 0 caller's action, 1 random legal, 2/3 baselineTeam
    int bot_first_best;                 // baselineTeam tie-break: first best action instead of a random one
    const uint16_t *dist; const int16_t *node_of_cell; int dist_nmax;   // maze-distance tables (pm_distances)
    int auto_reset;                     // step: terminal envs restart inside the launch
    int64_t n_base;                     // global index of env 0 (keys the random draws)
    int64_t n_begin, n_end;             // env range of this launch
    int a0, a1;                         // sub-moves [a0, a1) of the step played by this launch (0, 4 = the whole step)
    // step inputs / outputs (see pm_step_out in pacman_b200.h)
    const void *actions; int act_size;
    float *obs; int64_t obs_agent_stride;
    double *reward; uint8_t *done; uint8_t *legal; int32_t *score_change; double *shaping; float *shaped;
    void *packed;                       // [N] pm_host_record
    uint8_t *act_used;                  // [N][4] the action each agent actually played
    // learner views (pm_step_out.view_*): canonicalised actor observations of one team + the merged critic observation
    uint8_t *view_actor; int64_t view_actor_stride; uint8_t *view_critic; int view_red;
    const uint8_t *mask;                // reset only
    // ABI 7
    int layout_pool; uint32_t *layout_id_rw;     // a new board per episode: hash(seed, env, episode) % L, kept in layout_id
    const uint32_t *bot_draws;          // [N][4] externally supplied random.choice indices for the on-device bots
    uint32_t *obs_bits; int64_t obs_bits_stride; int NB;   // packed observations: NB words per observation
    int epc;                            // pm_env_desc.envs_per_cta (0 = auto)
    int obs_elide;                      // pm_step_out.obs_static_valid: skip the observation elements that cannot have changed
    int dist_stage_words;               // > 0: the (single) board's distance table + node map are staged in shared memory
    // ABI 8 (pm_step_chain): launch `chain_k` of a run of chained step launches; chain_flags[cta] = number of launches
    // of the run that have finished this CTA's env tile (see pm_env.cu)
    uint32_t *chain_flags; uint32_t chain_k;
};

#ifndef PM_TMA_STORE
#define PM_TMA_STORE 0                  // 1: the encoder stages 4 KB of fp32 in shared memory and lets the TMA engine
#endif                                  //    (cp.async.bulk shared -> global) write it instead of issuing STG.128 itself
#ifndef PM_TMA_WARPS
#define PM_TMA_WARPS 4
#endif

// ---- shared-memory map (u32 word offsets) ----------------------------------
// Per-env arrays are [..][E1] with E1 odd, so both the per-env phase (consecutive
// threads = consecutive envs) and the per-observation phase (consecutive lanes =
// consecutive words of one env) are bank-conflict free.
//   food[k], caps[k] : board bitsets (c-indexed, NW words) after sub-move k —
//                      slot k is the working copy during sub-move k and the
//                      snapshot agent k is observed from afterwards
//   posA/posB/patch[k] : the four agent bit positions and the own-cell patch
//   rec               : the CTA's pm_host_record block (E x 48 bytes), 16-byte aligned
struct PmSmem {
    int lut, masks, lay, blk0, blk_words, lid, posA, posB, patch, rec, food, caps, stage, total;
    int E1;                             // padded env stride
    int dist;                           // bots, one board: dist[nmax][nmax] u16 then node_of[HW] i16 (0 = not staged)
};
// words needed to stage one board's maze-distance table + cell->node map (baselineTeam look-ahead)
PM_HD int pm_dist_stage_words(int HW, int nmax) { return ((nmax * nmax * 2 + 3) >> 2) + ((HW * 2 + 3) >> 2); }
#define PM_DIST_STAGE_MAX_WORDS 3072    // 12 KB: tinyCapture 7.3 KB, smallCapture 9.6 KB; bloxCapture (72 KB) stays in L2

PM_HD PmSmem pm_smem_map(int HW, int LS, int E, int has_caps, int L, int nslots = 4, int dist_words = 0, int views = 0) {
    const int NW = pm_nw(HW), NCAT = pm_ncat(HW);
    PmSmem m;
    m.E1 = E | 1;
    int o = 0;
    // static tables (one copy per CTA)
    m.lut = o; o += 64;                 // 16 x float4
    m.masks = o; o += PM_G_NMASKS * NCAT;
    m.lay = o; o += (L == 1) ? LS : 0;
    o = (o + 3) & ~3;
    m.dist = 0;
    if (dist_words > 0) { m.dist = o; o += dist_words; o = (o + 3) & ~3; }
    (void)views;
    m.blk0 = o;
    // per-block buffer
    m.lid = o; o += E;
    m.posA = o; o += 4 * m.E1;
    m.posB = o; o += 4 * m.E1;
    m.patch = o; o += 4 * m.E1;
    o = (o + 3) & ~3;
    m.rec = o; o += 12 * E;             // pm_host_record staging: 12 words per env, written out coalesced
    m.food = o; o += nslots * NW * m.E1;   // slots 0-3: per-agent snapshots; slot 4 (bots): look-ahead scratch
    m.caps = o; o += has_caps ? nslots * NW * m.E1 : 0;
    o = (o + 3) & ~3;
    m.blk_words = o - m.blk0;
    m.stage = 0;
#if PM_TMA_STORE
    o = (o + 31) & ~31;                 // 128-byte aligned: one 4 KB fp32 staging buffer per warp for the bulk stores
    m.stage = o; o += PM_TMA_WARPS * 1024;
#endif
    m.total = o;
    return m;
}

PM_HD int pm_bit(const uint32_t *bits, int c) { return (bits[c >> 5] >> (c & 31)) & 1; }
PM_HD int pm_sbit(const uint32_t *base, int stride, int c) { return (base[(c >> 5) * stride] >> (c & 31)) & 1; }

// ---- food dump: AgentRules.dumpFoodFromDeath (capture.py:569-668) ------------
// Walks the precomputed visit order (pm_pack_geom) from the death cell and drops
// `todo` pellets on the first cells that are on the board interior, not wall,
// not food, on the death cell's side, not a capsule and not under any agent.
// food/caps: this env's working bitsets in shared memory (word stride `st`).
// Returns pellets added to (red half | blue half << 16).
PM_NOINLINE uint32_t pm_dump(int W, int H, int ND, const uint32_t *dumptab, const uint32_t *walls,
                             uint32_t *food, const uint32_t *caps, int st, int vx, int vy, int todo,
                             uint32_t cells01, uint32_t cells23) {
    const int side = 2 * vx < W;
    uint32_t added = 0;
    for (int t = 0; t < ND && todo > 0; ++t) {
        const uint32_t o = dumptab[t];
        const int x = vx + (int)(int8_t)(o & 0xff), y = vy + (int)(int8_t)((o >> 8) & 0xff);
        if (x <= 0 || y <= 0 || x >= W || y >= H) continue;          // capture.py:609
        const int c = y * W + x;
        if (pm_bit(walls, c)) continue;
        if (pm_sbit(food, st, c)) continue;
        if ((2 * x < W) != side) continue;                             // onRightSide
        if (caps && pm_sbit(caps, st, c)) continue;
        const uint32_t cc = (uint32_t)c;
        if (cc == (cells01 & 0xffff) || cc == (cells01 >> 16) || cc == (cells23 & 0xffff) || cc == (cells23 >> 16)) continue;
        food[(c >> 5) * st] |= 1u << (c & 31);
        added += (x < (W >> 1)) ? 1u : 0x10000u;
        --todo;
    }
    return added;
}

// 32 bits of an NW-word bitset (word stride st) starting at bit s; s may be
// negative or run past the end (missing bits read as zero).
PM_HD uint32_t pm_extract32(const uint32_t *w, int st, int NW, int s) {
    const int k = s >> 5, r = s & 31;
    const uint32_t lo = ((unsigned)k < (unsigned)NW) ? w[k * st] : 0u;
    const uint32_t hi = ((unsigned)(k + 1) < (unsigned)NW) ? w[(k + 1) * st] : 0u;
    return pm_funnel_r(lo, hi, r);
}

// ---- mirrored bitsets for the canonical red-learner view --------------------------
PM_HD uint32_t pm_brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
#endif
}
// dst = src with every row reversed (cell (x, y) -> (W-1-x, y)); both c-indexed, word stride st; W <= 32.
PM_HD void pm_mirror_bits(const uint32_t *src, uint32_t *dst, int st, int W, int H, int NW) {
    for (int w = 0; w < NW; ++w) dst[w * st] = 0;
    const uint32_t rowmask = W >= 32 ? 0xffffffffu : ((1u << W) - 1u);
    for (int y = 0; y < H; ++y) {
        const int s = y * W, k = s >> 5, r = s & 31;
        const uint32_t lo = src[k * st], hi = (k + 1 < NW) ? src[(k + 1) * st] : 0u;
        const uint32_t row = pm_funnel_r(lo, hi, r) & rowmask;
        if (!row) continue;
        const uint32_t m = pm_brev32(row) >> (32 - W);
        dst[k * st] |= m << r;
        if (r + W > 32) dst[(k + 1) * st] |= m >> (32 - r);
    }
}

// Agents of one env in registers, in the state blob's own packing (pm_format.h):
//   a[i] = x | y<<8 | dir<<16 | isPacman<<19 | scaredTimer<<20        b[i] = numCarrying | numReturned<<16
// Every array index in this file is a compile-time constant (loops over agents unroll), so the eight words stay in
// registers.  The agent that MOVES is a run-time index - the four sub-moves of a step, and the look-ahead of the
// on-device bots, share one copy of the rules code instead of one per agent (the per-agent template instantiations
// this replaces were 7 100 SASS instructions = 113 KB for the rules phase, and the warp that runs it spent 40 % of
// its time waiting for instruction fetches) - and is read / written through pm_sel4 / pm_put4.
struct PmAg { uint32_t a[4], b[4]; };

PM_HD uint32_t pm_sel4(const uint32_t (&v)[4], int i) { return (i & 2) ? ((i & 1) ? v[3] : v[2]) : ((i & 1) ? v[1] : v[0]); }
PM_HD void pm_put4(uint32_t (&v)[4], int i, uint32_t x) {
    v[0] = i == 0 ? x : v[0]; v[1] = i == 1 ? x : v[1]; v[2] = i == 2 ? x : v[2]; v[3] = i == 3 ? x : v[3];
}
PM_HD int pm_ax(uint32_t a) { return (int)(a & 0xff); }
PM_HD int pm_ay(uint32_t a) { return (int)((a >> 8) & 0xff); }
PM_HD int pm_adir(uint32_t a) { return (int)((a >> 16) & 7); }
PM_HD int pm_apac(uint32_t a) { return (int)((a >> 19) & 1); }
PM_HD int pm_ascared(uint32_t a) { return (int)(a >> 20); }
PM_HD int pm_acell(uint32_t a, int W) { return pm_ay(a) * W + pm_ax(a); }
PM_HD int pm_bcarry(uint32_t b) { return (int)(b & 0xffff); }
// configuration = start, not a pacman, heading Stop, not scared (capture.py:691-693 etc.)
PM_HD uint32_t pm_home_word(const uint32_t *tail, int i) { return (tail[PM_L_START + i] & 0xffffu) | 4u << 16; }

PM_HD int pm_legal_mask(const uint32_t *walls, int W, int c) {   // game.py:335-350, bit = gym action int
    return (pm_bit(walls, c + W) ? 0 : 1) | (pm_bit(walls, c + 1) ? 0 : 2) | (pm_bit(walls, c - W) ? 0 : 4) |
           (pm_bit(walls, c - 1) ? 0 : 8) | (pm_bit(walls, c) ? 0 : 16);
}

// Uniform draw from the set bits of `legal` (randomTeam: random.choice(legal actions),
// agents/randomTeam.py:94-100).  Counter-based so a trajectory depends only on
// (seed, global env index, episode, step, agent) and not on launch geometry or sharding.
PM_HD uint32_t pm_random_u32(unsigned long long seed, unsigned long long n, unsigned episode, unsigned ctr) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (n + 1) +
                           0xD1B54A32D192ED03ull * (((unsigned long long)episode << 32) + ctr + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}
// The board an env plays in episode `episode` when the batch draws from a layout pool (pm_env_desc.layout_pool).
PM_HD uint32_t pm_pool_layout(unsigned long long seed, unsigned long long n, unsigned episode, int L) {
    return pm_random_u32(seed ^ 0x6c61796f75747321ull, n, episode, 0xfffffff0u) % (uint32_t)L;
}
// random.choice over the legal-action LIST, which the reference builds in the order N, S, E, W, Stop
// (game.py:295-299, 344-348).  `draw`: a 32-bit uniform word (scaled to the list length), or with `is_index`
// the list index itself (pm_env_desc.bot_draws; wraps past the end).
PM_HD int pm_random_legal(int legal, uint32_t draw, bool is_index) {
    const int order[5] = {0, 2, 1, 3, 4};
    int cnt = 0;
    for (int b = 0; b < 5; ++b) cnt += (legal >> b) & 1;
    int pick = is_index ? (int)(draw % (uint32_t)cnt) : (int)(((unsigned long long)draw * (unsigned long long)cnt) >> 32);
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if ((legal >> order[k]) & 1) { if (pick == 0) return order[k]; --pick; }
    return 4;
}

// heuristic shaping (pacman_mappo_resnet.py:251-264); L1 distance is flip-invariant.  A turn only records WHICH of
// the formula's cases applies (case | carrying << 2); the fp64 value is formed where it is consumed.
#define PM_SHAPE_NONE 0xffffffffu      // agent has not moved in this step yet: 0.0
PM_HD uint32_t pm_shaping_code(int px, int py, int pcarry, int nx, int ny) {
    int dx = px - nx, dy = py - ny;
    const int dist = (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy);
    return (uint32_t)(dist >= 2 ? 2 : dist) | (uint32_t)pcarry << 2;
}
PM_HD double pm_shaping_value(uint32_t code) {
    if (code == PM_SHAPE_NONE) return 0.0;
    const double living = -0.15;
    const int cls = (int)(code & 3);
    if (cls == 2) return pm_dadd(pm_dadd(living, -0.4), -pm_dmul(0.15, (double)(code >> 2)));
    if (cls == 0) return pm_dadd(living, -0.05);
    return living;
}

PM_HD void pm_store_snapshot(uint32_t *sm, const PmSmem &M, const PmParams &P, int k, int el, const PmAg &g) {
    const int HW = P.HW, W = P.W;
    // gym teams by index parity, gymPacMan.py:210: me = k, ally = k ^ 2, enemies k ^ 1 and k ^ 3
    const uint32_t me = pm_sel4(g.a, k), carry = (uint32_t)pm_bcarry(pm_sel4(g.b, k));
    const uint32_t own = (uint32_t)(HW + pm_acell(me, W));
    sm[M.posA + k * M.E1 + el] = own | (uint32_t)(4 * HW + pm_acell(pm_sel4(g.a, k ^ 2), W)) << 16;
    sm[M.posB + k * M.E1 + el] = (uint32_t)(5 * HW + pm_acell(pm_sel4(g.a, k ^ 1), W)) | (uint32_t)(5 * HW + pm_acell(pm_sel4(g.a, k ^ 3), W)) << 16;
    sm[M.patch + k * M.E1 + el] = carry > 0 ? (own | (1u + carry) << 16) : 0xffffffffu;
}

// ============================================================================
// phase A, MODE_STEP
// ============================================================================
// Per-env constants of a step, and the mutable board counters a sub-move touches.
struct PmCtx {
    int W, H, NW, E1, ND, HW, NCAT;
    const uint32_t *walls, *dumptab;
    const uint32_t *tail;              // the board's 8 tail words: start[4] | teams | ... (pm_format.h)
    const uint32_t *halfmask;          // the staged plane masks M6 (blue food half) | M7 (red food half), NCAT words each
    uint32_t teams;
};
struct PmBoard {
    int rf, bf;                        // pellets on the red (x < W//2) / blue food half
    bool food_dirty, caps_dirty;
};

// One pacman/ghost pair on the same cell (capture.py:682-701 and :708-727): an
// unscared ghost kills the pacman (dump, respawn), a scared ghost is sent home.
PM_HD uint32_t pm_collide(const PmCtx &X, PmAg &g, int pacman, int ghost, uint32_t *fk, const uint32_t *ck) {
    uint32_t added = 0;
    if (pm_ascared(pm_sel4(g.a, ghost)) <= 0) {
        const uint32_t ap = pm_sel4(g.a, pacman), bp = pm_sel4(g.b, pacman);
        if (pm_bcarry(bp) > 0) {
            const int W = X.W;
            const uint32_t c01 = (uint32_t)pm_acell(g.a[0], W) | (uint32_t)pm_acell(g.a[1], W) << 16;
            const uint32_t c23 = (uint32_t)pm_acell(g.a[2], W) | (uint32_t)pm_acell(g.a[3], W) << 16;
            added = pm_dump(W, X.H, X.ND, X.dumptab, X.walls, fk, ck, X.E1, pm_ax(ap), pm_ay(ap), pm_bcarry(bp), c01, c23);
            pm_put4(g.b, pacman, bp & 0xffff0000u);
        }
        pm_put4(g.a, pacman, pm_home_word(X.tail, pacman));
    } else {
        pm_put4(g.a, ghost, pm_home_word(X.tail, ghost));
    }
    return added;
}

// GameState.generateSuccessor for agent A playing the (already legalised) gym action u:
// AgentRules.applyAction + consume + checkDeath + decrementTimer (capture.py:107-123, 468-728)
// on agents `g` and the bitsets fk / ck (word stride X.E1).  Returns the score change.
PM_HD int pm_submove(const PmCtx &X, PmAg &g, uint32_t *fk, uint32_t *ck, PmBoard &b, const int A, int u) {
    const int W = X.W, E1 = X.E1;
    const uint32_t teams = X.teams;
    // ---- AgentRules.applyAction (capture.py:468-517) ----
    uint32_t aA = pm_sel4(g.a, A), bA = pm_sel4(g.b, A);
    const int dx = (u == 1) - (u == 3), dy = (u == 0) - (u == 2);
    const int x = pm_ax(aA) + dx, y = pm_ay(aA) + dy;
    const int c = y * W + x;
    const int dir = (u != 4) ? u : pm_adir(aA);        // Stop keeps the heading, game.py:115-117
    const int is_red = (teams >> A) & 1;
    const int pac = is_red ^ (2 * x < W ? 1 : 0);
    aA = pm_agA_pack(x, y, dir, pac, pm_ascared(aA));
    pm_put4(g.a, A, aA);
    const uint32_t posA = aA & 0xffffu;
    int pacflag = pac;
    int sc = 0;
    if (pm_bcarry(bA) > 0 && !pac) {
        const int carry = pm_bcarry(bA);
        sc = is_red ? carry : -carry;
        pm_put4(g.b, A, (bA & 0xffff0000u) + ((uint32_t)carry << 16));     // numReturned += numCarrying; numCarrying = 0
        pacflag = pm_apac(g.a[3]);                     // the rebinding quirk: capture.py:505 -> :514
    }
    if (pacflag) {                                     // AgentRules.consume (capture.py:519-560)
        if (pm_sbit(fk, E1, c)) {
            bool given = false;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (!given && (int)((teams >> i) & 1) == is_red && (g.a[i] & 0xffffu) == posA) { g.b[i] += 1; given = true; }
            fk[(c >> 5) * E1] &= ~(1u << (c & 31));
            b.food_dirty = true;
            if (x < (W >> 1)) --b.rf; else --b.bf;
        }
        if (ck && pm_sbit(ck, E1, c) && (is_red ? (2 * x > W) : (2 * x <= W))) {
            ck[(c >> 5) * E1] &= ~(1u << (c & 31));
            b.caps_dirty = true;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if ((int)((teams >> i) & 1) != is_red) g.a[i] = (g.a[i] & 0x000fffffu) | (uint32_t)PM_SCARED_TIME << 20;
        }
    }
    // ---- AgentRules.checkDeath (capture.py:670-728) ----
    // The branch (mover is pacman / ghost) is chosen once; enemies are then
    // visited in index order against live positions and timers.
#pragma unroll 1
    for (int e = 0; e < 4; ++e) {
        if (e == A) continue;
        if ((int)((teams >> e) & 1) == is_red) continue;
        const uint32_t ae = pm_sel4(g.a, e);
        if ((ae & 0xffffu) != (pm_sel4(g.a, A) & 0xffffu)) continue;
        uint32_t added = 0;
        if (pac) {
            if (!pm_apac(ae)) added = pm_collide(X, g, A, e, fk, ck);      // A = pacman, e = ghost
        } else {
            if (pm_apac(ae)) added = pm_collide(X, g, e, A, fk, ck);       // e = pacman, A = ghost
        }
        if (added) { b.rf += (int)(added & 0xffff); b.bf += (int)(added >> 16); b.food_dirty = true; }
    }
    aA = pm_sel4(g.a, A);                              // decrementTimer, capture.py:562-567
    if (pm_ascared(aA) > 0) pm_put4(g.a, A, aA - (1u << 20));
    return sc;
}

// ---- on-device baselineTeam (agents/baselineTeam.py:56-187) -------------------
// Maze distances come from the table pm_distances built for this env's board
// (Distancer.getDistance, distanceCalculator.py:38-62): dist[node(a) * nmax + node(b)].
struct PmDist {
    const uint16_t *dist; const int16_t *node_of; int nmax;
};
PM_HD int pm_maze_dist(const PmDist &D, int ca, int cb) {
    return (int)D.dist[(size_t)D.node_of[ca] * D.nmax + D.node_of[cb]];
}

// Q(s, a) of ReflexCaptureAgent.evaluate for the successor state (g2, fs): integer features x integer weights.
PM_HD int pm_baseline_value(const PmCtx &X, const PmDist &D, const PmAg &g2, const uint32_t *fs, const PmBoard &b2,
                            const int A, int kind, int u, int dir_before) {
    const int W = X.W, is_red = (X.teams >> A) & 1;
    const uint32_t aA = pm_sel4(g2.a, A);
    const int me = pm_acell(aA, W);
    if (kind == 2) {                                   // OffensiveReflexAgent.getFeatures / getWeights (:132-150)
        const int n = is_red ? b2.bf : b2.rf;          // len(self.getFood(successor).asList()): the half this team eats
        int v = -100 * n;
        if (n > 0) {
            int best = 0x7fffffff;
            const uint16_t *row = D.dist + (size_t)D.node_of[me] * D.nmax;     // distances from my cell
            // the pellets this team eats: red eats x >= W//2, blue eats x < W//2 - the half-board mask of plane 6 / 7
            // (bits of the observation string, i.e. shifted by 6*HW / 7*HW) says which cells those are
            const uint32_t *hm = X.halfmask + (is_red ? 0 : X.NCAT);
            const int hshift = (is_red ? 6 : 7) * X.HW;
            for (int w = 0; w < X.NW; ++w) {
                uint32_t bits = fs[w * X.E1] & pm_extract32(hm, 1, X.NCAT, 32 * w + hshift);
                // four independent (node, distance) look-ups in flight per round instead of one dependent pair
                while (bits) {
                    int c[4], d[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
#if defined(__CUDA_ARCH__)
                        const int bpos = bits ? __ffs((int)bits) - 1 : -1;
#else
                        const int bpos = bits ? __builtin_ctz(bits) : -1;
#endif
                        bits &= bits - 1;
                        c[q] = bpos < 0 ? -1 : 32 * w + bpos;
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) c[q] = c[q] < 0 ? -1 : (int)D.node_of[c[q]];
#pragma unroll
                    for (int q = 0; q < 4; ++q) d[q] = c[q] < 0 ? 0x7fffffff : (int)row[c[q]];
#pragma unroll
                    for (int q = 0; q < 4; ++q) best = d[q] < best ? d[q] : best;
                }
            }
            v -= best;
        }
        return v;
    }
    // DefensiveReflexAgent.getFeatures / getWeights (:159-187)
    int v = pm_apac(aA) ? 0 : 100;                     // onDefense
    int ninv = 0, best = 0;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        if ((int)((X.teams >> o) & 1) == is_red || !pm_apac(g2.a[o])) continue;
        const int d = pm_maze_dist(D, me, pm_acell(g2.a[o], W));
        if (ninv == 0 || d < best) best = d;
        ++ninv;
    }
    v -= 1000 * ninv;
    if (ninv > 0) v -= 10 * best;
    if (u == 4) v -= 100;                              // stop
    const int rev = dir_before == 4 ? 4 : (dir_before ^ 2);   // Directions.REVERSE on gym ints: N0<->S2, E1<->W3
    if (u == rev) v -= 2;
    return v;
}

// ReflexCaptureAgent.chooseAction (:65-92): evaluate every legal action on a scratch copy of the
// state, keep the best set, break ties with `rnd` (random.choice) or take the first (test hook).
// fs / cs: scratch bitsets (slot 4).  Returns the gym action.
PM_HD int pm_baseline_choose(const PmCtx &X, const PmDist &D, const PmAg &g, const uint32_t *fk, const uint32_t *ck,
                             uint32_t *fs, uint32_t *cs, const PmBoard &b, const int A, int legal, int kind, uint32_t rnd,
                             bool first_best, bool rnd_is_index) {
    // Actions.getPossibleActions: N, S, E, W, Stop (game.py:295-299) = gym ints 0, 2, 1, 3, 4, one nibble each
    const uint32_t order = 0x43120u;
    const int is_red = (X.teams >> A) & 1;
    const bool go_home = (is_red ? b.bf : b.rf) <= 0;  // foodLeft <= 0: walk towards self.start (:79-90)
    const uint32_t st = X.tail[PM_L_START + A];
    const int start_c = (int)((st >> 8) & 0xff) * X.W + (int)(st & 0xff);
    const int dir_before = pm_adir(pm_sel4(g.a, A));
    int bestv = -0x7fffffff, home_a = 4, home_d = 9999;
    // the values of the (up to five) candidates: 5 x 12 bits would not hold them, so two passes over the best set use
    // a bitmask of the candidates that reached the best value so far instead of an array (no local memory)
    uint32_t best_set = 0;
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
        const int u = (int)((order >> (4 * k)) & 15u);
        if (!((legal >> u) & 1)) continue;
        for (int w = 0; w < X.NW; ++w) fs[w * X.E1] = fk[w * X.E1];
        if (ck) for (int w = 0; w < X.NW; ++w) cs[w * X.E1] = ck[w * X.E1];
        PmAg g2 = g;
        PmBoard b2 = b;
        pm_submove(X, g2, fs, ck ? cs : nullptr, b2, A, u);
        if (go_home) {
            const int d = pm_maze_dist(D, start_c, pm_acell(pm_sel4(g2.a, A), X.W));
            if (d < home_d) { home_d = d; home_a = u; }
        } else {
            const int v = pm_baseline_value(X, D, g2, fs, b2, A, kind, u, dir_before);
            if (v > bestv) { bestv = v; best_set = 1u << k; }
            else if (v == bestv) best_set |= 1u << k;
        }
    }
    if (go_home) return home_a;
    int nb = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) nb += (int)((best_set >> k) & 1u);
    int pick = first_best ? 0 : rnd_is_index ? (int)(rnd % (unsigned)nb) : (int)(((unsigned long long)rnd * (unsigned)nb) >> 32);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const int u = (int)((order >> (4 * k)) & 15u);
        if ((best_set >> k) & 1u) { if (pick == 0) return u; --pick; }
    }
    return 4;
}

// Everything a step accumulates across the four turns.
struct PmStepAcc {
    PmBoard b;
    int score, sc_total, red_sc, blue_sc;
    double red_r, blue_r;
    uint32_t shcode[4];                // pm_shaping_code of each agent's turn
    uint32_t snap[4];
    uint32_t act_used;                 // the action each agent actually played, one byte per agent
};

// One turn of gymPacMan.step's loop for agent A (gymPacMan.py:149-167): pick the action (caller's,
// random legal, or baselineTeam), get_reward bookkeeping, the sub-move, the observation snapshot.
template <int BOTS, int VIEWS>
PM_HD void pm_agent_turn(const PmParams &P, const PmCtx &X, uint32_t *sm, const PmSmem &M, int el, int64_t n, uint32_t lid,
                         PmAg &g, uint32_t *food, uint32_t *caps, PmStepAcc &S, const int A, uint32_t act4, int steps,
                         uint32_t episode) {
    const int NW = X.NW, E1 = X.E1, W = X.W;
    uint32_t *fk = food + (size_t)A * NW * E1;
    uint32_t *ck = caps ? caps + (size_t)A * NW * E1 : nullptr;
    if (A > 0) {                                       // working copy for this sub-move; slot A-1 stays as snapshot
        for (int w = 0; w < NW; ++w) fk[w * E1] = fk[(w - NW) * E1];
        if (ck) for (int w = 0; w < NW; ++w) ck[w * E1] = ck[(w - NW) * E1];
    }
    // ---- get_reward pre-state (gymPacMan.py:231-259) ----
    const bool gym_red = (A & 1) == 0;
    const int o1 = gym_red ? 1 : 0, o2 = o1 + 2;
    const int mine_before = gym_red ? S.b.rf : S.b.bf, theirs_before = gym_red ? S.b.bf : S.b.rf;
    const int pb1 = pm_apac(gym_red ? g.a[1] : g.a[0]), pb2 = pm_apac(gym_red ? g.a[3] : g.a[2]);

    const int legal = pm_legal_mask(X.walls, W, pm_acell(pm_sel4(g.a, A), W));
    int u = (act4 >> (8 * A)) & 0xff;
    const int kind = (P.bot_kind >> (4 * A)) & 15;
    const bool ext_draw = P.bot_draws != nullptr;
    if (kind == 1) {
        const uint32_t rnd = ext_draw ? P.bot_draws[4 * n + A]
                                      : pm_random_u32(P.bot_seed, (unsigned long long)(P.n_base + n), episode, (unsigned)steps * 4u + A);
        u = pm_random_legal(legal, rnd, ext_draw);
    } else if (BOTS && kind >= 2) {
        PmDist D;
        D.nmax = P.dist_nmax;
        if (M.dist) {                                  // one board: table + node map staged in shared memory
            D.dist = reinterpret_cast<const uint16_t *>(sm + M.dist);
            D.node_of = reinterpret_cast<const int16_t *>(sm + M.dist + ((P.dist_nmax * P.dist_nmax * 2 + 3) >> 2));
        } else {
            D.dist = P.dist + (size_t)lid * P.dist_nmax * P.dist_nmax;
            D.node_of = P.node_of_cell + (size_t)lid * X.HW;
        }
        const uint32_t rnd = ext_draw ? P.bot_draws[4 * n + A]
                                      : pm_random_u32(P.bot_seed, (unsigned long long)(P.n_base + n), episode, (unsigned)steps * 4u + A);
        u = pm_baseline_choose(X, D, g, fk, ck, food + (size_t)4 * NW * E1, caps ? caps + (size_t)4 * NW * E1 : nullptr,
                               S.b, A, legal, kind, rnd, P.bot_first_best != 0, ext_draw);
    }
    const bool req_legal = u <= 4 && ((legal >> u) & 1);
    if (!req_legal) u = 4;                             // illegal -> Stop, capture.py:472-475
    S.act_used |= (uint32_t)u << (8 * A);

    const int sc = pm_submove(X, g, fk, ck, S.b, A, u);
    S.score += sc;
    S.sc_total += sc;
    if (gym_red) S.red_sc += sc; else S.blue_sc -= sc;

    // ---- get_reward post-state ----
    double r = gym_red ? S.red_r : S.blue_r;
    const int mine_after = gym_red ? S.b.rf : S.b.bf, theirs_after = gym_red ? S.b.bf : S.b.rf;
    if (mine_before < mine_after) r = pm_dadd(r, 1.0);
    if (theirs_before > theirs_after) r = pm_dadd(r, 0.1);
    if (P.defence) {                                   // an enemy pacman is back on its start cell as a ghost
        const uint32_t a1 = gym_red ? g.a[1] : g.a[0], a2 = gym_red ? g.a[3] : g.a[2];
        if (pb1 && !pm_apac(a1) && (a1 & 0xffffu) == (X.tail[PM_L_START + o1] & 0xffffu)) r = pm_dadd(r, 0.25);
        if (pb2 && !pm_apac(a2) && (a2 & 0xffffu) == (X.tail[PM_L_START + o2] & 0xffffu)) r = pm_dadd(r, 0.25);
    }
    if (P.reward_legal && req_legal) r = pm_dadd(r, 0.01);
    if (gym_red) S.red_r = r; else S.blue_r = r;

    // ---- observation snapshot for agent A (gymPacMan.py:166) + shaping ----
    pm_store_snapshot(sm, M, P, A, el, g);
    if (VIEWS && P.view_red && (A & 1) == 0) {         // red learner: keep a mirrored copy of this snapshot's bitsets
        const int ms = 4 + BOTS + (A >> 1);
        pm_mirror_bits(fk, food + (size_t)ms * NW * E1, E1, W, X.H, NW);
        if (ck) pm_mirror_bits(ck, caps + (size_t)ms * NW * E1, E1, W, X.H, NW);
    }
    const uint32_t prev = pm_sel4(S.snap, A), aA = pm_sel4(g.a, A);
    pm_put4(S.shcode, A, pm_shaping_code((int)(prev & 0xff), (int)((prev >> 8) & 0xff), (int)(prev >> 16), pm_ax(aA), pm_ay(aA)));
    pm_put4(S.snap, A, (aA & 0xffffu) | (uint32_t)pm_bcarry(pm_sel4(g.b, A)) << 16);
}

// What a rules thread needs from global memory before it can start: issued by pm_phaseA_fetch at the very top of the
// kernel, BEFORE the CTA stages its tables, so that the (HBM-latency) state loads are in flight under the staging
// instead of stalling the first instructions of the rules phase (ncu, round 2: half of that phase's long-scoreboard
// samples sat on the first use of these loads).  The board words go straight to shared memory (cp.async on the device).
struct PmFetch {
    pm_u4 A, B, Sn, Mi;
    uint32_t lid, act4;
};
PM_HD void pm_phaseA_fetch(const PmParams &P, uint32_t *sm, const PmSmem &M, int el, int64_t n0, PmFetch &F) {
    const int64_t n = n0 + el;
    F.lid = 0xffffffffu;
    if (n >= P.n_end) return;
    F.lid = P.layout_id ? P.layout_id[n] : (P.L == 1 ? 0u : (uint32_t)(n % P.L));
    F.A = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agA(P.N))[n];
    F.B = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agB(P.N))[n];
    F.Sn = reinterpret_cast<const pm_u4 *>(P.state + pm_st_snap(P.N))[n];
    F.Mi = reinterpret_cast<const pm_u4 *>(P.state + pm_st_misc(P.N))[n];
    if (P.act_size == 1) F.act4 = reinterpret_cast<const uint32_t *>(P.actions)[n];
    else {
        const long long *a8 = reinterpret_cast<const long long *>(P.actions) + 4 * n;
        F.act4 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { long long v = a8[i]; F.act4 |= (uint32_t)((v < 0 || v > 4) ? 255 : (int)v) << (8 * i); }
    }
    const int NW = P.NW, E1 = M.E1;
    const int64_t NP = pm_np(P.N);
    const uint32_t *gfood = reinterpret_cast<const uint32_t *>(P.state + pm_st_food(P.N));
    const uint32_t *gcaps = reinterpret_cast<const uint32_t *>(P.state + pm_st_caps(P.N, P.HW));
    uint32_t *food = sm + M.food + el;                 // slot k, word w at food[(k*NW + w)*E1]
    uint32_t *caps = P.has_caps ? sm + M.caps + el : nullptr;
    // the board goes into the slot the first sub-move of this launch copies from (slot 0 for a whole step)
    const int s0 = P.a0 > 0 ? P.a0 - 1 : 0;
#if defined(__CUDA_ARCH__)
    for (int w = 0; w < NW; ++w)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(food + (s0 * NW + w) * E1)),
                     "l"(gfood + (size_t)w * NP + n) : "memory");
    if (caps)
        for (int w = 0; w < NW; ++w)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(caps + (s0 * NW + w) * E1)),
                         "l"(gcaps + (size_t)w * NP + n) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
#else
    for (int w = 0; w < NW; ++w) food[(s0 * NW + w) * E1] = gfood[(size_t)w * NP + n];
    if (caps) for (int w = 0; w < NW; ++w) caps[(s0 * NW + w) * E1] = gcaps[(size_t)w * NP + n];
#endif
}

template <int BOTS, int VIEWS>
PM_HD void pm_phaseA_step(const PmParams &P, const uint32_t *st, uint32_t *sm, const PmSmem &M, int el, int64_t n0, const PmFetch &F) {
    const int64_t n = n0 + el;
    if (n >= P.n_end) { sm[M.lid + el] = 0xffffffffu; return; }
    uint32_t lid = F.lid;
    sm[M.lid + el] = lid;
    const uint32_t *lay = (P.L == 1) ? st + M.lay : P.layouts + (size_t)lid * P.LS;
    const uint32_t *walls = lay + pm_l_walls(P.HW);
    const uint32_t *tail = lay + pm_l_tail(P.HW);
    const int W = P.W, NW = P.NW, E1 = M.E1;
    const int64_t NP = pm_np(P.N);
    PmCtx X;
    X.W = W; X.H = P.H; X.NW = NW; X.E1 = E1; X.ND = P.ND; X.HW = P.HW; X.NCAT = P.NCAT;
    X.walls = walls; X.dumptab = P.geom + pm_g_dump_off(P.HW); X.tail = tail; X.teams = tail[PM_L_TEAMS];
    X.halfmask = st + M.masks + 2 * P.NCAT;

    // ---- the state pm_phaseA_fetch loaded (128-bit loads, coalesced across the env dimension) ----
    pm_u4 A = F.A, B = F.B, Sn = F.Sn, Mi = F.Mi;
    uint32_t *gfood = reinterpret_cast<uint32_t *>(P.state + pm_st_food(P.N));
    uint32_t *gcaps = reinterpret_cast<uint32_t *>(P.state + pm_st_caps(P.N, P.HW));
    uint32_t *food = sm + M.food + el;                 // slot k, word w at food[(k*NW + w)*E1]
    uint32_t *caps = P.has_caps ? sm + M.caps + el : nullptr;
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_all;" ::: "memory");   // this thread's own board words have landed in shared memory
#endif

    PmAg g;
    g.a[0] = A.x; g.a[1] = A.y; g.a[2] = A.z; g.a[3] = A.w;
    g.b[0] = B.x; g.b[1] = B.y; g.b[2] = B.z; g.b[3] = B.w;
    int steps = (int)Mi.y;
    uint32_t episode = Mi.w;                                      // restarts of this env (keys the random draws)
    PmStepAcc S;
    S.score = (int)Mi.x;
    S.b.rf = (int)(Mi.z & 0xffff); S.b.bf = (int)(Mi.z >> 16);
    S.b.food_dirty = false; S.b.caps_dirty = false;
    S.sc_total = 0; S.red_sc = 0; S.blue_sc = 0; S.red_r = 0.0; S.blue_r = 0.0; S.act_used = 0;
    S.snap[0] = Sn.x; S.snap[1] = Sn.y; S.snap[2] = Sn.z; S.snap[3] = Sn.w;
    pm_u4 *gacc = reinterpret_cast<pm_u4 *>(P.state + pm_st_acc(P.N, P.HW));
#pragma unroll
    for (int i = 0; i < 4; ++i) S.shcode[i] = PM_SHAPE_NONE;
    if (P.a0 > 0) {                                    // continue a step that earlier launches started
        const pm_u4 c0 = gacc[n], c1 = gacc[NP + n], c2 = gacc[2 * NP + n];
        S.red_r = pm_dfrom(c0.x, c0.y); S.blue_r = pm_dfrom(c0.z, c0.w);
        S.sc_total = (int)c1.x; S.red_sc = (int)c1.y; S.blue_sc = (int)c1.z; S.act_used = c1.w;
        S.shcode[0] = c2.x; S.shcode[1] = c2.y; S.shcode[2] = c2.z; S.shcode[3] = c2.w;
    }

    const uint32_t act4 = F.act4;

    // the (up to) four ordered sub-moves of the step: ONE copy of the turn's code, the mover is a run-time index
#pragma unroll 1
    for (int k = P.a0; k < P.a1; ++k) pm_agent_turn<BOTS, VIEWS>(P, X, sm, M, el, n, lid, g, food, caps, S, k, act4, steps, episode);

    if (P.a1 < 4) {
        // ---- the step continues in a later launch (a CPU bot decides the next sub-move from the state as it is
        // now, gymPacMan.py:157): park the accumulators and the state; no epilogue, no step outputs ----
        pm_u4 c0, c1, c2;
        uint64_t u;
        u = pm_dbits(S.red_r); c0.x = (uint32_t)u; c0.y = (uint32_t)(u >> 32);
        u = pm_dbits(S.blue_r); c0.z = (uint32_t)u; c0.w = (uint32_t)(u >> 32);
        c1.x = (uint32_t)S.sc_total; c1.y = (uint32_t)S.red_sc; c1.z = (uint32_t)S.blue_sc; c1.w = S.act_used;
        c2.x = S.shcode[0]; c2.y = S.shcode[1]; c2.z = S.shcode[2]; c2.w = S.shcode[3];
        gacc[n] = c0; gacc[NP + n] = c1; gacc[2 * NP + n] = c2;
        A.x = g.a[0]; A.y = g.a[1]; A.z = g.a[2]; A.w = g.a[3];
        B.x = g.b[0]; B.y = g.b[1]; B.z = g.b[2]; B.w = g.b[3];
        reinterpret_cast<pm_u4 *>(P.state + pm_st_agA(P.N))[n] = A;
        reinterpret_cast<pm_u4 *>(P.state + pm_st_agB(P.N))[n] = B;
        Sn.x = S.snap[0]; Sn.y = S.snap[1]; Sn.z = S.snap[2]; Sn.w = S.snap[3];
        reinterpret_cast<pm_u4 *>(P.state + pm_st_snap(P.N))[n] = Sn;
        Mi.x = (uint32_t)S.score; Mi.z = (uint32_t)S.b.rf | (uint32_t)S.b.bf << 16;
        reinterpret_cast<pm_u4 *>(P.state + pm_st_misc(P.N))[n] = Mi;
        const int sl = P.a1 - 1;
        if (S.b.food_dirty) for (int w = 0; w < NW; ++w) gfood[(size_t)w * NP + n] = food[(sl * NW + w) * E1];
        if (S.b.caps_dirty) for (int w = 0; w < NW; ++w) gcaps[(size_t)w * NP + n] = caps[(sl * NW + w) * E1];
        if (P.act_used) reinterpret_cast<uint32_t *>(P.act_used)[n] = S.act_used;
        return;
    }

    int score = S.score, rf = S.b.rf, bf = S.b.bf;
    const int sc_total = S.sc_total;
    double red_r = S.red_r, blue_r = S.blue_r;
    bool food_dirty = S.b.food_dirty, caps_dirty = S.b.caps_dirty;
    uint32_t *snap = S.snap;

    // ---- step epilogue (gymPacMan.py:171-193) ----
    blue_r = pm_dadd(blue_r, (double)(S.blue_sc > 0 ? S.blue_sc : 0));
    red_r = pm_dadd(red_r, (double)(S.red_sc > 0 ? S.red_sc : 0));
    bool term = false;
    if (bf == 0 && pm_bcarry(g.b[0]) == 0 && pm_bcarry(g.b[2]) == 0) term = true;      // check_termination, :261-270
    if (rf == 0 && pm_bcarry(g.b[1]) == 0 && pm_bcarry(g.b[3]) == 0) term = true;
    if (steps >= P.length) term = true;
    if (term) {
        const double bonus = pm_dadd(20.0, pm_dmul(1.0 / 5, (double)(score < 0 ? -score : score)));
        if (score < 0) blue_r = pm_dadd(blue_r, bonus);
        else if (score > 0) red_r = pm_dadd(red_r, bonus);
    }
    steps += 1;

    // ---- fused restart (pm_env_desc.auto_reset): the trainer's `if done: env.reset()`
    // (pacman_mappo_resnet.py:524-538) without a second launch.  Rewards / done / shaping above
    // belong to the terminal step; state, the four observations and the legal masks below are
    // CaptureRules.newGame's (capture.py:369-382, game.py:490-508). ----
    if (term && P.auto_reset) {
        episode += 1;
        if (P.layout_pool) {                            // a new maze per episode (gymPacMan.py:98-100)
            lid = pm_pool_layout(P.bot_seed, (unsigned long long)(P.n_base + n), episode, P.L);
            P.layout_id_rw[n] = lid;
            sm[M.lid + el] = lid | 0x80000000u;         // new board: its observations are written in full (obs_static_valid)
            lay = P.layouts + (size_t)lid * P.LS;
            walls = lay + pm_l_walls(P.HW);
            tail = lay + pm_l_tail(P.HW);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t s0 = tail[PM_L_START + i];
            g.a[i] = pm_agA_pack((int)(s0 & 0xff), (int)((s0 >> 8) & 0xff), 4, (int)((s0 >> 16) & 1), 0);
            g.b[i] = 0;
            snap[i] = s0 & 0xffff;
        }
        score = 0; steps = 0;
        rf = (int)(tail[PM_L_FOODCNT] & 0xffff); bf = (int)(tail[PM_L_FOODCNT] >> 16);
        for (int w = 0; w < NW; ++w) {
            const uint32_t f = lay[pm_l_food(P.HW) + w];
#pragma unroll
            for (int k = 0; k < 4; ++k) food[(k * NW + w) * E1] = f;
            if (caps) {
                const uint32_t cp = lay[pm_l_caps(P.HW) + w];
#pragma unroll
                for (int k = 0; k < 4; ++k) caps[(k * NW + w) * E1] = cp;
            }
        }
        food_dirty = true; caps_dirty = caps != nullptr;
#pragma unroll
        for (int k = 0; k < 4; ++k) pm_store_snapshot(sm, M, P, k, el, g);
        if (VIEWS && P.view_red) {
            for (int l = 0; l < 2; ++l) {
                pm_mirror_bits(food, food + (size_t)(4 + BOTS + l) * NW * E1, E1, W, P.H, NW);
                if (caps) pm_mirror_bits(caps, caps + (size_t)(4 + BOTS + l) * NW * E1, E1, W, P.H, NW);
            }
        }
    }

    // ---- outputs ----
    if (P.reward) { P.reward[n] = red_r; P.reward[P.N + n] = blue_r; }
    if (P.done) P.done[n] = term ? 1 : 0;
    if (P.score_change) P.score_change[n] = sc_total;
    uint32_t lm = 0;
    if (P.legal || P.packed) {
#pragma unroll
        for (int i = 0; i < 4; ++i) lm |= (uint32_t)pm_legal_mask(walls, W, pm_acell(g.a[i], W)) << (8 * i);
        if (P.legal) reinterpret_cast<uint32_t *>(P.legal)[n] = lm;
    }
    float shaped[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double team = (i & 1) ? blue_r : red_r;          // sum over both learners = 2 * team reward
        const double shap = pm_shaping_value(S.shcode[i]);
        shaped[i] = (float)pm_dadd(pm_dadd(team, team), pm_dmul(P.shaping_scale, shap));
        if (P.shaping) P.shaping[(size_t)i * P.N + n] = shap;
        if (P.shaped) P.shaped[(size_t)i * P.N + n] = shaped[i];
    }
    if (P.act_used) reinterpret_cast<uint32_t *>(P.act_used)[n] = S.act_used;
    if (P.packed) {                                             // struct pm_host_record (48 B) -> smem staging
        pm_u4 r0, r1, r2;
        const uint64_t rr = pm_dbits(red_r), rb = pm_dbits(blue_r);
        r0.x = (uint32_t)rr; r0.y = (uint32_t)(rr >> 32); r0.z = (uint32_t)rb; r0.w = (uint32_t)(rb >> 32);
        r1.x = pm_fbits(shaped[0]); r1.y = pm_fbits(shaped[1]); r1.z = pm_fbits(shaped[2]); r1.w = pm_fbits(shaped[3]);
        r2.x = (uint32_t)sc_total; r2.y = lm; r2.z = term ? 1u : 0u; r2.w = 0u;
        pm_u4 *rec = reinterpret_cast<pm_u4 *>(sm + M.rec) + 3 * el;
        rec[0] = r0; rec[1] = r1; rec[2] = r2;
    }
    // ---- state write-back ----
    A.x = g.a[0]; A.y = g.a[1]; A.z = g.a[2]; A.w = g.a[3];
    B.x = g.b[0]; B.y = g.b[1]; B.z = g.b[2]; B.w = g.b[3];
    reinterpret_cast<pm_u4 *>(P.state + pm_st_agA(P.N))[n] = A;
    reinterpret_cast<pm_u4 *>(P.state + pm_st_agB(P.N))[n] = B;
    Sn.x = snap[0]; Sn.y = snap[1]; Sn.z = snap[2]; Sn.w = snap[3];
    reinterpret_cast<pm_u4 *>(P.state + pm_st_snap(P.N))[n] = Sn;
    Mi.x = (uint32_t)score; Mi.y = (uint32_t)steps; Mi.z = (uint32_t)rf | (uint32_t)bf << 16; Mi.w = episode;
    reinterpret_cast<pm_u4 *>(P.state + pm_st_misc(P.N))[n] = Mi;
    if (food_dirty) for (int w = 0; w < NW; ++w) gfood[(size_t)w * NP + n] = food[(3 * NW + w) * E1];
    if (caps_dirty) for (int w = 0; w < NW; ++w) gcaps[(size_t)w * NP + n] = caps[(3 * NW + w) * E1];
}

// ============================================================================
// phase A, MODE_RESET / MODE_OBSERVE: one state, four observations of it
// ============================================================================
template <int MODE, int VIEWS>
PM_HD void pm_phaseA_static(const PmParams &P, const uint32_t *st, uint32_t *sm, const PmSmem &M, int el, int64_t n0) {
    const int64_t n = n0 + el;
    if (n >= P.n_end || (MODE == PM_MODE_RESET && P.mask && !P.mask[n])) { sm[M.lid + el] = 0xffffffffu; return; }
    uint32_t lid = P.layout_id ? P.layout_id[n] : (P.L == 1 ? 0u : (uint32_t)(n % P.L));
    // episode counter: a full reset (mask == NULL) restarts it, a masked reset of an initialised batch advances it
    uint32_t episode = 0;
    if (MODE == PM_MODE_RESET) {
        episode = P.mask ? reinterpret_cast<const pm_u4 *>(P.state + pm_st_misc(P.N))[n].w + 1u : 0u;
        if (P.layout_pool) {                            // a new maze per episode (gymPacMan.py:98-100)
            lid = pm_pool_layout(P.bot_seed, (unsigned long long)(P.n_base + n), episode, P.L);
            P.layout_id_rw[n] = lid;
        }
    }
    sm[M.lid + el] = lid;
    const uint32_t *lay = (P.L == 1) ? st + M.lay : P.layouts + (size_t)lid * P.LS;
    const uint32_t *tail = lay + pm_l_tail(P.HW);
    const int NW = P.NW, E1 = M.E1, W = P.W;
    const int64_t NP = pm_np(P.N);
    uint32_t *gfood = reinterpret_cast<uint32_t *>(P.state + pm_st_food(P.N));
    uint32_t *gcaps = reinterpret_cast<uint32_t *>(P.state + pm_st_caps(P.N, P.HW));
    uint32_t *food = sm + M.food + el;
    uint32_t *caps = P.has_caps ? sm + M.caps + el : nullptr;
    PmAg g;
    if (MODE == PM_MODE_RESET) {
        // CaptureRules.newGame -> GameState.initialize (capture.py:369-382,311-323; game.py:490-508)
        pm_u4 A, B, S, Mi;
        uint32_t a[4], s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t st = tail[PM_L_START + i];
            a[i] = pm_agA_pack(st & 0xff, (st >> 8) & 0xff, 4, (st >> 16) & 1, 0);
            s[i] = st & 0xffff;
        }
        A.x = a[0]; A.y = a[1]; A.z = a[2]; A.w = a[3];
        B.x = B.y = B.z = B.w = 0;
        S.x = s[0]; S.y = s[1]; S.z = s[2]; S.w = s[3];
        Mi.x = 0; Mi.y = 0; Mi.z = tail[PM_L_FOODCNT];
        Mi.w = episode;
        reinterpret_cast<pm_u4 *>(P.state + pm_st_agA(P.N))[n] = A;
        reinterpret_cast<pm_u4 *>(P.state + pm_st_agB(P.N))[n] = B;
        reinterpret_cast<pm_u4 *>(P.state + pm_st_snap(P.N))[n] = S;
        reinterpret_cast<pm_u4 *>(P.state + pm_st_misc(P.N))[n] = Mi;
        for (int w = 0; w < NW; ++w) {
            const uint32_t f = lay[pm_l_food(P.HW) + w], cp = lay[pm_l_caps(P.HW) + w];
            food[w * E1] = f; gfood[(size_t)w * NP + n] = f;
            gcaps[(size_t)w * NP + n] = cp;
            if (caps) caps[w * E1] = cp;
        }
        g.a[0] = A.x; g.a[1] = A.y; g.a[2] = A.z; g.a[3] = A.w;
        g.b[0] = B.x; g.b[1] = B.y; g.b[2] = B.z; g.b[3] = B.w;
    } else {
        pm_u4 A = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agA(P.N))[n];
        pm_u4 B = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agB(P.N))[n];
        for (int w = 0; w < NW; ++w) food[w * E1] = gfood[(size_t)w * NP + n];
        if (caps) for (int w = 0; w < NW; ++w) caps[w * E1] = gcaps[(size_t)w * NP + n];
        g.a[0] = A.x; g.a[1] = A.y; g.a[2] = A.z; g.a[3] = A.w;
        g.b[0] = B.x; g.b[1] = B.y; g.b[2] = B.z; g.b[3] = B.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) pm_store_snapshot(sm, M, P, k, el, g);
    if (VIEWS && P.view_red) {
        for (int l = 0; l < 2; ++l) {
            pm_mirror_bits(food, food + (size_t)(4 + l) * NW * E1, E1, W, P.H, NW);
            if (caps) pm_mirror_bits(caps, caps + (size_t)(4 + l) * NW * E1, E1, W, P.H, NW);
        }
    }
    if (P.legal) {
        const uint32_t *walls = lay + pm_l_walls(P.HW);
        uint32_t lm = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) lm |= (uint32_t)pm_legal_mask(walls, W, pm_acell(g.a[i], W)) << (8 * i);
        reinterpret_cast<uint32_t *>(P.legal)[n] = lm;
    }
}

// Copy the CTA's record block from shared memory to P.packed with 128-bit stores, consecutive
// threads = consecutive 16-byte pieces (a warp writes 512 contiguous bytes).  P.packed may be
// device memory or MAPPED PINNED HOST memory (pm_step_host_mapped): coalescing is what turns the
// latter into full-size PCIe writes.
PM_HD void pm_records_out(const PmParams &P, const uint32_t *sm, const PmSmem &M, int tid, int64_t n0) {
    const int64_t left = P.n_end - n0;
    const int valid = (int)(left < P.E ? left : P.E);
    if (tid >= 3 * valid) return;
    reinterpret_cast<pm_u4 *>(P.packed)[3 * (size_t)n0 + tid] = reinterpret_cast<const pm_u4 *>(sm + M.rec)[tid];
}

// ============================================================================
// phase B: snapshot -> observation planes, one warp per observation ("unit")
// ============================================================================
// Per-observation constants, loaded once per unit.
struct PmUnit {
    const uint32_t *food, *caps;                // snapshot bitsets (caps == nullptr: batch has no capsules)
    uint32_t lid;
    uint32_t p_own, p_ally, p_e1, p_e2;         // bit positions of the four agents in the bit string (0xffffffff: none)
    uint32_t patch;                             // own-cell element | (1+carrying) << 16, or 0xffffffff
    uint32_t p_own2, patch2;                    // critic view: the second learner's own cell (plane 1) and its patch
    int mirror;                                 // canonical red-learner view: mirrored tables and bitsets
    int fresh;                                  // this env moved to another board inside the launch
};

template <int MODE>
PM_HD bool pm_unit_load(const PmParams &P, const uint32_t *sm, const PmSmem &M, int k, int el, PmUnit &U) {
    U.lid = sm[M.lid + el];
    if (U.lid == 0xffffffffu) return false;
    U.fresh = (int)(U.lid >> 31);
    U.lid &= 0x7fffffffu;
    if (MODE == PM_MODE_STEP && (k < P.a0 || k >= P.a1)) return false;     // only the agents that moved in this launch
    const int slot = (MODE == PM_MODE_STEP) ? k : 0;
    U.food = sm + M.food + (size_t)slot * P.NW * M.E1 + el;
    U.caps = P.has_caps ? sm + M.caps + (size_t)slot * P.NW * M.E1 + el : nullptr;
    const uint32_t pa = sm[M.posA + k * M.E1 + el], pb = sm[M.posB + k * M.E1 + el];
    U.p_own = pa & 0xffff; U.p_ally = pa >> 16; U.p_e1 = pb & 0xffff; U.p_e2 = pb >> 16;
    U.patch = sm[M.patch + k * M.E1 + el];
    U.p_own2 = 0xffffffffu; U.patch2 = 0xffffffffu; U.mirror = 0;
    return true;
}

// x -> W-1-x inside one plane of the bit string
PM_HD uint32_t pm_mirror_pos(uint32_t pos, int plane, int HW, int W) {
    if (pos == 0xffffffffu) return pos;
    const int c = (int)pos - plane * HW, x = c % W;
    return (uint32_t)(plane * HW + c + (W - 1 - 2 * x));
}
PM_HD uint32_t pm_mirror_patch(uint32_t patch, int HW, int W) {
    if (patch == 0xffffffffu) return patch;
    return (patch & 0xffff0000u) | pm_mirror_pos(patch & 0xffff, 1, HW, W);
}

// Units of the learner views (pacman_mappo_resnet.py:462-469): v = 0, 1: the two learner agents of team
// P.view_red as the actor sees them (canonicalize_obs); v = 2: merge_obs_for_critic of the two.
// mslot0: first of the two bitset slots that hold the mirrored snapshots (red learners only).
template <int MODE>
PM_HD bool pm_view_unit_load(const PmParams &P, const uint32_t *sm, const PmSmem &M, int v, int el, int mslot0, PmUnit &U) {
    U.lid = sm[M.lid + el];
    if (U.lid == 0xffffffffu) return false;
    U.fresh = (int)(U.lid >> 31);
    U.lid &= 0x7fffffffu;
    const int red = P.view_red != 0;
    const int l = v == 2 ? 0 : v;                                // whose snapshot the board comes from
    const int k = red ? 2 * l : 2 * l + 1;                       // agent index of that learner
    const int slot = red ? mslot0 + l : ((MODE == PM_MODE_STEP) ? k : 0);
    U.food = sm + M.food + (size_t)slot * P.NW * M.E1 + el;
    U.caps = P.has_caps ? sm + M.caps + (size_t)slot * P.NW * M.E1 + el : nullptr;
    const uint32_t pa = sm[M.posA + k * M.E1 + el], pb = sm[M.posB + k * M.E1 + el];
    U.p_own = pa & 0xffff; U.p_ally = pa >> 16; U.p_e1 = pb & 0xffff; U.p_e2 = pb >> 16;
    U.patch = sm[M.patch + k * M.E1 + el];
    U.p_own2 = 0xffffffffu; U.patch2 = 0xffffffffu; U.mirror = red;
    if (v == 2) {                                                // merged: both learners on plane 1, plane 4 cleared
        const int k2 = k + 2;
        U.p_ally = 0xffffffffu;
        U.p_own2 = sm[M.posA + k2 * M.E1 + el] & 0xffff;
        U.patch2 = sm[M.patch + k2 * M.E1 + el];
    }
    if (red) {
        U.p_own = pm_mirror_pos(U.p_own, 1, P.HW, P.W); U.p_own2 = pm_mirror_pos(U.p_own2, 1, P.HW, P.W);
        U.p_ally = pm_mirror_pos(U.p_ally, 4, P.HW, P.W);
        U.p_e1 = pm_mirror_pos(U.p_e1, 5, P.HW, P.W); U.p_e2 = pm_mirror_pos(U.p_e2, 5, P.HW, P.W);
        U.patch = pm_mirror_patch(U.patch, P.HW, P.W); U.patch2 = pm_mirror_patch(U.patch2, P.HW, P.W);
    }
    return true;
}

// Word j (bits 32j .. 32j+31) of the observation bit string of one unit:
// static wall plane | food planes 6,7 | capsule planes 2,3 | the four agent bits.
PM_HD uint32_t pm_cat_word(const PmParams &P, const uint32_t *st, const PmSmem &M, const PmUnit &U, int j) {
    const int NCAT = P.NCAT, NW = P.NW, HW = P.HW, E1 = M.E1;
    const uint32_t *m2 = st + M.masks + (U.mirror ? 4 * NCAT : 0), *m3 = m2 + NCAT, *m6 = m3 + NCAT, *m7 = m6 + NCAT;
    const int woff = (U.mirror ? pm_l_wallcatm(HW) : pm_l_wallcat(HW)) + j;
    uint32_t w;
    if (P.L == 1) w = st[M.lay + woff];
    else {
#if defined(__CUDA_ARCH__)
        w = __ldg(P.layouts + (size_t)U.lid * P.LS + woff);
#else
        w = P.layouts[(size_t)U.lid * P.LS + woff];
#endif
    }
    const int b0 = 32 * j;
    // a word overlaps at most two of the eight planes: the half-board masks are zero everywhere else, and consecutive
    // lanes hold consecutive words, so these branches are nearly uniform within a warp
    uint32_t mk;
    if ((mk = m6[j]) != 0) w |= pm_extract32(U.food, E1, NW, b0 - 6 * HW) & mk;
    if ((mk = m7[j]) != 0) w |= pm_extract32(U.food, E1, NW, b0 - 7 * HW) & mk;
    if (U.caps) {
        if ((mk = m2[j]) != 0) w |= pm_extract32(U.caps, E1, NW, b0 - 2 * HW) & mk;
        if ((mk = m3[j]) != 0) w |= pm_extract32(U.caps, E1, NW, b0 - 3 * HW) & mk;
    }
    unsigned d;
    d = U.p_own - (unsigned)b0;  if (d < 32u) w |= 1u << d;
    d = U.p_ally - (unsigned)b0; if (d < 32u) w |= 1u << d;
    d = U.p_e1 - (unsigned)b0;   if (d < 32u) w |= 1u << d;
    d = U.p_e2 - (unsigned)b0;   if (d < 32u) w |= 1u << d;
    d = U.p_own2 - (unsigned)b0; if (d < 32u) w |= 1u << d;
    return w;
}

#ifndef PM_ST256
#define PM_ST256 0
#endif
#ifndef PM_ALU_EXPAND
#define PM_ALU_EXPAND 0
#endif
// nibble -> four floats (bit i set -> 1.0f): 16-entry shared-memory table, or pure ALU (PM_ALU_EXPAND)
PM_HD pm_f4 pm_expand4(const pm_f4 *lut, uint32_t nib) {
#if PM_ALU_EXPAND && defined(__CUDA_ARCH__)
    pm_f4 v;
    v.x = __uint_as_float((nib & 1u) * 0x3f800000u);
    v.y = __uint_as_float((nib & 2u) * 0x1fc00000u);
    v.z = __uint_as_float((nib & 4u) * 0x0fe00000u);
    v.w = __uint_as_float((nib & 8u) * 0x07f00000u);
    return v;
#else
    return lut[nib];
#endif
}

// One warp turns whole observations into fp32 planes.  Per group of 32 bit-string
// words: lane L builds word 32g+L, then 8 store rounds follow in which lane L
// writes float4 number 256g + 32r + L of the observation; its four floats are
// nibble (L & 7) of word 32g + 4r + (L >> 3), fetched with one shuffle and turned
// into a float4 by a 16-entry shared-memory table.  A warp store covers 512
// contiguous bytes.  The own-cell value 1+numCarrying (gymPacMan.py:205) is
// patched by one lane after the warp's stores (ordered by __syncwarp).
template <int MODE>
PM_HD void pm_phaseB(const PmParams &P, const uint32_t *st, const uint32_t *sm, const PmSmem &M, int warp, int nwarps, int lane, int64_t n0) {
    const int NCAT = P.NCAT, E = P.E;
    const int nf4 = 2 * P.HW;                          // float4s per observation (8*HW/4)
    (void)nf4;
    const pm_f4 *lut = reinterpret_cast<const pm_f4 *>(st + M.lut);
    const int ngroups = (NCAT + 31) >> 5;
    for (int u = warp; u < 4 * E; u += nwarps) {
        const int k = u >> P.lgE, el = u & (E - 1);
        PmUnit U;
        if (!pm_unit_load<MODE>(P, sm, M, k, el, U)) continue;
        float *dstf = P.obs + (size_t)k * P.obs_agent_stride + (size_t)(n0 + el) * 8 * P.HW;
        pm_f4 *dst = reinterpret_cast<pm_f4 *>(dstf);
        (void)dst;
#if defined(__CUDA_ARCH__)
#if PM_ST256
        // 256-bit stores (STG.E.ENL2.256, sm_100+): lane L writes float8 number 128g + 32r + L = byte (L & 3) of word
        // 32g + 8r + (L >> 2); a warp store covers 1024 contiguous bytes, 4 rounds per 32 words.
        const int sh8 = (lane & 3) * 8, src8 = lane >> 2;
        const int nf8 = P.HW;
        for (int g = 0; g < ngroups; ++g) {
            const int j = 32 * g + lane;
            const uint32_t word = (j < NCAT) ? pm_cat_word(P, st, M, U, j) : 0u;
            const int fbase = 128 * g + lane;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const uint32_t w = __shfl_sync(0xffffffffu, word, 8 * r + src8) >> sh8;
                const int f = fbase + 32 * r;
                if (f < nf8) {
                    const pm_f4 lo = pm_expand4(lut, w & 15u), hi = pm_expand4(lut, (w >> 4) & 15u);
                    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dstf + 8 * (size_t)f), "f"(lo.x),
                                 "f"(lo.y), "f"(lo.z), "f"(lo.w), "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w)
                                 : "memory");
                }
            }
        }
        if (U.patch != 0xffffffffu) {
            __syncwarp();
            if (lane == 0) dstf[U.patch & 0xffff] = (float)(U.patch >> 16);
        }
#else
        const int sh = (lane & 7) * 4, src = lane >> 3;
#if PM_TMA_STORE
        // stage one group (32 words = 1024 floats = 4 KB) in this warp's shared-memory buffer, patch the own-cell
        // value there, and hand the buffer to the TMA engine: the LSU only sees shared-memory stores
        pm_f4 *buf = reinterpret_cast<pm_f4 *>(const_cast<uint32_t *>(sm) + M.stage + warp * 1024);
        const uint32_t buf_s = (uint32_t)__cvta_generic_to_shared(buf);
        const int pf = (U.patch != 0xffffffffu) ? (int)(U.patch & 0xffff) : -1;
        for (int g = 0; g < ngroups; ++g) {
            const int j = 32 * g + lane;
            const uint32_t word = (j < NCAT) ? pm_cat_word(P, st, M, U, j) : 0u;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // buffer free again?
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint32_t w = __shfl_sync(0xffffffffu, word, 4 * r + src);
                buf[32 * r + lane] = pm_expand4(lut, (w >> sh) & 15u);
            }
            __syncwarp();
            if (lane == 0 && (pf >> 10) == g) reinterpret_cast<float *>(buf)[pf & 1023] = (float)(U.patch >> 16);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const int left = nf4 - 256 * g;
                const uint32_t bytes = (uint32_t)(left >= 256 ? 256 : left) * 16u;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + 256 * g), "r"(buf_s), "r"(bytes)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
#else
        if (MODE == PM_MODE_STEP && P.obs_elide && !U.fresh) {
            // obs_static_valid: float4s that lie entirely inside plane 0 (walls) or, on capsule-free batches, planes
            // 2-3 already hold their value from the previous full write of this buffer - leave them alone
            // (ranges rounded inwards to whole 128-byte lines: a half-written 32-byte sector costs a read-modify-write)
            const int s0 = (P.HW >> 2) & ~7;            // float4 f is static if 4f+3 < HW ...
            const int s1 = (((2 * P.HW + 3) >> 2) + 7) & ~7, s2 = P.has_caps ? 0 : ((4 * P.HW) >> 2) & ~7;   // ... or 2HW <= 4f, 4f+3 < 4HW
            for (int g = 0; g < ngroups; ++g) {
                const int j = 32 * g + lane;
                const uint32_t word = (j < NCAT) ? pm_cat_word(P, st, M, U, j) : 0u;
                const int fbase = 256 * g + lane;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int f0 = 256 * g + 32 * r;    // warp-uniform: the whole round may be static
                    if (f0 + 32 <= s0 || (f0 >= s1 && f0 + 32 <= s2)) continue;
                    const uint32_t w = __shfl_sync(0xffffffffu, word, 4 * r + src);
                    const int f = fbase + 32 * r;
                    if (f < nf4 && !(f < s0 || (f >= s1 && f < s2))) dst[f] = pm_expand4(lut, (w >> sh) & 15u);
                }
            }
        } else {
            for (int g = 0; g < ngroups; ++g) {
                const int j = 32 * g + lane;
                const uint32_t word = (j < NCAT) ? pm_cat_word(P, st, M, U, j) : 0u;
                const int fbase = 256 * g + lane;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint32_t w = __shfl_sync(0xffffffffu, word, 4 * r + src);
                    const int f = fbase + 32 * r;
                    if (f < nf4) dst[f] = pm_expand4(lut, (w >> sh) & 15u);
                }
            }
        }
        if (U.patch != 0xffffffffu) {
            __syncwarp();
            if (lane == 0) dstf[U.patch & 0xffff] = (float)(U.patch >> 16);
        }
#endif
#endif
#else
        if (lane != 0) continue;                        // host emulation: lane 0 plays the whole warp
        for (int g = 0; g < ngroups; ++g) {
            uint32_t words[32];
            for (int L = 0; L < 32; ++L) { const int j = 32 * g + L; words[L] = (j < NCAT) ? pm_cat_word(P, st, M, U, j) : 0u; }
            const bool elide = MODE == PM_MODE_STEP && P.obs_elide && !U.fresh;
            const int s0 = (P.HW >> 2) & ~7, s1 = (((2 * P.HW + 3) >> 2) + 7) & ~7, s2 = P.has_caps ? 0 : ((4 * P.HW) >> 2) & ~7;
            for (int r = 0; r < 8; ++r)
                for (int L = 0; L < 32; ++L) {
                    const int f = 256 * g + 32 * r + L;
                    if (f < nf4 && !(elide && (f < s0 || (f >= s1 && f < s2))))
                        dst[f] = lut[(words[4 * r + (L >> 3)] >> ((L & 7) * 4)) & 15u];
                }
        }
        if (U.patch != 0xffffffffu) dstf[U.patch & 0xffff] = (float)(U.patch >> 16);
#endif
    }
#if PM_TMA_STORE && defined(__CUDA_ARCH__)
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all bulk stores of this warp have landed
#endif
}

// ============================================================================
// phase B': learner views as uint8 planes (f2: canonicalize_obs + merge_obs_for_critic
// fused into the encoder, pacman_mappo_resnet.py:215-229, 267-274)
// ============================================================================
// Three units per env: the two learners' actor observations and the merged critic observation, each
// 8*H*W bytes (0/1 planes, own cell = 1 + numCarrying: exact in uint8).  Lane L builds word 32g+L of the
// (possibly mirrored) bit string; four store rounds follow in which lane L writes bytes 8*(32r+L) ..+7 of
// the group = byte (L & 3) of word 32g + 8r + (L >> 2), spread to one bit per byte by a multiply-and-mask.
PM_HD uint32_t pm_spread4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }

// One (view, word) ITEM per lane: the CTA's E x 3 views x NCAT words form one flat list that the warps walk in
// chunks of 32, so no lane idles on the short tail of a view (35 words on tinyCapture = one full pass + 3 lanes in
// the per-view schedule this replaces).  Each lane expands its own word: 32 bits -> 32 bytes, written as ONE
// 256-bit store (two 128-bit / four 64-bit stores when the view's row pitch is not 32- / 16-byte aligned); the
// lanes of a warp cover a contiguous 1 KB run.  No shuffles, no second pass for the own-cell values: the lane that
// owns the word holding the patched element substitutes the byte before storing.
template <int MODE>
PM_HD void pm_phaseB_views(const PmParams &P, const uint32_t *st, const uint32_t *sm, const PmSmem &M, int warp, int nwarps,
                           int lane, int64_t n0, int mslot0) {
    const int NCAT = P.NCAT, E = P.E;
    const int nbytes = 8 * P.HW;                        // bytes per view
    const int nitems = 3 * E * NCAT;
    const uint32_t magic = 0xffffffffu / (uint32_t)NCAT + 1u;      // i / NCAT == (i * magic) >> 32 for i < 2^16
    uint8_t *const base0 = P.view_actor + (size_t)n0 * nbytes, *const base1 = base0 + P.view_actor_stride;
    uint8_t *const base2 = P.view_critic + (size_t)n0 * nbytes;
#if defined(__CUDA_ARCH__)
    const int lanes = 32, first = lane;
#else
    if (lane != 0) return;                              // host emulation: lane 0 plays the whole warp
    const int lanes = 1, first = 0;
#endif
    for (int i = warp * 32 + first; i < nitems; i += (lanes == 1 ? 1 : nwarps * 32)) {
#if !defined(__CUDA_ARCH__)
        if ((i >> 5) % nwarps != warp) continue;        // the chunks of 32 items this warp owns
#endif
        const int u = (int)(((unsigned long long)(uint32_t)i * magic) >> 32), j = i - u * NCAT;
        const int v = u >> P.lgE, el = u & (E - 1);
        PmUnit U;
        if (!pm_view_unit_load<MODE>(P, sm, M, v, el, mslot0, U)) continue;
        uint8_t *dstb = (v == 2 ? base2 : v == 1 ? base1 : base0) + (uint32_t)el * (uint32_t)nbytes;
        const uint32_t word = pm_cat_word(P, st, M, U, j);
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = pm_spread4((word >> (4 * q)) & 15u);     // (a 256 x 8-byte shared-memory table
                                                                                    //  was measured 5 % slower on B200)
        // own-cell values 1 + numCarrying (gymPacMan.py:205); merged view: elementwise max of the two learners' planes
        const uint32_t lo = 32u * (uint32_t)j;
        uint32_t pv = 0, pe = 0xffffffffu;
        if (U.patch != 0xffffffffu && (U.patch & 0xffff) - lo < 32u) { pe = (U.patch & 0xffff) - lo; pv = U.patch >> 16; }
        if (U.patch2 != 0xffffffffu && (U.patch2 & 0xffff) - lo < 32u) {
            const uint32_t e2 = (U.patch2 & 0xffff) - lo, v2 = U.patch2 >> 16;
            if (pe == e2) pv = v2 > pv ? v2 : pv;
            else if (pe == 0xffffffffu) { pe = e2; pv = v2; }
            else {                                      // two different cells inside one word: apply the first now
                const uint32_t sh0 = 8 * (pe & 3);
                o[pe >> 2] = (o[pe >> 2] & ~(0xffu << sh0)) | ((pv > 255 ? 255u : pv) << sh0);
                pe = e2; pv = v2;
            }
        }
        if (pe != 0xffffffffu) {
            const uint32_t sh0 = 8 * (pe & 3);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if ((int)(pe >> 2) == q) o[q] = (o[q] & ~(0xffu << sh0)) | ((pv > 255 ? 255u : pv) << sh0);
        }
        uint8_t *d = dstb + 32 * j;
        const int left = nbytes - 32 * j;               // >= 8 and a multiple of 8
#if defined(__CUDA_ARCH__)
        if (left >= 32 && ((size_t)d & 31) == 0) {
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(d), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                         "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                         : "memory");
        } else if (left >= 32 && ((size_t)d & 15) == 0) {
            reinterpret_cast<pm_u4 *>(d)[0] = pm_u4{o[0], o[1], o[2], o[3]};
            reinterpret_cast<pm_u4 *>(d)[1] = pm_u4{o[4], o[5], o[6], o[7]};
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (8 * q < left) reinterpret_cast<unsigned long long *>(d)[q] = (unsigned long long)o[2 * q] | (unsigned long long)o[2 * q + 1] << 32;
        }
#else
        for (int q = 0; q < 4; ++q)
            if (8 * q < left) reinterpret_cast<unsigned long long *>(d)[q] = (unsigned long long)o[2 * q] | (unsigned long long)o[2 * q + 1] << 32;
#endif
    }
}

// ============================================================================
// phase B'': packed observations (pm_step_out.obs_bits) for a host-side consumer
// ============================================================================
// The bit string the fp32 encoder expands is itself the output: NCAT words per observation, one patch word
// (own-cell element index | value << 16, or 0xFFFFFFFF), zero padding up to NB words.  One warp per
// observation, consecutive lanes = consecutive words (a 144-byte run on tinyCapture).
template <int MODE>
PM_HD void pm_phaseB_bits(const PmParams &P, const uint32_t *st, const uint32_t *sm, const PmSmem &M, int warp, int nwarps,
                          int lane, int64_t n0) {
    const int NCAT = P.NCAT, NB = P.NB, E = P.E;
    for (int u = warp; u < 4 * E; u += nwarps) {
        const int k = u >> P.lgE, el = u & (E - 1);
        PmUnit U;
        if (!pm_unit_load<MODE>(P, sm, M, k, el, U)) continue;
        uint32_t *dst = P.obs_bits + (size_t)k * P.obs_bits_stride + (size_t)(n0 + el) * NB;
#if defined(__CUDA_ARCH__)
        for (int j = lane; j < NB; j += 32)
#else
        if (lane != 0) continue;                        // host emulation: lane 0 plays the whole warp
        for (int j = 0; j < NB; ++j)
#endif
            dst[j] = j < NCAT ? pm_cat_word(P, st, M, U, j) : (j == NCAT ? U.patch : 0u);
    }
}

// ============================================================================
// GameState.makeObservation (capture.py:268-292): what agent `index` may know - SURVEY row f4
// ============================================================================
#define PM_SIGHT_RANGE 5            // capture.py:69, Manhattan
#define PM_SONAR_NOISE_RANGE 13     // capture.py:67; SONAR_NOISE_VALUES[k] = k - 6
// For env n and every observer i: visible[n][i] = bitmask over the four agents whose configuration survives (own
// team always; an enemy iff some member of i's team stands within SIGHT_RANGE of it), and
// noisy[n][i][j] = int(manhattan(pos_i, pos_j) + SONAR_NOISE_VALUES[k]) with k = random.choice's index: taken from
// draws[n][i][j] when given (wraps modulo 13), else hashed from (seed, global env, step, i, j).
PM_HD void pm_make_observation_env(const PmParams &P, int64_t n, const uint32_t *draws, uint8_t *visible, int32_t *noisy) {
    const pm_u4 A = reinterpret_cast<const pm_u4 *>(P.state + pm_st_agA(P.N))[n];
    const pm_u4 Mi = reinterpret_cast<const pm_u4 *>(P.state + pm_st_misc(P.N))[n];
    const uint32_t lid = P.layout_id ? P.layout_id[n] : (P.L == 1 ? 0u : (uint32_t)(n % P.L));
    const uint32_t teams = (P.layouts + (size_t)lid * P.LS + pm_l_tail(P.HW))[PM_L_TEAMS];
    const uint32_t a[4] = {A.x, A.y, A.z, A.w};
    int x[4], y[4];
    for (int i = 0; i < 4; ++i) { x[i] = a[i] & 0xff; y[i] = (a[i] >> 8) & 0xff; }
    for (int i = 0; i < 4; ++i) {
        const int ti = (teams >> i) & 1;
        uint32_t vis = 0;
        for (int e = 0; e < 4; ++e) {
            if ((int)((teams >> e) & 1) == ti) { vis |= 1u << e; continue; }
            for (int m = 0; m < 4; ++m) {
                if ((int)((teams >> m) & 1) != ti) continue;
                const int dx = x[e] - x[m], dy = y[e] - y[m];
                if ((dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) <= PM_SIGHT_RANGE) vis |= 1u << e;
            }
        }
        visible[4 * n + i] = (uint8_t)vis;
        for (int j = 0; j < 4; ++j) {
            const int dx = x[i] - x[j], dy = y[i] - y[j];
            const uint32_t k = draws ? draws[16 * n + 4 * i + j] % PM_SONAR_NOISE_RANGE
                                     : (uint32_t)(((unsigned long long)pm_random_u32(P.bot_seed ^ 0x736f6e6172ull, (unsigned long long)(P.n_base + n),
                                                                                    Mi.w, Mi.y * 16u + 4u * i + j) * PM_SONAR_NOISE_RANGE) >> 32);
            noisy[16 * n + 4 * i + j] = (dx < 0 ? -dx : dx) + (dy < 0 ? -dy : dy) + (int)k - (PM_SONAR_NOISE_RANGE - 1) / 2;
        }
    }
}

// ---- pm_env_desc -> PmParams (shared by the C ABI glue and the test emulation) ----
#ifdef PACMAN_B200_H
static inline int pm_fill_params(const pm_env_desc *d, PmParams &P, int E) {
    if (!d || !d->geom || !d->layouts || !d->state || d->n_envs <= 0 || d->n_layouts <= 0) return PM_EINVAL;
    if (d->W < 3 || d->H < 3) return PM_EINVAL;
    if (d->W > 255 || d->H > 255 || d->W * d->H > PM_MAX_CELLS) return PM_ETOOBIG;
    if (d->env_begin < 0 || d->env_count < 0 || (d->env_begin & 31) || d->env_begin + d->env_count > d->n_envs ||
        (d->env_count == 0 && d->env_begin != 0))
        return PM_EINVAL;
    P = PmParams();
    P.W = d->W; P.H = d->H; P.HW = d->W * d->H; P.NW = pm_nw(P.HW); P.NCAT = pm_ncat(P.HW);
    P.ND = pm_nd(P.W, P.H); P.LS = pm_layout_words_(P.HW);
    P.L = d->n_layouts; P.has_caps = d->has_caps != 0; P.N = d->n_envs;
    P.geom = d->geom; P.layouts = d->layouts; P.layout_id = d->layout_id; P.state = (uint8_t *)d->state;
    P.layout_pool = d->layout_pool != 0 && d->n_layouts > 1;
    if (d->layout_pool && !d->layout_id) return PM_EINVAL;
    P.layout_id_rw = const_cast<uint32_t *>(d->layout_id);
    P.bot_draws = d->bot_draws;
    P.NB = pm_obs_bits_words_(P.HW);
    if (d->envs_per_cta != 0 && d->envs_per_cta != 8 && d->envs_per_cta != 32) return PM_EINVAL;
    P.epc = d->envs_per_cta;
    P.length = d->length; P.reward_legal = d->reward_for_legal_action != 0; P.defence = d->defence_reward != 0;
    P.shaping_scale = d->shaping_scale;
    P.bot_kind = 0;
    for (int i = 0; i < 4; ++i) {
        int k = (d->bot_kind >> (4 * i)) & 15;
        if ((d->random_agent_mask >> i) & 1) k = 1;
        if (k > 3) return PM_EINVAL;
        if (k >= 2 && (!d->dist || !d->node_of_cell || d->dist_nmax <= 0)) return PM_EINVAL;
        P.bot_kind |= k << (4 * i);
    }
    P.bot_seed = d->random_seed; P.bot_first_best = d->bot_first_best;
    P.dist = d->dist; P.node_of_cell = d->node_of_cell; P.dist_nmax = d->dist_nmax;
    P.auto_reset = d->auto_reset != 0;
    P.n_base = d->env_index_base;
    P.n_begin = d->env_begin; P.n_end = d->env_count ? d->env_begin + d->env_count : d->n_envs;
    P.a0 = 0; P.a1 = 4;
    if (d->agent_begin || d->agent_end) {
        if (d->agent_begin < 0 || d->agent_end > 4 || d->agent_begin >= d->agent_end) return PM_EINVAL;
        P.a0 = d->agent_begin; P.a1 = d->agent_end;
    }
    P.E = E; P.lgE = 0;
    while ((1 << P.lgE) < E) ++P.lgE;
    return 0;
}
#endif
