// pm_format.h — bit-level formats shared by the host packer, the kernels and
// the C ABI.  Everything is u32 words; a "cell" is c = y*W + x (row-major,
// y = 0 at the BOTTOM row, like the reference's Grid[x][y]).
//
// Board bitsets ("c-indexed"): NW = ceil(HW/32) words, bit (c & 31) of word
// (c >> 5); bits >= HW are always zero.
//
// Observation bit string ("cat"): the reference's observation for one agent
// is 8 planes of H*W floats (gymPacMan.py:195-229).  Seven planes are 0/1 and
// the eighth (own cell) is 0 / 1+carrying, so one observation is fully
// described by an 8*HW-bit string S (bit p*HW + c = plane p, cell c) plus one
// (position, value) patch.  NCAT = ceil(8*HW/32) = ceil(HW/4) words.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define PM_HD __host__ __device__ __forceinline__
#else
#define PM_HD inline
#endif

#define PM_MAX_CELLS_ 1024

// ---- geometry record (per batch; depends only on W,H) ----------------------
// [0] W [1] H [2] HW [3] NW [4] NCAT [5] ND (dump offsets) [6] LS (layout stride) [7] R (dump radius)
// [8 .. 8+NCAT)            M2: plane-2 bits whose cell is a "blue capsule" cell (2x >  W)  capture.py:344-350
// [.. +NCAT)               M3: plane-3 bits, "red capsule" cells               (2x <= W)
// [.. +NCAT)               M6: plane-6 bits, blue food half (x >= W/2 int)     capture.py:332-342
// [.. +NCAT)               M7: plane-7 bits, red food half  (x <  W/2 int)
// [.. +4*NCAT)             M2m M3m M6m M7m: the same four masks for the MIRRORED (canonical red-learner) view of
//                          pacman_mappo_resnet.py:215-229: x -> W-1-x with the plane pairs (2,3) and (6,7) swapped
// [.. +ND)                 dump visit order: (dx & 0xff) | (dy & 0xff) << 8    capture.py:635-661
enum { PM_G_W = 0, PM_G_H, PM_G_HW, PM_G_NW, PM_G_NCAT, PM_G_ND, PM_G_LS, PM_G_R, PM_G_HDR = 8 };

PM_HD int pm_nw(int HW) { return (HW + 31) >> 5; }
PM_HD int pm_ncat(int HW) { return (8 * HW + 31) >> 5; }
// packed observation (pm_step_out.obs_bits): NCAT bit-string words + 1 patch word, padded to a multiple of 4 words
PM_HD int pm_obs_bits_words_(int HW) { return (pm_ncat(HW) + 1 + 3) & ~3; }
PM_HD int pm_dump_radius(int W, int H) { return (W > H ? W : H) - 1; }
PM_HD int pm_nd(int W, int H) { int r = pm_dump_radius(W, H); return (2 * r + 1) * (2 * r + 1); }
#define PM_G_NMASKS 8
PM_HD int pm_geom_words_(int W, int H) { return PM_G_HDR + PM_G_NMASKS * pm_ncat(W * H) + pm_nd(W, H); }
PM_HD int pm_g_mask_off(int HW, int which /*0:M2 1:M3 2:M6 3:M7, 4-7: mirrored*/) { return PM_G_HDR + which * pm_ncat(HW); }
PM_HD int pm_g_dump_off(int HW) { return PM_G_HDR + PM_G_NMASKS * pm_ncat(HW); }

// ---- layout record (per board) ---------------------------------------------
// [0 .. NW)        walls      (c-indexed)
// [NW .. 2NW)      initial food
// [2NW .. 3NW)     initial capsules
// [3NW .. +NCAT)   wallcat: the plane-0 part of the observation bit string
// [.. +NCAT)       wallcat of the mirrored board (canonical red-learner view)
// then 8 words:    start[4] = x | y<<8 | startIsPacman<<16 ; teams (bit i = agent i red,
//                  capture.py:316-319) ; totalFood ; redFood0 | blueFood0<<16 ; hasCaps
enum { PM_L_START = 0, PM_L_TEAMS = 4, PM_L_TOTALFOOD = 5, PM_L_FOODCNT = 6, PM_L_HASCAPS = 7, PM_L_TAILWORDS = 8 };
PM_HD int pm_l_walls(int HW) { (void)HW; return 0; }
PM_HD int pm_l_food(int HW) { return pm_nw(HW); }
PM_HD int pm_l_caps(int HW) { return 2 * pm_nw(HW); }
PM_HD int pm_l_wallcat(int HW) { return 3 * pm_nw(HW); }
PM_HD int pm_l_wallcatm(int HW) { return 3 * pm_nw(HW) + pm_ncat(HW); }
PM_HD int pm_l_tail(int HW) { return 3 * pm_nw(HW) + 2 * pm_ncat(HW); }
PM_HD int pm_layout_words_(int HW) { return (pm_l_tail(HW) + PM_L_TAILWORDS + 3) & ~3; }

// ---- per-env state (device blob, SoA over envs) -----------------------------
// agA  uint4[N]  per agent: x | y<<8 | dir<<16 | isPacman<<19 | scaredTimer<<20
// agB  uint4[N]  per agent: numCarrying | numReturned<<16
// snap uint4[N]  per agent: own cell of the last observation handed out: x | y<<8 | carrying<<16
// misc uint4[N]  score (i32) ; steps ; redFood | blueFood<<16 (board counts) ; reserved
// food u32[NW][NP] ; caps u32[NW][NP] ; acc uint4[4][NP]      NP = N rounded up to 32
PM_HD int64_t pm_np(int64_t N) { return (N + 31) & ~(int64_t)31; }
PM_HD size_t pm_st_agA(int64_t N) { (void)N; return 0; }
PM_HD size_t pm_st_agB(int64_t N) { return (size_t)pm_np(N) * 16; }
PM_HD size_t pm_st_snap(int64_t N) { return (size_t)pm_np(N) * 32; }
PM_HD size_t pm_st_misc(int64_t N) { return (size_t)pm_np(N) * 48; }
PM_HD size_t pm_st_food(int64_t N) { return (size_t)pm_np(N) * 64; }
PM_HD size_t pm_st_caps(int64_t N, int HW) { return pm_st_food(N) + (size_t)pm_np(N) * 4 * pm_nw(HW); }
// acc uint4[4][N]: the accumulators of a step that is played in several launches (pm_env_desc.agent_begin/agent_end):
//   [0] red reward, blue reward (fp64 bits)   [1] score change, red / blue score delta, actions used   [2],[3] shaping[4] (fp64)
PM_HD size_t pm_st_acc(int64_t N, int HW) { return pm_st_caps(N, HW) + (size_t)pm_np(N) * 4 * pm_nw(HW); }
PM_HD size_t pm_state_bytes_(int64_t N, int HW) { return pm_st_acc(N, HW) + (size_t)pm_np(N) * 64; }

PM_HD uint32_t pm_agA_pack(int x, int y, int dir, int pac, int scared) {
    return (uint32_t)x | (uint32_t)y << 8 | (uint32_t)dir << 16 | (uint32_t)(pac & 1) << 19 | (uint32_t)scared << 20;
}
