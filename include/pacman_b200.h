/*
 * pacman_b200.h — C ABI of the B200-native batched capture-the-flag Pac-Man
 * environment (libpacman_b200.so).
 *
 * The reference has no FFI: its "operator API" for this path is the Python class
 * gymPacMan.gymPacMan_parallel_env (gymPacMan.py:12-270) plus three trainer
 * helpers (pacman_mappo_resnet.py:241-290,577) and distanceCalculator.py:111-150.
 * Each entry point below names the reference interface it replaces.  All data
 * pointers marked [dev] are device pointers owned by the caller (PyTorch
 * tensors in the bundled host code), every launch goes to the caller's stream,
 * nothing here synchronises the device except the *_host entry points.  The
 * library keeps no per-batch state: everything a call needs is in its arguments,
 * so independent batches can be driven from different host threads / streams.
 * (The only process-wide data are per-device caches of the SM count and of the
 * kernels' dynamic shared-memory attribute; they are mutex-guarded.)
 * Return value: 0 = OK, > 0 = cudaError_t from the launch, < 0 = PM_E* argument
 * error.
 *
 * Batch layout ("SoA over envs"): N independent 4-agent games share one board
 * size W x H.  Boards come from a table of L packed layout records (L = 1: one
 * layout for all envs; L = N: one maze per env; env n uses layout_id[n], or
 * n % L when layout_id is NULL).
 */
#ifndef PACMAN_B200_H
#define PACMAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM_ABI_VERSION 8

/* argument errors */
#define PM_EINVAL   (-1)  /* null pointer / bad size                                   */
#define PM_ELAYOUT  (-2)  /* layout text malformed (ragged, != 4 agents, open border)  */
#define PM_ETOOBIG  (-3)  /* W*H > PM_MAX_CELLS or W,H > 255                           */
#define PM_ENOGPU   (-4)  /* no CUDA device / kernel image not loadable on this device */

#define PM_MAX_CELLS 1024

/* gym action ints, gymPacMan.py:66-89 */
enum { PM_NORTH = 0, PM_EAST = 1, PM_SOUTH = 2, PM_WEST = 3, PM_STOP = 4 };

typedef void *pm_stream_t; /* cudaStream_t */

/* ---- sizes ------------------------------------------------------------- */
int pm_abi_version(void);
/* u32 words of the per-batch geometry record / of one packed layout record */
int pm_geom_words(int W, int H);
int pm_layout_words(int W, int H);
/* bytes of the device state blob for N envs */
size_t pm_state_bytes(int W, int H, int64_t n_envs);

/* u32 words of one packed observation (pm_step_out.obs_bits): ceil(8*W*H/32) + 1 patch word, rounded up to 4 */
int pm_obs_bits_words(int W, int H);

/* ---- host-side packing (pure CPU, no CUDA call) -------------------------- */
/* Geometry record: plane masks for the three half-board formulas
 * (capture.py:325-350) and the food-dump visit order of
 * AgentRules.dumpFoodFromDeath (capture.py:635-661) as an offset table.      */
int pm_pack_geom(int W, int H, uint32_t *out);
/* One `.lay` board (layout.py:95-130): text = H rows of W chars, row 0 = top.
 * Writes pm_layout_words(W,H) words; *has_caps (optional) = board has capsules. */
int pm_pack_layout(const char *text, int W, int H, uint32_t *out, int *has_caps);
/* L boards of one size at once (a pool of mazeGenerator boards): texts = n_layouts x (H*W) characters back to back,
 * out = n_layouts x pm_layout_words(W,H) words.  Optional outputs: *has_caps = any board has a capsule, *max_open =
 * the largest number of open cells of any board (the nmax pm_distances needs), *bad = index of the first malformed
 * board when the call returns a PM_E* code.                                                                    */
int pm_pack_layouts(const char *texts, int W, int H, int n_layouts, uint32_t *out, int *has_caps, int *max_open, int *bad);

/* ---- batch descriptor ---------------------------------------------------- */
typedef struct pm_env_desc {
    int32_t W, H;
    int32_t n_layouts;            /* L                                             */
    int32_t has_caps;             /* any layout in the table has a capsule          */
    int64_t n_envs;               /* N                                             */
    const uint32_t *geom;         /* [dev] pm_geom_words words                      */
    const uint32_t *layouts;      /* [dev] L * pm_layout_words words                */
    const uint32_t *layout_id;    /* [dev] [N] or NULL                              */
    void *state;                  /* [dev] pm_state_bytes bytes                     */
    /* gymPacMan.__init__ kwargs (gymPacMan.py:15) */
    int32_t length;
    int32_t reward_for_legal_action;
    int32_t defence_reward;
    /* bit i set: agent i ignores `actions` and plays a uniformly random LEGAL action drawn in-kernel at
       its own sub-move (the randomTeam opponent, agents/randomTeam.py:94-100; gymPacMan.py:157)      */
    int32_t random_agent_mask;
    double shaping_scale;         /* SHAPING_SCALE, pacman_mappo_resnet.py:34       */
    uint64_t random_seed;         /* counter-based: draw = hash(seed, global env, episode, step, agent) */
    /* ---- ABI 2 ---- */
    /* pm_step only: an env whose step terminates is re-initialised inside the same launch
       (the caller-side `if done: env.reset()` of pacman_mappo_resnet.py:524-538 folded in).
       reward / done / score_change / shaping describe the terminal step; obs / legal and the
       stored state are those of the fresh episode.                                         */
    int32_t auto_reset;
    /* ---- ABI 7 ---- envs per CTA of the env kernels: 0 = chosen from the batch size (the large-batch geometry, or 8
       for batches that would leave SMs idle); 8 = force the small CTAs, 32 = force the large-batch geometry (32 envs
       per CTA, 64 for the launches bound by the rules phase).  Launch geometry only; results never depend on it.  */
    int32_t envs_per_cta;
    int64_t env_index_base;       /* global index of env 0 of this batch (rank offset when the batch is a
                                     shard): keys the in-kernel random draws, so trajectories do not
                                     depend on how envs are partitioned over GPUs                      */
    /* launch only envs [env_begin, env_begin + env_count) of the batch; env_count 0 = all.
       env_begin must be a multiple of 32.  Outputs are still indexed by the batch-wide env number. */
    int64_t env_begin, env_count;
    /* ---- ABI 4: on-device opponents (the team interface of gymPacMan.py:33-40,157) ---- */
    /* one nibble per agent: 0 = plays `actions`; 1 = randomTeam (same as random_agent_mask);
       2 / 3 = baselineTeam's OffensiveReflexAgent / DefensiveReflexAgent
       (agents/baselineTeam.py:56-187): one-step look-ahead over the legal actions with
       integer features and maze distances, evaluated at the agent's own sub-move.        */
    int32_t bot_kind;
    int32_t bot_first_best;        /* baselineTeam tie-break: 0 = uniform over the best actions (random.choice,
                                      counter-based draw), 1 = first best in N,S,E,W,Stop order (parity tests) */
    const uint16_t *dist;          /* [dev] pm_distances output [L][nmax][nmax]; required for kinds 2, 3        */
    const int16_t *node_of_cell;   /* [dev] pm_distances output [L][H*W]                                        */
    int32_t dist_nmax;
    /* ---- ABI 7: a new maze every episode (gymPacMan.py:98-100 with random_layout=True, capture.py:905-911) ----
     * non-zero: `layouts` is a POOL of L boards and `layout_id` ([N], required, written by the kernels) names the
     * board each env is playing.  Whenever an env (re)starts - pm_reset, or auto_reset inside pm_step - it moves to
     * board  hash(random_seed, global env index, episode) % L, so the sequence of boards an env sees depends on
     * nothing but (seed, env, episode number).                                                             */
    int32_t layout_pool;
    /* ---- ABI 6: a step played in several launches ----
     * pm_step plays only the sub-moves of agents [agent_begin, agent_end) (0, 0 = the whole step).  The launches
     * of one step must cover 0..4 in order; the state keeps the step's accumulators in between.  A launch writes
     * the observations of the agents it moved (and actions_used so far); rewards, done, legal masks, shaping and
     * the restart of finished envs come with the launch that ends at 4.  This is the hook for opponents that
     * decide on the host at their own sub-move from the live state (gymPacMan.py:157): step agent 0, hand the
     * state to the bot of agent 1, step agent 1, ...  Learner views need the whole step in one launch.        */
    int32_t agent_begin, agent_end;
    /* ---- ABI 7: externally supplied random draws for the on-device opponents ----
     * [dev] [N][4] uint32 or NULL.  When set, an on-device bot does not hash its own draw: agent a of env n uses
     * bot_draws[n][a] as the INDEX random.choice would have returned - into the legal-action list in the
     * reference's order N,S,E,W,Stop for randomTeam (agents/randomTeam.py:100), into the list of best actions
     * (same order) for baselineTeam (agents/baselineTeam.py:92); an index past the end wraps (modulo).  Feeding
     * the indices recorded from the reference's own MT19937 stream replays its trajectories bit for bit.     */
    const uint32_t *bot_draws;
} pm_env_desc;

enum { PM_BOT_NONE = 0, PM_BOT_RANDOM = 1, PM_BOT_BASELINE_OFFENSE = 2, PM_BOT_BASELINE_DEFENSE = 3 };

/* Outputs of one step; any pointer may be NULL to skip that output. */
typedef struct pm_step_out {
    float *obs;                   /* [dev] agent a, env n at obs + a*obs_agent_stride + n*8*H*W (fp32 [8][H][W]) */
    int64_t obs_agent_stride;     /* in floats; N*8*H*W for a dense [4][N][8][H][W] tensor                      */
    double *reward;               /* [dev] [2][N]: row 0 red team (agents 0,2), row 1 blue (1,3); fp64 like the
                                     reference's Python floats                                                  */
    uint8_t *done;                /* [dev] [N]                                                                   */
    uint8_t *legal;               /* [dev] [N][4] bit a = gym action a legal, from the FINAL state              */
    int32_t *score_change;        /* [dev] [N]                                                                   */
    double *shaping;              /* [dev] [4][N] compute_heuristic_shaping per agent (fp64)                    */
    float *shaped_reward;         /* [dev] [4][N] fp32(2*team_reward + shaping_scale*shaping), i.e. the value
                                     the trainer stores in reward_buf (pacman_mappo_resnet.py:518-522)          */
    /* ---- ABI 3 ---- */
    struct pm_host_record *packed; /* [dev] [N] every small per-env output in one 48-byte record, so a host-side
                                     consumer needs ONE device-to-host copy per step (pm_step_host_pipelined)  */
    /* ---- ABI 4 ---- */
    uint8_t *actions_used;         /* [dev] [N][4] the action every agent actually played (illegal requests
                                     appear as Stop, bot agents as the bot's choice)                          */
    /* ---- ABI 5: learner views, the trainer's network inputs written by the encoder itself ----
     * canonicalize_obs + merge_obs_for_critic (pacman_mappo_resnet.py:215-229, 267-274, 462-469) fused into the
     * step: for the learner team `view_red` (0: agents 1,3 as they are; 1: agents 0,2 mirrored in x with the
     * plane pairs (2,3) and (6,7) swapped) the kernel writes uint8 planes (0/1 and 1+numCarrying, exact):
     *   view_actor  [dev] learner l of env n at view_actor + l*view_actor_stride + n*8*H*W   (u8 [8][H][W])
     *   view_critic [dev] [N][8][H][W]: first learner's view, plane 1 = max over both learners, plane 4 = 0
     * Both NULL = off.  Pointers and stride must be 8-byte aligned; boards up to 32 columns.  With `obs` NULL
     * the step writes 3*8*H*W bytes per env instead of 4*32*H*W.                                              */
    uint8_t *view_actor;
    int64_t view_actor_stride;
    uint8_t *view_critic;
    int32_t view_red;
    /* ---- ABI 7: leave what cannot have changed untouched ----
     * non-zero: the caller guarantees that `obs` is the SAME buffer the previous pm_step / pm_reset / pm_observe of this
     * batch wrote in full (the host class's persistent observation tensor).  The step then skips the elements of
     * every observation that cannot differ from what is already there: plane 0 (walls, fixed per board) and, when no
     * board of the batch has a capsule, planes 2-3 (all zero) - 12.5 % / 37.5 % of the bytes.  An env that moved to
     * another board inside the launch (layout_pool) is written in full.  The tensors read back are identical.   */
    int32_t obs_static_valid;
    /* ---- ABI 7: packed observations for a HOST-side consumer ----
     * The eight planes of an observation hold 0/1 except the own cell (1 + numCarrying), so an observation is an
     * 8*H*W-bit string plus one patch.  obs_bits: [dev or mapped pinned host] agent a, env n at
     * obs_bits + a*obs_bits_agent_stride + n*pm_obs_bits_words(W,H) (u32 words): bit (p*H*W + y*W + x) & 31 of word
     * (p*H*W + y*W + x) >> 5 = plane p, row y, column x; then ONE patch word = element index | value << 16 of the
     * own cell (0xFFFFFFFF: value is 1, nothing to patch); then zero padding to a multiple of 4 words.
     * 144 B per observation on tinyCapture instead of 4 480 B of fp32 - what crosses PCIe when the policy lives on
     * the host.  NULL = off.  Pointer 16-byte aligned, stride a multiple of 4 words.                          */
    uint32_t *obs_bits;
    int64_t obs_bits_agent_stride;   /* in u32 words */
} pm_step_out;

/* What a CPU-side caller of gymPacMan.step reads back every step (gymPacMan.py:184-193), one record per env. */
typedef struct pm_host_record {
    double reward_red, reward_blue;   /* rewards dict: agents 0,2 / agents 1,3                                  */
    float shaped_reward[4];           /* trainer reward per agent (see pm_step_out.shaped_reward)               */
    int32_t score_change;             /* info['score_change']                                                   */
    uint8_t legal[4];                 /* info['legal_actions'] as bitmasks                                      */
    uint8_t done;                     /* terminations                                                           */
    uint8_t pad[7];
} pm_host_record;

/* ---- environment --------------------------------------------------------- */
/* gymPacMan_parallel_env.reset (gymPacMan.py:92-141) for every env with
 * mask[n] != 0 (mask NULL = all).  Writes obs/legal only for the envs it resets. */
int pm_reset(const pm_env_desc *d, const uint8_t *mask, float *obs, int64_t obs_agent_stride,
             uint8_t *legal, pm_stream_t stream);

/* gymPacMan_parallel_env.step (gymPacMan.py:143-193): four ordered sub-moves
 * (GameState.generateSuccessor, capture.py:107-123), per-agent mid-step
 * observation (get_Observation, :195-229), team rewards (get_reward, :231-259),
 * termination (:261-270).  actions: [dev] [N][4], one element per agent;
 * action_itemsize = 1 (uint8) or 8 (int64).  Values outside 0..4 act as an
 * illegal move (-> Stop, no legal-action bonus).                             */
int pm_step(const pm_env_desc *d, const void *actions, int action_itemsize, const pm_step_out *out,
            pm_stream_t stream);

/* ---- ABI 8: a run of steps whose actions are already resident ----
 * `for t in range(n_steps): env.step(actions[t])` (gymPacMan.py:143-193 driven by a policy that does not look at the
 * observations: the self-play random-action loop of BASELINE config 2, scripted opponents, a recorded move list) as
 * n_steps launches of the same kernel on `stream` that OVERLAP: launch t+1 is a programmatic dependent launch
 * (sm_90+: cudaLaunchAttributeProgrammaticStreamSerialization), so its CTAs take the SM slots launch t frees in its
 * tail and run their rules phase under launch t's last stores instead of after them.  Envs are independent, so the
 * only ordering that matters is per env tile: every CTA publishes a counter when it is done and the CTA of the next
 * launch that owns the same envs (same grid, same tile) waits for it (acquire / release at GPU scope) before it reads
 * the state or writes the outputs.  Results are identical to n_steps calls of pm_step.
 *   actions  [dev] step t reads the [N][4] block at actions + t*actions_step_stride elements (stride 0 = the same
 *            block every step); ALL of it must be valid before the call
 *   outs     n_outs descriptors; step t writes outs[t % n_outs] (1 = the same buffers every step, n_steps = a
 *            rollout buffer with one slot per step)
 *   scratch  [dev] pm_chain_scratch_bytes(N) bytes, 4-byte aligned, owned by the caller; contents irrelevant
 * Whole steps of the whole batch only (agent_begin/agent_end and env_begin/env_count must be 0).  Work queued on
 * `stream` after the call is ordered after the last step as usual.                                              */
size_t pm_chain_scratch_bytes(int64_t n_envs);
int pm_step_chain(const pm_env_desc *d, const void *actions, int action_itemsize, int64_t actions_step_stride,
                  int n_steps, const pm_step_out *outs, int n_outs, void *scratch, pm_stream_t stream);

/* pm_step_chain with HOST buffers and no copy engine (see pm_step_host_mapped): step t reads its actions from
 * actions_pinned[t] ([n_steps][N][4] uint8) and writes its records to records_pinned[t] ([n_steps][N]
 * pm_host_record), both page-locked, straight across PCIe from inside the kernel.  out_dev->packed is ignored;
 * the other non-NULL members of out_dev stay in device memory and are written by every step.  Asynchronous: the
 * records are complete once the caller has synchronised `stream`.                                             */
int pm_step_chain_host_mapped(const pm_env_desc *d, const uint8_t *actions_pinned, int n_steps, const pm_step_out *out_dev,
                              struct pm_host_record *records_pinned, void *scratch, pm_stream_t stream);

/* get_Observation(agentIndex) on the CURRENT state (gymPacMan.py:195-229) for
 * all four agents of every env. */
int pm_observe(const pm_env_desc *d, float *obs, int64_t obs_agent_stride, pm_stream_t stream);

/* The packed observations (pm_step_out.obs_bits) of the CURRENT state for all four agents. */
int pm_observe_bits(const pm_env_desc *d, uint32_t *obs_bits, int64_t obs_bits_agent_stride, pm_stream_t stream);

/* The learner views (see pm_step_out) of the CURRENT state, e.g. right after pm_reset. */
int pm_observe_views(const pm_env_desc *d, int view_red, uint8_t *view_actor, int64_t view_actor_stride,
                     uint8_t *view_critic, pm_stream_t stream);

/* pm_step with HOST buffers (the call a CPU-side consumer of the reference API makes):
 * copies actions_host ([N][4] uint8, ideally pinned) into actions_dev, runs pm_step
 * into the DEVICE buffers of out_dev, copies every member of out_dev that has a
 * non-NULL counterpart in out_host back to the host, and synchronises the stream.
 * out_host->obs may be NULL to keep observations resident in HBM for a policy
 * network on the same GPU.  out_host->view_actor / view_critic / obs_bits (ABI 7)
 * bring the uint8 learner views / the packed observations to the host instead. */
int pm_step_host(const pm_env_desc *d, const uint8_t *actions_host, uint8_t *actions_dev,
                 const pm_step_out *out_dev, const pm_step_out *out_host, pm_stream_t stream);

/* The same call, pipelined: the batch is cut into `n_chunks` env ranges that alternate between two
 * internal streams, so the device-to-host copy of one range's records overlaps the kernel of the
 * next.  Per step: one H2D (actions), one launch and one D2H (pm_host_record) per chunk.  Work is
 * ordered after everything already queued on `stream`, later work on `stream` is ordered after the
 * step, and the call returns when `records_host` ([N] pm_host_record, pinned) is complete.
 * out_dev->packed and actions_dev are required device scratch; other members of out_dev may be NULL
 * except obs, which stays in HBM for the policy.  A pipe belongs to the CUDA device that was current
 * when it was created and must not be used from two threads at once.                                */
typedef struct pm_pipe pm_pipe;
int pm_pipe_create(pm_pipe **pipe, int n_chunks);
int pm_pipe_destroy(pm_pipe *pipe);
int pm_step_host_pipelined(const pm_env_desc *d, pm_pipe *pipe, const uint8_t *actions_host, uint8_t *actions_dev,
                           const pm_step_out *out_dev, struct pm_host_record *records_host, pm_stream_t stream);

/* The same call with NO copy engine at all: `actions_pinned` ([N][4] uint8) and `records_pinned`
 * ([N] pm_host_record) are page-locked host buffers (cudaHostAlloc / torch pin_memory); the step
 * kernel reads the actions and writes the records straight across PCIe (coalesced 128-bit
 * accesses), so one step is ONE launch plus a stream synchronise.  out_dev->packed is ignored;
 * observations and any other non-NULL member of out_dev stay in device memory - except
 * out_dev->obs_bits, which may itself point at page-locked host memory: the packed observations
 * then cross PCIe from inside the kernel too.
 * PM_EINVAL if a buffer is not page-locked.                                                       */
int pm_step_host_mapped(const pm_env_desc *d, const uint8_t *actions_pinned, const pm_step_out *out_dev,
                        struct pm_host_record *records_pinned, pm_stream_t stream);
/* Same launch without the final synchronise: the records are valid once the caller has synchronised
 * `stream` (lets a caller queue more work, e.g. a copy of the observations, under the same wait). */
int pm_step_host_mapped_async(const pm_env_desc *d, const uint8_t *actions_pinned, const pm_step_out *out_dev,
                              struct pm_host_record *records_pinned, pm_stream_t stream);

/* Unpack the device state into plain arrays (GameState accessors,
 * capture.py:125-234): agents [N][4][7] int32 = x,y,dir,isPacman,scaredTimer,
 * numCarrying,numReturned; food/caps [N][H*W] uint8 (cell y*W+x);
 * misc [N][3] int32 = score, steps, episode (restarts of this env so far: it keys
 * the in-kernel opponent draws and the layout pool).  Any pointer may be NULL.
 * Export / import describe the state BETWEEN whole steps: the parked accumulators
 * of a step that is being played in several launches (agent_begin/agent_end) are
 * not part of it.                                                             */
int pm_get_state(const pm_env_desc *d, int32_t *agents, uint8_t *food, uint8_t *caps, int32_t *misc,
                 pm_stream_t stream);
/* Inverse of pm_get_state (checkpoint restore / crafted test states). All four required; misc is [N][3]. */
int pm_set_state(const pm_env_desc *d, const int32_t *agents, const uint8_t *food, const uint8_t *caps,
                 const int32_t *misc, pm_stream_t stream);

/* GameState.makeObservation (capture.py:268-292) for every env and every observer - what the tournament loop
 * (game.py:570-732) hands an agent instead of the full state:
 *   visible [dev] [N][4] uint8: bit e of visible[n][i] = agent e's configuration survives in agent i's view (own team
 *           always; an enemy iff a member of i's team stands within SIGHT_RANGE = 5 Manhattan cells of it)
 *   noisy_distances [dev] [N][4][4] int32: int(manhattanDistance(pos_i, pos_j) + SONAR_NOISE_VALUES[k]), k in 0..12
 *           (noise -6..+6, capture.py:67-78), i = observer
 *   draws [dev] [N][4][4] uint32 or NULL: the index k `random.choice(SONAR_NOISE_VALUES)` returned for each distance
 *           (wraps modulo 13); NULL = hashed from (random_seed, global env index, episode, step, i, j).          */
int pm_make_observation(const pm_env_desc *d, const uint32_t *draws, uint8_t *visible, int32_t *noisy_distances,
                        pm_stream_t stream);

/* ---- maze distances ------------------------------------------------------ */
/* distanceCalculator.computeDistances (distanceCalculator.py:111-150) for L
 * layouts at once.  Node order = walls.asList(False) (x-major, y-minor).
 *   n_open [dev] [L] int32        number of open cells per layout
 *   node_of_cell [dev] [L][H*W] int16   cell (y*W+x) -> node index, -1 for walls
 *   dist [dev] [L][nmax][nmax] uint16   dist[l][i][j] = distance(node i, node j);
 *        0xFFFF = unreachable (sys.maxsize in the reference) or i/j >= n_open[l] */
int pm_distances(int W, int H, const uint32_t *layouts, int n_layouts, int32_t *n_open,
                 int16_t *node_of_cell, uint16_t *dist, int nmax, pm_stream_t stream);

/* ---- rollout post-processing --------------------------------------------- */
/* compute_gae (pacman_mappo_resnet.py:277-290) over M independent columns:
 * rew/val/done/adv/ret are [T][M] fp32 row-major, last_value [M]; fp32
 * arithmetic in the reference's operation order (no FMA contraction).        */
int pm_gae(const float *rew, const float *val, const float *done, const float *last_value,
           double gamma, double lam, float *adv, float *ret, int T, int64_t M, pm_stream_t stream);
/* Minibatch advantage normalisation (pacman_mappo_resnet.py:577): for each
 * consecutive group of mb indices, out = (adv[idx]-mean)/(std_unbiased+1e-8). */
int pm_adv_norm(const float *adv, const int64_t *inds, float *out, int64_t n, int mb, pm_stream_t stream);

/* ---- ABI 7: the PPO minibatch pipeline around the networks -------- */
/* b_obs[mb].to(device) / b_merged[mb] (pacman_mappo_resnet.py:569-573) over the uint8 rollout buffer:
 * dst[r][:] = (float or bf16) src[idx[r]][:], rows of row_bytes bytes (a multiple of 8).  src [dev] 8-byte
 * aligned, dst [dev] 16-byte aligned, idx [dev] int64 [n_rows].                                          */
int pm_gather_rows(const uint8_t *src, const int64_t *idx, int64_t n_rows, int row_bytes, void *dst, int dst_bf16,
                   pm_stream_t stream);
/* The minibatch loss of pacman_mappo_resnet.py:575-585 and its gradient in one launch:
 *   lp = log_softmax(logits)[action]; ratio = exp(lp - old_logp);
 *   pg = -min(adv*ratio, adv*clamp(ratio, 1-eps, 1+eps)).mean(); vl = 0.5*((value - ret)^2).mean();
 *   loss = pg + vf_coef*vl - ent_coef*entropy.mean()
 * logits [B][n_actions <= 8] fp32; values [n_values] with sample i reading values[i % n_values] (n_values = B, or
 * B/2 when both learners of an env-step share one critic value); actions int64; old_logp / adv / ret fp32 [B].
 * Outputs: dlogits [B][n_actions], dvalues [B] (per sample) = d loss / d input; stats[8] = loss, pg, vl,
 * entropy mean, clip fraction (`:597-600`).  scratch: pm_ppo_scratch_bytes() bytes, zeroed once by the caller.  */
size_t pm_ppo_scratch_bytes(void);
int pm_ppo_loss(const float *logits, const float *values, const int64_t *actions, const float *old_logp,
                const float *adv, const float *ret, int64_t B, int64_t n_values, int n_actions, double clip_eps,
                double vf_coef, double ent_coef, float *dlogits, float *dvalues, float *stats, void *scratch,
                pm_stream_t stream);
/* nn.utils.clip_grad_norm_ + optimizer.step() (Adam, eps as given, no weight decay) + the EMA of the policy
 * (pacman_mappo_resnet.py:588-595) over FLAT fp32 buffers of n elements: grads are first scaled by grad_scale
 * (1/world after a sum all-reduce) and clipped to max_grad_norm; ema (may be NULL) = ema*decay + p*(1-decay).
 * step = 1 for the first update.  grad_norm_out [dev] float: the norm before clipping.  scratch:
 * pm_adam_scratch_bytes() bytes, zeroed once by the caller.  No host synchronisation.                      */
size_t pm_adam_scratch_bytes(void);
int pm_adam_ema(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, float *ema, int64_t n, int64_t step,
                double lr, double beta1, double beta2, double eps, double max_grad_norm, double grad_scale,
                double ema_decay, float *grad_norm_out, void *scratch, pm_stream_t stream);

/* ---- ABI 8: the critic's attention core ----
 * softmax(Q K^T / sqrt(8)) V of nn.MultiheadAttention inside the reference's centralised critic
 * (nn.TransformerEncoder(d_model 32, nhead 4) over one token per board cell, pacman_mappo_resnet.py:131-136,
 * 168-176) and its gradient, for heads of dimension 8 and S <= 1024 tokens, plain fp32.  Layouts are
 * nn.MultiheadAttention's (batch_first = False):
 *   qkv [dev] [S][B][3E] = in_proj(x), E = 8*n_heads (q | k | v along the last axis)
 *   out [dev] [S][B][E]  concatenated heads (input of out_proj);  lse [dev] [B][n_heads][S] log-sum-exp per row
 *   (forward: may be NULL when no backward pass follows);  dout like out, dqkv like qkv.  16-byte aligned. */
int pm_attn_fwd(const float *qkv, float *out, float *lse, int S, int64_t B, int n_heads, pm_stream_t stream);
int pm_attn_bwd(const float *qkv, const float *out, const float *dout, const float *lse, float *dqkv, int S, int64_t B,
                int n_heads, pm_stream_t stream);

/* ---- ABI 8: the networks' normalisations at the shapes PyTorch's kernels handle worst ----
 * y = LayerNorm_32(a + b) * gamma + beta over `rows` rows of 32 floats (b may be NULL): the post-norm residual of the
 * critic's nn.TransformerEncoderLayer (d_model 32, pacman_mappo_resnet.py:131-136).  stats [rows][2] = (mean, rstd),
 * may be NULL when no backward pass follows.  Backward: dx = d loss / d a = d loss / d b; partial is
 * [pm_layernorm32_blocks(rows)][64] floats whose column sums are dgamma (0..31) and dbeta (32..63).            */
int pm_layernorm32_blocks(int64_t rows);
int pm_add_layernorm32_fwd(const float *a, const float *b, const float *gamma, const float *beta, float *y, float *stats,
                           int64_t rows, double eps, pm_stream_t stream);
int pm_add_layernorm32_bwd(const float *a, const float *b, const float *gamma, const float *stats, const float *dy, float *dx,
                           float *partial, int64_t rows, pm_stream_t stream);
/* y = [gelu](GroupNorm(4, 32)(x) [+ res]) on [n_samples][32][HW] (NCHW, HW even, <= pm_groupnorm_4x8_max_hw() = 1024): the
 * ResidualBlock of the actor (`:49-68`): `act(gn1(conv1(x)))` (res NULL, gelu 1) and `act(gn2(conv2(h)) + x)` (res = x,
 * gelu 1).  stats [n_samples][4][2] = (mean, rstd) per group.  Backward: dx (w.r.t. x), dres (w.r.t. res, NULL when res
 * is NULL); partial is [pm_groupnorm_blocks(n_samples)][64] floats whose column sums are dgamma and dbeta.        */
int pm_groupnorm_blocks(int64_t n_samples);
int pm_groupnorm_4x8_max_hw(void);
int pm_groupnorm_4x8_fwd(const float *x, const float *res, const float *gamma, const float *beta, float *y, float *stats,
                         int64_t n_samples, int HW, double eps, int gelu, pm_stream_t stream);
int pm_groupnorm_4x8_bwd(const float *x, const float *res, const float *gamma, const float *beta, const float *stats,
                         const float *dy, float *dx, float *dres, float *partial, int64_t n_samples, int HW, int gelu,
                         pm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PACMAN_B200_H */
