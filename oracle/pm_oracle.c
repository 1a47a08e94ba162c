/*
 * pm_oracle.c — CPU restatement of the reference's capture-env hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (pacman_marl_2025_b200/) may import, link or call this file; it is the
 * checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  It deliberately uses the reference's own data
 * shapes (byte grids addressed [x][y], a capsule *list*, a literal FIFO BFS
 * for the food dump, a priority-queue UCS for maze distances) instead of the
 * bit-packed layout and offset tables of the CUDA path, so that agreement
 * between the two is evidence and not a tautology.
 *
 * Parity is PINNED: tests/test_oracle_golden.py replays the action sequences in
 * tests/golden/traj_*.npz (recorded from the unmodified Python reference by
 * tests/golden/make_golden.py) through this file and compares every field of
 * every step; dist_*.npz and gae.npz pin the distance and GAE functions.
 *
 * Reference files restated (all under /root/reference):
 *   gymPacMan.py:92-270        reset / step / get_Observation / get_reward / check_termination
 *   capture.py:107-123         GameState.generateSuccessor
 *   capture.py:311-350         initialize, isRed, halfGrid, halfList
 *   capture.py:453-728         AgentRules.*
 *   game.py:105-118,295-352    Configuration.generateSuccessor, Actions
 *   game.py:490-508            GameStateData.initialize
 *   layout.py:95-130           .lay parsing
 *   distanceCalculator.py:111-150  computeDistances
 *   pacman_mappo_resnet.py:241-290,577  shaping, compute_gae, advantage normalisation
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PMO_MAXW 64
#define PMO_MAXH 64
#define PMO_MAXCAPS 64

/* capture.py:66-75 */
#define KILL_POINTS 0
#define MIN_FOOD 2
#define SCARED_TIME 40

/* gym action ints, gymPacMan.py:66-89 */
enum { A_NORTH = 0, A_EAST = 1, A_SOUTH = 2, A_WEST = 3, A_STOP = 4 };
/* game.py:295-299 */
static const int VEC_X[5] = {0, 1, 0, -1, 0};
static const int VEC_Y[5] = {1, 0, -1, 0, 0};

typedef struct {
    int x, y;        /* configuration.pos (always integral: speed 1.0, capture.py:479) */
    int dir;         /* configuration.direction as a gym int; Stop until first move  */
    int is_pacman, scared, carrying, returned;
    int sx, sy;      /* start configuration                                           */
    int start_is_pacman;
} pmo_agent;

typedef struct {
    int W, H;
    uint8_t walls[PMO_MAXW][PMO_MAXH];
    uint8_t food[PMO_MAXW][PMO_MAXH];
    uint8_t food0[PMO_MAXW][PMO_MAXH];
    int ncaps, capx[PMO_MAXCAPS], capy[PMO_MAXCAPS];
    int ncaps0, capx0[PMO_MAXCAPS], capy0[PMO_MAXCAPS];
    pmo_agent ag[4];
    int team_red[4];           /* capture.py:316-319, teams by start column */
    int total_food;            /* capture.py:322-323                        */
    int score, score_change, win;
    int steps, length;
    int reward_legal, defence_reward;
    /* own-cell plane of the last observation handed out per agent (for shaping) */
    int snap_x[4], snap_y[4], snap_carry[4];
    /* bot support (agents/baselineTeam.py): shared all-pairs table, set by pmo_set_distances */
    const struct pmo_dist_table *dt;
    /* random.choice indices for this step's bot decisions (one per agent), or NULL: ties -> first best.
     * Lets a replay follow a trajectory the reference played with its real MT19937 stream. */
    const uint32_t *draws;
} pmo_env;

/* distanceCalculator.Distancer as the bots use it (distanceCalculator.py:38-62): a lookup
 * keyed by two grid points.  dist[i * n + j], nodes numbered x-major / y-minor.            */
typedef struct pmo_dist_table {
    int n;
    int16_t node_of[PMO_MAXW][PMO_MAXH];
    int32_t *dist;
} pmo_dist_table;

size_t pmo_sizeof_env(void) { return sizeof(pmo_env); }

/* ------------------------------------------------------------------ layout */
/* layout.py:95-130.  rows: H strings of W chars, text row 0 is y = H-1.      */
int pmo_env_init(pmo_env *e, const char *text, int W, int H, int length, int reward_legal,
                 int defence_reward) {
    if (W < 3 || H < 3 || W > PMO_MAXW || H > PMO_MAXH) return -1;
    memset(e, 0, sizeof(*e));
    e->W = W; e->H = H; e->length = length;
    e->reward_legal = reward_legal; e->defence_reward = defence_reward;
    int tag[16], tx[16], ty[16], nt = 0;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            char ch = text[(H - 1 - y) * W + x];
            if (ch == '%') e->walls[x][y] = 1;
            else if (ch == '.') e->food0[x][y] = 1;
            else if (ch == 'o') {
                if (e->ncaps0 >= PMO_MAXCAPS) return -2;
                e->capx0[e->ncaps0] = x; e->capy0[e->ncaps0] = y; e->ncaps0++;
            } else if (ch == 'P' || ch == 'G' || (ch >= '1' && ch <= '4')) {
                if (nt >= 16) return -3;
                tag[nt] = ch == 'P' ? 0 : ch == 'G' ? 1 : ch - '0';
                tx[nt] = x; ty[nt] = y; nt++;
            }
        }
    }
    if (nt != 4) return -4;                       /* the gym adapter needs exactly 4 */
    /* agentPositions.sort(): by (tag, (x, y)) */
    for (int i = 0; i < nt; ++i)
        for (int j = i + 1; j < nt; ++j) {
            int less = tag[j] < tag[i] || (tag[j] == tag[i] && (tx[j] < tx[i] || (tx[j] == tx[i] && ty[j] < ty[i])));
            if (less) {
                int t;
                t = tag[i]; tag[i] = tag[j]; tag[j] = t;
                t = tx[i]; tx[i] = tx[j]; tx[j] = t;
                t = ty[i]; ty[i] = ty[j]; ty[j] = t;
            }
        }
    for (int i = 0; i < 4; ++i) {
        e->ag[i].sx = tx[i]; e->ag[i].sy = ty[i];
        e->ag[i].start_is_pacman = tag[i] == 0;
    }
    /* closed border required: the reference would wrap (negative index) or raise */
    for (int x = 0; x < W; ++x) if (!e->walls[x][0] || !e->walls[x][H - 1]) return -5;
    for (int y = 0; y < H; ++y) if (!e->walls[0][y] || !e->walls[W - 1][y]) return -5;
    for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y) e->total_food += e->food0[x][y];
    return 0;
}

/* capture.py:325-330 */
static int is_red_x(const pmo_env *e, int x) { return (double)x < e->W / 2.0; }

/* gymPacMan.py:92-141 + capture.py:369-382,311-323 + game.py:490-508 */
void pmo_reset(pmo_env *e) {
    memcpy(e->food, e->food0, sizeof(e->food));
    e->ncaps = e->ncaps0;
    memcpy(e->capx, e->capx0, sizeof(e->capx));
    memcpy(e->capy, e->capy0, sizeof(e->capy));
    for (int i = 0; i < 4; ++i) {
        pmo_agent *a = &e->ag[i];
        a->x = a->sx; a->y = a->sy; a->dir = A_STOP;
        a->is_pacman = a->start_is_pacman;
        a->scared = 0; a->carrying = 0; a->returned = 0;
        e->team_red[i] = is_red_x(e, a->sx);
        e->snap_x[i] = a->x; e->snap_y[i] = a->y; e->snap_carry[i] = 0;
    }
    e->score = 0; e->score_change = 0; e->win = 0; e->steps = 0;
}

/* ------------------------------------------------------------------- rules */
/* capture.py:453-461 + game.py:335-350: bit a set <=> gym action a legal      */
static int legal_mask(const pmo_env *e, int i) {
    const pmo_agent *a = &e->ag[i];
    int m = 0;
    for (int act = 0; act < 5; ++act)
        if (!e->walls[a->x + VEC_X[act]][a->y + VEC_Y[act]]) m |= 1 << act;
    return m;
}

static int in_capsules(const pmo_env *e, int x, int y) {
    for (int k = 0; k < e->ncaps; ++k) if (e->capx[k] == x && e->capy[k] == y) return k;
    return -1;
}

/* capture.py:519-560 */
static void consume(pmo_env *e, int x, int y, int isRed) {
    if (e->food[x][y]) {
        for (int i = 0; i < 4; ++i) {
            if (e->team_red[i] != isRed) continue;         /* teamIndicesFunc() order */
            if (e->ag[i].x == x && e->ag[i].y == y) { e->ag[i].carrying += 1; break; }
        }
        e->food[x][y] = 0;
    }
    /* red eats "blue capsules" (x > W/2), blue eats "red capsules" (x <= W/2): capture.py:344-350 */
    int k = in_capsules(e, x, y);
    int mine = isRed ? ((double)x > e->W / 2.0) : ((double)x <= e->W / 2.0);
    if (k >= 0 && mine) {
        for (int j = k; j + 1 < e->ncaps; ++j) { e->capx[j] = e->capx[j + 1]; e->capy[j] = e->capy[j + 1]; }
        e->ncaps--;
        for (int i = 0; i < 4; ++i) if (e->team_red[i] != isRed) e->ag[i].scared = SCARED_TIME;
    }
}

/* capture.py:468-517 */
static void apply_action(pmo_env *e, int action, int idx) {
    int legal = legal_mask(e, idx);
    if (action < 0 || action > 4 || !((legal >> action) & 1)) action = A_STOP;
    pmo_agent *agentState = &e->ag[idx];
    /* Configuration.generateSuccessor: Stop keeps the old direction (game.py:115-117) */
    agentState->x += VEC_X[action];
    agentState->y += VEC_Y[action];
    if (action != A_STOP) agentState->dir = action;
    int nx = agentState->x, ny = agentState->y;

    int isRed = e->team_red[idx];
    agentState->is_pacman = (isRed + is_red_x(e, nx)) == 1;
    if (agentState->carrying > 0 && !agentState->is_pacman) {
        int score = isRed ? agentState->carrying : -agentState->carrying;
        e->score_change += score;
        agentState->returned += agentState->carrying;
        agentState->carrying = 0;
        int redCount = 0, blueCount = 0;
        for (int index = 0; index < 4; ++index) {
            agentState = &e->ag[index];               /* the rebinding: capture.py:505 */
            if (e->team_red[index]) redCount += agentState->returned;
            else blueCount += agentState->returned;
        }
        if (redCount >= (e->total_food / 2.0) - MIN_FOOD || blueCount >= (e->total_food / 2.0) - MIN_FOOD)
            e->win = 1;
    }
    if (agentState->is_pacman) consume(e, nx, ny, isRed);   /* capture.py:514-515 */
}

/* capture.py:569-668: literal FIFO graph search over 8-neighbourhoods */
static void dump_food(pmo_env *e, pmo_agent *victim) {
    if (victim->carrying == 0) return;
    int side = is_red_x(e, victim->x);
    enum { OFF = 80, DIM = 2 * OFF + PMO_MAXW };
    static _Thread_local uint8_t seen[DIM][DIM];
    memset(seen, 0, sizeof(seen));
    int cap = 1 << 16, head = 0, tail = 0;
    int *qx = (int *)malloc(sizeof(int) * cap), *qy = (int *)malloc(sizeof(int) * cap);
    qx[tail] = victim->x; qy[tail] = victim->y; tail++;
    int todo = victim->carrying;
    while (todo > 0) {
        if (head == tail) abort();                   /* 'Exhausted BFS! uh oh' */
        int x = qx[head], y = qy[head]; head++;
        if (x + OFF < 0 || y + OFF < 0 || x + OFF >= DIM || y + OFF >= DIM) abort();
        if (seen[x + OFF][y + OFF]) continue;
        seen[x + OFF][y + OFF] = 1;
        int ok = !(x >= e->W || y >= e->H || x <= 0 || y <= 0);
        if (ok && e->walls[x][y]) ok = 0;
        if (ok && e->food[x][y]) ok = 0;
        if (ok && is_red_x(e, x) != side) ok = 0;
        if (ok && in_capsules(e, x, y) >= 0) ok = 0;
        if (ok) for (int i = 0; i < 4; ++i) if (e->ag[i].x == x && e->ag[i].y == y) ok = 0;
        if (ok) { e->food[x][y] = 1; todo--; }
        if (tail + 9 > cap) {
            cap *= 2;
            qx = (int *)realloc(qx, sizeof(int) * cap); qy = (int *)realloc(qy, sizeof(int) * cap);
        }
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy) { qx[tail] = x + dx; qy[tail] = y + dy; tail++; }
    }
    free(qx); free(qy);
    victim->carrying = 0;
}

static void send_home(pmo_agent *a) {
    a->is_pacman = 0; a->x = a->sx; a->y = a->sy; a->dir = A_STOP; a->scared = 0;
}

/* capture.py:670-728 */
static void check_death(pmo_env *e, int idx) {
    pmo_agent *agentState = &e->ag[idx];
    int myRed = e->team_red[idx];
    if (agentState->is_pacman) {
        for (int index = 0; index < 4; ++index) {
            if (e->team_red[index] == myRed) continue;
            pmo_agent *other = &e->ag[index];
            if (other->is_pacman) continue;
            if (abs(other->x - agentState->x) + abs(other->y - agentState->y) <= 0.7) {
                if (other->scared <= 0) { dump_food(e, agentState); e->score_change += 0; send_home(agentState); }
                else { send_home(other); }
            }
        }
    } else {
        for (int index = 0; index < 4; ++index) {
            if (e->team_red[index] == myRed) continue;
            pmo_agent *other = &e->ag[index];
            if (!other->is_pacman) continue;
            if (abs(other->x - agentState->x) + abs(other->y - agentState->y) <= 0.7) {
                if (agentState->scared <= 0) { dump_food(e, other); send_home(other); }
                else { send_home(agentState); }
            }
        }
    }
}

/* capture.py:107-123 (in place; the caller keeps a copy when it needs the predecessor) */
static void successor(pmo_env *e, int idx, int action) {
    e->score_change = 0; e->win = 0;                  /* game.py:394-396 */
    apply_action(e, action, idx);
    check_death(e, idx);
    pmo_agent *a = &e->ag[idx];
    a->scared = a->scared - 1 > 0 ? a->scared - 1 : 0;   /* capture.py:562-567 */
    e->score += e->score_change;
}

/* capture.py:332-342 */
static int half_food_count(const pmo_env *e, int red) {
    int halfway = (int)(e->W / 2.0), n = 0;
    int x0 = red ? 0 : halfway, x1 = red ? halfway : e->W;
    for (int y = 0; y < e->H; ++y) for (int x = x0; x < x1; ++x) n += e->food[x][y];
    return n;
}

/* gymPacMan.py:231-259 */
static double get_reward(const pmo_env *s, const pmo_env *next, int idx, int action, double reward) {
    int red_mover = (idx == 0 || idx == 2);
    int mine = red_mover, theirs = !red_mover;
    if (half_food_count(s, mine) < half_food_count(next, mine)) reward += 1;
    if (half_food_count(s, theirs) > half_food_count(next, theirs)) reward += 0.1;
    if (s->defence_reward) {
        for (int k = 0; k < 2; ++k) {
            int o = red_mover ? 1 + 2 * k : 2 * k;
            if (s->ag[o].is_pacman && !next->ag[o].is_pacman)
                if (next->ag[o].x == next->ag[o].sx && next->ag[o].y == next->ag[o].sy) reward += 0.25;
        }
    }
    if (s->reward_legal) {
        int legal = legal_mask(s, idx);
        if (action >= 0 && action <= 4 && ((legal >> action) & 1)) reward += 0.01;
    }
    return reward;
}

/* gymPacMan.py:195-229 -> out[8][H][W] fp32 */
void pmo_observe(const pmo_env *e, int me, float *out) {
    int W = e->W, H = e->H, HW = W * H;
    memset(out, 0, sizeof(float) * 8 * HW);
    for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y) out[0 * HW + y * W + x] = e->walls[x][y];
    out[1 * HW + e->ag[me].y * W + e->ag[me].x] = (float)(1 + e->ag[me].carrying);
    for (int o = 0; o < 4; ++o) {
        if (o == me) continue;
        int ally = ((o == 0 || o == 2) && (me == 0 || me == 2)) || ((o == 1 || o == 3) && (me == 1 || me == 3));
        if (ally) {
            out[4 * HW + e->ag[o].y * W + e->ag[o].x] = 1;
            for (int k = 0; k < e->ncaps; ++k)            /* getBlueCapsules: x > W/2 */
                if ((double)e->capx[k] > W / 2.0) out[2 * HW + e->capy[k] * W + e->capx[k]] = 1;
        } else {
            out[5 * HW + e->ag[o].y * W + e->ag[o].x] = 1;
            for (int k = 0; k < e->ncaps; ++k)            /* getRedCapsules: x <= W/2 */
                if ((double)e->capx[k] <= W / 2.0) out[3 * HW + e->capy[k] * W + e->capx[k]] = 1;
        }
    }
    int halfway = (int)(W / 2.0);
    for (int y = 0; y < H; ++y) {
        for (int x = halfway; x < W; ++x) out[6 * HW + y * W + x] = e->food[x][y];
        for (int x = 0; x < halfway; ++x) out[7 * HW + y * W + x] = e->food[x][y];
    }
}

/* pacman_mappo_resnet.py:241-264 (flip-invariant, so canonicalisation is moot) */
static double shaping_value(int py, int px, int pcarry, int ny, int nx) {
    double living = -0.15;
    double dist = abs(py - ny) + abs(px - nx);
    if (dist > 1.5) return living - 0.4 - (0.15 * (double)pcarry);
    if (dist < 0.1) return living - 0.05;
    return living;
}

/* gymPacMan.py:143-193.
 * obs: [4][8][H][W] fp32 or NULL; reward[2] = {red, blue}; legal[4] bitmasks from
 * the FINAL state; shaping[4] vs. the previous observation of each agent.       */
static int baseline_choose(const pmo_env *e, int idx, int kind, int *best_mask);
static int random_choose(const pmo_env *e, int idx);

/* bot_kind[4] (or NULL): 0 = agent plays actions_in[idx]; 2 / 3 = baselineTeam Offensive / Defensive reflex agent
 * choosing at its own sub-move from the live state (gymPacMan.py:157), ties -> FIRST best action in the
 * reference's action order.  actions_used[4] / best_masks[4] (optional) report what was played.            */
void pmo_step_bots(pmo_env *e, const int32_t *actions_in, const uint8_t *bot_kind, float *obs, double *reward,
                   uint8_t *done, uint8_t *legal, int32_t *score_change, double *shaping, int32_t *actions_used,
                   uint8_t *best_masks) {
    double blue_reward = 0, red_reward = 0;
    int sc = 0, blue_sc = 0, red_sc = 0;
    int HW = e->W * e->H;
    int32_t actions[4];
    for (int idx = 0; idx < 4; ++idx) {
        actions[idx] = actions_in[idx];
        int bm = 0;
        if (bot_kind && bot_kind[idx] >= 2) actions[idx] = baseline_choose(e, idx, bot_kind[idx], &bm);
        else if (bot_kind && bot_kind[idx] == 1) actions[idx] = random_choose(e, idx);
        if (actions_used) actions_used[idx] = actions[idx];
        if (best_masks) best_masks[idx] = (uint8_t)bm;
        pmo_env next = *e;
        successor(&next, idx, actions[idx]);
        if (idx == 1 || idx == 3) {
            blue_reward = get_reward(e, &next, idx, actions[idx], blue_reward);
            *e = next;
            sc += e->score_change; blue_sc -= e->score_change;
        } else {
            red_reward = get_reward(e, &next, idx, actions[idx], red_reward);
            *e = next;
            sc += e->score_change; red_sc += e->score_change;
        }
        if (obs) pmo_observe(e, idx, obs + (size_t)idx * 8 * HW);
        if (shaping)
            shaping[idx] = shaping_value(e->snap_y[idx], e->snap_x[idx], e->snap_carry[idx], e->ag[idx].y, e->ag[idx].x);
        e->snap_x[idx] = e->ag[idx].x; e->snap_y[idx] = e->ag[idx].y; e->snap_carry[idx] = e->ag[idx].carrying;
    }
    blue_reward = blue_reward + (blue_sc > 0 ? blue_sc : 0);
    red_reward = red_reward + (red_sc > 0 ? red_sc : 0);
    /* check_termination, gymPacMan.py:261-270 (steps not yet incremented) */
    int term = 0;
    if (half_food_count(e, 0) == 0 && e->ag[0].carrying == 0 && e->ag[2].carrying == 0) term = 1;
    if (half_food_count(e, 1) == 0 && e->ag[1].carrying == 0 && e->ag[3].carrying == 0) term = 1;
    if (e->steps >= e->length) term = 1;
    if (term) {
        int fs = e->score;
        if (fs < 0) blue_reward += 20.0 + (1.0 / 5) * abs(fs);
        else if (fs > 0) red_reward += 20.0 + (1.0 / 5) * abs(fs);
    }
    e->steps += 1;
    reward[0] = red_reward; reward[1] = blue_reward;
    *done = (uint8_t)term;
    for (int i = 0; i < 4; ++i) legal[i] = (uint8_t)legal_mask(e, i);
    *score_change = sc;
}

void pmo_step(pmo_env *e, const int32_t *actions, float *obs, double *reward, uint8_t *done,
              uint8_t *legal, int32_t *score_change, double *shaping) {
    pmo_step_bots(e, actions, NULL, obs, reward, done, legal, score_change, shaping, NULL, NULL);
}

/* ------------------------------------------------------------------- bots */
/* agents/baselineTeam.py:56-187 on top of captureAgents.py (getFood :176-187, getOpponents :211-220,
 * getMazeDistance :250-258).  All feature values and weights are integers, so Q-values are exact.   */
static int maze_dist(const pmo_env *e, int x1, int y1, int x2, int y2) {
    const pmo_dist_table *t = e->dt;
    int a = t->node_of[x1][y1], b = t->node_of[x2][y2];
    if (a < 0 || b < 0) abort();                       /* 'Positions not in grid' */
    int d = t->dist[a * t->n + b];
    return d < 0 ? 65535 : d;                          /* unreachable: sys.maxsize in the reference; only the order matters */
}

/* self.getFood(state): the food this agent's team eats = the OTHER side's half grid */
static int food_to_eat(const pmo_env *s, int red, int *fx, int *fy) {
    int halfway = (int)(s->W / 2.0), n = 0;
    int x0 = red ? halfway : 0, x1 = red ? s->W : halfway;       /* red eats getBlueFood (x >= W//2) */
    for (int x = x0; x < x1; ++x) for (int y = 0; y < s->H; ++y)   /* Grid.asList order: x-major */
        if (s->food[x][y]) { if (fx) { fx[n] = x; fy[n] = y; } n++; }
    return n;
}

static const int REF_ORDER[5] = {A_NORTH, A_SOUTH, A_EAST, A_WEST, A_STOP};   /* game.py:295-299 */
static const int REVERSE[5] = {A_SOUTH, A_WEST, A_NORTH, A_EAST, A_STOP};      /* game.py:60-64   */

static long evaluate_action(const pmo_env *e, int idx, int kind, int action) {
    pmo_env s = *e;
    successor(&s, idx, action);                        /* getSuccessor: positions are always grid points */
    int red = e->team_red[idx];
    const pmo_agent *me = &s.ag[idx];
    if (kind == 2) {                                   /* OffensiveReflexAgent :132-150 */
        static _Thread_local int fx[PMO_MAXW * PMO_MAXH], fy[PMO_MAXW * PMO_MAXH];
        int n = food_to_eat(&s, red, fx, fy);
        long v = 100L * -n;
        if (n > 0) {
            int best = maze_dist(&s, me->x, me->y, fx[0], fy[0]);
            for (int k = 1; k < n; ++k) { int d = maze_dist(&s, me->x, me->y, fx[k], fy[k]); if (d < best) best = d; }
            v += -1L * best;
        }
        return v;
    }
    /* DefensiveReflexAgent :159-187 */
    long v = 0;
    v += 100L * (me->is_pacman ? 0 : 1);
    int ninv = 0, best = 0;
    for (int o = 0; o < 4; ++o) {
        if (s.team_red[o] == red || !s.ag[o].is_pacman) continue;
        int d = maze_dist(&s, me->x, me->y, s.ag[o].x, s.ag[o].y);
        if (ninv == 0 || d < best) best = d;
        ninv++;
    }
    v += -1000L * ninv;
    if (ninv > 0) v += -10L * best;
    if (action == A_STOP) v += -100;
    if (action == REVERSE[e->ag[idx].dir]) v += -2;
    return v;
}

/* ReflexCaptureAgent.chooseAction :65-92 */
static int baseline_choose(const pmo_env *e, int idx, int kind, int *best_mask) {
    if (!e->dt) abort();
    int legal = legal_mask(e, idx), red = e->team_red[idx];
    int acts[5], na = 0;
    for (int k = 0; k < 5; ++k) if ((legal >> REF_ORDER[k]) & 1) acts[na++] = REF_ORDER[k];
    long vals[5], mx = 0;
    for (int k = 0; k < na; ++k) { vals[k] = evaluate_action(e, idx, kind, acts[k]); if (k == 0 || vals[k] > mx) mx = vals[k]; }
    int first = -1, mask = 0;
    for (int k = 0; k < na; ++k) if (vals[k] == mx) { if (first < 0) first = acts[k]; mask |= 1 << acts[k]; }
    if (food_to_eat(e, red, NULL, NULL) <= 0) {        /* nothing left to eat: head for the start cell */
        int bestd = 9999, besta = -1;
        for (int k = 0; k < na; ++k) {
            pmo_env s = *e;
            successor(&s, idx, acts[k]);
            int d = maze_dist(&s, e->ag[idx].sx, e->ag[idx].sy, s.ag[idx].x, s.ag[idx].y);
            if (d < bestd) { besta = acts[k]; bestd = d; }
        }
        if (best_mask) *best_mask = 1 << besta;
        return besta;
    }
    if (best_mask) *best_mask = mask;
    if (e->draws) {                                    /* random.choice(bestActions), baselineTeam.py:92 */
        int best[5], nb = 0;
        for (int k = 0; k < na; ++k) if (vals[k] == mx) best[nb++] = acts[k];
        return best[e->draws[idx] % (uint32_t)nb];
    }
    return first;
}

/* agents/randomTeam.py:94-100: random.choice(gameState.getLegalActions(self.index)); the index comes from
 * e->draws (recorded from the reference's own stream) - without draws the first legal action is taken. */
static int random_choose(const pmo_env *e, int idx) {
    int legal = legal_mask(e, idx), acts[5], na = 0;
    for (int k = 0; k < 5; ++k) if ((legal >> REF_ORDER[k]) & 1) acts[na++] = REF_ORDER[k];
    return acts[e->draws ? e->draws[idx] % (uint32_t)na : 0];
}

void pmo_set_draws(pmo_env *e, const uint32_t *draws4) { e->draws = draws4; }

/* reset outputs: obs[4][8][H][W] (all from the initial state) and legal masks */
void pmo_reset_outputs(const pmo_env *e, float *obs, uint8_t *legal) {
    int HW = e->W * e->H;
    for (int i = 0; i < 4; ++i) {
        if (obs) pmo_observe(e, i, obs + (size_t)i * 8 * HW);
        legal[i] = (uint8_t)legal_mask(e, i);
    }
}

/* state export for comparisons: agents[4][7] = x,y,dir,isPacman,scared,carrying,returned;
 * food/caps as byte-per-cell, cell index c = y*W + x; misc = score, steps, win          */
void pmo_get_state(const pmo_env *e, int32_t *agents, uint8_t *food, uint8_t *caps, int32_t *misc) {
    for (int i = 0; i < 4; ++i) {
        const pmo_agent *a = &e->ag[i];
        int32_t *o = agents + 7 * i;
        o[0] = a->x; o[1] = a->y; o[2] = a->dir; o[3] = a->is_pacman; o[4] = a->scared;
        o[5] = a->carrying; o[6] = a->returned;
    }
    for (int y = 0; y < e->H; ++y) for (int x = 0; x < e->W; ++x) {
        food[y * e->W + x] = e->food[x][y];
        caps[y * e->W + x] = 0;
    }
    for (int k = 0; k < e->ncaps; ++k) caps[e->capy[k] * e->W + e->capx[k]] = 1;
    misc[0] = e->score; misc[1] = e->steps; misc[2] = e->win;
}

/* ---------------------------------------------------------- batched driver */
/* N independent envs (array of pmo_env), optional auto-reset after done.
 * actions[N][4]; reward[N][2]; done[N]; legal[N][4]; score_change[N]; shaping[N][4];
 * obs[N][4][8][H][W] or NULL.  Used for parity at scale and as the CPU baseline. */
void pmo_step_batch(pmo_env *envs, int64_t N, const int32_t *actions, float *obs, double *reward,
                    uint8_t *done, uint8_t *legal, int32_t *score_change, double *shaping,
                    int auto_reset, int nthreads) {
    if (N <= 0) return;
    size_t per = (size_t)4 * 8 * envs[0].W * envs[0].H;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t n = 0; n < N; ++n) {
        pmo_step(&envs[n], actions + 4 * n, obs ? obs + per * n : NULL, reward + 2 * n, done + n,
                 legal + 4 * n, score_change + n, shaping ? shaping + 4 * n : NULL);
        if (auto_reset && done[n]) pmo_reset(&envs[n]);
    }
}

void pmo_step_batch_bots(pmo_env *envs, int64_t N, const int32_t *actions, const uint8_t *bot_kind, float *obs,
                         double *reward, uint8_t *done, uint8_t *legal, int32_t *score_change, double *shaping,
                         int32_t *actions_used, uint8_t *best_masks, int auto_reset, int nthreads) {
    if (N <= 0) return;
    size_t per = (size_t)4 * 8 * envs[0].W * envs[0].H;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
    for (int64_t n = 0; n < N; ++n) {
        pmo_step_bots(&envs[n], actions + 4 * n, bot_kind, obs ? obs + per * n : NULL, reward + 2 * n, done + n,
                      legal + 4 * n, score_change + n, shaping ? shaping + 4 * n : NULL,
                      actions_used ? actions_used + 4 * n : NULL, best_masks ? best_masks + 4 * n : NULL);
        if (auto_reset && done[n]) pmo_reset(&envs[n]);
    }
}

/* Observation digests for parity at full batch size (65,536 envs x 4 observations per step never cross PCIe):
 * digest(o) = sum_i bits(o[i]) * K[i] mod 2^64 over the 8*H*W fp32 elements, K = a table of pseudo-random 64-bit
 * weights supplied by the caller (position-dependent and non-linear in i, so moves of different agents cannot
 * cancel).  The GPU side computes the same sum with integer tensor ops on the kernel's output
 * (tests/test_gpu_full_size.py).  With auto_reset the digests (and legal masks) of a finished env are those of its
 * freshly reset state, like the fused restart of pm_step.  digest[N][4].                                        */
static uint64_t obs_digest(const float *o, int n, const uint64_t *K) {
    uint64_t d = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t u; memcpy(&u, &o[i], 4);
        if (u) d += (uint64_t)u * K[i];
    }
    return d;
}

void pmo_step_batch_digest(pmo_env *envs, int64_t N, const int32_t *actions, const uint64_t *K, uint64_t *digest,
                           double *reward, uint8_t *done, uint8_t *legal, int32_t *score_change, double *shaping,
                           int auto_reset, int nthreads) {
    if (N <= 0) return;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        /* one scratch block per worker thread, kept for the life of the thread */
        static _Thread_local float *obs = NULL;
        if (!obs) obs = (float *)malloc(sizeof(float) * 4 * 8 * PMO_MAXW * PMO_MAXH);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int64_t n = 0; n < N; ++n) {
            pmo_env *e = &envs[n];
            int per = 8 * e->W * e->H;
            pmo_step(e, actions + 4 * n, obs, reward + 2 * n, done + n, legal + 4 * n, score_change + n,
                     shaping ? shaping + 4 * n : NULL);
            if (auto_reset && done[n]) { pmo_reset(e); pmo_reset_outputs(e, obs, legal + 4 * n); }
            for (int i = 0; i < 4; ++i) digest[4 * n + i] = obs_digest(obs + (size_t)i * per, per, K);
        }
    }
}

int pmo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------ maze distances */
/* distanceCalculator.py:111-150 with util.PriorityQueue (heap of (priority, count, item)).
 * nodes = walls.asList(False): x-major, y-minor.  out[n][n] int32, out[i][j] =
 * distances[(nodes[i], nodes[j])] = dist from source j to target i; unreachable = -1
 * (sys.maxsize in the reference).  Returns n_open, or -1 if out_cap is too small.   */
typedef struct { int pri; long cnt; int node; } pq_item;
static int pq_less(const pq_item *a, const pq_item *b) {
    return a->pri < b->pri || (a->pri == b->pri && a->cnt < b->cnt);
}
int pmo_distances(const pmo_env *e, int32_t *out, int out_cap, int16_t *nodes_xy) {
    int W = e->W, H = e->H, n = 0;
    static _Thread_local int node_of[PMO_MAXW][PMO_MAXH];
    int nx[PMO_MAXW * PMO_MAXH], ny[PMO_MAXW * PMO_MAXH];
    for (int x = 0; x < W; ++x) for (int y = 0; y < H; ++y) {
        node_of[x][y] = -1;
        if (!e->walls[x][y]) { node_of[x][y] = n; nx[n] = x; ny[n] = y; n++; }
    }
    if (n > out_cap) return -1;
    if (nodes_xy) for (int i = 0; i < n; ++i) { nodes_xy[2 * i] = (int16_t)nx[i]; nodes_xy[2 * i + 1] = (int16_t)ny[i]; }
    long *dist = (long *)malloc(sizeof(long) * n);
    uint8_t *closed = (uint8_t *)malloc((size_t)n + 1);
    pq_item *heap = (pq_item *)malloc(sizeof(pq_item) * (4 * n + 8));
    static const int ADJX[4] = {0, 0, 1, -1}, ADJY[4] = {1, -1, 0, 0};
    for (int src = 0; src < n; ++src) {
        for (int i = 0; i < n; ++i) { dist[i] = -1; closed[i] = 0; }   /* -1 == sys.maxsize */
        int hn = 0; long count = 0;
        heap[hn++] = (pq_item){0, count++, src};
        dist[src] = 0;
        while (hn > 0) {
            pq_item top = heap[0];
            heap[0] = heap[--hn];
            for (int i = 0;;) {                       /* sift down */
                int l = 2 * i + 1, r = l + 1, m = i;
                if (l < hn && pq_less(&heap[l], &heap[m])) m = l;
                if (r < hn && pq_less(&heap[r], &heap[m])) m = r;
                if (m == i) break;
                pq_item t = heap[i]; heap[i] = heap[m]; heap[m] = t; i = m;
            }
            int node = top.node;
            if (closed[node]) continue;
            closed[node] = 1;
            long nd = dist[node];
            for (int k = 0; k < 4; ++k) {
                int ax = nx[node] + ADJX[k], ay = ny[node] + ADJY[k];
                if (ax < 0 || ay < 0 || ax >= W || ay >= H) continue;
                if (e->walls[ax][ay]) continue;
                int o = node_of[ax][ay];
                long old = dist[o], nw = nd + 1;
                if (old < 0 || nw < old) {
                    dist[o] = nw;
                    heap[hn] = (pq_item){(int)nw, count++, o};
                    for (int i = hn++; i > 0;) {      /* sift up */
                        int p = (i - 1) / 2;
                        if (!pq_less(&heap[i], &heap[p])) break;
                        pq_item t = heap[i]; heap[i] = heap[p]; heap[p] = t; i = p;
                    }
                }
            }
        }
        for (int tgt = 0; tgt < n; ++tgt) out[(size_t)tgt * n + src] = (int32_t)dist[tgt];
    }
    free(dist); free(closed); free(heap);
    return n;
}

/* ---------------------------------------------------------------- GAE etc. */
/* pacman_mappo_resnet.py:277-290, one column at a time, with torch's fp32
 * semantics: a Python float times a 0-dim fp32 tensor is computed in fp32 with
 * the scalar rounded to fp32 first; Python-float * Python-float stays double.
 * rew/val/done/adv/ret are [T][M] row-major; last_value[M].                   */
void pmo_gae(const float *rew, const float *val, const float *done, const float *last_value,
             double gamma, double lam, float *adv, float *ret, int T, int M) {
    float gamma_f = (float)gamma;
    float gl_f = (float)(gamma * lam);
    for (int m = 0; m < M; ++m) {
        float lastgaelam = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            float nonterm = 1.0f - done[(size_t)t * M + m];
            float gv;
            if (t == T - 1) gv = (float)(gamma * (double)last_value[m]);   /* float*float in Python */
            else gv = gamma_f * val[(size_t)(t + 1) * M + m];
            float delta = rew[(size_t)t * M + m] + gv * nonterm - val[(size_t)t * M + m];
            lastgaelam = delta + (gl_f * nonterm) * lastgaelam;
            adv[(size_t)t * M + m] = lastgaelam;
        }
        for (int t = 0; t < T; ++t) ret[(size_t)t * M + m] = adv[(size_t)t * M + m] + val[(size_t)t * M + m];
    }
}

/* pacman_mappo_resnet.py:577 per minibatch: (a - mean) / (std_unbiased + 1e-8).
 * (double accumulation; the reference sums in fp32 with torch's own reduction
 * order, so this function is compared with a tolerance, not bit-exactly)       */
void pmo_adv_norm(const float *adv, const int64_t *inds, float *out, int n, int mb) {
    for (int s = 0; s < n; s += mb) {
        int k = n - s < mb ? n - s : mb;
        double sum = 0, sq = 0;
        for (int i = 0; i < k; ++i) sum += adv[inds[s + i]];
        double mean = sum / k;
        for (int i = 0; i < k; ++i) { double d = adv[inds[s + i]] - mean; sq += d * d; }
        float mean_f = (float)mean;
        float std_f = (float)sqrt(sq / (k - 1));
        for (int i = 0; i < k; ++i) out[s + i] = (adv[inds[s + i]] - mean_f) / (std_f + 1e-8f);
    }
}

/* ---------------------------------------------------------------- bot distance tables */
int pmo_distances(const pmo_env *e, int32_t *out, int out_cap, int16_t *nodes_xy);
size_t pmo_sizeof_dist_table(void) { return sizeof(pmo_dist_table); }
/* Fill `t` (caller-allocated, pmo_sizeof_dist_table bytes) for env e's board; `buf` holds n*n int32 and must
 * outlive every env that points at the table. */
int pmo_build_dist_table(const pmo_env *e, pmo_dist_table *t, int32_t *buf, int cap) {
    int16_t *nodes = (int16_t *)malloc(sizeof(int16_t) * 2 * PMO_MAXW * PMO_MAXH);
    int n = pmo_distances(e, buf, cap, nodes);
    if (n < 0) { free(nodes); return -1; }
    for (int x = 0; x < PMO_MAXW; ++x) for (int y = 0; y < PMO_MAXH; ++y) t->node_of[x][y] = -1;
    for (int k = 0; k < n; ++k) t->node_of[nodes[2 * k]][nodes[2 * k + 1]] = (int16_t)k;
    t->n = n; t->dist = buf;
    free(nodes);
    return n;
}
void pmo_set_distances(pmo_env *e, const pmo_dist_table *t) { e->dt = t; }
