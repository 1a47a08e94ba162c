#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz|json

Everything here *imports and executes the unmodified reference*
(`gymPacMan.gymPacMan_parallel_env`, `distanceCalculator.computeDistances`,
`pacman_mappo_resnet.compute_gae / compute_heuristic_shaping / canonicalize_obs`,
`mazeGenerator.generateMaze`) and records inputs + outputs.  The oracle
(`oracle/`) and the CUDA path are then tested against these files.

Fixture kinds
  traj_<name>.npz   action sequences + per-step full state / rewards / done /
                    legal masks / obs digests (+ full obs for the first steps)
  dist_<name>.npz   all-pairs maze distance matrix in reference node order
  gae.npz           compute_gae + minibatch advantage normalisation vectors
  mazes.json        mazeGenerator text for a list of seeds
  digests.json      the SURVEY A.4 trajectory / distance digests, regenerated
"""
import contextlib
import hashlib
import io
import json
import os
import random
import struct
import sys
import tempfile
import types

import numpy as np

REF = os.environ.get("PM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
import torch  # noqa: E402

torch.set_num_threads(1)

# the trainer imports matplotlib at module top; it is not installed here
_mpl = types.ModuleType("matplotlib")
_plt = types.ModuleType("matplotlib.pyplot")
_mpl.pyplot = _plt
sys.modules.setdefault("matplotlib", _mpl)
sys.modules.setdefault("matplotlib.pyplot", _plt)

_quiet = contextlib.redirect_stdout(io.StringIO())
with _quiet:
    import capture  # noqa: E402
    import distanceCalculator  # noqa: E402
    import layout as ref_layout  # noqa: E402
    import mazeGenerator as ref_maze  # noqa: E402
    import pacman_mappo_resnet as ref_trainer  # noqa: E402
    from gymPacMan import gymPacMan_parallel_env  # noqa: E402

# coverage probe only (does not change behaviour): count consume() calls made
# for a mover that is NOT a pacman, i.e. the `capture.py:505 -> :514` rebinding
_REBIND = {"n": 0, "ate": 0}
_orig_consume = capture.AgentRules.consume


def _probe_consume(position, state, isRed):
    movers = [a for a in state.data.agentStates if a.getPosition() == position]
    if movers and not any(a.isPacman for a in movers):
        _REBIND["n"] += 1
        if state.data.food[position[0]][position[1]] or position in state.data.capsules:
            _REBIND["ate"] += 1
    return _orig_consume(position, state, isRed)


capture.AgentRules.consume = staticmethod(_probe_consume)

DIR_ID = {"North": 0, "East": 1, "South": 2, "West": 3, "Stop": 4}
ACT_VEC = {0: (0, 1), 1: (1, 0), 2: (0, -1), 3: (-1, 0), 4: (0, 0)}

# ---------------------------------------------------------------------------
# crafted layouts (the three shipped ones have no capsules)
# ---------------------------------------------------------------------------
CRAFTED = {
    # smallCapture with capsules, one of them on the x == W/2 column
    "capsSmall": [
        "%%%%%%%%%%%%%%",
        "%   o3%%2    %",
        "% %%%%%%%%%% %",
        "%....  o ....%",
        "%.%%.%%%%.%%.%",
        "%. .  o   . .%",
        "%.%%.%%%%.%%.%",
        "%....    ....%",
        "% %%%%%%%%%% %",
        "%    1%%4o   %",
        "%%%%%%%%%%%%%%",
    ],
    # odd width: W/2 = 7.5, int(W/2) = 7 -> the three half formulas disagree at x = 7
    "odd15": [
        "%%%%%%%%%%%%%%%",
        "%1 ..  . o .. %",
        "% %%.%.%.%.%%2%",
        "%3 .o  .   .  %",
        "% %%.%.%.%.%% %",
        "%  ..  .  o..4%",
        "%%%%%%%%%%%%%%%",
    ],
    # two pellets per side: episodes end by food exhaustion
    "mini": [
        "%%%%%%%%%%",
        "%1 .  . 2%",
        "%  %  %  %",
        "%3 .  . 4%",
        "%%%%%%%%%%",
    ],
    # open arena with many pellets + capsules: lots of kills / dumps near borders
    "arena": [
        "%%%%%%%%%%%%",
        "%1...o.....%",
        "%..........%",
        "%3........2%",
        "%..........%",
        "%.....o...4%",
        "%%%%%%%%%%%%",
    ],
    # pellets on both boundary columns: agents returning food step onto own
    # pellets while agent 3 raids -> exercises the agentState rebinding quirk
    "rebind": [
        "%%%%%%%%%%%%",
        "%1  ....  2%",
        "%   .%%.   %",
        "%  ......  %",
        "%   .%%.   %",
        "%3  ....  4%",
        "%%%%%%%%%%%%",
    ],
    # digits 1/3 start on the right: capture.py teams (by start column) differ
    # from gymPacMan's hard-coded red = {0, 2}
    "swapped": [
        "%%%%%%%%%%%%",
        "%2 ..  .. 1%",
        "%  %o  o%  %",
        "%4 ..  .. 3%",
        "%%%%%%%%%%%%",
    ],
}


def layout_text(name):
    """-> (list of text rows, tag).  name: shipped layout, crafted name or maze<seed>."""
    if name in CRAFTED:
        return list(CRAFTED[name])
    if name.startswith("maze"):
        with _quiet:
            return ref_maze.generateMaze(int(name[4:])).split("\n")
    with open(os.path.join(REF, "layouts", name + ".lay")) as f:
        return [ln.strip() for ln in f]


def make_env(rows, enemie=None, **kw):
    """enemie=None: self-play (all four agents driven by the recorded actions).  enemie='baselineTeam': the red
    side (agents 0, 2) is played by the reference's bots (`gymPacMan.py:33-40, 157`)."""
    fd, path = tempfile.mkstemp(suffix=".lay")
    with os.fdopen(fd, "w") as f:
        f.write("\n".join(rows))
    try:
        with _quiet:
            cwd = os.getcwd()
            os.chdir(REF)                 # loadAgents resolves agents/<name>.py next to gymPacMan.py
            try:
                env = gymPacMan_parallel_env(layout_file=path, enemieName=enemie or "randomTeam",
                                             self_play=enemie is None, **kw)
            finally:
                os.chdir(cwd)
    finally:
        os.unlink(path)
    return env


class BotSpy:
    """Records what the reference's bots decide.  `random.choice` (the tie-break of
    `baselineTeam.py:92`) is pinned to the FIRST best action while recording, and the set of best
    actions it was offered is kept, so a replay can either follow the same trajectory (tie-break =
    first) or check that a randomly tie-broken choice lies inside the recorded set."""

    def __init__(self, env, real_rng=False):
        """real_rng=True: `random.choice` is NOT pinned - it draws from the global MT19937 exactly like the stock
        function (one `_randbelow(len(seq))`), and the index it returned is recorded per bot decision, so a replay
        can be fed the reference's own draws (`pm_env_desc.bot_draws`)."""
        self.env = env
        self.real_rng = real_rng
        self.choice_sets = []
        self.choice_idx = []
        self.taken = {}
        self.draw = {}
        self._orig_choice = random.choice
        random.choice = self._choice
        for i, ag in enumerate(env.agents):
            if not isinstance(ag, int):
                ag.getAction = self._wrap(i, ag.getAction)

    def _choice(self, seq):
        self.choice_sets.append(list(seq))
        if not self.real_rng:
            self.choice_idx.append(0)
            return seq[0]
        i = random.randrange(len(seq))          # == random.choice: seq[self._randbelow(len(seq))]
        self.choice_idx.append(i)
        return seq[i]

    def _wrap(self, i, fn):
        def inner(state):
            self.choice_sets, self.choice_idx = [], []
            a = fn(state)
            m = 0
            for name in (self.choice_sets[-1] if self.choice_sets else [a]):
                m |= 1 << DIR_ID[name]
            self.taken[i] = (DIR_ID[a], m)
            self.draw[i] = self.choice_idx[-1] if self.choice_idx else 0
            return a
        return inner

    def close(self):
        random.choice = self._orig_choice


# ---------------------------------------------------------------------------
# state extraction
# ---------------------------------------------------------------------------
def grid_bits(grid):
    """Grid[x][y] -> uint8 bit string, cell index c = y*W + x, little bit order."""
    a = np.array(grid.data, dtype=bool)          # [W, H]
    return np.packbits(a.T.reshape(-1), bitorder="little")


def caps_bits(caps, W, H):
    a = np.zeros(W * H, dtype=bool)
    for (x, y) in caps:
        a[y * W + x] = True
    return np.packbits(a, bitorder="little")


def obs_digest(t):
    b = t.numpy().astype("<f4").tobytes()
    return np.frombuffer(hashlib.blake2b(b, digest_size=8).digest(), dtype="<u8")[0]


def legal_mask(lst):
    m = 0
    for k in lst:
        m |= 1 << k
    return m


def snapshot(env):
    d = env.game.state.data
    pos = np.zeros((4, 2), np.int16)
    misc = np.zeros((4, 5), np.int16)   # dir, isPacman, scared, carrying, returned
    for i, a in enumerate(d.agentStates):
        x, y = a.configuration.pos
        assert x == int(x) and y == int(y)
        pos[i] = (int(x), int(y))
        misc[i] = (DIR_ID[a.configuration.direction], int(a.isPacman), a.scaredTimer,
                   a.numCarrying, a.numReturned)
    W, H = d.layout.width, d.layout.height
    return pos, misc, grid_bits(d.food), caps_bits(d.capsules, W, H), int(d.score)


def shaping_all(prev_obs, next_obs):
    """compute_heuristic_shaping for every agent, on canonicalised views exactly
    as the trainer does for its learner side (`pacman_mappo_resnet.py:462-516`)."""
    out = np.zeros(4, np.float64)
    for i in range(4):
        red = i in (0, 2)
        a = ref_trainer.canonicalize_obs(prev_obs[i].float(), red)
        b = ref_trainer.canonicalize_obs(next_obs[i].float(), red)
        out[i] = ref_trainer.compute_heuristic_shaping(a, b)
    return out


# ---------------------------------------------------------------------------
# action policies
# ---------------------------------------------------------------------------
class SmartPolicy:
    """Scripted attackers / defenders so that eating, returning, capsules,
    kills, scared kills, dumps and the rebinding quirk all occur.  Only the
    *recorded actions* matter for parity; this is just a generator."""

    def __init__(self, env, rng, eps=0.15):
        self.rng = rng
        self.eps = eps
        lay = env.layout
        self.W, self.H = lay.width, lay.height
        with _quiet:
            self.dist = distanceCalculator.computeDistances(lay)
        self.new_episode()

    def new_episode(self):
        r = self.rng
        self.role = [("A" if r.random() < 0.65 else "D") for _ in range(4)]
        self.thresh = [int(r.integers(1, 7)) for _ in range(4)]

    def md(self, a, b):
        return self.dist.get((a, b), 10 ** 6)

    def act(self, env):
        st = env.game.state
        d = st.data
        acts = []
        for i in range(4):
            if self.rng.random() < self.eps:
                acts.append(int(self.rng.integers(0, 5)))
                continue
            me = d.agentStates[i]
            pos = tuple(int(v) for v in me.configuration.pos)
            red = st.isOnRedTeam(i)
            enemies = st.getBlueTeamIndices() if red else st.getRedTeamIndices()
            target = None
            if me.numCarrying >= self.thresh[i]:
                target = tuple(int(v) for v in me.start.pos)
            elif self.role[i] == "D":
                inv = [tuple(int(v) for v in d.agentStates[e].configuration.pos)
                       for e in enemies if d.agentStates[e].isPacman]
                if inv:
                    target = min(inv, key=lambda p: self.md(pos, p))
            if target is None:
                food = (st.getBlueFood() if red else st.getRedFood()).asList()
                caps = st.getBlueCapsules() if red else st.getRedCapsules()
                goals = food + list(caps)
                if goals:
                    target = min(goals, key=lambda p: self.md(pos, p))
                else:
                    target = tuple(int(v) for v in me.start.pos)
            legal = [env.reversed_action_mapping[a] for a in st.getLegalActions(i)]
            best, bestd = 4, None
            for a in legal:
                dx, dy = ACT_VEC[a]
                nd = self.md((pos[0] + dx, pos[1] + dy), target)
                if bestd is None or nd < bestd:
                    best, bestd = a, nd
            acts.append(best)
        return acts


class GlobalRngBluePolicy:
    """BASELINE config 1's caller: the red side is played by the env's bots, blue picks uniformly from the legal
    lists of the last info dict with the GLOBAL `random` (what randomTeam.chooseAction does, randomTeam.py:94-100)."""

    def __init__(self, env, rng):
        self.info = None

    def new_episode(self):
        pass

    def act(self, env):
        la = self.info["legal_actions"]
        return [0, random.choice(la[1]), 0, random.choice(la[3])]


class RandomPolicy:
    def __init__(self, env, rng):
        self.rng = rng

    def new_episode(self):
        pass

    def act(self, env):
        return [int(v) for v in self.rng.integers(0, 5, size=4)]


# ---------------------------------------------------------------------------
# trajectory recorder
# ---------------------------------------------------------------------------
def record(name, lay_name, policy="smart", episodes=4, seed=0, full_obs_steps=48, max_steps=None, enemie=None,
           real_rng=False, **envkw):
    rows = layout_text(lay_name)
    kw = dict(length=300, reward_forLegalAction=True, defenceReward=True)
    kw.update(envkw)
    if real_rng:
        random.seed(seed)                          # BASELINE config 1: random.seed(1)
    env = make_env(rows, enemie=enemie, **kw)
    spy = BotSpy(env, real_rng=real_rng) if enemie else None
    byidx = lambda d: {i: d[env.agents[i]] for i in range(4)}      # bot objects key the dicts of the red side
    rng = np.random.default_rng(seed)
    pol = (SmartPolicy(env, rng) if policy == "smart" else GlobalRngBluePolicy(env, rng) if policy == "global_rng"
           else RandomPolicy(env, rng))
    W, H = env.layout.width, env.layout.height

    R = dict(actions=[], pos=[], misc=[], food=[], caps=[], score=[], score_change=[], reward=[],
             done=[], legal=[], obs_dig=[], shaping=[], ep_start=[], reset_legal=[], reset_obs_dig=[],
             full_obs=[], bot_best=[], draws=[])
    cover = dict(steps=0, eats=0, returns=0, kills=0, dumps=0, scared_kills=0, caps=0, food_end=0,
                 rebinding=0, nonzero_final=0)
    _REBIND["n"] = _REBIND["ate"] = 0
    for ep in range(episodes):
        pol.new_episode()
        with _quiet:
            obs, info = env.reset()
        obs, info["legal_actions"] = byidx(obs), byidx(info["legal_actions"])
        pol.info = info
        R["ep_start"].append(len(R["actions"]))
        R["reset_legal"].append([legal_mask(info["legal_actions"][i]) for i in range(4)])
        R["reset_obs_dig"].append([obs_digest(obs[i]) for i in range(4)])
        done = False
        while not done:
            acts = pol.act(env)
            pre = snapshot(env)
            with _quiet:
                nobs, rew, term, info = env.step({i: acts[i] for i in range(4)})
            nobs, rew, info["legal_actions"] = byidx(nobs), byidx(rew), byidx(info["legal_actions"])
            pol.info = info
            best, draws = [0, 0, 0, 0], [0, 0, 0, 0]
            if spy:                                   # what the bots really played replaces the ignored red entries
                acts = list(acts)
                for i, (a, m) in spy.taken.items():
                    acts[i], best[i], draws[i] = a, m, spy.draw.get(i, 0)
            R["bot_best"].append(best)
            R["draws"].append(draws)
            done = any(term.values())
            pos, misc, food, caps, score = snapshot(env)
            R["actions"].append(acts)
            R["pos"].append(pos); R["misc"].append(misc); R["food"].append(food); R["caps"].append(caps)
            R["score"].append(score); R["score_change"].append(info["score_change"])
            R["reward"].append([rew[0], rew[1]])
            assert rew[0] == rew[2] and rew[1] == rew[3]
            R["done"].append(done)
            R["legal"].append([legal_mask(info["legal_actions"][i]) for i in range(4)])
            R["obs_dig"].append([obs_digest(nobs[i]) for i in range(4)])
            R["shaping"].append(shaping_all(obs, nobs))
            if len(R["full_obs"]) < full_obs_steps:
                R["full_obs"].append(np.stack([nobs[i].numpy() for i in range(4)]).astype(np.float32))
            # coverage bookkeeping (informational)
            cover["steps"] += 1
            pm, m = pre[1], misc
            cover["eats"] += int((m[:, 3] > pm[:, 3]).sum())
            cover["returns"] += int((m[:, 4] > pm[:, 4]).sum())
            died = (pm[:, 1] == 1) & (m[:, 1] == 0) & np.array(
                [tuple(pos[i]) == tuple(int(v) for v in env.game.state.data.agentStates[i].start.pos) for i in range(4)])
            cover["kills"] += int(died.sum())
            cover["dumps"] += int((died & (pm[:, 3] > 0)).sum())
            cover["caps"] += int(np.unpackbits(pre[3]).sum() - np.unpackbits(caps).sum())
            cover["scared_kills"] += int(((pm[:, 2] > 1) & (m[:, 2] == 0)).sum())
            if done and env.steps <= env.length:
                cover["food_end"] += 1
            if done and score != 0:
                cover["nonzero_final"] += 1
            obs = nobs
            if max_steps and cover["steps"] >= max_steps:
                break
    # every sub-move is executed twice by the reference (peek in get_reward)
    cover["rebinding"] = _REBIND["n"] // 2
    cover["rebinding_ate"] = _REBIND["ate"] // 2
    out = {k: np.array(v) for k, v in R.items()}
    out["actions"] = out["actions"].astype(np.uint8)
    out["reward"] = out["reward"].astype(np.float64)
    out["obs_dig"] = out["obs_dig"].astype(np.uint64)
    out["reset_obs_dig"] = out["reset_obs_dig"].astype(np.uint64)
    out["legal"] = out["legal"].astype(np.uint8)
    out["reset_legal"] = out["reset_legal"].astype(np.uint8)
    out["done"] = out["done"].astype(np.uint8)
    out["score"] = out["score"].astype(np.int32)
    out["score_change"] = out["score_change"].astype(np.int32)
    out["layout"] = np.array(rows)
    out["length"] = np.int32(kw["length"])
    out["reward_forLegalAction"] = np.uint8(kw["reward_forLegalAction"])
    out["defenceReward"] = np.uint8(kw["defenceReward"])
    out["WH"] = np.array([W, H], np.int32)
    out["bot_best"] = out["bot_best"].astype(np.uint8)
    # 0 = driven by `actions`, 2 / 3 = baselineTeam Offensive / Defensive reflex agent (createTeam order, baselineTeam.py:34-51)
    out["bot_kind"] = np.array([2, 0, 3, 0] if enemie == "baselineTeam" else [1, 0, 1, 0] if enemie == "randomTeam"
                               else [0, 0, 0, 0], np.uint8)
    out["draws"] = out["draws"].astype(np.uint32)
    out["real_rng"] = np.uint8(real_rng)
    if spy:
        spy.close()
    np.savez_compressed(os.path.join(HERE, f"traj_{name}.npz"), **out)
    print(f"traj_{name}: {cover}", file=sys.stderr)
    return cover


# ---------------------------------------------------------------------------
# SURVEY A.4 digests
# ---------------------------------------------------------------------------
def survey_digest(lay_name):
    with _quiet:
        env = gymPacMan_parallel_env(layout_file=os.path.join(REF, "layouts", lay_name + ".lay"),
                                     enemieName="randomTeam", self_play=True, length=300)
    rng = np.random.default_rng(12345)
    h = hashlib.sha256()
    for _ in range(10):
        with _quiet:
            env.reset()
        done = False
        while not done:
            a = rng.integers(0, 5, size=4)
            with _quiet:
                obs, rew, term, info = env.step({i: int(a[i]) for i in range(4)})
            for i in range(4):
                h.update(obs[i].numpy().astype("<f4").tobytes())
                h.update(struct.pack("<d", rew[i]))
                h.update(bytes([legal_mask(info["legal_actions"][i])]))
            done = any(term.values())
            h.update(struct.pack("<i?", env.game.state.data.score, done))
    return h.hexdigest()


def distance_fixture(name, lay_name):
    rows = layout_text(lay_name)
    lay = ref_layout.Layout(rows)
    with _quiet:
        dist = distanceCalculator.computeDistances(lay)
    nodes = lay.walls.asList(False)
    n = len(nodes)
    M = np.zeros((n, n), np.int64)
    for i, a in enumerate(nodes):
        for j, b in enumerate(nodes):
            M[i, j] = dist[(a, b)]
    sha = hashlib.sha256(np.where(M == sys.maxsize, np.iinfo(np.int32).max, M).astype("<i4").tobytes()).hexdigest()
    M32 = np.where(M == sys.maxsize, -1, M).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, f"dist_{name}.npz"), layout=np.array(rows),
                        nodes=np.array(nodes, np.int16), dist=M32)
    return dict(n_open=n, max=int(M32.max()), sum=int(M32[M32 >= 0].sum()), sha256=sha,
                unreachable=int((M32 < 0).sum()))


def gae_fixture():
    g = torch.Generator().manual_seed(1234)
    out = {}
    cases = [(16, 3), (301, 4), (2048, 2), (5, 2), (1, 2)]
    for ci, (T, M) in enumerate(cases):
        r = torch.randn(T, M, generator=g)
        v = torch.randn(T, M, generator=g)
        d = (torch.rand(T, M, generator=g) < 0.05).float()
        d[:, 1] = (torch.rand(T, generator=g) < 0.5).float()
        lv = torch.randn(M, generator=g)
        adv = torch.zeros(T, M)
        ret = torch.zeros(T, M)
        for m in range(M):
            a, rt = ref_trainer.compute_gae(r[:, m], v[:, m], d[:, m], lv[m].item(), ref_trainer.GAMMA)
            adv[:, m], ret[:, m] = a, rt
        out[f"c{ci}_rew"], out[f"c{ci}_val"], out[f"c{ci}_done"] = r.numpy(), v.numpy(), d.numpy()
        out[f"c{ci}_last"], out[f"c{ci}_adv"], out[f"c{ci}_ret"] = lv.numpy(), adv.numpy(), ret.numpy()
    out["n_cases"] = np.int32(len(cases))
    out["gamma"], out["lam"] = np.float64(ref_trainer.GAMMA), np.float64(ref_trainer.GAE_LAMBDA)
    # SURVEY A.4 KAT
    torch.manual_seed(0)
    r = torch.randn(16); v = torch.randn(16); d = (torch.rand(16) < 0.2).float()
    a, rt = ref_trainer.compute_gae(r, v, d, 0.37, 0.99)
    out["kat_rew"], out["kat_val"], out["kat_done"] = r.numpy(), v.numpy(), d.numpy()
    out["kat_adv"], out["kat_ret"] = a.numpy(), rt.numpy()
    # minibatch advantage normalisation, `pacman_mappo_resnet.py:577`
    b_adv = torch.randn(4096, generator=g) * 3 + 0.5
    inds = torch.randperm(4096, generator=g).numpy()
    norm = np.zeros(4096, np.float32)
    for s in range(0, 4096, ref_trainer.BATCH_SIZE):
        mb = inds[s:s + ref_trainer.BATCH_SIZE]
        norm[s:s + ref_trainer.BATCH_SIZE] = ((b_adv[mb] - b_adv[mb].mean()) / (b_adv[mb].std() + 1e-8)).numpy()
    out["norm_adv"], out["norm_inds"], out["norm_out"] = b_adv.numpy(), inds.astype(np.int64), norm
    out["norm_mb"] = np.int32(ref_trainer.BATCH_SIZE)
    np.savez_compressed(os.path.join(HERE, "gae.npz"), **out)


def partial_fixture(lay_name, seed, episodes):
    """GameState.makeObservation (capture.py:268-292) of the live reference at every step of scripted episodes: which
    configurations survive in each agent's view and the noisy sonar distances, with the index every
    random.choice(SONAR_NOISE_VALUES) returned (real global MT19937) so that a replay can be fed the same noise."""
    rows = layout_text(lay_name)
    env = make_env(rows, length=120)
    rng = np.random.default_rng(seed)
    pol = SmartPolicy(env, rng)
    random.seed(seed)
    picks = []
    orig = random.choice

    def choice(seq):
        i = random.randrange(len(seq))
        picks.append(i)
        return seq[i]
    R = dict(actions=[], ep_start=[], visible=[], noisy=[], draws=[])
    for ep in range(episodes):
        pol.new_episode()
        with _quiet:
            env.reset()
        R["ep_start"].append(len(R["actions"]))
        done = False
        while not done:
            acts = pol.act(env)
            with _quiet:
                _, _, term, _ = env.step({i: acts[i] for i in range(4)})
            done = any(term.values())
            st = env.game.state
            vis, noisy, draws = [], [], []
            random.choice = choice
            try:
                for i in range(4):
                    del picks[:]
                    o = st.makeObservation(i)
                    assert len(picks) == 4 and list(capture.SONAR_NOISE_VALUES) == [k - 6 for k in range(13)]
                    vis.append(sum(1 << e for e in range(4) if o.data.agentStates[e].configuration is not None))
                    noisy.append(list(o.getAgentDistances()))
                    draws.append(list(picks))
            finally:
                random.choice = orig
            R["actions"].append(acts); R["visible"].append(vis); R["noisy"].append(noisy); R["draws"].append(draws)
    np.savez_compressed(os.path.join(HERE, f"partial_{lay_name}.npz"), layout=np.array(rows), length=np.int32(120),
                        actions=np.array(R["actions"], np.uint8), ep_start=np.array(R["ep_start"], np.int32),
                        visible=np.array(R["visible"], np.uint8), noisy=np.array(R["noisy"], np.int32),
                        draws=np.array(R["draws"], np.uint32))
    v = np.array(R["visible"])
    print(f"partial_{lay_name}: {len(R['actions'])} steps, hidden enemies in {int((v != 15).sum())} of {v.size} views", file=sys.stderr)


def replay_fixture():
    """A whole game played by the reference's own tournament loop (`Game.run`, game.py:570-732: baselineTeam red vs
    randomTeam blue, partial observations, global RNG) and recorded the way `capture.runGames(record=True)` does
    (`capture.py:988-996`: pickle of layout / agents / moveHistory / length / team names), plus the state after every
    move obtained with `capture.replayGame`'s own loop (`capture.py:952-968`)."""
    import pickle
    import game as ref_game
    import textDisplay
    random.seed(77)
    rows = layout_text("smallCapture")
    lay = ref_layout.Layout(list(rows))
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        with _quiet:
            red = capture.loadAgents(True, os.path.join(REF, "agents", "baselineTeam"), None, "")
            blue = capture.loadAgents(False, os.path.join(REF, "agents", "randomTeam"), None, "")
    finally:
        os.chdir(cwd)
    agents = [red[0], blue[0], red[1], blue[1]]
    length = 400
    rules = capture.CaptureRules()
    with _quiet:
        g = rules.newGame(lay, agents, textDisplay.NullGraphics(), length, True, False)
        g.run()
    comp = {"layout": lay, "agents": [ref_game.Agent() for a in agents], "actions": g.moveHistory, "length": length,
            "redTeamName": "baselineTeam", "blueTeamName": "randomTeam"}
    with open(os.path.join(HERE, "replay_smallCapture.pkl"), "wb") as f:
        f.write(pickle.dumps(comp))
    # the states capture.replayGame walks through
    with _quiet:
        g2 = capture.CaptureRules().newGame(lay, agents, textDisplay.NullGraphics(), length, False, False)
    state = g2.state
    pos, misc, food, score = [], [], [], []
    for action in g.moveHistory:
        state = state.generateSuccessor(*action)
        d = state.data
        pos.append([[int(a.configuration.pos[0]), int(a.configuration.pos[1])] for a in d.agentStates])
        misc.append([[DIR_ID[a.configuration.direction], int(a.isPacman), a.scaredTimer, a.numCarrying, a.numReturned]
                     for a in d.agentStates])
        food.append(grid_bits(d.food))
        score.append(int(d.score))
    assert score[-1] == g.state.data.score
    np.savez_compressed(os.path.join(HERE, "replay_smallCapture_states.npz"), pos=np.array(pos, np.int16),
                        misc=np.array(misc, np.int16), food=np.array(food), score=np.array(score, np.int32),
                        moves=np.array([[i, DIR_ID[a]] for i, a in g.moveHistory], np.int16))
    print(f"replay_smallCapture: {len(g.moveHistory)} moves, final score {score[-1]}", file=sys.stderr)


def main():
    only = set(sys.argv[1:])
    want = lambda k: not only or k in only
    digests = {}
    dpath = os.path.join(HERE, "digests.json")
    if os.path.exists(dpath):
        digests = json.load(open(dpath))

    if want("mazes"):
        seeds = [1, 2, 3, 7, 23, 42, 99, 1001, 4242, 9999, 31337, 123456, 99999999]
        mz = {}
        for s in seeds:
            with _quiet:
                mz[str(s)] = ref_maze.generateMaze(s)
        # randomLayout() with no seed draws the seed from the global RNG
        random.seed(2025)
        with _quiet:
            mz["global2025"] = capture.randomLayout()
        json.dump(mz, open(os.path.join(HERE, "mazes.json"), "w"), indent=0)

    if want("traj"):
        cov = {}
        for lay in ("tinyCapture", "smallCapture", "bloxCapture"):
            cov[f"{lay}_random"] = record(f"{lay}_random", lay, "random", episodes=2, seed=11)
            cov[f"{lay}_smart"] = record(f"{lay}_smart", lay, "smart", episodes=5, seed=12)
        cov["capsSmall"] = record("capsSmall", "capsSmall", "smart", episodes=6, seed=13)
        cov["odd15"] = record("odd15", "odd15", "smart", episodes=6, seed=14)
        cov["mini"] = record("mini", "mini", "smart", episodes=12, seed=15, length=60)
        cov["arena"] = record("arena", "arena", "smart", episodes=6, seed=16, length=200)
        cov["rebind"] = record("rebind", "rebind", "smart", episodes=8, seed=21, length=150)
        cov["swapped"] = record("swapped", "swapped", "smart", episodes=6, seed=17, length=120)
        for s in (7, 23, 1001):
            cov[f"maze{s}"] = record(f"maze{s}", f"maze{s}", "smart", episodes=3, seed=18 + s)
        cov["small_noflags"] = record("small_noflags", "smallCapture", "smart", episodes=3, seed=19, length=80,
                                      reward_forLegalAction=False, defenceReward=False)
        digests["coverage"] = cov

    if want("bots"):
        cov = digests.get("coverage", {})
        for lay, eps in (("tinyCapture", 3), ("smallCapture", 3), ("bloxCapture", 2), ("capsSmall", 2), ("maze7", 1)):
            cov[f"bots_{lay}"] = record(f"bots_{lay}", lay, "smart", episodes=eps, seed=31, enemie="baselineTeam",
                                        full_obs_steps=8)
        digests["coverage"] = cov

    if want("rngbots"):
        # BASELINE config 1 and its randomTeam twin with the bots' REAL random.choice (global MT19937, random.seed(1)):
        # red = the env's bots deciding at their own sub-move, blue = uniform over the legal lists from the same stream
        cov = digests.get("coverage", {})
        cov["rng_baselineTeam_tinyCapture"] = record("rng_baselineTeam_tinyCapture", "tinyCapture", "global_rng", episodes=3,
                                                      seed=1, enemie="baselineTeam", real_rng=True, full_obs_steps=8, length=299)
        cov["rng_randomTeam_tinyCapture"] = record("rng_randomTeam_tinyCapture", "tinyCapture", "global_rng", episodes=3,
                                                    seed=1, enemie="randomTeam", real_rng=True, full_obs_steps=8, length=299)
        cov["rng_baselineTeam_capsSmall"] = record("rng_baselineTeam_capsSmall", "capsSmall", "global_rng", episodes=2,
                                                    seed=1, enemie="baselineTeam", real_rng=True, full_obs_steps=8, length=150)
        digests["coverage"] = cov

    if want("partial"):
        partial_fixture("smallCapture", seed=41, episodes=2)
        partial_fixture("tinyCapture", seed=42, episodes=2)

    if want("replay"):
        replay_fixture()

    if want("survey"):
        digests["survey_traj_sha256"] = {l: survey_digest(l) for l in ("tinyCapture", "smallCapture", "bloxCapture")}

    if want("dist"):
        dd = {}
        for lay in ("tinyCapture", "smallCapture", "bloxCapture", "capsSmall", "odd15", "mini", "arena", "rebind", "swapped",
                    "maze7", "maze23", "maze1001"):
            dd[lay] = distance_fixture(lay, lay)
        digests["dist"] = dd

    if want("gae"):
        gae_fixture()

    json.dump(digests, open(dpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
