"""Batched, GPU-resident drop-in for ``gymPacMan.gymPacMan_parallel_env``.

Reference interface: `/root/reference/gymPacMan.py:12-270` (constructor kwargs
`:15`, ``reset`` `:92-141`, ``step`` `:143-193`, ``get_Observation`` `:195-229`).
The class keeps those names, argument meanings and return structure; every
tensor gains a leading env dimension when ``num_envs > 1``:

  ========================  ==========================  =============================
  reference (1 env, CPU)    here, ``num_envs == 1``      here, ``num_envs == N > 1``
  ========================  ==========================  =============================
  obs[k]  f32 [8,H,W]       f32 [8,H,W]   (cuda)         f32 [N,8,H,W]   (cuda)
  rewards[k]  float         float                        f64 [N]         (cuda)
  terminations[k]  bool     bool                         bool [N]        (cuda)
  legal_actions[k]  [int]   [int] (order N,S,E,W,Stop)   uint8 [N] bitmask (bit a)
  score_change  int         int                          int32 [N]       (cuda)
  ========================  ==========================  =============================

All rule / observation / reward arithmetic runs in the hand-written sm_100a
kernels behind the C ABI (``include/pacman_b200.h``); this file only owns
tensors, argument marshalling and the dict-shaped return values.  There is no
CPU path: constructing the env without the CUDA library or a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import layout as _layout
from . import mazeGenerator as _maze

file_path = os.path.dirname(os.path.abspath(__file__)) + "/"

ACTION_NAMES = ["North", "East", "South", "West", "Stop"]
# Actions.getPossibleActions iterates N, S, E, W, Stop (game.py:295-299,344-348)
LEGAL_LIST_ORDER = [0, 2, 1, 3, 4]
# opponents that run inside the step kernel -> pm_env_desc.bot_kind of the red agents (0, 2), in the order
# createTeam returns them (randomTeam.py: two random agents; baselineTeam.py:34-51: Offensive, Defensive)
ON_DEVICE_TEAMS = {"randomTeam": (1, 1), "baselineTeam": (2, 3)}


def legal_mask_to_list(mask: int) -> List[int]:
    """Bitmask (bit a = gym action a) -> the reference's list, in its order."""
    return [a for a in LEGAL_LIST_ORDER if (int(mask) >> a) & 1]


def decode_obs_bits(bits, H: int, W: int):
    """Packed observations (``set_bits_output`` / ``observe_bits``; numpy or torch int32 [..., NB]) -> float32
    [..., 8, H, W], the tensors ``get_Observation`` returns (gymPacMan.py:195-229).  Host-side helper for consumers
    of the packed format; runs on whatever device / library the input lives on."""
    HW = H * W
    ncat = (8 * HW + 31) // 32
    if isinstance(bits, torch.Tensor):
        w = bits[..., :ncat].to(torch.int64) & 0xFFFFFFFF
        sh = torch.arange(32, device=bits.device, dtype=torch.int64)
        planes = ((w.unsqueeze(-1) >> sh) & 1).reshape(*bits.shape[:-1], ncat * 32)[..., :8 * HW].to(torch.float32)
        patch = bits[..., ncat].to(torch.int64) & 0xFFFFFFFF
        has = patch != 0xFFFFFFFF
        idx = torch.where(has, patch & 0xFFFF, torch.zeros_like(patch))
        val = torch.where(has, (patch >> 16).to(torch.float32), planes.gather(-1, idx.unsqueeze(-1)).squeeze(-1))
        planes.scatter_(-1, idx.unsqueeze(-1), val.unsqueeze(-1))
        return planes.reshape(*bits.shape[:-1], 8, H, W)
    b = np.ascontiguousarray(bits).view(np.uint32)
    planes = np.unpackbits(np.ascontiguousarray(b[..., :ncat]).view(np.uint8), axis=-1, bitorder="little")[..., :8 * HW]
    planes = planes.astype(np.float32)
    patch = b[..., ncat]
    flat = planes.reshape(-1, 8 * HW)
    pf = patch.reshape(-1)
    rows = np.nonzero(pf != 0xFFFFFFFF)[0]
    flat[rows, pf[rows] & 0xFFFF] = (pf[rows] >> 16).astype(np.float32)
    return flat.reshape(*b.shape[:-1], 8, H, W)


def pack_tables(layouts: Sequence[_layout.Layout], with_max_open: bool = False):
    """Layout objects -> (geom[u32], table[L, LS] u32, has_caps) via the library's host packer (one call for all L
    boards).  ``with_max_open``: a fourth value, the largest open-cell count of any board (what pm_distances sizes its
    tables with)."""
    L = _lib.load()
    W, H = layouts[0].width, layouts[0].height
    gw = L.pm_geom_words(W, H)
    _lib.check(min(gw, 0), "pm_geom_words")
    geom = np.zeros(gw, np.uint32)
    _lib.check(L.pm_pack_geom(W, H, geom.ctypes.data), "pm_pack_geom")
    LS = L.pm_layout_words(W, H)
    nl = len(layouts)
    table = np.zeros((nl, LS), np.uint32)
    parts = []
    for lay in layouts:
        if (lay.width, lay.height) != (W, H):
            raise ValueError("all layouts of one batch must share the same width and height")
        text = "".join(lay.layoutText)
        if len(text) != W * H:
            raise ValueError("ragged layout text")
        parts.append(text)
    hc, mo, bad = C.c_int(0), C.c_int(0), C.c_int(-1)
    rc = L.pm_pack_layouts("".join(parts).encode("ascii"), W, H, nl, table.ctypes.data, C.byref(hc), C.byref(mo), C.byref(bad))
    _lib.check(rc, "pm_pack_layouts" + (f" (layout {bad.value})" if bad.value >= 0 else ""))
    if with_max_open:
        return geom, table, hc.value, mo.value
    return geom, table, hc.value


class OnDeviceAgent:
    """Stands where the reference keeps the bot object of a red agent (``env.agents[0]``, ``env.agents[2]`` in
    non-self-play, gymPacMan.py:33-40).  The bot itself runs inside the step kernel at its own sub-move; what it
    played comes back in ``actions_used``."""

    def __init__(self, index: int, team: str, kind: int):
        self.index, self.team, self.kind = index, team, kind
        self.red = True

    def registerInitialState(self, state):
        pass

    def getAction(self, state):
        raise RuntimeError(f"agent {self.index} ({self.team}) decides inside the step kernel; its moves are reported "
                           "in the step's actions_used output")

    def __repr__(self):
        return f"OnDeviceAgent(index={self.index}, team={self.team!r})"


class _AgentDict(dict):
    """The dicts ``reset`` / ``step`` return are keyed like the reference's: by ``env.agents[i]`` - ints, or the bot
    objects of the red agents in non-self-play (gymPacMan.py:186).  Plain agent indices resolve too."""

    def __init__(self, agents, items):
        super().__init__(items)
        self._agents = agents

    def __missing__(self, k):
        if isinstance(k, int) and not isinstance(k, bool) and 0 <= k < len(self._agents) and self._agents[k] is not k \
                and dict.__contains__(self, self._agents[k]):
            return dict.__getitem__(self, self._agents[k])
        raise KeyError(k)


class _AgentStateView:
    def __init__(self, row, start):
        self.configuration = _ConfigView((int(row[0]), int(row[1])), ACTION_NAMES[int(row[2])])
        self.start = _ConfigView(start, "Stop")
        self.isPacman = bool(row[3])
        self.scaredTimer = int(row[4])
        self.numCarrying = int(row[5])
        self.numReturned = int(row[6])

    def getPosition(self):
        return self.configuration.pos

    def getDirection(self):
        return self.configuration.direction


class _ConfigView:
    def __init__(self, pos, direction):
        self.pos, self.direction = pos, direction

    def getPosition(self):
        return self.pos

    def getDirection(self):
        return self.direction


class _StateView:
    """Read-only GameState-like view of env ``n`` (capture.py:101-234 accessors).
    Pulls the packed device state on first use of each step (one small D2H)."""

    def __init__(self, env, n=0):
        self._env, self._n = env, n
        self.data = self

    def _pull(self):
        return self._env.cached_state()

    # GameStateData-ish
    @property
    def score(self):
        return int(self._pull()["score"][self._n])

    @property
    def agentStates(self):
        s = self._pull()
        lay = self._env._layout_of(self._n)
        return [_AgentStateView(s["agents"][self._n, i], lay.starts[i]) for i in range(4)]

    @property
    def layout(self):
        return self._env._layout_of(self._n)

    @property
    def food(self):
        lay = self.layout
        a = self._pull()["food"][self._n].reshape(lay.height, lay.width).T.astype(bool)
        return _layout.BoolGrid(lay.width, lay.height, a)

    @property
    def capsules(self):
        lay = self.layout
        cells = np.nonzero(self._pull()["caps"][self._n])[0]
        return [(int(c % lay.width), int(c // lay.width)) for c in cells]

    # GameState-ish
    def getScore(self):
        return self.score

    def getAgentState(self, i):
        return self.agentStates[i]

    def getAgentPosition(self, i):
        r = self._pull()["agents"][self._n, i]
        return (int(r[0]), int(r[1]))

    def getNumAgents(self):
        return 4

    def getWalls(self):
        return self.layout.walls

    def hasWall(self, x, y):
        return bool(self.layout.walls[x][y])

    def hasFood(self, x, y):
        return bool(self.food[x][y])

    def getCapsules(self):
        return self.capsules

    def _half(self, red):
        f = self.food
        half = int(f.width / 2)
        a = f.array.copy()
        if red:
            a[half:, :] = False
        else:
            a[:half, :] = False
        return _layout.BoolGrid(f.width, f.height, a)

    def getRedFood(self):
        return self._half(True)

    def getBlueFood(self):
        return self._half(False)

    def getRedCapsules(self):
        w = self.layout.width
        return [(x, y) for x, y in self.capsules if x <= w / 2]

    def getBlueCapsules(self):
        w = self.layout.width
        return [(x, y) for x, y in self.capsules if x > w / 2]

    def getLegalActions(self, i=0):
        x, y = self.getAgentPosition(i)
        walls = self.layout.walls
        out = []
        for name, (dx, dy) in (("North", (0, 1)), ("South", (0, -1)), ("East", (1, 0)), ("West", (-1, 0)), ("Stop", (0, 0))):
            if not walls[x + dx][y + dy]:
                out.append(name)
        return out

    def isOnRedTeam(self, i):
        lay = self.layout
        return lay.starts[i][0] < lay.width / 2

    def getRedTeamIndices(self):
        return [i for i in range(4) if self.isOnRedTeam(i)]

    def getBlueTeamIndices(self):
        return [i for i in range(4) if not self.isOnRedTeam(i)]

    def getInitialAgentPosition(self, i):
        return self.layout.starts[i]


class _PartialStateView(_StateView):
    """`GameState.makeObservation(index)` (capture.py:268-292): enemies nobody on the observer's team can see have no
    configuration, and the sonar readings come along."""

    def __init__(self, env, n, visible_mask, distances):
        super().__init__(env, n)
        self._visible, self.agentDistances = visible_mask, distances

    def getAgentDistances(self):
        return self.agentDistances

    def getAgentPosition(self, i):
        return super().getAgentPosition(i) if (self._visible >> i) & 1 else None

    def getAgentState(self, i):
        a = super().getAgentState(i)
        if not (self._visible >> i) & 1:
            a.configuration = None
        return a


class _GameView:
    def __init__(self, env):
        self.state = _StateView(env, 0)
        self.length = env.length


class gymPacMan_parallel_env:
    metadata = {"render_modes": [], "name": "b200_capture_v1"}

    def __init__(self, layout_file=f"{file_path}/layouts/smallCapture.lay", display=False, length=299,
                 reward_forLegalAction=True, defenceReward=True, random_layout=False, enemieName="randomTeam",
                 self_play=False, *, num_envs: int = 1, device=None, layouts: Optional[Sequence] = None,
                 shaping_scale: float = 0.1, opponent_seed: int = 0, auto_reset: bool = False,
                 env_index_base: int = 0, opponent_tiebreak: str = "random", layout_ids=None,
                 bot_team=None, state_factory=None, layout_pool: int = 0, envs_per_cta: int = 0,
                 elide_static_planes: bool = False):
        if display:
            raise NotImplementedError("display=True (tkinter viewer) is outside the batched GPU path")
        if not torch.cuda.is_available():
            raise RuntimeError("pacman_marl_2025_b200 needs a CUDA device; there is no CPU fallback")
        self._L = _lib.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.num_envs = int(num_envs)
        if self.num_envs < 1:
            raise ValueError("num_envs must be >= 1")
        self.random_layout = random_layout
        self.enemieName = enemieName
        self.self_play = self_play
        self.display = None
        self.display_bool = False
        self.reward_forLegalAction = reward_forLegalAction
        self.defenceReward = defenceReward
        self.length = length
        self.shaping_scale = float(shaping_scale)
        self.num_moves = 0
        self.steps = 0
        self.action_mapping = {0: "North", 1: "East", 2: "South", 3: "West", 4: "Stop", "North": "North",
                               "East": "East", "South": "South", "West": "West", "Stop": "Stop"}
        self.reversed_action_mapping = {"North": 0, "East": 1, "South": 2, "West": 3, "Stop": 4,
                                        0: 0, 1: 1, 2: 2, 3: 3, 4: 4}
        self._name_to_int = {"North": 0, "East": 1, "South": 2, "West": 3, "Stop": 4}
        # agent keys (gymPacMan.py:33-40): ints in self-play; without self-play agents 0 and 2 are the red team's bot
        # objects - the caller's host bots, or OnDeviceAgent placeholders for the teams that run inside the kernel
        self.agents = [0, 1, 2, 3]
        # host-side opponents (the reference's agents/*.py teams): two objects with getAction(state) [and
        # registerInitialState(state)] playing agents 0 and 2 at their own sub-move; `state_factory(env, n)` builds
        # what they are handed (default: this module's read-only GameState view; reference_bridge.state_factory
        # builds the reference's own GameState for bots that call generateSuccessor)
        self._bot_team = list(bot_team) if bot_team is not None else None
        self._state_factory = state_factory
        if self._bot_team is not None:
            if self_play or len(self._bot_team) != 2:
                raise ValueError("bot_team needs self_play=False and exactly two bots (agents 0 and 2)")
        if not self_play and self._bot_team is None and enemieName not in ON_DEVICE_TEAMS:
            raise NotImplementedError(
                f"enemieName={enemieName!r}: the on-device opponents are {sorted(ON_DEVICE_TEAMS)}; the other CPU "
                "bot teams (agents/*.py search agents) are outside the GPU path (SURVEY §8 row f1)")
        if opponent_tiebreak not in ("random", "first"):
            raise ValueError("opponent_tiebreak must be 'random' (random.choice, baselineTeam.py:92) or 'first'")
        self._opponent_tiebreak = opponent_tiebreak
        self._opponent_seed = int(opponent_seed)
        self._envs_per_cta = int(envs_per_cta)
        # the step may leave the observation elements that cannot have changed (walls; capsule planes of capsule-free
        # batches) as the previous write of the env's own observation tensor left them (pm_step_out.obs_static_valid)
        self.elide_static_planes = bool(elide_static_planes)
        self._obs_static_ok = False
        self._submoves_in_step = 0
        # batched extensions: restart finished envs inside the step launch; global index of env 0 when
        # this batch is one shard of a larger one (keys the on-device opponent's random draws)
        self.auto_reset = bool(auto_reset)
        self.env_index_base = int(env_index_base)
        # with several `layouts`: env n plays layouts[layout_ids[n]] (default: n % len(layouts))
        self._layout_ids_arg = layout_ids

        # random_layout: the reference draws a NEW maze on every reset (gymPacMan.py:98-100 -> capture.py:905-911).
        # One env does exactly that (host generator, same global-RNG draw).  A batch keeps a POOL of `layout_pool`
        # host-generated mazes on the device and every env moves to pool[hash(seed, env, episode) % pool] whenever
        # it (re)starts - inside pm_reset / the auto_reset of pm_step (pm_env_desc.layout_pool).
        self._pool_mode = False
        self._pool_epoch = 0
        if layouts is not None:
            lays = [l if isinstance(l, _layout.Layout) else _layout.Layout(list(l)) for l in layouts]
            self._pool_mode = bool(layout_pool) and len(lays) > 1
        elif os.path.exists(layout_file) and not random_layout:
            lays = [_layout.getLayout(layout_file)]
        elif self.num_envs > 1:
            lays = [_layout.Layout(_maze.randomLayout().split("\n")) for _ in range(int(layout_pool) or 256)]
            self._pool_mode = len(lays) > 1
        else:
            lays = [_layout.Layout(_maze.randomLayout().split("\n"))]
        self._version = 0
        self._state_cache_version = -1
        self._state_cache = None
        self._install_layouts(lays)
        self.game = _GameView(self)
        self.reset(_first=True)
        self.max_score = int(self.game.state.getBlueFood().count())

    # ------------------------------------------------------------------ setup
    _CACHED = ("_mapped_out", "_rec_mapped", "_rec_mapped_views", "_rec_np", "_pipe_out", "_rec_dev", "_rec_host",
               "_rec_views", "_host", "_host_out", "_act_dev", "_views", "_bits")

    def _install_layouts(self, lays: List[_layout.Layout]):
        """(Re)build the device tables for `lays`.  Boards of the size already installed are copied into the existing
        tables, so every device buffer (and every cached descriptor that points into one) stays where it is; a new
        board size reallocates and drops the cached descriptors and host staging buffers."""
        W, H = lays[0].width, lays[0].height
        geom, table, has_caps = pack_tables(lays)
        dev = self.device
        N = self.num_envs
        self._layouts = lays
        self.layout = lays[0]
        key = (W, H, len(lays))
        if getattr(self, "_alloc_key", None) == key:
            self._table.copy_(torch.from_numpy(table.view(np.int32)))
            self._desc.has_caps = int(has_caps)
        else:
            for name in self._CACHED:                  # they hold pointers / strides of the buffers replaced below
                self.__dict__.pop(name, None)
            self._alloc_key = key
            self._geom = torch.from_numpy(geom.view(np.int32)).to(dev)
            self._table = torch.from_numpy(table.view(np.int32)).to(dev)
            nbytes = int(self._L.pm_state_bytes(W, H, N))
            self._state = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self._desc = _lib.PmEnvDesc(W, H, len(lays), int(has_caps), N, self._geom.data_ptr(), self._table.data_ptr(),
                                        None, self._state.data_ptr(), int(self.length),
                                        int(bool(self.reward_forLegalAction)), int(bool(self.defenceReward)), 0,
                                        self.shaping_scale, int(self._opponent_seed), int(self.auto_reset),
                                        self._envs_per_cta, self.env_index_base, 0, 0)
            self.obs_shape = (8, H, W)
            self._obs = torch.empty((4, N, 8, H, W), dtype=torch.float32, device=dev)
            self._reward = torch.zeros((2, N), dtype=torch.float64, device=dev)
            self._done = torch.zeros(N, dtype=torch.uint8, device=dev)
            self._legal = torch.zeros((N, 4), dtype=torch.uint8, device=dev)
            self._score_change = torch.zeros(N, dtype=torch.int32, device=dev)
            self._shaping = torch.zeros((4, N), dtype=torch.float64, device=dev)
            self._shaped = torch.zeros((4, N), dtype=torch.float32, device=dev)
            self._actions_used = torch.zeros((N, 4), dtype=torch.uint8, device=dev)
            self._step_out = _lib.PmStepOut(self._obs.data_ptr(), self._obs[0].numel(), self._reward.data_ptr(),
                                            self._done.data_ptr(), self._legal.data_ptr(), self._score_change.data_ptr(),
                                            self._shaping.data_ptr(), self._shaped.data_ptr(), None,
                                            self._actions_used.data_ptr())
            self._layout_id = None
        self._desc.layout_pool = int(self._pool_mode and len(lays) > 1)
        if self._pool_mode and len(lays) > 1:
            if self._layout_id is None:
                self._layout_id = torch.zeros(N, dtype=torch.int32, device=dev)      # written by the kernels
            self._layout_id_host = None
            self._desc.layout_id = self._layout_id.data_ptr()
        elif self._layout_ids_arg is not None and len(lays) > 1:
            ids = torch.as_tensor(np.asarray(self._layout_ids_arg), dtype=torch.int32)
            if ids.numel() != N or int(ids.min()) < 0 or int(ids.max()) >= len(lays):
                raise ValueError("layout_ids must hold one index into `layouts` per env")
            self._layout_id = ids.to(dev).contiguous()
            self._layout_id_host = ids.numpy().copy()
            self._desc.layout_id = self._layout_id.data_ptr()
        else:
            self._layout_id_host = None
            self._desc.layout_id = None
        self._set_opponent(self.enemieName)
        self._version += 1

    def _set_opponent(self, name):
        """Red side = the named on-device team (loadAgents(True, ...) puts bots on agents 0 and 2,
        gymPacMan.py:33-40).  baselineTeam needs the maze-distance tables: every reset of the reference
        rebuilds them per bot (captureAgents.py:91-109); here one BFS launch per layout set."""
        d = self._desc
        if self.self_play:
            d.bot_kind = 0
            self.agents = [0, 1, 2, 3]
            return
        if self._bot_team is not None:
            d.bot_kind = 0
            self.agents = [self._bot_team[0], 1, self._bot_team[1], 3]
            return
        k0, k2 = ON_DEVICE_TEAMS[name]
        self.agents = [OnDeviceAgent(0, name, k0), 1, OnDeviceAgent(2, name, k2), 3]
        d.bot_kind = k0 | (k2 << 8)
        d.bot_first_best = int(self._opponent_tiebreak == "first")
        if max(k0, k2) >= 2 and getattr(self, "_dist_for", None) is not self._layouts:
            from .distanceCalculator import compute_distances
            self._dist_tables = compute_distances(self._layouts, device=self.device, table=self._table)
            self._dist_for = self._layouts
        if max(k0, k2) >= 2:
            t = self._dist_tables
            d.dist, d.node_of_cell, d.dist_nmax = t.dist.data_ptr(), t.node_of_cell.data_ptr(), int(t.dist.shape[1])

    def _layout_of(self, n):
        if self._pool_mode and len(self._layouts) > 1:
            return self._layouts[int(self.cached_state()["layout_id"][n])]
        if getattr(self, "_layout_id_host", None) is not None:
            return self._layouts[int(self._layout_id_host[n])]
        return self._layouts[n % len(self._layouts)]

    def cached_state(self):
        """``export_state()`` of the current state, pulled from the device at most once per step / sub-move launch."""
        if self._state_cache_version != self._version:
            self._state_cache = self.export_state()
            self._state_cache_version = self._version
        return self._state_cache

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------ batched API
    def reset_batched(self, mask: Optional[torch.Tensor] = None):
        """Reset every env (mask None) or the envs with mask[n] != 0.  Returns
        (obs [4,N,8,H,W], legal [N,4] bitmasks); rows of envs that were not
        reset keep their previous contents."""
        mptr = None
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            mptr = mask.data_ptr()
        elif self._pool_mode:
            # a full reset restarts the episode counters; a fresh epoch in the seed keeps the boards (and the
            # on-device opponents' draws) of successive rollouts from repeating
            self._desc.random_seed = (self._opponent_seed + 0x9E3779B97F4A7C15 * self._pool_epoch) & 0xFFFFFFFFFFFFFFFF
            self._pool_epoch += 1
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_reset(C.byref(self._desc), mptr, self._obs.data_ptr(), self._obs[0].numel(),
                                        self._legal.data_ptr(), self._stream()), "pm_reset")
        self._version += 1
        if mask is None:
            self._obs_static_ok = True                  # every observation of the batch has just been written in full
        return self._obs, self._legal

    def step_batched(self, actions: torch.Tensor):
        """actions: [N,4] uint8 or int64 on the env's device.  Returns a dict of
        views of the env's output buffers (overwritten by the next call)."""
        if actions.device != self.device:
            actions = actions.to(self.device, non_blocking=True)
        if actions.dtype not in (torch.uint8, torch.int64):
            actions = actions.to(torch.int64)
        actions = actions.contiguous()
        if actions.numel() != 4 * self.num_envs:
            raise ValueError(f"expected {self.num_envs}x4 actions, got {tuple(actions.shape)}")
        return self._launch_step(actions)

    def step_chain(self, actions: torch.Tensor, collect: bool = False):
        """``for t in range(T): step_batched(actions[t])`` as ONE call (pm_step_chain): T launches that overlap, each
        env tile of step t+1 starting as soon as the same tile of step t is done.  For loops whose actions do not depend
        on the observations (random-action self-play, scripted policies, recorded move lists).  actions: [T, N, 4]
        uint8 / int64 on the env's device.  Returns the dict of the LAST step; with ``collect`` the per-step small
        outputs are kept as well: ``reward`` [T,2,N], ``done`` [T,N], ``legal`` [T,N,4], ``score_change`` [T,N],
        ``shaped_reward`` [T,4,N], ``actions_used`` [T,N,4] (observations are always the last step's)."""
        if actions.device != self.device:
            actions = actions.to(self.device)
        if actions.dtype not in (torch.uint8, torch.int64):
            actions = actions.to(torch.int64)
        actions = actions.contiguous()
        N = self.num_envs
        if actions.dim() != 3 or tuple(actions.shape[1:]) != (N, 4):
            raise ValueError(f"expected [T, {N}, 4] actions, got {tuple(actions.shape)}")
        if self._bot_team is not None:
            raise ValueError("host-side bots decide from the live state: use step()")
        T = int(actions.shape[0])
        if getattr(self, "_chain_scratch", None) is None:
            self._chain_scratch = torch.empty(max(4, int(self._L.pm_chain_scratch_bytes(N))), dtype=torch.uint8,
                                              device=self.device)
        self._mark_static()
        if collect:
            dev = self.device
            coll = dict(reward=torch.empty((T, 2, N), dtype=torch.float64, device=dev),
                        done=torch.empty((T, N), dtype=torch.uint8, device=dev),
                        legal=torch.empty((T, N, 4), dtype=torch.uint8, device=dev),
                        score_change=torch.empty((T, N), dtype=torch.int32, device=dev),
                        shaped_reward=torch.empty((T, 4, N), dtype=torch.float32, device=dev),
                        actions_used=torch.empty((T, N, 4), dtype=torch.uint8, device=dev))
            outs = (_lib.PmStepOut * T)()
            for t in range(T):
                C.memmove(C.byref(outs[t]), C.byref(self._step_out), C.sizeof(_lib.PmStepOut))
                o = outs[t]
                o.reward, o.done, o.legal = coll["reward"][t].data_ptr(), coll["done"][t].data_ptr(), coll["legal"][t].data_ptr()
                o.score_change, o.shaped_reward = coll["score_change"][t].data_ptr(), coll["shaped_reward"][t].data_ptr()
                o.actions_used = coll["actions_used"][t].data_ptr()
            n_outs = T
        else:
            outs, n_outs = (_lib.PmStepOut * 1)(), 1
            C.memmove(C.byref(outs[0]), C.byref(self._step_out), C.sizeof(_lib.PmStepOut))
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_step_chain(C.byref(self._desc), actions.data_ptr(), actions.element_size(), 4 * N, T,
                                             outs, n_outs, self._chain_scratch.data_ptr(), self._stream()), "pm_step_chain")
        self._version += 1
        self.steps += T
        out = self._out_dict()
        if collect:
            # the env's own small-output buffers were not written by this call: hand back the last step's slices
            out.update({k: v[-1] for k, v in coll.items()})
            out["collected"] = coll
        return out

    # ---- learner views: the trainer's network inputs written by the encoder (pm_step_out.view_*) ----
    def set_view_outputs(self, red: Optional[bool], actor: Optional[torch.Tensor] = None,
                         critic: Optional[torch.Tensor] = None, raw_obs: bool = True):
        """Make every following ``step_batched`` also write the learner views of team ``red`` (False: agents
        1, 3 as they are; True: agents 0, 2 mirrored — ``canonicalize_obs``) into ``actor`` (uint8
        [2, N, 8, H, W], or any tensor whose two leading blocks are contiguous [N, 8, H, W]) and ``critic`` (uint8
        [N, 8, H, W], ``merge_obs_for_critic``).  ``red=None`` switches the views off.  ``raw_obs=False`` skips
        the four fp32 observations altogether (the step then writes 3*8*H*W bytes per env instead of
        4*32*H*W).  Returns (actor, critic)."""
        so = self._step_out
        so.obs = self._obs.data_ptr() if raw_obs else None
        self.__dict__.pop("_mapped_out", None)           # the host-step descriptors mirror these pointers: rebuild them
        self.__dict__.pop("_pipe_out", None)
        if red is None:
            so.view_actor, so.view_critic = None, None
            self._views = None
            return None, None
        N, shp = self.num_envs, self.obs_shape
        if actor is None:
            actor = torch.empty((2, N) + shp, dtype=torch.uint8, device=self.device)
        if critic is None:
            critic = torch.empty((N,) + shp, dtype=torch.uint8, device=self.device)
        for t_, lead in ((actor, 2), (critic, 1)):
            if t_.dtype != torch.uint8 or t_.device != self.device:
                raise ValueError("view buffers must be uint8 tensors on the env's device")
        if tuple(actor.shape) != (2, N) + shp or not actor[0].is_contiguous() or not actor[1].is_contiguous():
            raise ValueError(f"actor view buffer must be [2, {N}, {shp}] with contiguous [N,8,H,W] blocks")
        if tuple(critic.shape) != (N,) + shp or not critic.is_contiguous():
            raise ValueError(f"critic view buffer must be a contiguous [{N}, {shp}] tensor")
        so.view_actor, so.view_actor_stride = actor.data_ptr(), actor.stride(0)
        so.view_critic, so.view_red = critic.data_ptr(), int(bool(red))
        self._views = (actor, critic, bool(red))
        return actor, critic

    # ---- packed observations: the 8*H*W-bit string + own-cell patch per observation (pm_step_out.obs_bits) ----
    def obs_bits_words(self) -> int:
        return int(self._L.pm_obs_bits_words(self.layout.width, self.layout.height))

    def set_bits_output(self, on: bool = True, raw_obs: bool = True, host_mapped: bool = False):
        """Make every following step also write the PACKED observations: int32 [4, N, NB] words (NB =
        ``obs_bits_words()``; layout in include/pacman_b200.h).  144 B per observation on tinyCapture instead of
        4 480 B of fp32 - the format for a consumer on the host.  ``host_mapped=True`` puts the buffer in pinned
        host memory that the kernel writes itself across PCIe (use with ``step_host_mapped``).  ``raw_obs=False``
        skips the fp32 observations.  Returns the buffer (``decode_obs_bits`` expands it)."""
        so = self._step_out
        so.obs = self._obs.data_ptr() if raw_obs else None
        if not on:
            so.obs_bits = None
            self._bits = None
            self.__dict__.pop("_mapped_out", None)
            return None
        NB = self.obs_bits_words()
        if host_mapped:
            bits = torch.zeros((4, self.num_envs, NB), dtype=torch.int32).pin_memory()
        else:
            bits = torch.zeros((4, self.num_envs, NB), dtype=torch.int32, device=self.device)
        so.obs_bits, so.obs_bits_agent_stride = bits.data_ptr(), bits.stride(0)
        self._bits = bits
        self.__dict__.pop("_mapped_out", None)           # rebuilt with the new pointers on the next mapped step
        return bits

    def observe_bits(self, out: Optional[torch.Tensor] = None):
        """Packed observations of the CURRENT state (e.g. right after a reset): int32 [4, N, NB] on the device."""
        if out is None:
            out = torch.zeros((4, self.num_envs, self.obs_bits_words()), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_observe_bits(C.byref(self._desc), out.data_ptr(), out.stride(0), self._stream()),
                       "pm_observe_bits")
        return out

    def set_bot_draws(self, draws: Optional[torch.Tensor]):
        """``pm_env_desc.bot_draws``: [N,4] int32 tensor on the env's device holding, for each agent's on-device bot,
        the index ``random.choice`` returns at its next decision (None = the kernel hashes its own draws).  The
        tensor is read by every following step; update it in place between steps.  Feeding indices recorded from
        the reference's MT19937 stream replays the reference's trajectories exactly (tests/golden/traj_rng_*)."""
        if draws is None:
            self._bot_draws = None
            self._desc.bot_draws = None
            return
        if draws.dtype != torch.int32 or draws.device != self.device or draws.numel() != 4 * self.num_envs \
                or not draws.is_contiguous():
            raise ValueError("bot draws must be a contiguous int32 [N,4] tensor on the env's device")
        self._bot_draws = draws
        self._desc.bot_draws = draws.data_ptr()

    def observe_views(self, red: bool, actor: Optional[torch.Tensor] = None, critic: Optional[torch.Tensor] = None):
        """The learner views of the CURRENT state (e.g. right after a reset)."""
        N, shp = self.num_envs, self.obs_shape
        if actor is None:
            actor = torch.empty((2, N) + shp, dtype=torch.uint8, device=self.device)
        if critic is None:
            critic = torch.empty((N,) + shp, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_observe_views(C.byref(self._desc), int(bool(red)), actor.data_ptr(), actor.stride(0),
                                                critic.data_ptr(), self._stream()), "pm_observe_views")
        return actor, critic

    def step_host(self, actions_host: torch.Tensor, want_obs: bool = False, want_views: bool = False,
                  want_bits: bool = False):
        """One step with HOST buffers through the C ABI's ``pm_step_host``: ``actions_host``
        is a CPU uint8 [N,4] tensor (pinned memory makes the copy asynchronous);
        rewards / done / legal / score_change / shaped rewards come back in pinned
        host tensors; observations stay in HBM (``self._obs``) unless ``want_obs``.
        ``want_views`` / ``want_bits`` also bring the uint8 learner views (``set_view_outputs``) / the packed
        observations (``set_bits_output``) to pinned host tensors ``view_actor``, ``view_critic`` / ``obs_bits``.
        The call returns after the stream has been synchronised."""
        if actions_host.device.type != "cpu" or actions_host.dtype != torch.uint8 or not actions_host.is_contiguous():
            raise ValueError("actions_host must be a contiguous CPU uint8 [N,4] tensor")
        if actions_host.numel() != 4 * self.num_envs:
            raise ValueError(f"expected {self.num_envs}x4 actions")
        if not hasattr(self, "_host"):
            N = self.num_envs
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            self._host = dict(reward=pin((2, N), torch.float64), done=pin((N,), torch.uint8),
                              legal=pin((N, 4), torch.uint8), score_change=pin((N,), torch.int32),
                              shaped_reward=pin((4, N), torch.float32))
            self._act_dev = torch.empty((N, 4), dtype=torch.uint8, device=self.device)
            h = self._host
            self._host_out = _lib.PmStepOut(None, 0, h["reward"].data_ptr(), h["done"].data_ptr(), h["legal"].data_ptr(),
                                            h["score_change"].data_ptr(), None, h["shaped_reward"].data_ptr())
        ho = self._host_out
        if want_obs or want_views or want_bits:
            ho = _lib.PmStepOut(None, 0, ho.reward, ho.done, ho.legal, ho.score_change, None, ho.shaped_reward)
        if want_obs:
            if "obs" not in self._host:
                self._host["obs"] = torch.empty(self._obs.shape, dtype=torch.float32).pin_memory()
            ho.obs, ho.obs_agent_stride = self._host["obs"].data_ptr(), self._obs[0].numel()
        if want_views:
            if not getattr(self, "_views", None):
                raise RuntimeError("want_views needs set_view_outputs(...) first")
            if "view_actor" not in self._host:
                self._host["view_actor"] = torch.empty(self._views[0].shape, dtype=torch.uint8).pin_memory()
                self._host["view_critic"] = torch.empty(self._views[1].shape, dtype=torch.uint8).pin_memory()
            ho.view_actor, ho.view_actor_stride = self._host["view_actor"].data_ptr(), self._host["view_actor"].stride(0)
            ho.view_critic = self._host["view_critic"].data_ptr()
        if want_bits:
            if getattr(self, "_bits", None) is None or self._bits.device.type != "cuda":
                raise RuntimeError("want_bits needs set_bits_output(on=True, host_mapped=False) first")
            if "obs_bits" not in self._host:
                self._host["obs_bits"] = torch.empty(self._bits.shape, dtype=torch.int32).pin_memory()
            ho.obs_bits, ho.obs_bits_agent_stride = self._host["obs_bits"].data_ptr(), self._host["obs_bits"].stride(0)
        self._mark_static()
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_step_host(C.byref(self._desc), actions_host.data_ptr(), self._act_dev.data_ptr(),
                                            C.byref(self._step_out), C.byref(ho), self._stream()), "pm_step_host")
        self._version += 1
        self.steps += 1
        return self._host

    def step_host_pipelined(self, actions_host: torch.Tensor, n_chunks: int = 4):
        """One step with HOST buffers through ``pm_step_host_pipelined``: the batch is cut into
        ``n_chunks`` env ranges on two internal streams so the copy-back of one range overlaps the
        kernel of the next, and every small output travels in ONE 48-byte record per env
        (``struct pm_host_record``).  Returns a dict of views into the pinned record buffer:
        reward [N,2] f64 (red, blue), shaped_reward [N,4] f32, score_change [N] i32, legal [N,4] u8,
        done [N] u8.  Observations stay in HBM (``self._obs``)."""
        if actions_host.device.type != "cpu" or actions_host.dtype != torch.uint8 or not actions_host.is_contiguous():
            raise ValueError("actions_host must be a contiguous CPU uint8 [N,4] tensor")
        if actions_host.numel() != 4 * self.num_envs:
            raise ValueError(f"expected {self.num_envs}x4 actions")
        N = self.num_envs
        if getattr(self, "_pipe_chunks", None) != n_chunks:
            if getattr(self, "_pipe", None):
                self._L.pm_pipe_destroy(self._pipe)
            pipe = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self._L.pm_pipe_create(C.byref(pipe), int(n_chunks)), "pm_pipe_create")
            self._pipe, self._pipe_chunks = pipe, n_chunks
        if not hasattr(self, "_rec_host"):
            self._rec_dev = torch.empty((N, 48), dtype=torch.uint8, device=self.device)
            self._rec_host = torch.empty((N, 48), dtype=torch.uint8).pin_memory()
            self._rec_views = self._record_views(self._rec_host)
        if not hasattr(self, "_act_dev"):
            self._act_dev = torch.empty((N, 4), dtype=torch.uint8, device=self.device)
        if not hasattr(self, "_pipe_out"):
            so = self._step_out
            self._pipe_out = _lib.PmStepOut(so.obs, so.obs_agent_stride, None, None, None, None, None, None,
                                            self._rec_dev.data_ptr())
        self._mark_static()
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_step_host_pipelined(C.byref(self._desc), self._pipe, actions_host.data_ptr(),
                                                      self._act_dev.data_ptr(), C.byref(self._pipe_out),
                                                      self._rec_host.data_ptr(), self._stream()),
                       "pm_step_host_pipelined")
        self._version += 1
        self.steps += 1
        return self._rec_views

    def _record_views(self, r):
        return dict(reward=r.view(torch.float64)[:, 0:2], shaped_reward=r.view(torch.float32)[:, 4:8],
                    score_change=r.view(torch.int32)[:, 8], legal=r[:, 36:40], done=r[:, 40])

    def step_host_mapped(self, actions_host: torch.Tensor):
        """One step with HOST buffers and no copy engine (``pm_step_host_mapped``): the kernel reads the
        pinned ``actions_host`` ([N,4] uint8) and writes one ``pm_host_record`` per env into a pinned host
        buffer straight across PCIe; the call returns after a stream synchronise.  Same return value as
        :meth:`step_host_pipelined`."""
        if (actions_host.device.type != "cpu" or actions_host.dtype != torch.uint8 or not actions_host.is_contiguous()
                or not actions_host.is_pinned()):
            raise ValueError("actions_host must be a contiguous, pinned CPU uint8 [N,4] tensor")
        if actions_host.numel() != 4 * self.num_envs:
            raise ValueError(f"expected {self.num_envs}x4 actions")
        self._ensure_mapped()
        self._mark_static()
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_step_host_mapped(C.byref(self._desc), actions_host.data_ptr(),
                                                   C.byref(self._mapped_out), self._rec_mapped.data_ptr(),
                                                   self._stream()), "pm_step_host_mapped")
        self._version += 1
        self.steps += 1
        return self._rec_mapped_views

    def step_chain_host_mapped(self, actions_host: torch.Tensor, records: Optional[torch.Tensor] = None, sync: bool = True):
        """``for t: step_host_mapped(actions_host[t])`` as one chained run (``pm_step_chain_host_mapped``): the T
        launches overlap, each reading its [N,4] actions from the pinned ``actions_host`` [T,N,4] and writing its
        records into the pinned ``records`` [T,N,48] (allocated and cached when not given) across PCIe.  Returns the
        per-step views (``reward`` [T,N,2], ``shaped_reward`` [T,N,4], ``score_change`` [T,N], ``legal`` [T,N,4],
        ``done`` [T,N]) after a stream synchronise (``sync=False``: the caller synchronises)."""
        N = self.num_envs
        if (actions_host.device.type != "cpu" or actions_host.dtype != torch.uint8 or not actions_host.is_contiguous()
                or not actions_host.is_pinned() or actions_host.dim() != 3 or tuple(actions_host.shape[1:]) != (N, 4)):
            raise ValueError(f"actions_host must be a contiguous, pinned CPU uint8 [T, {N}, 4] tensor")
        T = int(actions_host.shape[0])
        if records is None:
            records = getattr(self, "_rec_chain", None)
            if records is None or records.shape[0] < T:
                records = self._rec_chain = torch.zeros((T, N, 48), dtype=torch.uint8).pin_memory()
            records = records[:T]
        elif not records.is_pinned() or records.dtype != torch.uint8 or tuple(records.shape) != (T, N, 48):
            raise ValueError(f"records must be a pinned uint8 [T, {N}, 48] tensor")
        self._ensure_mapped()
        self._mark_static()
        if getattr(self, "_chain_scratch", None) is None:
            self._chain_scratch = torch.empty(max(4, int(self._L.pm_chain_scratch_bytes(N))), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_step_chain_host_mapped(C.byref(self._desc), actions_host.data_ptr(), T,
                                                         C.byref(self._mapped_out), records.data_ptr(),
                                                         self._chain_scratch.data_ptr(), self._stream()),
                       "pm_step_chain_host_mapped")
            if sync:
                torch.cuda.current_stream(self.device).synchronize()
        self._version += 1
        self.steps += T
        r = records.view(T * N, 48)
        return {k: v.view((T, N) + tuple(v.shape[1:])) for k, v in self._record_views(r).items()}

    def _ensure_mapped(self):
        """Pinned record buffer + the pm_step_out the zero-copy entry uses (observations / views / packed bits as the
        env is currently configured; the small outputs travel in the records)."""
        if not hasattr(self, "_rec_mapped"):
            self._rec_mapped = torch.zeros((self.num_envs, 48), dtype=torch.uint8).pin_memory()
            self._rec_mapped_views = self._record_views(self._rec_mapped)
            dt = np.dtype([("reward", "<f8", 2), ("shaped", "<f4", 4), ("score_change", "<i4"), ("legal", "u1", 4),
                           ("done", "u1"), ("pad", "u1", 7)])
            self._rec_np = self._rec_mapped.numpy().view(dt)[:, 0]
        if not hasattr(self, "_mapped_out"):
            so = self._step_out
            self._mapped_out = _lib.PmStepOut(so.obs, so.obs_agent_stride, None, None, None, None, None, None, None,
                                              self._actions_used.data_ptr(), so.view_actor, so.view_actor_stride,
                                              so.view_critic, so.view_red, 0, so.obs_bits, so.obs_bits_agent_stride)

    def __del__(self):
        try:
            if getattr(self, "_pipe", None):
                self._L.pm_pipe_destroy(self._pipe)
                self._pipe = None
        except Exception:
            pass

    def observe_batched(self, out: Optional[torch.Tensor] = None):
        """get_Observation for all four agents of every env on the CURRENT state."""
        if out is None:
            out = torch.empty_like(self._obs)
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_observe(C.byref(self._desc), out.data_ptr(), out[0].numel(), self._stream()),
                       "pm_observe")
        return out

    # ---- partial observability: GameState.makeObservation (capture.py:268-292) ----
    def make_observation_batched(self, draws: Optional[torch.Tensor] = None):
        """What every agent of every env may know of the CURRENT state, as the tournament loop hands it out
        (`game.py:570-732` -> `capture.py:268-292`): ``visible`` uint8 [N,4] (bit e of ``visible[n, i]``: agent e's
        configuration survives in agent i's view - own team always, an enemy iff a team-mate of i stands within 5
        Manhattan cells of it) and ``noisy_distances`` int32 [N,4,4] (row i: agent i's sonar readings).  ``draws``
        int32 [N,4,4]: the index `random.choice(SONAR_NOISE_VALUES)` returned for every reading (default: hashed
        in-kernel from the env's seed, env index, episode and step)."""
        N, dev = self.num_envs, self.device
        vis = torch.empty((N, 4), dtype=torch.uint8, device=dev)
        noisy = torch.empty((N, 4, 4), dtype=torch.int32, device=dev)
        dptr = None
        if draws is not None:
            draws = draws.to(device=dev, dtype=torch.int32).contiguous()
            if draws.numel() != 16 * N:
                raise ValueError("draws must hold one index per (env, observer, agent)")
            dptr = draws.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(self._L.pm_make_observation(C.byref(self._desc), dptr, vis.data_ptr(), noisy.data_ptr(),
                                                   self._stream()), "pm_make_observation")
        return vis, noisy

    def make_observation(self, index: int, n: int = 0, draws=None):
        """``env.game.state.makeObservation(index)`` for env ``n``: a GameState-like view in which unseen enemies have no
        position (``getAgentPosition`` -> None) and ``getAgentDistances()`` returns the noisy readings."""
        d = None if draws is None else torch.as_tensor(np.asarray(draws, np.int32))
        vis, noisy = self.make_observation_batched(d)
        return _PartialStateView(self, n, int(vis[n, index].item()), [int(v) for v in noisy[n, index].tolist()])

    def export_state(self) -> Dict[str, np.ndarray]:
        """Unpacked copy of the device state (GameState accessors; one D2H)."""
        N, HW = self.num_envs, self.layout.width * self.layout.height
        dev = self.device
        ag = torch.empty((N, 4, 7), dtype=torch.int32, device=dev)
        food = torch.empty((N, HW), dtype=torch.uint8, device=dev)
        caps = torch.empty((N, HW), dtype=torch.uint8, device=dev)
        misc = torch.empty((N, 3), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(self._L.pm_get_state(C.byref(self._desc), ag.data_ptr(), food.data_ptr(), caps.data_ptr(),
                                            misc.data_ptr(), self._stream()), "pm_get_state")
        misc = misc.cpu().numpy()
        out = dict(agents=ag.cpu().numpy(), food=food.cpu().numpy(), caps=caps.cpu().numpy(),
                   score=misc[:, 0].copy(), steps=misc[:, 1].copy(), episode=misc[:, 2].copy())
        if self._layout_id is not None:
            out["layout_id"] = self._layout_id.cpu().numpy()
        return out

    def import_state(self, agents, food, caps, score, steps, episode=None, layout_id=None):
        """Inverse of :meth:`export_state` (checkpoint restore / crafted test states).  Valid between whole steps:
        the parked accumulators of a step that is being played in several launches are not part of the state.
        ``episode`` (restarts so far) keys the on-device opponents' draws and the layout pool; default 0."""
        dev = self.device
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev).contiguous()
        ag, fo, ca = t(agents, torch.int32), t(food, torch.uint8), t(caps, torch.uint8)
        ep = np.zeros_like(np.asarray(score)) if episode is None else np.asarray(episode)
        misc = t(np.stack([np.asarray(score), np.asarray(steps), ep], axis=1), torch.int32)
        if layout_id is not None and self._layout_id is not None:
            self._layout_id.copy_(torch.as_tensor(np.asarray(layout_id), dtype=torch.int32))
        with torch.cuda.device(dev):
            _lib.check(self._L.pm_set_state(C.byref(self._desc), ag.data_ptr(), fo.data_ptr(), ca.data_ptr(),
                                            misc.data_ptr(), self._stream()), "pm_set_state")
            torch.cuda.current_stream(dev).synchronize()
        self._version += 1

    # -------------------------------------------------------- reference API
    def _legal_dict(self):
        if self.num_envs == 1:
            m = self._legal[0].cpu().tolist()
            return self._by_agent(lambda i: legal_mask_to_list(m[i]))
        return self._by_agent(lambda i: self._legal[:, i])

    def _by_agent(self, fn):
        return _AgentDict(self.agents, ((self.agents[i], fn(i)) for i in range(4)))

    def _obs_dict(self):
        if self.num_envs == 1:
            return self._by_agent(lambda i: self._obs[i, 0].clone())
        return self._by_agent(lambda i: self._obs[i].clone())

    def reset(self, layout_file="None", enemieName="None", mask=None, _first=False):
        """gymPacMan.reset (gymPacMan.py:92-141).  ``mask`` (batched extension)
        resets only the selected envs."""
        self.steps = 0
        self.num_moves = 0
        self._submoves_in_step = 0
        new_lays = None
        if os.path.exists(layout_file):
            new_lays = [_layout.getLayout(layout_file)]
            self._pool_mode = False
        if self.random_layout and not _first and not self._pool_mode:
            new_lays = [_layout.Layout(_maze.randomLayout().split("\n"))]      # one env: a new maze per reset, like the reference
        if new_lays is not None:
            self._install_layouts(new_lays)
        if enemieName != "None":
            if not self.self_play and self._bot_team is None and enemieName not in ON_DEVICE_TEAMS:
                raise NotImplementedError(f"enemieName={enemieName!r} is outside the GPU path")
            self.enemieName = enemieName
            self._set_opponent(enemieName)
        self.reset_batched(mask)
        if self._bot_team is not None:                 # gymPacMan.py:130-132: every bot sees the initial state
            if len(self._layouts) > 1:
                raise NotImplementedError("host-side bots keep one start position / distance table: bot_team mode "
                                          "needs a single layout for the whole batch")
            for b in self._bot_team:
                if hasattr(b, "registerInitialState"):
                    b.registerInitialState(self._bot_state(0))
        return self._obs_dict(), {"legal_actions": self._legal_dict()}

    def _gather_actions(self, actions) -> torch.Tensor:
        N = self.num_envs
        if isinstance(actions, torch.Tensor) and actions.dim() == 2:
            a = actions
        else:
            cols = []
            for i in range(4):
                if not self.self_play and i in (0, 2):
                    cols.append(None)
                    continue
                v = actions[self.agents[i]]
                if isinstance(v, str):
                    v = self._name_to_int[v]
                if isinstance(v, torch.Tensor):
                    cols.append(v.to(self.device).reshape(N).to(torch.int64))
                else:
                    cols.append(torch.full((N,), int(v), dtype=torch.int64, device=self.device))
            zero = torch.zeros((N,), dtype=torch.int64, device=self.device)
            a = torch.stack([c if c is not None else zero for c in cols], dim=1)   # bot columns are ignored in-kernel
        return a

    def _step_single(self, actions):
        """num_envs == 1 with scalar actions (the reference's own calling convention): the actions go into a pinned
        host word the kernel reads itself, the scalars come back in one pinned ``pm_host_record``
        (``pm_step_host_mapped``) - one launch, one clone of the observation block and one stream synchronise per
        step instead of a handful of small copies."""
        if not hasattr(self, "_single_act"):
            self._single_act = torch.zeros((1, 4), dtype=torch.uint8).pin_memory()
            self._single_np = self._single_act.numpy()
        for i in range(4):
            if not self.self_play and i in (0, 2):
                continue
            v = actions[self.agents[i]]
            if isinstance(v, str):
                v = self._name_to_int.get(v, 255)
            v = int(v)
            self._single_np[0, i] = v if 0 <= v <= 4 else 255          # out of range = illegal -> Stop, like the kernel
        # launch (the kernel reads the pinned action word and writes the pinned record), clone the observation block
        # on the same stream, and only then synchronise: one wait covers both
        self._ensure_mapped()
        self._mark_static()
        with torch.cuda.device(self.device):
            _lib.check(self._L.pm_step_host_mapped_async(C.byref(self._desc), self._single_act.data_ptr(),
                                                         C.byref(self._mapped_out), self._rec_mapped.data_ptr(),
                                                         self._stream()), "pm_step_host_mapped_async")
            o = self._obs.clone()
            torch.cuda.current_stream(self.device).synchronize()
        self._version += 1
        self.steps += 1
        r = self._rec_np[0]
        obs = self._by_agent(lambda i: o[i, 0])
        rb = (float(r["reward"][0]), float(r["reward"][1]))
        rewards = self._by_agent(lambda i: rb[i & 1])
        done = bool(r["done"])
        terms = self._by_agent(lambda i: done)
        legal = r["legal"]
        info = {"legal_actions": self._by_agent(lambda i: legal_mask_to_list(legal[i])),
                "score_change": int(r["score_change"])}
        return obs, rewards, terms, info

    # ---- host-side opponents: a step in two launches, each bot asked at its own sub-move ----
    def _bot_state(self, n):
        return self._state_factory(self, n) if self._state_factory is not None else _StateView(self, n)

    def _step_with_host_bots(self, actions):
        """gymPacMan.step with `bot.getAction(state)` for the red agents (gymPacMan.py:149-167): agent 0's bot sees the
        state before the step; after the launch that plays agents 0 and 1, agent 2's bot sees the state as it is
        then; the second launch plays agents 2 and 3 and finishes the step.  The device state is exported ONCE per
        sub-move launch for the whole batch (``cached_state``); every env's bot call reads from that copy."""
        N = self.num_envs
        a = self._gather_actions(actions).to(torch.int64).clone()             # blue columns; red filled below
        name_to_int = self._name_to_int
        out = None
        for first, (lo, hi) in ((0, (0, 2)), (2, (2, 4))):
            bot = self._bot_team[first >> 1]
            self._submoves_in_step = lo
            red = [bot.getAction(self._bot_state(n)) for n in range(N)]
            a[:, first] = torch.as_tensor([name_to_int[v] if isinstance(v, str) else int(v) for v in red],
                                          dtype=torch.int64, device=a.device)
            out = self._launch_step(a, (lo, hi))
            self.num_moves += hi - lo
        self._submoves_in_step = 0
        return out

    def _launch_step(self, actions, sub_moves=None):
        """One pm_step launch; ``sub_moves=(begin, end)`` plays only those agents' sub-moves (default: what the
        descriptor says, normally the whole step).  Counts a step when the launch finishes one."""
        d = self._desc
        keep = (d.agent_begin, d.agent_end)
        if sub_moves is not None:
            d.agent_begin, d.agent_end = sub_moves
        end = d.agent_end
        self._mark_static()
        try:
            with torch.cuda.device(self.device):
                _lib.check(self._L.pm_step(C.byref(d), actions.data_ptr(), actions.element_size(),
                                           C.byref(self._step_out), self._stream()), "pm_step")
        finally:
            d.agent_begin, d.agent_end = keep
        self._version += 1
        if end in (0, 4):
            self.steps += 1
        return self._out_dict()

    def _mark_static(self):
        """Set pm_step_out.obs_static_valid for the coming launch.  Valid while self._obs holds a full write for every
        env's CURRENT board: a step that skips the fp32 observations while boards may change (layout pool) ends it
        until the next full reset."""
        so = self._step_out
        if so.obs is None and self._pool_mode:
            self._obs_static_ok = False
        flag = int(self.elide_static_planes and self._obs_static_ok and so.obs is not None)
        so.obs_static_valid = flag
        for name in ("_mapped_out", "_pipe_out"):
            o = self.__dict__.get(name)
            if o is not None:
                o.obs_static_valid = flag

    def _out_dict(self):
        out = dict(obs=self._obs, reward=self._reward, done=self._done, legal=self._legal,
                   score_change=self._score_change, shaping=self._shaping, shaped_reward=self._shaped,
                   actions_used=self._actions_used)
        if getattr(self, "_views", None):
            out["view_actor"], out["view_critic"] = self._views[0], self._views[1]
        if getattr(self, "_bits", None) is not None:
            out["obs_bits"] = self._bits
        return out

    def step(self, actions):
        """gymPacMan.step (gymPacMan.py:143-193)."""
        if self._bot_team is not None:
            out = self._step_with_host_bots(actions)
        elif self.num_envs == 1 and not isinstance(actions, torch.Tensor) and not any(
                isinstance(actions[self.agents[i]], torch.Tensor) for i in range(4)
                if self.self_play or i in (1, 3)):
            return self._step_single(actions)
        else:
            out = self.step_batched(self._gather_actions(actions))
        obs = self._obs_dict()
        if self.num_envs == 1:
            rew = out["reward"][:, 0].cpu().tolist()
            done = bool(out["done"][0].item())
            rewards = self._by_agent(lambda i: rew[i & 1])
            terms = self._by_agent(lambda i: done)
            sc = int(out["score_change"][0].item())
        else:
            rb = (out["reward"][0].clone(), out["reward"][1].clone())
            rewards = self._by_agent(lambda i: rb[i & 1])
            d = out["done"].bool()
            terms = self._by_agent(lambda i: d)
            sc = out["score_change"].clone()
        return obs, rewards, terms, {"legal_actions": self._legal_dict(), "score_change": sc}

    def get_Observation(self, agentIndex):
        """gymPacMan.get_Observation (gymPacMan.py:195-229) on the current state."""
        o = self.observe_batched()
        return o[agentIndex, 0].clone() if self.num_envs == 1 else o[agentIndex].clone()
