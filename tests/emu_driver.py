"""TEST-ONLY numpy driver of tests/emu/libpm_emu.so (host emulation of the CUDA
kernels' per-thread logic, see tests/emu/pm_emu.cpp).  Not a product path."""
import ctypes as C
import os
import subprocess

import numpy as np

from pacman_marl_2025_b200 import _lib as pmlib

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "emu", "pm_emu.cpp")
_SO = os.path.join(_HERE, "emu", "libpm_emu.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "pacman_marl_2025_b200", "csrc")
_L = None


def lib():
    global _L
    if _L is None:
        deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("pm_env_core.cuh", "pm_format.h", "pm_pack.cpp")]
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(d) for d in deps):
            subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared",
                                   "-Wno-unknown-pragmas", "-o", _SO, _SRC, os.path.join(_CSRC, "pm_pack.cpp")])
        _L = pmlib.declare_pack_api(C.CDLL(_SO))
    return _L


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class EmuEnv:
    def __init__(self, layouts, n_envs, length=299, reward_forLegalAction=True, defenceReward=True,
                 shaping_scale=0.1, E=64, KB=4, layout_id=None, random_agent_mask=0, random_seed=0,
                 auto_reset=False, env_index_base=0, bot_kind=None, bot_first_best=True, layout_ids=None,
                 layout_pool=False, bot_draws=False):
        layout_id = layout_ids if layout_ids is not None else layout_id
        L = lib()
        per_env = not isinstance(layouts[0], str)
        lays = layouts if per_env else [layouts]
        self.W, self.H = len(lays[0][0]), len(lays[0])
        self.N, self.E, self.KB = int(n_envs), E, KB
        self.geom = np.zeros(L.pm_geom_words(self.W, self.H), np.uint32)
        assert L.pm_pack_geom(self.W, self.H, _p(self.geom)) == 0
        LS = L.pm_layout_words(self.W, self.H)
        self.layouts = np.zeros((len(lays), LS), np.uint32)
        has_caps = 0
        for i, rows in enumerate(lays):
            hc = C.c_int(0)
            rc = L.pm_pack_layout("".join(rows).encode(), self.W, self.H, _p(self.layouts[i]), C.byref(hc))
            assert rc == 0, rc
            has_caps |= hc.value
        self.state = np.zeros(L.pm_state_bytes(self.W, self.H, self.N), np.uint8)
        self.layout_id = None if layout_id is None else np.ascontiguousarray(layout_id, np.uint32)
        if layout_pool:
            self.layout_id = np.zeros(self.N, np.uint32)
        self.desc = pmlib.PmEnvDesc(self.W, self.H, len(lays), has_caps, self.N, self.geom.ctypes.data,
                                    self.layouts.ctypes.data,
                                    None if self.layout_id is None else self.layout_id.ctypes.data,
                                    self.state.ctypes.data, int(length), int(reward_forLegalAction),
                                    int(defenceReward), int(random_agent_mask), float(shaping_scale), int(random_seed),
                                    int(auto_reset), 0, int(env_index_base), 0, 0)
        self.desc.layout_pool = int(bool(layout_pool))
        self.draws = np.zeros((self.N, 4), np.uint32) if bot_draws else None
        if bot_draws:
            self.desc.bot_draws = self.draws.ctypes.data
        self.want_bits = False
        if bot_kind is not None and any(k >= 2 for k in bot_kind):
            # maze-distance tables in pm_distances' output format, built here by the ORACLE (the CUDA BFS kernel
            # cannot run in the CPU emulation; it is checked against the same oracle in the GPU tests)
            from oracle import oracle
            HW = self.W * self.H
            tabs = [oracle.distances(rows) for rows in lays]
            nmax = (max(len(nodes) for nodes, _ in tabs) + 7) & ~7
            self.dist = np.full((len(lays), nmax, nmax), 0xFFFF, np.uint16)
            self.node_of = np.full((len(lays), HW), -1, np.int16)
            for l, (nodes, d) in enumerate(tabs):
                n = len(nodes)
                self.dist[l, :n, :n] = np.where(d < 0, 0xFFFF, d).astype(np.uint16)
                self.node_of[l, nodes[:, 1].astype(int) * self.W + nodes[:, 0].astype(int)] = np.arange(n)
            self.desc.dist, self.desc.node_of_cell, self.desc.dist_nmax = self.dist.ctypes.data, self.node_of.ctypes.data, nmax
        if bot_kind is not None:
            self.desc.bot_kind = sum(int(k) << (4 * i) for i, k in enumerate(bot_kind))
            self.desc.bot_first_best = int(bot_first_best)
        self.obs_shape = (8, self.H, self.W)

    def make_observation(self, draws=None):
        vis = np.zeros((self.N, 4), np.uint8)
        noisy = np.zeros((self.N, 4, 4), np.int32)
        d = None if draws is None else np.ascontiguousarray(np.asarray(draws).reshape(self.N, 4, 4), np.uint32)
        rc = lib().emu_make_observation(C.byref(self.desc), _p(d), _p(vis), _p(noisy))
        assert rc == 0, rc
        return vis, noisy

    def set_elide(self, on=True):
        """pm_step_out.obs_static_valid: steps write into ONE persistent observation buffer (first filled by reset) and
        may leave the static planes untouched."""
        self.elide = bool(on)

    def reset(self, mask=None):
        if getattr(self, "elide", False):
            if getattr(self, "_obs_persist", None) is None:
                self._obs_persist = np.full((4, self.N) + self.obs_shape, np.nan, np.float32)
            obs = self._obs_persist
            legal = np.zeros((self.N, 4), np.uint8)
            m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
            lib().emu_reset(C.byref(self.desc), _p(m), _p(obs), C.c_int64(obs[0].size), _p(legal), self.E, self.KB)
            return obs.copy(), legal
        obs = np.full((4, self.N) + self.obs_shape, np.nan, np.float32)
        legal = np.zeros((self.N, 4), np.uint8)
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        lib().emu_reset(C.byref(self.desc), _p(m), _p(obs), C.c_int64(obs[0].size), _p(legal), self.E, self.KB)
        return obs, legal

    def observe(self):
        obs = np.full((4, self.N) + self.obs_shape, np.nan, np.float32)
        lib().emu_observe(C.byref(self.desc), _p(obs), C.c_int64(obs[0].size), self.E, self.KB)
        return obs

    def set_views(self, red=None):
        """red=None: views off; False / True: learner team blue (agents 1,3) / red (agents 0,2, mirrored)."""
        self.view_red = red

    def observe_views(self, red):
        va = np.full((2, self.N) + self.obs_shape, 0xEE, np.uint8)
        vc = np.full((self.N,) + self.obs_shape, 0xEE, np.uint8)
        rc = lib().emu_observe_views(C.byref(self.desc), int(red), _p(va), C.c_int64(va[0].size), _p(vc), self.E, self.KB)
        assert rc == 0, rc
        return va, vc

    def set_agents(self, begin=0, end=0):
        """Play only the sub-moves of agents [begin, end) in the following step launches (0, 0 = whole step)."""
        self.desc.agent_begin, self.desc.agent_end = int(begin), int(end)

    def set_range(self, begin=0, count=0):
        self.desc.env_begin, self.desc.env_count = int(begin), int(count)

    def pool_seed(self):
        return int(self.desc.random_seed)

    def set_draws(self, draws):
        """random.choice indices for the on-device bots of the next step (pm_env_desc.bot_draws)."""
        self.draws[:] = np.asarray(draws, np.uint32).reshape(self.N, 4)

    def set_bits(self, on=True):
        self.want_bits = bool(on)

    def observe_bits(self):
        NB = lib().pm_obs_bits_words(self.W, self.H)
        bits = np.full((4, self.N, NB), 0xEEEEEEEE, np.uint32)
        rc = lib().emu_observe_bits(C.byref(self.desc), _p(bits), C.c_int64(bits[0].size), self.E, self.KB)
        assert rc == 0, rc
        return bits

    def step(self, actions, dtype=np.uint8):
        a = np.ascontiguousarray(np.asarray(actions).reshape(self.N, 4), dtype)
        elide = getattr(self, "elide", False)
        o = dict(obs=self._obs_persist if elide else np.full((4, self.N) + self.obs_shape, np.nan, np.float32), reward=np.zeros((2, self.N)),
                 done=np.zeros(self.N, np.uint8), legal=np.zeros((self.N, 4), np.uint8),
                 score_change=np.zeros(self.N, np.int32), shaping=np.zeros((4, self.N)),
                 shaped=np.zeros((4, self.N), np.float32))
        self.last_packed = np.zeros((self.N, 48), np.uint8)           # struct pm_host_record per env
        o["actions_used"] = np.zeros((self.N, 4), np.uint8)
        so = pmlib.PmStepOut(o["obs"].ctypes.data, o["obs"][0].size, o["reward"].ctypes.data, o["done"].ctypes.data,
                             o["legal"].ctypes.data, o["score_change"].ctypes.data, o["shaping"].ctypes.data,
                             o["shaped"].ctypes.data, self.last_packed.ctypes.data, o["actions_used"].ctypes.data)
        if getattr(self, "view_red", None) is not None:
            o["view_actor"] = np.full((2, self.N) + self.obs_shape, 0xEE, np.uint8)
            o["view_critic"] = np.full((self.N,) + self.obs_shape, 0xEE, np.uint8)
            so.view_actor, so.view_actor_stride = o["view_actor"].ctypes.data, o["view_actor"][0].size
            so.view_critic, so.view_red = o["view_critic"].ctypes.data, int(self.view_red)
        if self.want_bits:
            NB = lib().pm_obs_bits_words(self.W, self.H)
            o["obs_bits"] = np.full((4, self.N, NB), 0xEEEEEEEE, np.uint32)
            so.obs_bits, so.obs_bits_agent_stride = o["obs_bits"].ctypes.data, o["obs_bits"][0].size
        so.obs_static_valid = int(elide)
        rc = lib().emu_step(C.byref(self.desc), _p(a), a.dtype.itemsize, C.byref(so), self.E, self.KB)
        assert rc == 0, rc
        if elide:
            o["obs"] = o["obs"].copy()
        return o

    def get_state(self):
        HW = self.W * self.H
        ag = np.zeros((self.N, 4, 7), np.int32)
        food = np.zeros((self.N, HW), np.uint8)
        caps = np.zeros((self.N, HW), np.uint8)
        misc = np.zeros((self.N, 3), np.int32)
        lib().emu_get_state(C.byref(self.desc), _p(ag), _p(food), _p(caps), _p(misc))
        out = dict(agents=ag, food=food, caps=caps, score=misc[:, 0], steps=misc[:, 1], episode=misc[:, 2])
        if self.layout_id is not None:
            out["layout_id"] = self.layout_id.astype(np.int32)
        return out


class SingleEnvDriver:
    """Adapts a batched env (emu or CUDA) to helpers.replay, looking at env `n`
    of the batch while every env receives the same actions."""

    def __init__(self, env, n=0):
        self.env, self.n = env, n

    def reset(self):
        obs, legal = self.env.reset()
        return np.asarray(obs)[:, self.n], np.asarray(legal)[self.n]

    def step(self, a):
        o = self.env.step(np.tile(np.asarray(a, np.uint8), (self.env.N, 1)))
        n = self.n
        return dict(obs=np.asarray(o["obs"])[:, n], reward=np.asarray(o["reward"])[:, n], done=o["done"][n],
                    legal=np.asarray(o["legal"])[n], score_change=o["score_change"][n],
                    shaping=np.asarray(o["shaping"])[:, n], shaped=np.asarray(o["shaped"])[:, n],
                    actions_used=None if o.get("actions_used") is None else np.asarray(o["actions_used"])[n])

    def set_draws(self, d4):
        self.env.set_draws(np.tile(np.asarray(d4, np.uint32), (self.env.N, 1)))

    def state(self):
        s = self.env.get_state()
        n = self.n
        return dict(agents=s["agents"][n], food=s["food"][n], caps=s["caps"][n], score=int(s["score"][n]))
