"""ctypes front-end of the CPU oracle (oracle/pm_oracle.c).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpm_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.pmo_sizeof_env.restype = C.c_size_t
        L.pmo_env_init.restype = C.c_int
        L.pmo_env_init.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.pmo_reset.argtypes = [C.c_void_p]
        L.pmo_observe.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.pmo_step.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.pmo_reset_outputs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.pmo_get_state.argtypes = [C.c_void_p] * 5
        L.pmo_step_batch.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 7 + [C.c_int, C.c_int]
        L.pmo_step_batch_digest.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 8 + [C.c_int, C.c_int]
        L.pmo_step_batch_bots.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 10 + [C.c_int, C.c_int]
        L.pmo_sizeof_dist_table.restype = C.c_size_t
        L.pmo_build_dist_table.restype = C.c_int
        L.pmo_build_dist_table.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.pmo_set_distances.argtypes = [C.c_void_p, C.c_void_p]
        L.pmo_set_draws.argtypes = [C.c_void_p, C.c_void_p]
        L.pmo_max_threads.restype = C.c_int
        L.pmo_distances.restype = C.c_int
        L.pmo_distances.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.pmo_gae.argtypes = [C.c_void_p] * 4 + [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.pmo_adv_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _rows(layout_rows: Sequence[str]):
    rows = [str(r) for r in layout_rows]
    W, H = len(rows[0]), len(rows)
    assert all(len(r) == W for r in rows), "ragged layout text"
    return "".join(rows).encode("ascii"), W, H


class OracleBatch:
    """N independent environments on one layout (or one layout per env)."""

    def __init__(self, layout_rows, n_envs=1, length=299, reward_forLegalAction=True, defenceReward=True):
        L = lib()
        self.N = int(n_envs)
        self.esz = int(L.pmo_sizeof_env())
        self.buf = np.zeros(self.N * self.esz, np.uint8)
        per_env = isinstance(layout_rows[0], (list, tuple, np.ndarray)) and not isinstance(layout_rows[0], str)
        first = layout_rows[0] if per_env else layout_rows
        _, self.W, self.H = _rows(first)
        for n in range(self.N):
            text, W, H = _rows(layout_rows[n] if per_env else layout_rows)
            assert (W, H) == (self.W, self.H)
            rc = L.pmo_env_init(self._env(n), text, W, H, int(length), int(bool(reward_forLegalAction)),
                                int(bool(defenceReward)))
            if rc != 0:
                raise ValueError(f"pmo_env_init failed with {rc}")
        self.obs_shape = (8, self.H, self.W)

    def _env(self, n):
        return C.c_void_p(self.buf.ctypes.data + n * self.esz)

    def reset(self, with_obs=True):
        L = lib()
        obs = np.zeros((self.N, 4) + self.obs_shape, np.float32) if with_obs else None
        legal = np.zeros((self.N, 4), np.uint8)
        for n in range(self.N):
            L.pmo_reset(self._env(n))
            L.pmo_reset_outputs(self._env(n), _p(obs[n]) if with_obs else None, _p(legal[n]))
        return obs, legal

    def reset_one(self, n):
        lib().pmo_reset(self._env(n))

    def reset_outputs_for(self, idx):
        """obs [len(idx),4,8,H,W] and legal [len(idx),4] of the given (freshly reset) envs."""
        obs = np.zeros((len(idx), 4) + self.obs_shape, np.float32)
        legal = np.zeros((len(idx), 4), np.uint8)
        for k, n in enumerate(idx):
            lib().pmo_reset_outputs(self._env(int(n)), _p(obs[k]), _p(legal[k]))
        return obs, legal

    def step(self, actions, with_obs=True, auto_reset=False, nthreads=1):
        L = lib()
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int32).reshape(self.N, 4))
        out = dict(
            obs=np.empty((self.N, 4) + self.obs_shape, np.float32) if with_obs else None,
            reward=np.zeros((self.N, 2), np.float64),
            done=np.zeros(self.N, np.uint8),
            legal=np.zeros((self.N, 4), np.uint8),
            score_change=np.zeros(self.N, np.int32),
            shaping=np.zeros((self.N, 4), np.float64),
        )
        L.pmo_step_batch(C.c_void_p(self.buf.ctypes.data), self.N, _p(a), _p(out["obs"]), _p(out["reward"]),
                         _p(out["done"]), _p(out["legal"]), _p(out["score_change"]), _p(out["shaping"]),
                         int(auto_reset), int(nthreads))
        return out

    def step_digest(self, actions, weights, auto_reset=True, nthreads=0):
        """Like step(), but instead of the observations it returns their digests [N,4] uint64 (see pm_oracle.c,
        obs_digest; `weights` = digest_weights(8*H*W)) - for parity at batch sizes whose observations cannot be
        hauled to the host every step."""
        L = lib()
        assert weights.dtype == np.uint64 and weights.size >= 8 * self.W * self.H
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int32).reshape(self.N, 4))
        out = dict(digest=np.zeros((self.N, 4), np.uint64), reward=np.zeros((self.N, 2), np.float64),
                   done=np.zeros(self.N, np.uint8), legal=np.zeros((self.N, 4), np.uint8),
                   score_change=np.zeros(self.N, np.int32), shaping=np.zeros((self.N, 4), np.float64))
        L.pmo_step_batch_digest(C.c_void_p(self.buf.ctypes.data), self.N, _p(a), _p(weights), _p(out["digest"]), _p(out["reward"]),
                                _p(out["done"]), _p(out["legal"]), _p(out["score_change"]), _p(out["shaping"]),
                                int(auto_reset), int(nthreads))
        return out

    def set_bots(self, bot_kind):
        """bot_kind[4]: 0 = driven by the caller's actions, 2 / 3 = baselineTeam Offensive / Defensive reflex
        agent.  Builds the maze-distance table(s) the bots look up (one per distinct board)."""
        L = lib()
        self.bot_kind = np.ascontiguousarray(bot_kind, np.uint8)
        self._tables = {}
        if not any(int(k) >= 2 for k in self.bot_kind):
            return
        cap = self.W * self.H
        for n in range(self.N):
            key = bytes(self.buf[n * self.esz:(n + 1) * self.esz][:8 + 64 * 64])       # W, H, walls
            if key not in self._tables:
                t = np.zeros(int(L.pmo_sizeof_dist_table()), np.uint8)
                d = np.zeros(cap * cap, np.int32)
                assert L.pmo_build_dist_table(self._env(n), _p(t), _p(d), cap) > 0
                self._tables[key] = (t, d)
            L.pmo_set_distances(self._env(n), _p(self._tables[key][0]))

    def set_draws(self, draws):
        """draws [N,4] uint32 (or None): the index random.choice returned for each agent's bot decision of the NEXT
        step(s) - into the legal list (kind 1, randomTeam) or the best-action list (kinds 2/3, baselineTeam)."""
        L = lib()
        if draws is None:
            self._draws = None
            for n in range(self.N):
                L.pmo_set_draws(self._env(n), None)
            return
        self._draws = np.ascontiguousarray(np.asarray(draws).reshape(self.N, 4), np.uint32)
        for n in range(self.N):
            L.pmo_set_draws(self._env(n), C.c_void_p(self._draws.ctypes.data + 16 * n))

    def step_bots(self, actions, with_obs=True, auto_reset=False, nthreads=1):
        """Like step(), with the agents flagged by set_bots() choosing their own action at their sub-move
        (ties -> first best).  Adds actions_used [N,4] and best_masks [N,4] (bit a = action a is a best action)."""
        L = lib()
        a = np.ascontiguousarray(np.asarray(actions, dtype=np.int32).reshape(self.N, 4))
        out = dict(
            obs=np.empty((self.N, 4) + self.obs_shape, np.float32) if with_obs else None,
            reward=np.zeros((self.N, 2), np.float64), done=np.zeros(self.N, np.uint8),
            legal=np.zeros((self.N, 4), np.uint8), score_change=np.zeros(self.N, np.int32),
            shaping=np.zeros((self.N, 4), np.float64), actions_used=np.zeros((self.N, 4), np.int32),
            best_masks=np.zeros((self.N, 4), np.uint8))
        L.pmo_step_batch_bots(C.c_void_p(self.buf.ctypes.data), self.N, _p(a), _p(self.bot_kind), _p(out["obs"]),
                              _p(out["reward"]), _p(out["done"]), _p(out["legal"]), _p(out["score_change"]),
                              _p(out["shaping"]), _p(out["actions_used"]), _p(out["best_masks"]), int(auto_reset),
                              int(nthreads))
        return out

    def observe(self, n, agent):
        o = np.zeros(self.obs_shape, np.float32)
        lib().pmo_observe(self._env(n), int(agent), _p(o))
        return o

    def state(self, n=0):
        agents = np.zeros((4, 7), np.int32)
        food = np.zeros(self.W * self.H, np.uint8)
        caps = np.zeros(self.W * self.H, np.uint8)
        misc = np.zeros(3, np.int32)
        lib().pmo_get_state(self._env(n), _p(agents), _p(food), _p(caps), _p(misc))
        return dict(agents=agents, food=food, caps=caps, score=int(misc[0]), steps=int(misc[1]), win=int(misc[2]))


def digest_weights(n: int) -> np.ndarray:
    """Pseudo-random 64-bit weight per observation element (splitmix64 of the index)."""
    z = (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return np.ascontiguousarray(z ^ (z >> np.uint64(31)))


def max_threads() -> int:
    return int(lib().pmo_max_threads())


def distances(layout_rows):
    """-> (nodes[n,2] int16 in walls.asList(False) order, dist[n,n] int32, -1 = unreachable)."""
    env = OracleBatch(layout_rows, 1)
    cap = env.W * env.H
    out = np.zeros(cap * cap, np.int32)
    nodes = np.zeros((cap, 2), np.int16)
    n = lib().pmo_distances(env._env(0), _p(out), cap, _p(nodes))
    assert n >= 0
    return nodes[:n].copy(), out[: n * n].reshape(n, n).copy()


def gae(rew, val, done, last_value, gamma=0.99, lam=0.95):
    rew = np.ascontiguousarray(rew, np.float32)
    val = np.ascontiguousarray(val, np.float32)
    done = np.ascontiguousarray(done, np.float32)
    T, M = rew.shape
    last = np.ascontiguousarray(np.broadcast_to(np.asarray(last_value, np.float32), (M,)))
    adv = np.zeros((T, M), np.float32)
    ret = np.zeros((T, M), np.float32)
    lib().pmo_gae(_p(rew), _p(val), _p(done), _p(last), float(gamma), float(lam), _p(adv), _p(ret), T, M)
    return adv, ret


def adv_norm(adv, inds, mb):
    adv = np.ascontiguousarray(adv, np.float32)
    inds = np.ascontiguousarray(inds, np.int64)
    out = np.zeros(len(inds), np.float32)
    lib().pmo_adv_norm(_p(adv), _p(inds), _p(out), len(inds), int(mb))
    return out
