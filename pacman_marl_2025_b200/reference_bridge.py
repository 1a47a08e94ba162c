"""Bridge to opponents that run on the HOST (the reference's `agents/*.py` teams).

The reference's search bots (MCTSTeam, AstarTeam, heuristicTeam, approxQTeam, ...) are CPU programs that take a
`capture.GameState` at their own sub-move (`gymPacMan.py:157`: ``bot.getAction(state)``) and call
``generateSuccessor`` on it.  They are outside the GPU path, but the env can still *serve* them:
``gymPacMan_parallel_env(bot_team=..., state_factory=...)`` plays a step in several launches
(``pm_env_desc.agent_begin / agent_end``) and asks each host bot for its move from the live state in between.

This module builds the object those bots expect — the reference's OWN ``GameState`` — from the device state, using
the reference's own classes.  It therefore needs a checkout of the reference on disk (``reference_root``); nothing of
the reference is vendored here.  The rules the bots run inside ``generateSuccessor`` are then the reference's Python
code on the reference's object, not a CPU path of this package.

    mods = load_reference("/path/to/pacman-marl-2025")
    team = make_bot_team(mods, "heuristicTeam")                       # capture.loadAgents(True, ...): agents 0 and 2
    env = gymPacMan_parallel_env(layout_file=..., self_play=False, bot_team=team, state_factory=state_factory(mods))
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
from types import SimpleNamespace
from typing import Sequence

import numpy as np

ACTION_NAMES = ["North", "East", "South", "West", "Stop"]
NAME_TO_INT = {n: i for i, n in enumerate(ACTION_NAMES)}


def load_reference(reference_root: str) -> SimpleNamespace:
    """Import the reference's modules from ``reference_root`` (its top-level directory)."""
    root = os.path.abspath(reference_root)
    if not os.path.exists(os.path.join(root, "capture.py")):
        raise FileNotFoundError(f"{root} does not look like a pacman-marl-2025 checkout (no capture.py)")
    if root not in sys.path:
        sys.path.insert(0, root)
    with contextlib.redirect_stdout(io.StringIO()):
        mods = SimpleNamespace(root=root, capture=importlib.import_module("capture"), game=importlib.import_module("game"),
                               layout=importlib.import_module("layout"))
    return mods


def make_bot_team(mods: SimpleNamespace, name: str):
    """The two red bots of team ``name`` (agents 0 and 2), created like gymPacMan.py:31 does."""
    with contextlib.redirect_stdout(io.StringIO()):
        bots = mods.capture.loadAgents(True, os.path.join(mods.root, "agents", name), None, "")
    if bots is None or any(b is None for b in bots):
        raise ImportError(f"team {name!r} could not be loaded from {mods.root}/agents")
    return list(bots)


def to_reference_state(mods: SimpleNamespace, layout_rows: Sequence[str], agents: np.ndarray, food: np.ndarray,
                       caps: np.ndarray, score: int, timeleft: int):
    """A reference ``capture.GameState`` holding the given env state.

    agents [4,7] = x, y, dir, isPacman, scaredTimer, numCarrying, numReturned (``export_state()['agents'][n]``);
    food / caps [H*W] bytes, cell y*W + x.  ``timeleft`` as the reference counts it: ``length`` minus the
    sub-moves played so far (`capture.py:122`, `:377`)."""
    lay = mods.layout.Layout(list(layout_rows))
    st = mods.capture.GameState()
    st.initialize(lay, 4)
    W, H = lay.width, lay.height
    d = st.data
    f = np.asarray(food).reshape(H, W)
    for x in range(W):
        col = d.food.data[x]
        for y in range(H):
            col[y] = bool(f[y, x])
    c = np.asarray(caps).reshape(H, W)
    d.capsules = [p for p in lay.capsules if c[p[1], p[0]]]
    for i, a in enumerate(d.agentStates):
        x, y, direction, pac, scared, carry, ret = (int(v) for v in agents[i])
        a.configuration = mods.game.Configuration((x, y), ACTION_NAMES[direction])
        a.isPacman = bool(pac)
        a.scaredTimer = scared
        a.numCarrying = carry
        a.numReturned = ret
    d.score = int(score)
    d.timeleft = int(timeleft)
    return st


def state_factory(mods: SimpleNamespace):
    """``state_factory(env, n) -> capture.GameState`` for :class:`gymPacMan_parallel_env` (host-bot mode).
    Reads env ``n`` out of ``env.cached_state()`` - ONE export of the device state per sub-move launch for the whole
    batch - so asking every env's bot costs O(N), not O(N^2).  ``timeleft`` counts down from ``length`` by the
    sub-moves this env has played in its current episode (`capture.py:122`, `:377`)."""
    def make(env, n=0):
        s = env.cached_state()
        lay = env._layout_of(n)
        played = 4 * int(s["steps"][n]) + int(getattr(env, "_submoves_in_step", 0))
        return to_reference_state(mods, lay.layoutText, s["agents"][n], s["food"][n], s["caps"][n], int(s["score"][n]),
                                  env.length - played)
    return make


def action_to_int(a) -> int:
    return NAME_TO_INT[a] if isinstance(a, str) else int(a)
