"""Record / replay of whole games - the move-list format of the reference's tournament loop.

Reference: ``capture.runGames(record=True)`` pickles ``{'layout', 'agents', 'actions', 'length', 'redTeamName',
'blueTeamName'}`` where ``actions`` is ``Game.moveHistory``, a list of ``(agentIndex, direction name)`` in the order the
moves were played (`capture.py:988-996`, `game.py:570-732`); ``capture.replayGame`` (`capture.py:952-968`) applies them
one ``generateSuccessor`` at a time.  Here a recorded game is replayed on the GPU env one sub-move per launch
(``pm_env_desc.agent_begin / agent_end``), for one env or a whole batch of copies, and games played on the GPU env can
be written in the same format.

Reading a replay written by the reference needs no reference checkout: its pickled ``layout.Layout`` / ``game.Grid`` /
``game.Agent`` objects are mapped to attribute bags and the board is rebuilt from ``Layout.layoutText``.
"""
from __future__ import annotations

import io
import pickle
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

NAME_TO_INT = {"North": 0, "East": 1, "South": 2, "West": 3, "Stop": 4}
INT_TO_NAME = {v: k for k, v in NAME_TO_INT.items()}


class _Bag:
    """Stands in for any class of the reference while unpickling: keeps the attributes, runs no code."""

    def __init__(self, *a, **kw):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self.__dict__["state"] = state


class _TolerantUnpickler(pickle.Unpickler):
    SAFE_BUILTINS = {"list", "dict", "tuple", "set", "frozenset", "int", "float", "str", "bool", "bytes", "complex", "object"}

    def find_class(self, module, name):
        if module == "builtins" and name in self.SAFE_BUILTINS:
            return super().find_class(module, name)
        if module == "copyreg" and name == "_reconstructor":
            return lambda cls, base, state: cls()
        return type(name, (_Bag,), {"__module__": module})     # layout.Layout, game.Grid, game.Agent, ...


def load_replay(source) -> Dict:
    """``source``: a path or the bytes of a pickle written by ``capture.runGames(record=True)`` or by
    :func:`save_replay`.  Returns ``dict(layout_rows, actions, length, redTeamName, blueTeamName)`` with ``actions`` a
    list of ``(agentIndex, gym action int)``."""
    data = source if isinstance(source, (bytes, bytearray)) else open(source, "rb").read()
    rec = _TolerantUnpickler(io.BytesIO(data)).load()
    if not isinstance(rec, dict) or "actions" not in rec or "layout" not in rec:
        raise ValueError("not a capture replay: expected a dict with 'layout' and 'actions'")
    lay = rec["layout"]
    rows = lay if isinstance(lay, (list, tuple)) else getattr(lay, "layoutText", None)
    if rows is None:
        raise ValueError("replay layout carries no layoutText")
    acts = []
    for idx, a in rec["actions"]:
        acts.append((int(idx), NAME_TO_INT[a] if isinstance(a, str) else int(a)))
    return dict(layout_rows=[str(r) for r in rows], actions=acts, length=int(rec.get("length", 1200)),
                redTeamName=rec.get("redTeamName", "red"), blueTeamName=rec.get("blueTeamName", "blue"))


def save_replay(path: Optional[str], layout_rows: Sequence[str], actions: Sequence[Tuple[int, int]], length: int,
                redTeamName: str = "red", blueTeamName: str = "blue", reference_layout=None) -> bytes:
    """Write a game in the reference's replay format.  ``actions``: ``(agentIndex, action)`` with gym ints or direction
    names.  ``reference_layout``: a ``layout.Layout`` object of the reference (pass it when the file has to be opened by
    the reference's own ``capture.py --replay``); by default the board travels as its text rows, which is what
    :func:`load_replay` needs."""
    comp = {"layout": reference_layout if reference_layout is not None else [str(r) for r in layout_rows],
            "agents": [None] * 4,
            "actions": [(int(i), a if isinstance(a, str) else INT_TO_NAME[int(a)]) for i, a in actions],
            "length": int(length), "redTeamName": redTeamName, "blueTeamName": blueTeamName}
    data = pickle.dumps(comp)
    if path:
        with open(path, "wb") as f:
            f.write(data)
    return data


def replay_game(replay: Dict, num_envs: int = 1, device=None, on_move=None):
    """``capture.replayGame`` on the GPU env: every recorded move is one launch that plays exactly that agent's
    sub-move (GameState.generateSuccessor, `capture.py:107-123`).  ``on_move(k, agentIndex, action, env)`` is called
    after move k (e.g. to look at ``env.export_state()``).  Returns the env; ``env.game.state.data.score`` is the
    final score."""
    from .gymPacMan import gymPacMan_parallel_env
    from .layout import Layout
    env = gymPacMan_parallel_env(layouts=[Layout(list(replay["layout_rows"]))], self_play=True, length=10 ** 9,
                                 num_envs=num_envs, device=device, reward_forLegalAction=False, defenceReward=False)
    env.reset_batched()
    a = torch.zeros((num_envs, 4), dtype=torch.uint8, device=env.device)
    # CaptureRules.newGame draws the first mover (capture.py:372-374: `starter = random.randint(0,1)`), so a recorded
    # game may open with agent 1; after that the turn order is cyclic (game.py:700-708)
    expect = replay["actions"][0][0] if replay["actions"] else 0
    for k, (idx, act) in enumerate(replay["actions"]):
        if idx != expect:
            raise ValueError(f"move {k}: agent {idx} moves out of turn (expected agent {expect})")
        a[:, idx] = act
        env._launch_step(a, (idx, idx + 1))
        expect = (idx + 1) % 4
        if on_move is not None:
            on_move(k, idx, act, env)
    return env


class Recorder:
    """Collects the move history of ONE env of a batched run (``actions_used`` of every step) in replay order."""

    def __init__(self, env, n: int = 0):
        self.env, self.n, self.moves = env, n, []

    def after_step(self, out):
        used = out["actions_used"][self.n].tolist()
        self.moves += [(i, int(used[i])) for i in range(4)]

    def save(self, path=None, **kw):
        return save_replay(path, self.env._layout_of(self.n).layoutText, self.moves, self.env.length, **kw)


if __name__ == "__main__":
    import sys
    rp = load_replay(sys.argv[1])
    e = replay_game(rp)
    print(f"{len(rp['actions'])} moves on a {e.layout.width}x{e.layout.height} board: final score {e.game.state.data.score}")
