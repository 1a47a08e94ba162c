"""Property tests on the kernel logic, CPU-only through the emulation: for random boards
(mazeGenerator seeds and the shipped layouts), random action streams and random option mixes, every step keeps
  * pellet conservation: food on board + carried + returned == total food of the layout,
  * legal mask == open neighbours of the agent's cell (+ Stop),
  * observation structure: plane 0 == walls, exactly one own cell with value 1 + carrying, one ally cell,
    <= 2 enemy cells, food planes == board food split at W//2, rewards >= 0 and finite."""
import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from emu_driver import EmuEnv
from parity_bodies import lay_rows

BOARDS = {}


def board(kind):
    if kind not in BOARDS:
        if isinstance(kind, int):
            from pacman_marl_2025_b200 import mazeGenerator
            BOARDS[kind] = mazeGenerator.generateMaze(kind).split("\n")
        else:
            BOARDS[kind] = lay_rows(kind)
    return BOARDS[kind]


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(kind=st.one_of(st.sampled_from(["tinyCapture", "smallCapture"]), st.integers(1, 40)),
       seed=st.integers(0, 2 ** 31 - 1), auto_reset=st.booleans(), bots=st.sampled_from([0, 0b0101, 0b1010]),
       p_fwd=st.floats(0.2, 0.8))
def test_step_invariants(kind, seed, auto_reset, bots, p_fwd):
    rows = board(kind)
    H, W = len(rows), len(rows[0])
    walls = np.array([[rows[H - 1 - y][x] == "%" for x in range(W)] for y in range(H)])
    total = sum(r.count(".") for r in rows)
    N = 6
    env = EmuEnv(rows, N, length=40, auto_reset=auto_reset, random_agent_mask=bots, random_seed=seed, E=8, KB=1)
    env.reset()
    rng = np.random.default_rng(seed)
    for t in range(60):
        a = rng.integers(0, 5, size=(N, 4))
        fwd = rng.random((N, 4)) < p_fwd
        a[:, 0::2] = np.where(fwd[:, 0::2], 1, a[:, 0::2])
        a[:, 1::2] = np.where(fwd[:, 1::2], 3, a[:, 1::2])
        o = env.step(a.astype(np.uint8))
        s = env.get_state()
        ag = s["agents"]
        assert (s["food"].sum(1) + ag[:, :, 5].sum(1) + ag[:, :, 6].sum(1) == total).all()
        assert np.isfinite(o["reward"]).all() and (o["reward"] >= 0).all()
        for n in range(N):
            food = s["food"][n].reshape(H, W)
            for i in range(4):
                x, y = int(ag[n, i, 0]), int(ag[n, i, 1])
                want = (not walls[y + 1, x]) | (not walls[y, x + 1]) << 1 | (not walls[y - 1, x]) << 2 | \
                       (not walls[y, x - 1]) << 3 | 16
                assert o["legal"][n, i] == want
            done = bool(o["done"][n])
            if done and not auto_reset:
                continue
            for i in range(4):
                ob = o["obs"][i, n]
                assert (ob[0] == walls).all()
                assert (ob[1] > 0).sum() == 1 and (ob[4] > 0).sum() == 1 and 1 <= (ob[5] > 0).sum() <= 2
            # the LAST agent's snapshot is the final state: food planes and own cell must match it
            ob = o["obs"][3, n]
            assert (ob[6][:, W // 2:] == food[:, W // 2:]).all() and (ob[7][:, :W // 2] == food[:, :W // 2]).all()
            assert ob[6][:, :W // 2].sum() == 0 and ob[7][:, W // 2:].sum() == 0
            x, y = int(ag[n, 3, 0]), int(ag[n, 3, 1])
            assert ob[1][y, x] == 1 + ag[n, 3, 5]
        if o["done"].any() and not auto_reset:
            env.reset(mask=o["done"])
