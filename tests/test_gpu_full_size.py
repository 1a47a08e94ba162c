"""Parity at the batch sizes BASELINE.json names — the CUDA path through the C ABI against the CPU oracle on the
same seeded action streams, at every step:

  config 2  smallCapture, 65,536 envs x 301 steps (length 300, auto_reset)
  config 3  bloxCapture, 16,384 envs x 301 steps
  config 4  1,024 distinct mazeGenerator boards over 8,192 envs x 150 steps (explicit layout ids), and all 1,024 BFS
            distance tables against distanceCalculator.computeDistances as restated by the oracle

Rewards (fp64 bit patterns), done, legal masks, score changes and shaping are compared in full at every step.  The
observations of such a batch are 1.3 - 3.4 GB per step and never cross PCIe: both sides reduce every observation to
a 64-bit digest  sum_i bits(o[i]) * K[i]  (K = oracle.digest_weights, pseudo-random per element) - the oracle in C
over its own float planes, the GPU side with integer tensor ops over the kernel's output - and the [N,4] digests are
compared; a sample of envs is also compared element by element.  The packed state is compared in full at checkpoints.
"""
import numpy as np
import pytest
import torch

from oracle import oracle
from parity_bodies import biased_actions, lay_rows

pytestmark = pytest.mark.gpu


def gpu_obs_digest(obs4, Kt):
    """obs4 [4,N,8,H,W] f32 on the device -> [N,4] int64 digests (two's-complement wrap == mod 2^64)."""
    out = []
    for k in range(4):
        bits = obs4[k].reshape(obs4.shape[1], -1).view(torch.int32).to(torch.int64)
        out.append((bits * Kt).sum(dim=1))
    return torch.stack(out, dim=1)


def run_full_size(layouts, N, T, length, layout_ids=None, state_every=50, sample=61):
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    per_env = not isinstance(layouts[0], str)
    lays = layouts if per_env else [layouts]
    env = gymPacMan_parallel_env(layouts=lays, self_play=True, length=length, num_envs=N, auto_reset=True,
                                 layout_ids=layout_ids)
    if per_env:
        ids = np.asarray(layout_ids) if layout_ids is not None else np.arange(N) % len(lays)
        ref = oracle.OracleBatch([lays[i] for i in ids], N, length=length)
    else:
        ref = oracle.OracleBatch(layouts, N, length=length)
    H, W = len(lays[0]), len(lays[0][0])
    K = oracle.digest_weights(8 * H * W)
    Kt = torch.from_numpy(K.view(np.int64)).cuda()
    obs, legal = env.reset_batched()
    robs, rlegal = ref.reset(with_obs=False)
    np.testing.assert_array_equal(legal.cpu().numpy(), rlegal)
    rng = np.random.default_rng(N + T)
    idx = np.arange(0, N, sample)
    events = dict(done=0, score=0, eaten=0)
    for t in range(T):
        a = biased_actions(rng, N, 0.6, t)
        o = env.step_batched(torch.from_numpy(a).cuda())
        dig = gpu_obs_digest(o["obs"], Kt)
        r = ref.step_digest(a, K, auto_reset=True, nthreads=0)
        np.testing.assert_array_equal(o["reward"].cpu().numpy().view(np.uint64), r["reward"].T.copy().view(np.uint64),
                                      err_msg=f"t={t} reward")
        np.testing.assert_array_equal(o["done"].cpu().numpy(), r["done"], err_msg=f"t={t} done")
        np.testing.assert_array_equal(o["legal"].cpu().numpy(), r["legal"], err_msg=f"t={t} legal")
        np.testing.assert_array_equal(o["score_change"].cpu().numpy(), r["score_change"], err_msg=f"t={t} score_change")
        np.testing.assert_array_equal(o["shaping"].cpu().numpy().view(np.uint64), r["shaping"].T.copy().view(np.uint64),
                                      err_msg=f"t={t} shaping")
        np.testing.assert_array_equal(dig.cpu().numpy().view(np.uint64), r["digest"], err_msg=f"t={t} observation digests")
        events["done"] += int(r["done"].sum())
        events["score"] += int((r["score_change"] != 0).sum())
        if t % state_every == state_every - 1 or t == T - 1:
            s = env.export_state()
            for n in idx:
                rs = ref.state(int(n))
                np.testing.assert_array_equal(s["agents"][n], rs["agents"], err_msg=f"t={t} env {n}")
                np.testing.assert_array_equal(s["food"][n], rs["food"])
                np.testing.assert_array_equal(s["caps"][n], rs["caps"])
                assert s["score"][n] == rs["score"] and s["steps"][n] == rs["steps"]
            # and element by element for the sampled envs (current state of every agent's view)
            cur = env.observe_batched()[:, torch.from_numpy(idx).cuda()].cpu().numpy()
            for j, n in enumerate(idx[:64]):
                for i in range(4):
                    np.testing.assert_array_equal(cur[i, j], ref.observe(int(n), i), err_msg=f"t={t} env {n} agent {i}")
            events["eaten"] += int((s["agents"][:, :, 5].sum(1) + s["agents"][:, :, 6].sum(1) > 0).sum())
    return events


def test_config2_smallCapture_65536_envs_full_episodes_vs_oracle():
    ev = run_full_size(lay_rows("smallCapture"), 65536, 301, 300, state_every=75)
    assert ev["done"] >= 65536 and ev["score"] > 0 and ev["eaten"] > 0, ev


def test_config3_bloxCapture_16384_envs_full_episodes_vs_oracle():
    ev = run_full_size(lay_rows("bloxCapture"), 16384, 301, 300, state_every=75)
    assert ev["done"] >= 16384 and ev["eaten"] > 0, ev


def _mazes(L, first_seed=5000):
    from pacman_marl_2025_b200 import mazeGenerator
    return [mazeGenerator.generateMaze(first_seed + s).split("\n") for s in range(L)]


def test_config4_1024_distinct_mazes_8192_envs_vs_oracle():
    mazes = _mazes(1024)
    ids = np.random.default_rng(4).permutation(8192) % 1024           # every board is played by eight envs
    ev = run_full_size(mazes, 8192, 150, 100, layout_ids=ids, state_every=50, sample=37)
    assert ev["done"] >= 8192 and ev["eaten"] > 0, ev


def test_config4_all_1024_bfs_tables_vs_oracle():
    """One pm_distances launch over the 1,024 boards the bench times; every table against the oracle's
    priority-queue UCS (distanceCalculator.py:111-150)."""
    from pacman_marl_2025_b200.distanceCalculator import compute_distances
    from pacman_marl_2025_b200.layout import Layout
    mazes = _mazes(1024)
    t = compute_distances([Layout(m) for m in mazes])
    n_open = t.n_open.cpu().numpy()
    dist = (t.dist.to(torch.int32) & 0xFFFF).cpu().numpy()
    node_of = t.node_of_cell.cpu().numpy()
    W = len(mazes[0][0])
    for l, m in enumerate(mazes):
        nodes, od = oracle.distances(m)
        n = len(nodes)
        assert n_open[l] == n
        got = dist[l, :n, :n]
        np.testing.assert_array_equal(np.where(got == 0xFFFF, -1, got), od, err_msg=f"maze {l}")
        np.testing.assert_array_equal(node_of[l, nodes[:, 1].astype(int) * W + nodes[:, 0].astype(int)], np.arange(n))
