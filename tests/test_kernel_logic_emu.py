"""Kernel LOGIC parity without a GPU: the per-thread phase functions of the
CUDA kernels (csrc/pm_env_core.cuh) are compiled for the host by tests/emu and
replayed against the golden fixtures recorded from the reference.  The GPU
tests (test_gpu_parity.py) repeat this through the real C ABI on a B200."""
import numpy as np
import pytest

import helpers
from emu_driver import EmuEnv, SingleEnvDriver
from parity_bodies import (body_int64_out_of_range, body_large_batch, body_observe_current_state,
                           body_per_env_mazes, lay_rows)


@pytest.mark.parametrize("name", helpers.traj_names())
def test_emulated_kernels_match_reference(name):
    fx = helpers.load_traj(name)
    # 3 envs in a CTA of 2 -> exercises the tail CTA and a non-zero env slot
    env = EmuEnv(fx["layout"], 3, length=int(fx["length"]), reward_forLegalAction=bool(fx["reward_forLegalAction"]),
                 defenceReward=bool(fx["defenceReward"]), E=2, KB=2)
    helpers.replay(fx, SingleEnvDriver(env, n=2), where=name)


def test_emulated_batch_vs_oracle_with_masked_reset():
    body_large_batch(EmuEnv, "smallCapture", 96, 260, 110, state_stride=7, obs_stride=5)


def test_emulated_per_env_mazes():
    body_per_env_mazes(lambda lays, n, **kw: EmuEnv(lays, n, E=8, KB=1, **kw), 12, 110)


def test_emulated_int64_actions_and_out_of_range():
    body_int64_out_of_range(EmuEnv)


def test_emulated_observe_current_state():
    body_observe_current_state(EmuEnv)


def test_emulated_random_agents_are_legal_and_reproducible():
    rows = lay_rows("smallCapture")
    mk = lambda: EmuEnv(rows, 40, random_agent_mask=0b0101, random_seed=123)
    e1, e2 = mk(), mk()
    e1.reset(); e2.reset()
    rng = np.random.default_rng(0)
    moved = 0
    for t in range(40):
        a = rng.integers(0, 5, size=(40, 4)).astype(np.uint8)
        p0 = e1.get_state()["agents"][:, 0, :2].copy()
        o1, o2 = e1.step(a), e2.step(a)
        np.testing.assert_array_equal(o1["obs"], o2["obs"])
        assert (o1["reward"][0] >= 0.02 - 1e-12).all()
        moved += int((e1.get_state()["agents"][:, 0, :2] != p0).any(1).sum())
    assert moved > 40


def test_emulated_auto_reset_matches_oracle():
    from parity_bodies import body_auto_reset
    body_auto_reset(lambda rows, n, **kw: EmuEnv(rows, n, E=8, KB=1, **kw), N=40, T=120)
    body_auto_reset(lambda rows, n, **kw: EmuEnv(rows, n, E=8, KB=1, **kw), layname=helpers.load_traj("capsSmall")["layout"], N=24, T=100)


def test_emulated_env_ranges_and_shard_invariance():
    from parity_bodies import body_env_ranges, body_shard_invariant_random_agents
    body_env_ranges(lambda rows, n, **kw: EmuEnv(rows, n, E=32, KB=1, **kw), lambda e, b, c: e.set_range(b, c))
    body_shard_invariant_random_agents(lambda rows, n, **kw: EmuEnv(rows, n, E=16, KB=1, **kw))


def test_emulated_packed_host_record_matches_separate_outputs():
    """struct pm_host_record (one 48-byte record per env) carries exactly the separate small outputs."""
    import ctypes as C
    from pacman_marl_2025_b200 import _lib
    assert C.sizeof(_lib.PmHostRecord) == 48
    rows = lay_rows("tinyCapture")
    N = 70
    env = EmuEnv(rows, N, length=30, auto_reset=True, E=16, KB=1)
    env.reset()
    rng = np.random.default_rng(2)
    from parity_bodies import biased_actions
    for t in range(60):
        o = env.step(biased_actions(rng, N, 0.6, t))
        p = env.last_packed
        np.testing.assert_array_equal(p[:, 0:16].copy().view(np.float64), o["reward"].T)
        np.testing.assert_array_equal(p[:, 16:32].copy().view(np.float32), o["shaped"].T)
        np.testing.assert_array_equal(p[:, 32:36].copy().view(np.int32)[:, 0], o["score_change"])
        np.testing.assert_array_equal(p[:, 36:40], o["legal"])
        np.testing.assert_array_equal(p[:, 40], o["done"])
        assert not p[:, 41:].any()


@pytest.mark.parametrize("name", [n for n in helpers.traj_names() if n.startswith("bots_")])
def test_emulated_baseline_team_matches_reference(name):
    """On-device baselineTeam (Offensive + Defensive reflex agents): the kernel logic picks exactly the actions
    the reference's bots picked, and the whole trajectory follows."""
    fx = helpers.load_traj(name)
    env = EmuEnv(fx["layout"], 2, length=int(fx["length"]), E=2, KB=1, bot_kind=list(fx["bot_kind"]), bot_first_best=True)
    helpers.replay(fx, helpers.BotReplayDriver(SingleEnvDriver(env, n=1), fx), where=name)


@pytest.mark.parametrize("name", [n for n in helpers.traj_names() if n.startswith("rng_")])
def test_emulated_bots_follow_the_references_real_rng_stream(name):
    """pm_env_desc.bot_draws: fed the indices the reference's own `random.choice` returned (global MT19937,
    random.seed(1)), the in-kernel randomTeam / baselineTeam replay BASELINE config 1 bit for bit."""
    fx = helpers.load_traj(name)
    env = EmuEnv(fx["layout"], 2, length=int(fx["length"]), E=2, KB=1, bot_kind=list(fx["bot_kind"]),
                 bot_first_best=False, bot_draws=True)
    helpers.replay(fx, helpers.BotReplayDriver(SingleEnvDriver(env, n=1), fx), where=name)


def test_emulated_max_size_board():
    from parity_bodies import body_max_size_board
    body_max_size_board(lambda rows, n, **kw: EmuEnv(rows, n, E=8, KB=1, **kw), N=10, T=90)


@pytest.mark.parametrize("lay", ["tinyCapture", "capsSmall", "odd15"])
def test_emulated_learner_views(lay):
    from parity_bodies import body_learner_views
    rows = lay_rows(lay) if lay.endswith("Capture") else helpers.load_traj(lay)["layout"]
    body_learner_views(lambda r, n, **kw: EmuEnv(r, n, E=8, KB=1, **kw), rows, N=20, T=110)


def test_emulated_learner_views_with_baseline_bots():
    from parity_bodies import body_learner_views
    body_learner_views(lambda r, n, **kw: EmuEnv(r, n, E=4, KB=1, **kw), lay_rows("smallCapture"), N=12, T=80,
                       bot_kind=[2, 0, 3, 0])


def test_emulated_explicit_layout_ids():
    from parity_bodies import body_layout_ids
    body_layout_ids(lambda lays, n, **kw: EmuEnv(lays, n, E=8, KB=1, **kw), N=20, T=70)


def test_emulated_bots_and_views_on_distinct_mazes():
    from parity_bodies import body_bots_and_views_on_distinct_mazes
    body_bots_and_views_on_distinct_mazes(lambda lays, n, **kw: EmuEnv(lays, n, E=4, KB=1, **kw), N=6, T=50)


def test_emulated_step_in_several_launches():
    from parity_bodies import body_step_in_several_launches
    mk = lambda rows, n, **kw: EmuEnv(rows, n, E=8, KB=1, **kw)
    sa = lambda e, b, c: e.set_agents(b, c)
    body_step_in_several_launches(mk, sa, lay_rows("tinyCapture"), N=16, T=110)
    body_step_in_several_launches(mk, sa, helpers.load_traj("capsSmall")["layout"], N=12, T=90, splits=((0, 2), (2, 3), (3, 4)))
    body_step_in_several_launches(mk, sa, lay_rows("smallCapture"), N=10, T=60, splits=((0, 1), (1, 4)), bot_kind=[0, 0, 3, 0])


@pytest.mark.parametrize("lay", ["tinyCapture", "smallCapture", "capsSmall", "odd15"])
def test_emulated_packed_observation_bits(lay):
    from parity_bodies import body_obs_bits
    rows = lay_rows(lay) if lay.endswith("Capture") else helpers.load_traj(lay)["layout"]
    body_obs_bits(lambda r, n, **kw: EmuEnv(r, n, E=8, KB=1, **kw), rows, N=20, T=130)


def test_emulated_layout_pool_new_maze_every_episode():
    from parity_bodies import body_layout_pool
    body_layout_pool(lambda lays, n, **kw: EmuEnv(lays, n, E=8, KB=1, **kw), N=12, T=100, L=5)
    body_layout_pool(lambda lays, n, **kw: EmuEnv(lays, n, E=4, KB=1, **kw), N=6, T=70, L=3, seed=5, base=1000)


def test_emulated_static_plane_elision():
    from parity_bodies import body_elide_static_planes
    from pacman_marl_2025_b200 import mazeGenerator
    mk = lambda lays, n, **kw: EmuEnv(lays, n, E=8, KB=1, **kw)
    body_elide_static_planes(mk, lay_rows("tinyCapture"), N=20, T=100)                      # no capsules: planes 0, 2, 3 skipped
    body_elide_static_planes(mk, helpers.load_traj("capsSmall")["layout"], N=12, T=90)      # capsules: plane 0 only
    body_elide_static_planes(mk, helpers.load_traj("odd15")["layout"], N=12, T=90)          # odd width: ragged float4 edges
    mazes = [mazeGenerator.generateMaze(8000 + s).split("\n") for s in range(4)]
    body_elide_static_planes(mk, mazes, N=10, T=90, layout_pool=True, random_seed=3)         # boards change at restarts


@pytest.mark.parametrize("name", ["smallCapture", "tinyCapture"])
def test_emulated_make_observation_matches_reference(name):
    from parity_bodies import body_make_observation
    body_make_observation(lambda r, n, **kw: EmuEnv(r, n, E=2, KB=1, **kw), name)


def test_replay_file_of_the_reference_loads_and_replays_move_by_move():
    """SURVEY row f4, record / replay: a game played by the reference's tournament loop (Game.run) and pickled the way
    capture.runGames(record=True) does is read WITHOUT the reference's classes and replayed one sub-move per launch;
    the state after every move equals the one capture.replayGame walks through."""
    import os
    from pacman_marl_2025_b200 import replay as R
    from parity_bodies import check_replay_move
    rp = R.load_replay(os.path.join(helpers.GOLDEN, "replay_smallCapture.pkl"))
    z = np.load(os.path.join(helpers.GOLDEN, "replay_smallCapture_states.npz"))
    assert rp["layout_rows"] == lay_rows("smallCapture") and rp["length"] == 400 and rp["redTeamName"] == "baselineTeam"
    assert [list(m) for m in rp["actions"]] == z["moves"].tolist()
    env = EmuEnv(rp["layout_rows"], 2, length=10 ** 9, reward_forLegalAction=False, defenceReward=False, E=2, KB=1)
    env.reset()
    a = np.zeros((2, 4), np.uint8)
    for k, (idx, act) in enumerate(rp["actions"]):
        a[:, idx] = act
        env.set_agents(idx, idx + 1)
        env.step(a)
        check_replay_move(z, k, env.get_state(), n=1)
    # and our own writer round-trips through the same loader
    again = R.load_replay(R.save_replay(None, rp["layout_rows"], rp["actions"], rp["length"], "a", "b"))
    assert again["actions"] == rp["actions"] and again["layout_rows"] == rp["layout_rows"] and again["blueTeamName"] == "b"
