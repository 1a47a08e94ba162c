"""Parity tests proper: the CUDA path through the C ABI (libpacman_b200.so) vs
the golden fixtures recorded from the reference and vs the CPU oracle on the
same seeded inputs.  Bit-exact for integer state, masks, fp64 rewards, fp32 obs
and GAE; 1e-5 for advantage normalisation.  Run on the B200 box: -m gpu."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import helpers
from emu_driver import SingleEnvDriver
from oracle import oracle
from parity_bodies import (biased_actions, body_int64_out_of_range, body_large_batch, body_observe_current_state,
                           body_per_env_mazes, lay_rows, LAYDIR)

pytestmark = pytest.mark.gpu

class CudaEnv:
    """Thin numpy-returning adapter over the product class (same surface as tests/emu_driver.EmuEnv)."""

    def __init__(self, layouts, n_envs, length=299, reward_forLegalAction=True, defenceReward=True,
                 random_agent_mask=0, random_seed=0, bot_kind=None, bot_first_best=True, **kw):
        from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
        per_env = not isinstance(layouts[0], str)
        lays = layouts if per_env else [layouts]
        assert random_agent_mask in (0, 0b0101)
        team, self_play = "randomTeam", random_agent_mask == 0
        if bot_kind is not None and list(bot_kind) == [2, 0, 3, 0]:
            team, self_play = "baselineTeam", False
            kw["opponent_tiebreak"] = "first" if bot_first_best else "random"
        elif bot_kind is not None and list(bot_kind) == [1, 0, 1, 0]:
            team, self_play = "randomTeam", False
        want_draws = kw.pop("bot_draws", False)
        self.env = gymPacMan_parallel_env(length=length, reward_forLegalAction=reward_forLegalAction,
                                          defenceReward=defenceReward, self_play=self_play, enemieName=team,
                                          num_envs=n_envs, layouts=lays, opponent_seed=random_seed, **kw)
        self.N = n_envs
        if want_draws:
            self.draws = torch.zeros((n_envs, 4), dtype=torch.int32, device="cuda")
            self.env.set_bot_draws(self.draws)

    def set_draws(self, draws):
        self.draws.copy_(torch.as_tensor(np.asarray(draws, np.int64).reshape(self.N, 4).astype(np.int32)))

    def pool_seed(self):
        return int(self.env._desc.random_seed)

    def set_elide(self, on=True):
        self.env.elide_static_planes = bool(on)

    def make_observation(self, draws=None):
        d = None if draws is None else torch.as_tensor(np.asarray(draws).astype(np.int32))
        vis, noisy = self.env.make_observation_batched(d)
        return vis.cpu().numpy(), noisy.cpu().numpy()

    def set_bits(self, on=True):
        self.env.set_bits_output(on)

    def observe_bits(self):
        return self.env.observe_bits().cpu().numpy().view(np.uint32)

    def reset(self, mask=None):
        m = None if mask is None else torch.as_tensor(np.asarray(mask, np.uint8))
        obs, legal = self.env.reset_batched(m)
        return obs.cpu().numpy(), legal.cpu().numpy()

    def step(self, actions, dtype=np.uint8):
        a = torch.as_tensor(np.ascontiguousarray(np.asarray(actions).reshape(self.N, 4), dtype)).cuda()
        o = self.env.step_batched(a)
        return dict(obs=o["obs"].cpu().numpy(), reward=o["reward"].cpu().numpy(), done=o["done"].cpu().numpy(),
                    legal=o["legal"].cpu().numpy(), score_change=o["score_change"].cpu().numpy(),
                    shaping=o["shaping"].cpu().numpy(), shaped=o["shaped_reward"].cpu().numpy(),
                    actions_used=o["actions_used"].cpu().numpy(),
                    **({"obs_bits": o["obs_bits"].cpu().numpy().view(np.uint32)} if "obs_bits" in o else {}),
                    **({k: o[k].cpu().numpy() for k in ("view_actor", "view_critic")} if "view_actor" in o else {}))

    def set_views(self, red=None):
        self.env.set_view_outputs(red)

    def observe_views(self, red):
        a, c = self.env.observe_views(red)
        return a.cpu().numpy(), c.cpu().numpy()

    def set_agents(self, begin=0, end=0):
        self.env._desc.agent_begin, self.env._desc.agent_end = int(begin), int(end)

    def set_range(self, begin=0, count=0):
        self.env._desc.env_begin, self.env._desc.env_count = int(begin), int(count)

    def get_state(self):
        return self.env.export_state()

    def observe(self):
        return self.env.observe_batched().cpu().numpy()


@pytest.mark.parametrize("name", helpers.traj_names())
def test_cuda_matches_reference_fixture(name):
    fx = helpers.load_traj(name)
    env = CudaEnv(fx["layout"], 3, length=int(fx["length"]), reward_forLegalAction=bool(fx["reward_forLegalAction"]),
                  defenceReward=bool(fx["defenceReward"]))
    helpers.replay(fx, SingleEnvDriver(env, n=2), where=name)


@pytest.mark.parametrize("layname,N,T,length", [("tinyCapture", 2048, 220, 150), ("smallCapture", 4096, 310, 300),
                                                ("bloxCapture", 1024, 130, 120)])
def test_large_batch_vs_oracle(layname, N, T, length):
    body_large_batch(CudaEnv, layname, N, T, length, expect_score=layname != "bloxCapture")


def test_per_env_distinct_mazes_vs_oracle():
    body_per_env_mazes(CudaEnv, 96, 150)


def test_auto_reset_fused_into_step_vs_oracle():
    from parity_bodies import body_auto_reset
    body_auto_reset(CudaEnv, "tinyCapture", N=1500, T=200, length=40)
    body_auto_reset(CudaEnv, helpers.load_traj("capsSmall")["layout"], N=300, T=150, length=60)
    body_auto_reset(CudaEnv, "bloxCapture", N=200, T=90, length=25)


def test_env_ranges_and_shard_invariant_random_agents():
    from parity_bodies import body_env_ranges, body_shard_invariant_random_agents
    body_env_ranges(CudaEnv, lambda e, b, c: e.set_range(b, c), N=96)
    body_shard_invariant_random_agents(CudaEnv, N=256)


def test_max_size_board_and_oversize_refusal():
    from parity_bodies import big_board, body_max_size_board
    body_max_size_board(CudaEnv, N=200, T=160)
    with pytest.raises(RuntimeError, match="PM_ETOOBIG"):
        CudaEnv(big_board(33, 32), 4)


def test_explicit_layout_ids():
    from parity_bodies import body_layout_ids
    body_layout_ids(CudaEnv, N=300, T=120)


def test_bots_and_views_on_distinct_mazes():
    from parity_bodies import body_bots_and_views_on_distinct_mazes
    body_bots_and_views_on_distinct_mazes(CudaEnv, N=70, T=90)


def test_step_in_several_launches():
    from parity_bodies import body_step_in_several_launches
    sa = lambda e, b, c: e.set_agents(b, c)
    body_step_in_several_launches(CudaEnv, sa, lay_rows("tinyCapture"), N=700, T=130)
    body_step_in_several_launches(CudaEnv, sa, helpers.load_traj("capsSmall")["layout"], N=200, T=100, splits=((0, 2), (2, 3), (3, 4)))
    body_step_in_several_launches(CudaEnv, sa, lay_rows("smallCapture"), N=100, T=70, splits=((0, 1), (1, 4)), bot_kind=[2, 0, 3, 0])


def test_int64_actions_and_out_of_range():
    body_int64_out_of_range(CudaEnv)


def test_observe_current_state_and_set_state_roundtrip():
    env, rng = body_observe_current_state(CudaEnv)
    s = env.get_state()
    env2 = CudaEnv(lay_rows("smallCapture"), env.N)
    env2.env.import_state(s["agents"], s["food"], s["caps"], s["score"], s["steps"])
    s2 = env2.get_state()
    for k in s:
        np.testing.assert_array_equal(s[k], s2[k])
    a = biased_actions(rng, env.N)
    o1, o2 = env.step(a), env2.step(a)
    for k in ("obs", "reward", "done", "legal", "score_change"):
        np.testing.assert_array_equal(o1[k], o2[k])


def test_step_host_entry_matches_device_entry():
    """pm_step_host (host buffers, copies inside the call) == pm_step (device buffers)."""
    rows = lay_rows("tinyCapture")
    N = 257
    e1, e2 = CudaEnv(rows, N), CudaEnv(rows, N)
    e1.reset(); e2.reset()
    rng = np.random.default_rng(1)
    for t in range(30):
        a = biased_actions(rng, N)
        o1 = e1.step(a)
        h = e2.env.step_host(torch.as_tensor(a).pin_memory(), want_obs=True)
        for k in ("obs", "reward", "done", "legal", "score_change"):
            np.testing.assert_array_equal(h[k].numpy(), o1[k], err_msg=f"t={t} {k}")
        np.testing.assert_array_equal(h["shaped_reward"].numpy(), o1["shaped"])


def test_step_host_pipelined_matches_device_entry():
    """pm_step_host_pipelined (chunked, two streams, packed 48-byte records) == pm_step."""
    rows = lay_rows("tinyCapture")
    for N, chunks in ((1000, 4), (257, 3), (64, 1), (4096, 8), (3001, 1)):
        e1, e2 = CudaEnv(rows, N, length=40, auto_reset=True), CudaEnv(rows, N, length=40, auto_reset=True)
        e1.reset(); e2.reset()
        rng = np.random.default_rng(N)
        for t in range(60):
            a = biased_actions(rng, N, 0.6, t)
            o1 = e1.step(a)
            if chunks == 1:          # the zero-copy entry: kernel reads pinned actions / writes pinned records over PCIe
                h = e2.env.step_host_mapped(torch.as_tensor(a).pin_memory())
            else:
                h = e2.env.step_host_pipelined(torch.as_tensor(a).pin_memory(), n_chunks=chunks)
            np.testing.assert_array_equal(h["reward"].numpy().T, o1["reward"], err_msg=f"N={N} t={t}")
            np.testing.assert_array_equal(h["shaped_reward"].numpy().T, o1["shaped"])
            np.testing.assert_array_equal(h["score_change"].numpy(), o1["score_change"])
            np.testing.assert_array_equal(h["legal"].numpy(), o1["legal"])
            np.testing.assert_array_equal(h["done"].numpy(), o1["done"])
            assert torch.equal(e2.env._obs, e1.env._obs)
        s1, s2 = e1.get_state(), e2.get_state()
        for k in s1:
            np.testing.assert_array_equal(s1[k], s2[k])


def test_step_host_mapped_rejects_pageable_memory():
    from pacman_marl_2025_b200 import _lib
    env = CudaEnv(lay_rows("tinyCapture"), 64).env
    a = torch.zeros((64, 4), dtype=torch.uint8)
    with pytest.raises(ValueError):
        env.step_host_mapped(a)
    rec = torch.zeros((64, 48), dtype=torch.uint8).pin_memory()
    rc = env._L.pm_step_host_mapped(C.byref(env._desc), a.data_ptr(), C.byref(env._step_out), rec.data_ptr(), None)
    assert rc == -1                                     # PM_EINVAL: actions not page-locked
    torch.cuda.synchronize()
    env.step_host_mapped(a.pin_memory())                # and the context is still healthy


def test_properties_at_full_size():
    """BASELINE config 2 size (65,536 envs, smallCapture): size-independent
    invariants — pellet conservation, legal mask == open neighbours, obs plane
    sums, and shard invariance (two half batches == one full batch)."""
    rows = lay_rows("smallCapture")
    N = 65536
    env = CudaEnv(rows, N, length=300)
    half = [CudaEnv(rows, N // 2, length=300) for _ in range(2)]
    env.reset(); [h.reset() for h in half]
    from pacman_marl_2025_b200.layout import Layout
    lay = Layout(rows)
    walls = lay.walls.array
    total = lay.totalFood
    rng = np.random.default_rng(77)
    for t in range(48):
        a = biased_actions(rng, N, 0.6)
        A = torch.as_tensor(a).cuda()
        o = env.env.step_batched(A)
        oh = [half[i].env.step_batched(A[i * N // 2:(i + 1) * N // 2].contiguous()) for i in range(2)]
        for k in ("obs", "reward", "done", "legal", "score_change", "shaped_reward"):
            full = o[k]
            cat = torch.cat([oh[0][k], oh[1][k]], dim=1 if k in ("obs", "reward", "shaped_reward") else 0)
            assert torch.equal(full, cat), f"shard invariance broken for {k} at t={t}"
        if t % 12 == 11:
            s = env.get_state()
            on_board = s["food"].sum(1)
            carried = s["agents"][:, :, 5].sum(1)
            returned = s["agents"][:, :, 6].sum(1)
            np.testing.assert_array_equal(on_board + carried + returned, total)
            x, y = s["agents"][:, :, 0], s["agents"][:, :, 1]
            want = ((~walls[x, y + 1]).astype(np.uint8) | (~walls[x + 1, y]).astype(np.uint8) << 1
                    | (~walls[x, y - 1]).astype(np.uint8) << 2 | (~walls[x - 1, y]).astype(np.uint8) << 3 | 16)
            np.testing.assert_array_equal(o["legal"].cpu().numpy(), want)
            obs = o["obs"]
            assert torch.equal(obs[:, :, 0].sum((2, 3)), torch.full((4, N), float(walls.sum()), device="cuda"))
            assert torch.equal(obs[:, :, 4].sum((2, 3)), torch.ones((4, N), device="cuda"))
            food_planes = (obs[3, :, 6] + obs[3, :, 7]).sum((1, 2)).cpu().numpy()
            np.testing.assert_array_equal(food_planes, on_board)


def test_random_team_opponent_in_kernel():
    """self_play=False: agents 0 and 2 draw uniform legal actions in-kernel (randomTeam).
    Legal by construction -> the red team always earns the legal-action bonus twice."""
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    N = 512
    mk = lambda: gymPacMan_parallel_env(layout_file=os.path.join(LAYDIR, "smallCapture.lay"), self_play=False,
                                        num_envs=N, opponent_seed=123)
    e1, e2 = mk(), mk()
    g = torch.Generator().manual_seed(0)
    moved = 0
    for t in range(50):
        a = torch.randint(0, 5, (N, 4), generator=g).cuda()
        p0 = torch.as_tensor(e1.export_state()["agents"][:, 0, :2])
        o1, o2 = e1.step_batched(a), e2.step_batched(a.clone())
        assert torch.equal(o1["obs"], o2["obs"])            # counter-based RNG: reproducible
        assert (o1["reward"][0].cpu().numpy() >= 0.02 - 1e-12).all()     # two legal red moves per step
        moved += int((torch.as_tensor(e1.export_state()["agents"][:, 0, :2]) != p0).any(1).sum())
    assert moved > N    # bots do move


@pytest.mark.parametrize("name", helpers.dist_names())
def test_bfs_distances_match_reference(name):
    from pacman_marl_2025_b200.distanceCalculator import compute_distances, Distancer
    from pacman_marl_2025_b200.layout import Layout
    rows, nodes, dist = helpers.load_dist(name)
    lay = Layout(rows)
    t = compute_distances([lay])
    np.testing.assert_array_equal(t.matrix(0), dist)
    assert int(t.n_open[0]) == len(nodes)
    node_of = t.node_of_cell[0].cpu().numpy()
    for i, (x, y) in enumerate(nodes):
        assert node_of[y * lay.width + x] == i
    d = Distancer(lay)
    d.getMazeDistances()
    a, b = tuple(int(v) for v in nodes[0]), tuple(int(v) for v in nodes[-1])
    assert d.getDistance(a, b) == dist[0, -1]


def test_bfs_batch_of_random_mazes_vs_oracle():
    from oracle import oracle
    from pacman_marl_2025_b200 import mazeGenerator
    from pacman_marl_2025_b200.distanceCalculator import compute_distances
    from pacman_marl_2025_b200.layout import Layout
    texts = [mazeGenerator.generateMaze(s).split("\n") for s in range(300, 364)]
    t = compute_distances([Layout(x) for x in texts])
    for l in range(0, 64, 7):
        _, od = oracle.distances(texts[l])
        np.testing.assert_array_equal(t.matrix(l), od)
    # symmetry + zero diagonal for every table
    n = t.n_open.cpu().numpy()
    D = (t.dist.to(torch.int32) & 0xFFFF)
    assert torch.equal(D, D.transpose(1, 2))
    for l in range(64):
        assert int(torch.diagonal(D[l])[: n[l]].sum()) == 0


def test_gae_bit_exact_vs_reference_and_oracle():
    from oracle import oracle
    from pacman_marl_2025_b200.mappo_ops import compute_gae
    z = np.load(helpers.GOLDEN + "/gae.npz")
    for ci in range(int(z["n_cases"])):
        adv, ret = compute_gae(torch.as_tensor(z[f"c{ci}_rew"]).cuda(), torch.as_tensor(z[f"c{ci}_val"]).cuda(),
                               torch.as_tensor(z[f"c{ci}_done"]).cuda(), torch.as_tensor(z[f"c{ci}_last"]).cuda(),
                               float(z["gamma"]), float(z["lam"]))
        np.testing.assert_array_equal(adv.cpu().numpy().view(np.uint32), z[f"c{ci}_adv"].view(np.uint32))
        np.testing.assert_array_equal(ret.cpu().numpy().view(np.uint32), z[f"c{ci}_ret"].view(np.uint32))
    adv, ret = compute_gae(torch.as_tensor(z["kat_rew"]).cuda(), torch.as_tensor(z["kat_val"]).cuda(),
                           torch.as_tensor(z["kat_done"]).cuda(), 0.37, 0.99, 0.95)
    np.testing.assert_array_equal(adv.cpu().numpy().view(np.uint32), z["kat_adv"].view(np.uint32))
    # rollout-sized case vs the oracle: T = 2048, 4096 columns
    rng = np.random.default_rng(0)
    T, M = 2048, 4096
    r = rng.standard_normal((T, M)).astype(np.float32)
    v = rng.standard_normal((T, M)).astype(np.float32)
    d = (rng.random((T, M)) < 0.01).astype(np.float32)
    lv = rng.standard_normal(M).astype(np.float32)
    oa, orr = oracle.gae(r, v, d, lv)
    adv, ret = compute_gae(torch.as_tensor(r).cuda(), torch.as_tensor(v).cuda(), torch.as_tensor(d).cuda(),
                           torch.as_tensor(lv).cuda())
    np.testing.assert_array_equal(adv.cpu().numpy().view(np.uint32), oa.view(np.uint32))
    np.testing.assert_array_equal(ret.cpu().numpy().view(np.uint32), orr.view(np.uint32))
    # wide buffers take four columns per thread (128-bit accesses); a width that is not a multiple of four, or an
    # unaligned view, falls back to one: the same bits either way
    for T, M, off in ((48, 65536, 0), (33, 65538, 0), (40, 65536, 1)):
        r = rng.standard_normal((T, M + off)).astype(np.float32)
        v = rng.standard_normal((T, M + off)).astype(np.float32)
        d = (rng.random((T, M + off)) < 0.05).astype(np.float32)
        lv = rng.standard_normal(M + off).astype(np.float32)
        rt, vt, dt, lt = (torch.as_tensor(a).cuda() for a in (r, v, d, lv))
        if off:
            rt, vt, dt, lt = rt[:, off:], vt[:, off:], dt[:, off:], lt[off:]
            r, v, d, lv = r[:, off:], v[:, off:], d[:, off:], lv[off:]
        oa, orr = oracle.gae(np.ascontiguousarray(r), np.ascontiguousarray(v), np.ascontiguousarray(d), np.ascontiguousarray(lv))
        adv, ret = compute_gae(rt, vt, dt, lt)
        np.testing.assert_array_equal(adv.cpu().numpy().view(np.uint32), oa.view(np.uint32), err_msg=f"{T}x{M}+{off}")
        np.testing.assert_array_equal(ret.cpu().numpy().view(np.uint32), orr.view(np.uint32))


def test_adv_norm_vs_reference():
    from pacman_marl_2025_b200.mappo_ops import normalize_advantages
    z = np.load(helpers.GOLDEN + "/gae.npz")
    out = normalize_advantages(torch.as_tensor(z["norm_adv"]).cuda(), torch.as_tensor(z["norm_inds"]).cuda(),
                               int(z["norm_mb"]))
    np.testing.assert_allclose(out.cpu().numpy(), z["norm_out"], rtol=1e-5, atol=1e-6)   # north_star tolerance
    # larger shuffles against the reference's expression (`:577`), minibatches that fit the kernel's shared-memory cache
    # (8 192, 512) and one that does not (50 000), ragged tails included
    g = torch.Generator(device="cuda").manual_seed(4)
    adv = torch.randn(300_001, generator=g, device="cuda") * 3 + 0.7
    idx = torch.randperm(adv.numel(), generator=g, device="cuda")
    for mb in (8192, 512, 50_000):
        got = normalize_advantages(adv, idx, mb)
        for s0 in range(0, adv.numel(), mb):
            a = adv[idx[s0:s0 + mb]].double()
            if a.numel() > 1:
                want = (a - a.mean()) / (a.std() + 1e-8)
                assert float((got[s0:s0 + mb].double() - want).abs().max()) <= 2e-5, (mb, s0)


def test_reference_shaped_single_env_api():
    """num_envs=1 returns the reference's shapes: dict of [8,H,W] tensors, float rewards, bool dones, int lists."""
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    env = gymPacMan_parallel_env(layout_file=os.path.join(LAYDIR, "tinyCapture.lay"), self_play=True, length=300)
    obs, info = env.reset()
    assert info["legal_actions"] == {0: [1, 3, 4], 1: [1, 4], 2: [3, 4], 3: [1, 3, 4]}
    assert obs[0].shape == (8, 7, 20) and obs[0].dtype == torch.float32
    obs, rew, term, info = env.step({0: 1, 1: "East", 2: 1, 3: 1})
    assert isinstance(rew[0], float) and isinstance(term[0], bool) and isinstance(info["score_change"], int)
    # obs[0] is a mid-step snapshot (enemies not yet moved)
    assert obs[0][5].nonzero().tolist() == [[1, 12], [1, 13]]
    assert env.game.state.getAgentPosition(1) == (13, 1)
    assert env.game.state.data.score == 0 and env.max_score == 11
    assert env.get_Observation(0).shape == (8, 7, 20)
    # pacman.py:352-356 passes the actions as a LIST indexed by agent and reads info["legal_actions"][agent]
    obs2, rew2, term2, info2 = env.step([4, 4, 4, 4])
    assert all(isinstance(info2["legal_actions"][i], list) for i in range(4)) and rew2[0] == 0.02 and rew2[1] == 0.02
    assert torch.equal(env.get_Observation(3), obs2[3])              # last mover's snapshot == current state


def test_mappo_trainer_rollout_gae_and_update():
    """BASELINE config 5 in miniature: a batched rollout on smallCapture, rewards / dones in the buffers
    equal what the reference loop would store, pm_gae over the buffers equals the oracle's compute_gae,
    and PPO updates move the weights with finite losses."""
    from oracle import oracle
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    cfg = M.TrainConfig(num_envs=96, num_steps=24, num_minibatches=4, update_epochs=2, total_updates=6, length=12,
                        seed=3, easy_until=0, selfplay_after=2, checkpoint_every=0)
    tr = M.MAPPOTrainer(cfg, device="cuda:0")
    before = torch.cat([p.detach().reshape(-1).clone() for p in tr.agent.parameters()])
    modes = set()
    for upd in range(6):
        rec = tr.train_one(upd)
        modes.add(rec["mode"])
        assert all(np.isfinite(rec[k]) for k in ("pg_loss", "v_loss", "entropy", "clip_frac", "sps", "reward"))
        assert rec["episodes"] >= 96                        # length 12 -> every env finishes at least once in 24 steps
        ro = tr.rollout
        # dones: identical for both learners; obs buffers hold exactly one own cell per learner
        assert torch.equal(ro.done[:, 0], ro.done[:, 1])
        assert bool(((ro.obs[:, :, :, 1] > 0).sum((-1, -2)) == 1).all())
        assert bool((ro.merged[:, :, 4] == 0).all()) and bool(((ro.merged[:, :, 1] > 0).sum((-1, -2)) <= 2).all())
        # GAE over the stored buffers == oracle restatement of compute_gae, per column, bit for bit
        T, N = cfg.num_steps, cfg.num_envs
        with torch.no_grad():
                lv = tr.agent.get_value(ro.merged[-1].float()) * 0 + 0.25          # any [N] vector
        adv, ret = M.compute_gae(ro.reward.view(T, 2 * N), ro.value.view(T, 2 * N), ro.done.view(T, 2 * N), lv.repeat(2))
        oadv, oret = oracle.gae(ro.reward.view(T, 2 * N).cpu().numpy(), ro.value.view(T, 2 * N).cpu().numpy(),
                                ro.done.view(T, 2 * N).cpu().numpy(), np.full(2 * N, 0.25, np.float32))
        np.testing.assert_array_equal(adv.cpu().numpy().view(np.uint32), oadv.view(np.uint32))
        np.testing.assert_array_equal(ret.cpu().numpy().view(np.uint32), oret.view(np.uint32))
    after = torch.cat([p.detach().reshape(-1) for p in tr.agent.parameters()])
    assert not torch.equal(before, after) and bool(torch.isfinite(after).all())
    assert "randomTeam" in modes and "baselineTeam" in modes, modes       # updates 0 / 1-2 of the curriculum
    for team in ("randomTeam", "baselineTeam"):
        mean_ret, std_ret, win = tr.evaluate_vs_bots(num_envs=64, team=team)
        assert np.isfinite(mean_ret) and 0.0 <= win <= 1.0


def test_mappo_rollout_rewards_match_reference_formula():
    """reward_buf[t, i] = team_reward + SHAPING_SCALE * shaping_i with team_reward summed over both learners
    (pacman_mappo_resnet.py:513-522), checked against the oracle driven by the actions the trainer chose."""
    from oracle import oracle
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    cfg = M.TrainConfig(num_envs=64, num_steps=40, num_minibatches=2, update_epochs=1, total_updates=1, length=15,
                        seed=5, checkpoint_every=0)
    tr = M.MAPPOTrainer(cfg, device="cuda:0")
    env = tr.env_selfplay
    rows = [ln.strip() for ln in open(os.path.join(LAYDIR, "smallCapture.lay"))]
    ref = oracle.OracleBatch(rows, cfg.num_envs, length=cfg.length)
    ref.reset()
    # drive the self-play env (learners blue, opponent = current policy) and mirror every step into the oracle
    recorded = []
    orig = env.step_batched

    def spy(actions):
        recorded.append(actions.clone())
        return orig(actions)
    env.step_batched = spy
    tr.collect("Self(Curr)", play_as_red=False)
    env.step_batched = orig
    ro = tr.rollout
    for t, a in enumerate(recorded):
        r = ref.step(a.cpu().numpy(), with_obs=False, auto_reset=True, nthreads=0)
        want = (2.0 * r["reward"][:, 1][:, None] + 0.1 * r["shaping"][:, [1, 3]]).astype(np.float32).T.copy()
        np.testing.assert_array_equal(ro.reward[t].cpu().numpy().view(np.uint32), want.view(np.uint32), err_msg=f"t={t}")
        np.testing.assert_array_equal(ro.done[t, 0].cpu().numpy(), r["done"].astype(np.float32))


# ---------------------------------------------------------------------------------------------------------
# on-device baselineTeam
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in helpers.traj_names() if n.startswith("bots_")])
def test_cuda_baseline_team_matches_reference_fixture(name):
    """Red side = baselineTeam inside the step kernel (distance tables from the BFS kernel): it plays exactly
    what the reference's Offensive/Defensive reflex agents played, and every output of every step follows."""
    fx = helpers.load_traj(name)
    env = CudaEnv(fx["layout"], 3, length=int(fx["length"]), bot_kind=list(fx["bot_kind"]), bot_first_best=True)
    helpers.replay(fx, helpers.BotReplayDriver(SingleEnvDriver(env, n=1), fx), where=name)


@pytest.mark.parametrize("layname,N,T", [("tinyCapture", 1024, 120), ("bloxCapture", 256, 80)])
def test_cuda_baseline_team_large_batch_vs_oracle(layname, N, T):
    rows = lay_rows(layname)
    env = CudaEnv(rows, N, length=60, auto_reset=True, bot_kind=[2, 0, 3, 0], bot_first_best=True)
    ref = oracle.OracleBatch(rows, N, length=60)
    ref.set_bots([2, 0, 3, 0])
    env.reset(); ref.reset()
    rng = np.random.default_rng(17)
    for t in range(T):
        a = biased_actions(rng, N, 0.5, t)
        o = env.step(a)
        r = ref.step_bots(a, with_obs=(t % 10 == 0), auto_reset=True, nthreads=0)
        np.testing.assert_array_equal(o["actions_used"][:, [0, 2]], r["actions_used"][:, [0, 2]], err_msg=f"t={t} bot actions")
        np.testing.assert_array_equal(o["reward"].view(np.uint64), r["reward"].T.copy().view(np.uint64), err_msg=f"t={t}")
        np.testing.assert_array_equal(o["done"], r["done"])
        np.testing.assert_array_equal(o["score_change"], r["score_change"])
        if t % 10 == 0:
            keep = ~r["done"].astype(bool)               # finished envs already show the next episode on the GPU side
            np.testing.assert_array_equal(o["obs"].transpose(1, 0, 2, 3, 4)[keep], r["obs"][keep], err_msg=f"t={t} obs")


def test_cuda_baseline_team_random_tiebreak_stays_in_best_set():
    """opponent_tiebreak='random' (random.choice over the best actions, baselineTeam.py:92): agent 0's choice
    is always inside the oracle's best set for the same pre-step state, and ties really get broken both ways."""
    rows = lay_rows("smallCapture")
    N = 512
    env = CudaEnv(rows, N, length=80, auto_reset=True, bot_kind=[2, 0, 3, 0], bot_first_best=False, random_seed=5)
    mirror = oracle.OracleBatch(rows, N, length=80)      # follows the GPU trajectory through explicit actions
    probe = oracle.OracleBatch(rows, N, length=80)       # asked for agent 0's best set on a copy of that state
    probe.set_bots([2, 0, 0, 0])
    env.reset(); mirror.reset(); probe.reset()
    rng = np.random.default_rng(23)
    seen_multi, not_first = 0, 0
    for t in range(60):
        a = biased_actions(rng, N, 0.5, t)
        probe.buf[:] = mirror.buf
        for n in range(N):
            oracle.lib().pmo_set_distances(probe._env(n), oracle._p(next(iter(probe._tables.values()))[0]))
        best = probe.step_bots(a, with_obs=False)["best_masks"][:, 0]
        o = env.step(a)
        used = o["actions_used"]
        chosen = used[:, 0].astype(int)
        assert (((best >> chosen) & 1) == 1).all(), f"t={t}: a chosen action is outside the best set"
        multi = np.array([bin(int(b)).count("1") > 1 for b in best])
        seen_multi += int(multi.sum())
        order = [0, 2, 1, 3, 4]
        first_best = np.array([next(u for u in order if (int(b) >> u) & 1) for b in best])
        not_first += int((multi & (chosen != first_best)).sum())
        mirror.step(used.astype(np.int32), with_obs=False, auto_reset=True, nthreads=0)
    assert seen_multi > 50 and not_first > 10, (seen_multi, not_first)


# ---------------------------------------------------------------------------------------------------------
# learner views fused into the encoder
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lay", ["tinyCapture", "smallCapture", "bloxCapture", "capsSmall", "odd15"])
def test_cuda_learner_views_match_canonicalize_and_merge(lay):
    from parity_bodies import body_learner_views
    rows = lay_rows(lay) if lay.endswith("Capture") else helpers.load_traj(lay)["layout"]
    body_learner_views(CudaEnv, rows, N=300, T=140)


def test_cuda_learner_views_with_bots_and_without_raw_obs():
    """Views with the baselineTeam kernel variant; and with raw_obs=False the views are unchanged while the fp32
    observation buffer is left untouched."""
    from parity_bodies import body_learner_views, np_views
    body_learner_views(CudaEnv, lay_rows("smallCapture"), N=200, T=100, bot_kind=[2, 0, 3, 0])
    rows = lay_rows("tinyCapture")
    N = 130
    a_env, b_env = CudaEnv(rows, N, length=50, auto_reset=True), CudaEnv(rows, N, length=50, auto_reset=True)
    a_env.reset(); b_env.reset()
    a_env.env.set_view_outputs(True)
    b_env.env.set_view_outputs(True, raw_obs=False)
    b_env.env._obs.fill_(-7.0)
    rng = np.random.default_rng(2)
    for t in range(80):
        act = biased_actions(rng, N, 0.6, t)
        oa, ob = a_env.step(act), b_env.step(act)
        np.testing.assert_array_equal(oa["view_actor"], ob["view_actor"])
        np.testing.assert_array_equal(oa["view_critic"], ob["view_critic"])
        np.testing.assert_array_equal(oa["reward"], ob["reward"])
        assert (ob["obs"] == -7.0).all()
    # torch-side functions the trainer would otherwise run give the same tensors
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    o = a_env.env.step_batched(torch.as_tensor(biased_actions(rng, N)).cuda())
    ca, cb = M.canonicalize_obs(o["obs"][0], True), M.canonicalize_obs(o["obs"][2], True)
    assert torch.equal(o["view_actor"][0].float(), ca) and torch.equal(o["view_actor"][1].float(), cb)
    assert torch.equal(o["view_critic"].float(), M.merge_obs_for_critic(ca, cb))


def test_mappo_rollout_identical_with_kernel_views_and_torch_views():
    """use_views=True (network inputs written by the step kernel into the rollout buffer) and use_views=False
    (torch flip / plane swap / merge on the fp32 observations, as the reference does) give the same rollout,
    for blue and red learners, against bots and in self-play."""
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    for mode, red in (("randomTeam", False), ("baselineTeam", False), ("Self(Curr)", True), ("Self(Curr)", False)):
        ros = []
        for use_views in (True, False):
            cfg = M.TrainConfig(num_envs=64, num_steps=30, num_minibatches=2, update_epochs=1, total_updates=1,
                                length=20, seed=11, checkpoint_every=0, use_views=use_views)
            tr = M.MAPPOTrainer(cfg, device="cuda:0")
            torch.manual_seed(1234)
            lv, stats = tr.collect(mode, play_as_red=red)
            ros.append((tr.rollout, lv, stats))
        (a, lva, sa), (b, lvb, sb) = ros
        for name in ("obs", "merged", "action", "logp", "reward", "done", "value"):
            assert torch.equal(getattr(a, name), getattr(b, name)), f"{mode} red={red}: {name} differs"
        assert torch.equal(lva, lvb) and torch.equal(sa, sb)


def test_mappo_rollout_from_cuda_graphs_equals_eager_launches():
    """graph_rollout: from its second use on, a rollout step (both network forwards, sampling, the env launch, the
    bookkeeping) is replayed from a CUDA graph per t.  Same seeds -> the same rollouts as with eager launches, collect after
    collect (the later ones start from masked resets: new episode numbers, new opponent draws), and the weights after an
    update in between agree too."""
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    flags = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    # plain fp32 and no autotuning: both trainers then run the very same network kernels, captured or not
    torch.backends.cudnn.benchmark = torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        for mode, red in (("randomTeam", False), ("Self(Curr)", True)):
            trs = []
            for graph in (True, False):
                cfg = M.TrainConfig(num_envs=64, num_steps=12, num_minibatches=2, update_epochs=1, total_updates=4,
                                    length=20, seed=3, checkpoint_every=0, graph_rollout=graph, fast_math=False)
                trs.append(M.MAPPOTrainer(cfg, device="cuda:0"))
            prev_actions = None
            for it in range(4):
                outs = []
                for tr in trs:
                    torch.manual_seed(100 + it)
                    lv, stats = tr.collect(mode, play_as_red=red)
                    outs.append((lv.clone(), stats.clone()))
                a, b = trs[0].rollout, trs[1].rollout
                for name in ("obs", "merged", "action", "reward", "done"):
                    assert torch.equal(getattr(a, name), getattr(b, name)), f"{mode} collect {it}: {name} differs"
                for name in ("logp", "value"):
                    torch.testing.assert_close(getattr(a, name), getattr(b, name), rtol=1e-4, atol=1e-5)
                assert torch.equal(outs[0][1], outs[1][1])
                if prev_actions is not None and mode == "randomTeam":
                    assert not torch.equal(prev_actions, a.action)      # a new rollout is not a replay of the previous one
                prev_actions = a.action.clone()
            assert len(trs[0]._graphs) == 1 and len(next(iter(trs[0]._graphs.values()))) == 12 and not trs[1]._graphs
            for tr in trs:                                              # and an update on top of a replayed rollout
                torch.manual_seed(7)                                    # (the minibatch shuffle draws from torch's generator)
                tr.update(outs[0][0], tr.schedule(1))
            for pa, pb in zip(trs[0].agent.parameters(), trs[1].agent.parameters()):
                torch.testing.assert_close(pa, pb, rtol=1e-3, atol=2e-5)
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = flags


# ---------------------------------------------------------------------------------------------------------
# the reference-shaped API itself (1 env, dicts, Python scalars, lists of legal actions)
# ---------------------------------------------------------------------------------------------------------
class RefApiDriver:
    """Drives gymPacMan_parallel_env exactly like the reference's callers do (pacman_mappo_resnet.py:457-538,
    pacman.py:342-358): reset() / step({agent: action}) on ONE env, reading dicts of tensors, floats, bools and
    legal-action lists, and the GameState accessors (env.game.state...)."""

    def __init__(self, fx, bots=False):
        import tempfile
        from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
        with tempfile.NamedTemporaryFile("w", suffix=".lay", delete=False) as f:
            f.write("\n".join(fx["layout"]))
        kw = dict(enemieName="baselineTeam", self_play=False, opponent_tiebreak="first") if bots else dict(self_play=True)
        self.env = gymPacMan_parallel_env(layout_file=f.name, length=int(fx["length"]),
                                          reward_forLegalAction=bool(fx["reward_forLegalAction"]),
                                          defenceReward=bool(fx["defenceReward"]), **kw)
        os.unlink(f.name)
        self.bots, self.prev = bots, None

    @staticmethod
    def _mask(lst):
        m = 0
        for a in lst:
            m |= 1 << a
        return m

    def reset(self):
        obs, info = self.env.reset()
        self.prev = obs
        ag = self.env.agents
        assert all(isinstance(a, int) for a in info["legal_actions"][ag[0]])
        return np.stack([obs[ag[i]].cpu().numpy() for i in range(4)]), [self._mask(info["legal_actions"][ag[i]]) for i in range(4)]

    def step(self, a):
        ag = self.env.agents
        acts = {ag[i]: int(a[i]) for i in range(4) if not (self.bots and i in (0, 2))}
        if not self.bots:
            acts[ag[1]] = self.env.action_mapping[int(a[1])]          # direction names are accepted too (gymPacMan.py:66-77)
        obs, rew, term, info = self.env.step(acts)
        assert isinstance(rew[ag[0]], float) and isinstance(term[ag[0]], bool) and isinstance(info["score_change"], int)
        assert rew[ag[0]] == rew[ag[2]] and rew[ag[1]] == rew[ag[3]] and len(set(term.values())) == 1
        from pacman_marl_2025_b200 import pacman_mappo_resnet as M
        shaping = [0.0] * 4
        for i in range(4):                                            # the trainer's own shaping on the returned tensors
            red = i in (0, 2)
            pa = M.canonicalize_obs(self.prev[ag[i]], red)[1].nonzero()[0]
            pb_t = M.canonicalize_obs(obs[ag[i]], red)[1]
            pb = pb_t.nonzero()[0]
            carry = float(M.canonicalize_obs(self.prev[ag[i]], red)[1][pa[0], pa[1]]) - 1
            d = abs(int(pa[0]) - int(pb[0])) + abs(int(pa[1]) - int(pb[1]))
            shaping[i] = (-0.15 - 0.4 - 0.15 * carry) if d > 1.5 else ((-0.15 - 0.05) if d < 0.1 else -0.15)
        self.prev = obs
        used = self.env._actions_used[0].cpu().numpy()
        return dict(obs=np.stack([obs[ag[i]].cpu().numpy() for i in range(4)]), reward=[rew[ag[0]], rew[ag[1]]],
                    done=term[ag[0]], legal=[self._mask(info["legal_actions"][ag[i]]) for i in range(4)],
                    score_change=info["score_change"], shaping=shaping, actions_used=used)

    def state(self):
        st = self.env.game.state
        ags = np.zeros((4, 7), np.int64)
        names = {"North": 0, "East": 1, "South": 2, "West": 3, "Stop": 4}
        for i in range(4):
            s = st.getAgentState(i)
            x, y = s.getPosition()
            ags[i] = (x, y, names[s.configuration.direction], int(s.isPacman), s.scaredTimer, s.numCarrying, s.numReturned)
        lay = self.env.layout
        food = np.zeros(lay.width * lay.height, np.uint8)
        f = st.data.food
        for x in range(lay.width):
            for y in range(lay.height):
                food[y * lay.width + x] = bool(f[x][y])
        caps = np.zeros(lay.width * lay.height, np.uint8)
        for x, y in st.getCapsules():
            caps[y * lay.width + x] = 1
        return dict(agents=ags, food=food, caps=caps, score=st.data.score)


@pytest.mark.parametrize("name", ["tinyCapture_smart", "capsSmall", "rebind", "odd15", "maze7", "small_noflags"])
def test_reference_api_single_env_replays_reference_fixture(name):
    fx = helpers.load_traj(name)
    helpers.replay(fx, RefApiDriver(fx), where=name)


@pytest.mark.parametrize("name", ["bots_tinyCapture", "bots_capsSmall"])
def test_reference_api_single_env_with_baseline_team(name):
    """The reference's config-1 shape: one env, red = baselineTeam bots, blue driven through the dict API."""
    fx = helpers.load_traj(name)
    helpers.replay(fx, helpers.BotReplayDriver(RefApiDriver(fx, bots=True), fx), where=name)


def test_envs_per_cta_choice_does_not_change_results():
    """Launch geometry only: small batches run 8 envs per CTA, large ones 32 (pm_env.cu); forcing either through
    pm_env_desc.envs_per_cta must give identical outputs, states and learner views."""
    rows = lay_rows("smallCapture")
    N = 1000
    res = {}
    with pytest.raises(RuntimeError, match="PM_EINVAL"):
        CudaEnv(rows, 8, envs_per_cta=16)
    for epc in ("8", "32"):
        env = CudaEnv(rows, N, length=35, auto_reset=True, bot_kind=[2, 0, 3, 0], envs_per_cta=int(epc))
        env.reset()
        env.set_views(True)
        rng = np.random.default_rng(8)
        outs = []
        for t in range(45):
            o = env.step(biased_actions(rng, N, 0.6, t))
            outs.append({k: o[k].copy() for k in ("obs", "reward", "done", "legal", "score_change", "shaped", "actions_used",
                                                  "view_actor", "view_critic")})
        res[epc] = (outs, env.get_state())
    for a, b in zip(res["8"][0], res["32"][0]):
        for k in a:
            np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    for k in res["8"][1]:
        np.testing.assert_array_equal(res["8"][1][k], res["32"][1][k])


# ---------------------------------------------------------------------------------------------------------
# host-side opponents: the step in two launches, each bot asked at its own sub-move (rest of row f1)
# ---------------------------------------------------------------------------------------------------------
class _GreedyEastBot:
    """A stand-in for an agents/*.py bot that only uses the GameState accessors (no reference checkout on the GPU box)."""

    def __init__(self, index):
        self.index, self.registered, self.seen_positions = index, 0, []

    def registerInitialState(self, state):
        self.registered += 1

    def getAction(self, state):
        self.seen_positions.append(tuple(state.getAgentPosition(i) for i in range(4)))
        legal = state.getLegalActions(self.index)
        x, y = state.getAgentPosition(self.index)
        if state.getAgentState(self.index).numCarrying >= 2 and "West" in legal:
            return "West"
        for a in ("East", "North", "South", "West"):
            if a in legal:
                return a
        return "Stop"


def test_host_side_bots_play_at_their_own_sub_move():
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    rows = lay_rows("tinyCapture")
    N = 3
    bots = [_GreedyEastBot(0), _GreedyEastBot(2)]
    env = gymPacMan_parallel_env(layout_file=os.path.join(LAYDIR, "tinyCapture.lay"), self_play=False, length=40,
                                 num_envs=N, bot_team=bots, enemieName="someHostTeam")
    ref = oracle.OracleBatch(rows, N, length=40)
    obs, info = env.reset()
    ref.reset()
    assert bots[0].registered >= 1 and bots[1].registered >= 1
    rng = np.random.default_rng(4)
    for t in range(41):
        blue = rng.integers(0, 5, size=(N, 2))
        before = ref.state(0)["agents"][:, :2].copy()
        obs, rew, term, info = env.step({1: torch.as_tensor(blue[:, 0]), 3: torch.as_tensor(blue[:, 1])})
        used = env._actions_used.cpu().numpy().astype(np.int32)
        # what the bots saw: agent 0's bot the state before the step, agent 2's bot the state after agents 0 and 1 moved
        assert bots[0].seen_positions[-N] == tuple(map(tuple, before))
        r = ref.step(used, with_obs=True, nthreads=0)
        # 'used' holds the EFFECTIVE actions (illegal blue requests became Stop), so drop the legal-action bonus
        # difference by comparing everything except the 0.01 bonuses of illegal requests: compare obs / done / score
        np.testing.assert_array_equal(torch.stack([obs[i] for i in range(4)]).cpu().numpy().transpose(1, 0, 2, 3, 4), r["obs"])
        np.testing.assert_array_equal(term[1].cpu().numpy().astype(np.uint8), r["done"])
        np.testing.assert_array_equal(info["score_change"].cpu().numpy(), r["score_change"])
        req_illegal = np.stack([(used[:, 1] != blue[:, 0]), (used[:, 3] != blue[:, 1])], 1).sum(1)
        np.testing.assert_allclose(rew[1].cpu().numpy(), r["reward"][:, 1] - 0.01 * req_illegal, rtol=0, atol=1e-12)
        np.testing.assert_array_equal(rew[0].cpu().numpy().view(np.uint64), r["reward"][:, 0].copy().view(np.uint64))
        # agent 2's bot saw env 0 AFTER agents 0 and 1 had moved: their cells are already the ones of the final state
        mid, after = bots[1].seen_positions[-N], ref.state(0)["agents"][:, :2]
        if not r["done"][0] and tuple(after[0]) != tuple(ref.state(0)["agents"][0, :2] * 0 - 1):
            assert mid[0] == tuple(after[0]) or tuple(after[0]) == (env.layout.starts[0][0], env.layout.starts[0][1])
        if bool(term[1].any()):
            break
    assert env.steps == t + 1 and env.num_moves == 4 * (t + 1)


# ---------------------------------------------------------------------------------------------------------
# ABI 7: the reference's real RNG stream, packed observations, a new maze every episode, layout changes
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", [n for n in helpers.traj_names() if n.startswith("rng_")])
def test_cuda_bots_follow_the_references_real_rng_stream(name):
    """BASELINE config 1 exactly as the reference plays it: tinyCapture, red = baselineTeam (or randomTeam) bots
    whose `random.choice` draws from the global MT19937 under `random.seed(1)`, blue = uniform over the legal lists
    from the same stream.  Fed the recorded choice indices (`pm_env_desc.bot_draws`), the in-kernel bots reproduce
    the reference's trajectory bit for bit - every state field, reward, mask and observation of every step."""
    fx = helpers.load_traj(name)
    env = CudaEnv(fx["layout"], 3, length=int(fx["length"]), bot_kind=list(fx["bot_kind"]), bot_first_best=False,
                  bot_draws=True)
    helpers.replay(fx, helpers.BotReplayDriver(SingleEnvDriver(env, n=1), fx), where=name)


@pytest.mark.parametrize("lay", ["tinyCapture", "smallCapture", "bloxCapture", "capsSmall", "odd15"])
def test_cuda_packed_observation_bits(lay):
    from parity_bodies import body_obs_bits
    rows = lay_rows(lay) if lay.endswith("Capture") else helpers.load_traj(lay)["layout"]
    body_obs_bits(CudaEnv, rows, N=300, T=130)


def test_cuda_packed_bits_reach_the_host_three_ways():
    """The packed observations and the uint8 learner views through the host-buffer entries: pm_step_host (device
    buffer + copies inside the call) and pm_step_host_mapped with the bits buffer itself in pinned host memory
    (the kernel writes it across PCIe) - all equal to the device-side outputs of an identical env."""
    from pacman_marl_2025_b200.gymPacMan import decode_obs_bits
    rows = lay_rows("tinyCapture")
    N = 513
    mk = lambda: CudaEnv(rows, N, length=45, auto_reset=True).env
    dev, cpy, mapped = mk(), mk(), mk()
    for e in (dev, cpy, mapped):
        e.reset_batched()
    dev.set_bits_output(True); dev.set_view_outputs(False)
    cpy.set_bits_output(True); cpy.set_view_outputs(False)
    hbits = mapped.set_bits_output(True, host_mapped=True)
    assert hbits.is_pinned()
    rng = np.random.default_rng(6)
    for t in range(70):
        a = torch.as_tensor(biased_actions(rng, N, 0.6, t))
        o = dev.step_batched(a.cuda())
        h = cpy.step_host(a.pin_memory(), want_bits=True, want_views=True)
        mapped.step_host_mapped(a.pin_memory())
        want = o["obs_bits"].cpu()
        assert torch.equal(h["obs_bits"], want), f"t={t} pm_step_host bits"
        assert torch.equal(hbits, want), f"t={t} zero-copy bits"
        assert torch.equal(h["view_actor"], o["view_actor"].cpu()) and torch.equal(h["view_critic"], o["view_critic"].cpu())
        if t % 10 == 0:
            np.testing.assert_array_equal(decode_obs_bits(hbits.numpy(), 7, 20), o["obs"].cpu().numpy())
            assert torch.equal(decode_obs_bits(o["obs_bits"], 7, 20), o["obs"])      # torch decoder, on the device


def test_cuda_layout_pool_new_maze_every_episode():
    from parity_bodies import body_layout_pool
    body_layout_pool(CudaEnv, N=600, T=150, L=37)
    body_layout_pool(CudaEnv, N=100, T=90, L=5, seed=9, base=123456)


def test_reset_to_another_layout_then_step_every_host_entry(tmp_path):
    """reset(layout_file=...) to a board of the same and of a different size, then step through the single-env
    reference API, pm_step_host_mapped and pm_step_host_pipelined: outputs follow the oracle on the NEW board
    (cached descriptors / staging buffers must not point at the buffers of the old one)."""
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    names = ["tinyCapture", "smallCapture", "tinyCapture", "bloxCapture"]
    env = gymPacMan_parallel_env(layout_file=os.path.join(LAYDIR, "tinyCapture.lay"), self_play=True, length=300)
    rng = np.random.default_rng(12)
    for name in names + ["capsSmall"]:
        if name == "capsSmall":                                    # same size as smallCapture? no: a crafted board file
            rows = helpers.load_traj("capsSmall")["layout"]
            f = tmp_path / "capsSmall.lay"
            f.write_text("\n".join(rows))
            path = str(f)
        else:
            rows, path = lay_rows(name), os.path.join(LAYDIR, name + ".lay")
        ref = oracle.OracleBatch(rows, 1, length=300)
        obs, info = env.reset(layout_file=path)
        robs, rlegal = ref.reset()
        assert tuple(obs[0].shape) == (8, len(rows), len(rows[0]))
        np.testing.assert_array_equal(torch.stack([obs[i] for i in range(4)]).cpu().numpy(), robs[0])
        for t in range(25):
            a = biased_actions(rng, 1, 0.6, t)[0]
            o, rew, term, info = env.step({i: int(a[i]) for i in range(4)})        # the fast single-env path
            r = ref.step(a[None])
            np.testing.assert_array_equal(torch.stack([o[i] for i in range(4)]).cpu().numpy(), r["obs"][0], err_msg=f"{name} t={t}")
            assert [rew[0], rew[1]] == list(r["reward"][0]) and info["score_change"] == int(r["score_change"][0])
            assert [sum(1 << x for x in info["legal_actions"][i]) for i in range(4)] == list(r["legal"][0])
    # batched env, host entries, after a change of board size
    N = 200
    benv = gymPacMan_parallel_env(layout_file=os.path.join(LAYDIR, "tinyCapture.lay"), self_play=True, length=300, num_envs=N)
    a0 = torch.as_tensor(biased_actions(rng, N)).pin_memory()
    benv.step_host_mapped(a0); benv.step_host_pipelined(a0, n_chunks=2); benv.step_host(a0)     # warm every cached descriptor
    for name in ("smallCapture", "tinyCapture"):
        rows = lay_rows(name)
        benv.reset(layout_file=os.path.join(LAYDIR, name + ".lay"))
        ref = oracle.OracleBatch(rows, N, length=300)
        ref.reset()
        for t in range(12):
            a = biased_actions(rng, N, 0.6, t)
            ah = torch.as_tensor(a).pin_memory()
            h = (benv.step_host_mapped, lambda x: benv.step_host_pipelined(x, n_chunks=3), None)[t % 3]
            r = ref.step(a)
            if h is None:
                out = benv.step_host(ah, want_obs=True)
                np.testing.assert_array_equal(out["reward"].numpy().T, r["reward"])
                np.testing.assert_array_equal(out["obs"].numpy().transpose(1, 0, 2, 3, 4), r["obs"])
            else:
                out = h(ah)
                np.testing.assert_array_equal(out["reward"].numpy(), r["reward"], err_msg=f"{name} t={t}")
                np.testing.assert_array_equal(out["legal"].numpy(), r["legal"])
            np.testing.assert_array_equal(benv._obs.cpu().numpy().transpose(1, 0, 2, 3, 4), r["obs"], err_msg=f"{name} t={t} obs")


def test_random_layout_single_env_new_maze_every_reset():
    """random_layout=True, one env (the reference's own shape): every reset() draws a fresh mazeGenerator board from
    the global `random` stream (gymPacMan.py:98-100, capture.py:905-911) and the episode that follows matches the
    oracle on that board - across several episodes on one env object."""
    import random
    from pacman_marl_2025_b200 import mazeGenerator
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    random.seed(99)
    env = gymPacMan_parallel_env(random_layout=True, self_play=True, length=40)
    random.seed(99)
    first = mazeGenerator.randomLayout().split("\n")
    assert env.layout.layoutText == first
    rng = np.random.default_rng(3)
    seen = set()
    for ep in range(4):
        random.seed(1000 + ep)
        want_rows = mazeGenerator.randomLayout().split("\n")
        random.seed(1000 + ep)
        obs, info = env.reset()
        assert env.layout.layoutText == want_rows
        seen.add("".join(want_rows))
        ref = oracle.OracleBatch(want_rows, 1, length=40)
        robs, _ = ref.reset()
        np.testing.assert_array_equal(torch.stack([obs[i] for i in range(4)]).cpu().numpy(), robs[0])
        done = False
        t = 0
        while not done:
            a = biased_actions(rng, 1, 0.6, t)[0]
            o, rew, term, info = env.step({i: int(a[i]) for i in range(4)})
            r = ref.step(a[None])
            np.testing.assert_array_equal(torch.stack([o[i] for i in range(4)]).cpu().numpy(), r["obs"][0], err_msg=f"ep {ep} t={t}")
            assert [rew[0], rew[1]] == list(r["reward"][0])
            done = term[0]
            t += 1
        assert t == 41
    assert len(seen) == 4


def test_random_layout_batched_pool_and_agent_keys():
    """random_layout=True with num_envs > 1: a pool of host-generated mazes on the device, a new board per episode
    picked in-kernel; envs really play different boards, and all invariants hold.  Also: in non-self-play the red
    agents are keyed by objects like in the reference (gymPacMan.py:33-40), and plain indices still resolve."""
    from pacman_marl_2025_b200.gymPacMan import OnDeviceAgent, gymPacMan_parallel_env
    N = 512
    env = gymPacMan_parallel_env(random_layout=True, self_play=False, enemieName="baselineTeam", length=30, num_envs=N,
                                 auto_reset=True, layout_pool=16, opponent_seed=3)
    assert isinstance(env.agents[0], OnDeviceAgent) and env.agents[1] == 1 and isinstance(env.agents[2], OnDeviceAgent)
    obs, info = env.reset()
    assert torch.equal(obs[env.agents[0]], obs[0]) and torch.equal(info["legal_actions"][2], info["legal_actions"][env.agents[2]])
    ids0 = env.export_state()["layout_id"].copy()
    assert len(set(ids0.tolist())) > 8
    g = torch.Generator().manual_seed(0)
    changed = 0
    for t in range(70):
        a = torch.randint(0, 5, (N, 4), generator=g)
        o, rew, term, info = env.step({1: a[:, 1], 3: a[:, 3]})
        assert set(o.keys()) == set(env.agents)
    s = env.export_state()
    changed = int((s["layout_id"] != ids0).sum())
    assert changed > N // 2 and int(s["episode"].min()) >= 2
    # walls of each env's CURRENT board show in plane 0 of its observation
    walls = np.stack([env._layouts[i].walls.array.T for i in s["layout_id"]]).astype(np.float32)
    np.testing.assert_array_equal(o[1][:, 0].cpu().numpy(), walls)


def test_cuda_static_plane_elision():
    """pm_step_out.obs_static_valid through the host class (elide_static_planes): identical tensors, fewer bytes."""
    from parity_bodies import body_elide_static_planes
    from pacman_marl_2025_b200 import mazeGenerator
    body_elide_static_planes(CudaEnv, lay_rows("tinyCapture"), N=700, T=120)
    body_elide_static_planes(CudaEnv, lay_rows("bloxCapture"), N=300, T=80)
    body_elide_static_planes(CudaEnv, helpers.load_traj("capsSmall")["layout"], N=200, T=90)
    body_elide_static_planes(CudaEnv, helpers.load_traj("odd15")["layout"], N=100, T=90)
    mazes = [mazeGenerator.generateMaze(8000 + s).split("\n") for s in range(9)]
    body_elide_static_planes(CudaEnv, mazes, N=150, T=90, layout_pool=True, random_seed=3)
    # the host entries carry the flag too
    rows = lay_rows("smallCapture")
    a_env, b_env = CudaEnv(rows, 257, length=40, auto_reset=True).env, CudaEnv(rows, 257, length=40, auto_reset=True).env
    b_env.elide_static_planes = True
    a_env.reset_batched(); b_env.reset_batched()
    rng = np.random.default_rng(1)
    for t in range(60):
        a = torch.as_tensor(biased_actions(rng, 257, 0.6, t)).pin_memory()
        a_env.step_host_mapped(a)
        (b_env.step_host_mapped, b_env.step_host, lambda x: b_env.step_host_pipelined(x, n_chunks=3))[t % 3](a)
        assert b_env._step_out.obs_static_valid == 1
        assert torch.equal(a_env._obs, b_env._obs), t


@pytest.mark.parametrize("name", ["smallCapture", "tinyCapture"])
def test_cuda_make_observation_matches_reference(name):
    """SURVEY row f4, partial observability: pm_make_observation vs GameState.makeObservation of the live reference."""
    from parity_bodies import body_make_observation
    body_make_observation(CudaEnv, name)
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    env = gymPacMan_parallel_env(layout_file=os.path.join(LAYDIR, name + ".lay"), self_play=True)
    env.reset()
    o = env.make_observation(0)
    pos = [env.game.state.getAgentPosition(i) for i in range(4)]
    assert o.getAgentPosition(0) == pos[0] and o.getAgentPosition(2) == pos[2]
    for e in (1, 3):                                             # an enemy shows iff a red agent stands within 5 cells of it
        seen = any(abs(pos[e][0] - pos[m][0]) + abs(pos[e][1] - pos[m][1]) <= 5 for m in (0, 2))
        assert (o.getAgentPosition(e) == pos[e]) if seen else (o.getAgentPosition(e) is None and o.getAgentState(e).configuration is None)
    assert len(o.getAgentDistances()) == 4 and all(abs(d - (abs(pos[0][0] - p[0]) + abs(pos[0][1] - p[1]))) <= 6
                                                    for d, p in zip(o.getAgentDistances(), pos))


def test_cuda_replay_of_a_reference_game_and_recorder_round_trip(tmp_path):
    """capture.replayGame on the GPU env: the reference's recorded game (tests/golden/replay_smallCapture.pkl) move by
    move, for a batch of copies; then a game played on the GPU env is recorded, saved in the same format and replayed
    to the same final state."""
    from pacman_marl_2025_b200 import replay as R
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    from parity_bodies import check_replay_move
    rp = R.load_replay(os.path.join(helpers.GOLDEN, "replay_smallCapture.pkl"))
    z = np.load(os.path.join(helpers.GOLDEN, "replay_smallCapture_states.npz"))
    seen = []
    env = R.replay_game(rp, num_envs=5, on_move=lambda k, i, a, e: (check_replay_move(z, k, e.export_state(), n=3), seen.append(k)))
    assert len(seen) == len(rp["actions"]) == 132 and env.game.state.data.score == int(z["score"][-1]) == 14
    # record a GPU game (baselineTeam vs random blue), write it, replay it
    g = gymPacMan_parallel_env(layout_file=os.path.join(LAYDIR, "tinyCapture.lay"), self_play=False, enemieName="baselineTeam",
                               length=60, num_envs=4, opponent_seed=9)
    g.reset_batched()
    rec = R.Recorder(g, n=2)
    gen = torch.Generator().manual_seed(3)
    for t in range(61):
        rec.after_step(g.step_batched(torch.randint(0, 5, (4, 4), generator=gen)))
    final = g.export_state()
    path = str(tmp_path / "replay-0")
    rec.save(path, redTeamName="baselineTeam", blueTeamName="caller")
    e2 = R.replay_game(R.load_replay(path))
    s2 = e2.export_state()
    for k in ("agents", "food", "caps", "score"):
        np.testing.assert_array_equal(s2[k][0], final[k][2], err_msg=k)
