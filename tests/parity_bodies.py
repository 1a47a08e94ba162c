"""Backend-agnostic bodies of the batched parity tests.  `make_env(layouts, n_envs,
**kw)` returns an object with reset(mask) / step(actions, dtype) / get_state()
returning numpy arrays: tests/emu_driver.EmuEnv (host emulation of the kernel
logic, CPU CI) or tests/test_gpu_parity.CudaEnv (the real C ABI on a B200)."""
import os

import numpy as np

import helpers
from oracle import oracle

LAYDIR = os.path.join(os.path.dirname(helpers.GOLDEN), "..", "pacman_marl_2025_b200", "layouts")


def lay_rows(name):
    return [ln.strip() for ln in open(os.path.join(LAYDIR, name + ".lay"))]


def biased_actions(rng, N, p_forward=0.45, t=None):
    """Random actions that drift red east and blue west so teams meet, eat and die;
    with `t`, every env alternates raid / retreat phases so pellets get returned."""
    a = rng.integers(0, 5, size=(N, 4))
    p_home = 0.12
    if t is not None and (t % 50) >= 30:
        p_forward, p_home = 0.08, 0.6
    fwd = rng.random((N, 4)) < p_forward
    a[:, 0::2] = np.where(fwd[:, 0::2], 1, a[:, 0::2])
    a[:, 1::2] = np.where(fwd[:, 1::2], 3, a[:, 1::2])
    home = rng.random((N, 4)) < p_home
    a[:, 0::2] = np.where(home[:, 0::2], 3, a[:, 0::2])
    a[:, 1::2] = np.where(home[:, 1::2], 1, a[:, 1::2])
    return a.astype(np.uint8)


def compare_batch(o, ref, t, obs_envs):
    np.testing.assert_array_equal(o["reward"].view(np.uint64), ref["reward"].T.copy().view(np.uint64), err_msg=f"t={t} reward")
    np.testing.assert_array_equal(o["done"], ref["done"], err_msg=f"t={t} done")
    np.testing.assert_array_equal(o["legal"], ref["legal"], err_msg=f"t={t} legal")
    np.testing.assert_array_equal(o["score_change"], ref["score_change"], err_msg=f"t={t} score_change")
    np.testing.assert_array_equal(o["shaping"].view(np.uint64), ref["shaping"].T.copy().view(np.uint64), err_msg=f"t={t} shaping")
    want = (2.0 * ref["reward"][:, [0, 1, 0, 1]] + 0.1 * ref["shaping"]).astype(np.float32).T
    np.testing.assert_array_equal(o["shaped"].view(np.uint32), want.copy().view(np.uint32), err_msg=f"t={t} shaped")
    for n in obs_envs:
        np.testing.assert_array_equal(o["obs"][:, n], ref["obs"][n], err_msg=f"t={t} obs env {n}")


def body_large_batch(make_env, layname, N, T, length, state_stride=101, obs_stride=37, expect_score=True):
    """Many envs with independent action streams, reset-by-mask after done, every
    output compared with the oracle at every step."""
    rows = lay_rows(layname)
    rng = np.random.default_rng(sum(map(ord, layname)))
    env = make_env(rows, N, length=length)
    ref = oracle.OracleBatch(rows, N, length=length)
    obs, legal = env.reset()
    robs, rlegal = ref.reset()
    np.testing.assert_array_equal(legal, rlegal)
    np.testing.assert_array_equal(obs.transpose(1, 0, 2, 3, 4), robs)
    events = dict(done=0, score=0)
    for t in range(T):
        a = biased_actions(rng, N, 0.6, t)
        o = env.step(a)
        r = ref.step(a, with_obs=True, nthreads=0)
        compare_batch(o, r, t, range(N) if t % 16 == 0 else range(0, N, obs_stride))
        events["done"] += int(o["done"].sum())
        events["score"] += int((o["score_change"] != 0).sum())
        if t % 20 == 0 or t == T - 1:
            s = env.get_state()
            for n in range(0, N, state_stride):
                rs = ref.state(n)
                np.testing.assert_array_equal(s["agents"][n], rs["agents"])
                np.testing.assert_array_equal(s["food"][n], rs["food"])
                assert s["score"][n] == rs["score"] and s["steps"][n] == rs["steps"]
        d = o["done"].astype(bool)
        if d.any():
            obs2, legal2 = env.reset(mask=d)
            idx = np.nonzero(d)[0]
            for n in idx:
                ref.reset_one(int(n))
            ro, rl = ref.reset_outputs_for(idx)
            np.testing.assert_array_equal(legal2[d], rl)
            np.testing.assert_array_equal(obs2[:, d].transpose(1, 0, 2, 3, 4), ro)
    assert events["done"] > 0 and (events["score"] > 0 or not expect_score), events


def body_per_env_mazes(make_env, N, T):
    """BASELINE config 4: one mazeGenerator board per env (L = N)."""
    from pacman_marl_2025_b200 import mazeGenerator
    mazes = [mazeGenerator.generateMaze(1000 + s).split("\n") for s in range(N)]
    env = make_env(mazes, N, length=100)
    ref = oracle.OracleBatch(mazes, N, length=100)
    obs, legal = env.reset()
    robs, rlegal = ref.reset()
    np.testing.assert_array_equal(legal, rlegal)
    np.testing.assert_array_equal(obs.transpose(1, 0, 2, 3, 4), robs)
    rng = np.random.default_rng(5)
    for t in range(T):
        a = biased_actions(rng, N, 0.6)
        o = env.step(a)
        r = ref.step(a, with_obs=True, nthreads=0, auto_reset=True)
        compare_batch(o, r, t, range(N))
        d = o["done"].astype(bool)
        if d.any():
            env.reset(mask=d)


def body_int64_out_of_range(make_env):
    rows = lay_rows("smallCapture")
    N = 64
    env = make_env(rows, N)
    ref = oracle.OracleBatch(rows, N)
    env.reset(); ref.reset()
    rng = np.random.default_rng(3)
    for t in range(40):
        a = rng.integers(-2, 8, size=(N, 4))
        o = env.step(a, dtype=np.int64)
        r = ref.step(a, with_obs=True)
        compare_batch(o, r, t, range(N))


def body_observe_current_state(make_env):
    rows = lay_rows("smallCapture")
    N = 33
    env = make_env(rows, N)
    ref = oracle.OracleBatch(rows, N)
    env.reset(); ref.reset()
    rng = np.random.default_rng(9)
    for t in range(25):
        a = biased_actions(rng, N, 0.7)
        env.step(a); ref.step(a, with_obs=False)
    cur = env.observe()
    for n in range(N):
        for i in range(4):
            np.testing.assert_array_equal(cur[i, n], ref.observe(n, i))
    return env, rng


def body_auto_reset(make_env, layname="tinyCapture", N=80, T=170, length=40):
    """pm_env_desc.auto_reset: an env whose step terminates restarts inside the same launch.
    Terminal-step outputs (reward, done, score_change, shaping) are the oracle's step outputs;
    obs / legal / state of a finished env are the oracle's reset outputs."""
    rows = layname if isinstance(layname, list) else lay_rows(layname)
    rng = np.random.default_rng(11)
    env = make_env(rows, N, length=length, auto_reset=True)
    ref = oracle.OracleBatch(rows, N, length=length)
    env.reset(); ref.reset()
    n_done = 0
    for t in range(T):
        a = biased_actions(rng, N, 0.6, t)
        o = env.step(a)
        r = ref.step(a, with_obs=True, nthreads=0, auto_reset=True)     # oracle: reset after done, outputs = terminal step
        d = r["done"].astype(bool)
        np.testing.assert_array_equal(o["done"], r["done"], err_msg=f"t={t}")
        np.testing.assert_array_equal(o["reward"].view(np.uint64), r["reward"].T.copy().view(np.uint64), err_msg=f"t={t}")
        np.testing.assert_array_equal(o["score_change"], r["score_change"])
        np.testing.assert_array_equal(o["shaping"].view(np.uint64), r["shaping"].T.copy().view(np.uint64), err_msg=f"t={t}")
        want_obs, want_legal = r["obs"].copy(), r["legal"].copy()
        if d.any():
            idx = np.nonzero(d)[0]
            ro, rl = ref.reset_outputs_for(idx)
            want_obs[idx], want_legal[idx] = ro, rl
            n_done += len(idx)
        np.testing.assert_array_equal(o["legal"], want_legal, err_msg=f"t={t} legal")
        np.testing.assert_array_equal(o["obs"].transpose(1, 0, 2, 3, 4), want_obs, err_msg=f"t={t} obs")
        if t % 13 == 0 or d.any():
            s = env.get_state()
            for n in list(np.nonzero(d)[0][:4]) + [0, N - 1]:
                rs = ref.state(int(n))
                np.testing.assert_array_equal(s["agents"][n], rs["agents"])
                np.testing.assert_array_equal(s["food"][n], rs["food"])
                np.testing.assert_array_equal(s["caps"][n], rs["caps"])
                assert s["score"][n] == rs["score"] and s["steps"][n] == rs["steps"]
    assert n_done >= N


def body_env_ranges(make_env, set_range, N=96):
    """pm_env_desc.env_begin / env_count: launching the batch as consecutive ranges gives
    the same outputs and state as one launch over all envs."""
    rows = lay_rows("smallCapture")
    rng = np.random.default_rng(21)
    whole, parts = make_env(rows, N, length=30, auto_reset=True), make_env(rows, N, length=30, auto_reset=True)
    whole.reset(); parts.reset()
    cuts = [0, 32, 64, N]
    for t in range(70):
        a = biased_actions(rng, N, 0.6, t)
        o = whole.step(a)
        acc = None
        for b, e in zip(cuts[:-1], cuts[1:]):
            set_range(parts, b, e - b)
            p = parts.step(a)
            if acc is None:
                acc = {k: np.array(v, copy=True) for k, v in p.items()}
            else:
                for k, v in p.items():
                    if k in ("obs", "reward", "shaping", "shaped"):
                        acc[k][:, b:e] = v[:, b:e]
                    else:
                        acc[k][b:e] = v[b:e]
        set_range(parts, 0, 0)
        for k in o:
            np.testing.assert_array_equal(o[k], acc[k], err_msg=f"t={t} {k}")
    s1, s2 = whole.get_state(), parts.get_state()
    for k in s1:
        np.testing.assert_array_equal(s1[k], s2[k])


def body_shard_invariant_random_agents(make_env, N=64):
    """env_index_base: the on-device randomTeam draws are keyed by the GLOBAL env index, so
    two shards [0,N/2) and [N/2,N) reproduce one batch of N."""
    rows = lay_rows("smallCapture")
    kw = dict(length=25, auto_reset=True, random_agent_mask=0b0101, random_seed=99)
    full = make_env(rows, N, **kw)
    sh = [make_env(rows, N // 2, env_index_base=i * (N // 2), **kw) for i in range(2)]
    full.reset(); [s.reset() for s in sh]
    rng = np.random.default_rng(4)
    for t in range(80):
        a = biased_actions(rng, N, 0.6, t)
        o = full.step(a)
        os_ = [sh[i].step(a[i * N // 2:(i + 1) * N // 2]) for i in range(2)]
        for k in ("obs", "reward", "shaped"):
            np.testing.assert_array_equal(o[k], np.concatenate([os_[0][k], os_[1][k]], axis=1), err_msg=f"t={t} {k}")
        for k in ("done", "legal", "score_change"):
            np.testing.assert_array_equal(o[k], np.concatenate([os_[0][k], os_[1][k]], axis=0), err_msg=f"t={t} {k}")
    # and a different base gives a different stream
    other = make_env(rows, N // 2, env_index_base=10_000, **kw)
    other.reset()
    fresh = make_env(rows, N // 2, env_index_base=0, **kw)
    fresh.reset()
    a = biased_actions(rng, N // 2)
    differs = False
    for t in range(20):
        differs |= not np.array_equal(other.step(a)["obs"], fresh.step(a)["obs"])
    assert differs


def big_board(W=32, H=32, seed=0):
    """A crafted W x H capture board (up to PM_MAX_CELLS = 1024 cells): closed border, random inner walls kept
    sparse enough to stay connected around the agents, mirrored food, two capsules, four starts."""
    rng = np.random.default_rng(seed)
    g = [[" "] * W for _ in range(H)]
    for x in range(W):
        g[0][x] = g[H - 1][x] = "%"
    for y in range(H):
        g[y][0] = g[y][W - 1] = "%"
    for y in range(2, H - 2, 2):
        for x in range(2, W // 2, 1):
            if rng.random() < 0.45:
                g[y][x] = "%"
                g[H - 1 - y][W - 1 - x] = "%"
    for y in range(1, H - 1):
        for x in range(1, W // 2):
            if g[y][x] == " " and rng.random() < 0.18:
                g[y][x] = "."
                if g[H - 1 - y][W - 1 - x] == " ":
                    g[H - 1 - y][W - 1 - x] = "."
    g[1][1], g[2][1] = "1", "3"
    g[H - 2][W - 2], g[H - 3][W - 2] = "2", "4"
    g[1][W // 2 - 2], g[H - 2][W // 2 + 1] = "o", "o"
    return ["".join(r) for r in g]


def body_max_size_board(make_env, N=40, T=150):
    """32 x 32 = PM_MAX_CELLS cells: the largest board the ABI accepts (and 33 x 32 is refused)."""
    rows = big_board()
    env = make_env(rows, N, length=70, auto_reset=True)
    ref = oracle.OracleBatch(rows, N, length=70)
    obs, legal = env.reset()
    robs, rlegal = ref.reset()
    np.testing.assert_array_equal(legal, rlegal)
    np.testing.assert_array_equal(obs.transpose(1, 0, 2, 3, 4), robs)
    rng = np.random.default_rng(6)
    for t in range(T):
        a = biased_actions(rng, N, 0.7, t)
        o = env.step(a)
        r = ref.step(a, with_obs=True, nthreads=0, auto_reset=True)
        keep = ~r["done"].astype(bool)
        np.testing.assert_array_equal(o["reward"].view(np.uint64), r["reward"].T.copy().view(np.uint64), err_msg=f"t={t}")
        np.testing.assert_array_equal(o["done"], r["done"])
        np.testing.assert_array_equal(o["score_change"], r["score_change"])
        np.testing.assert_array_equal(o["legal"][keep], r["legal"][keep])
        np.testing.assert_array_equal(o["obs"].transpose(1, 0, 2, 3, 4)[keep], r["obs"][keep], err_msg=f"t={t} obs")


# ---- learner views (f2): numpy restatement of the trainer's view functions, themselves pinned to the reference
# by tests/test_mappo_host.py (canonicalize_obs / merge_obs_for_critic, pacman_mappo_resnet.py:215-229, 267-274) ----
def np_canonicalize(obs, red):
    if not red:
        return obs
    return obs[..., ::-1][..., [0, 1, 3, 2, 4, 5, 7, 6], :, :]


def np_views(obs4, red):
    """obs4 [4,N,8,H,W] f32 -> (actor [2,N,8,H,W], critic [N,8,H,W]) as uint8."""
    ids = (0, 2) if red else (1, 3)
    a, b = np_canonicalize(obs4[ids[0]], red), np_canonicalize(obs4[ids[1]], red)
    merged = a.copy()
    merged[:, 1] = np.maximum(a[:, 1], b[:, 1])
    merged[:, 4] = 0
    return np.stack([a, b]).astype(np.uint8), merged.astype(np.uint8)


def body_learner_views(make_env, rows, N=48, T=120, length=90, **kw):
    """view_actor / view_critic written by the encoder == canonicalize + merge applied to the raw observations of
    the same step, for both learner teams, across auto-resets, kills, carrying and capsules."""
    rng = np.random.default_rng(33)
    carried = 0
    for red in (False, True):
        env = make_env(rows, N, length=length, auto_reset=True, **kw)
        env.set_views(red)
        obs, _ = env.reset()
        va, vc = env.observe_views(red)
        wa, wc = np_views(obs, red)
        np.testing.assert_array_equal(va, wa, err_msg=f"reset views red={red}")
        np.testing.assert_array_equal(vc, wc)
        for t in range(T):
            o = env.step(biased_actions(rng, N, 0.6, t))
            wa, wc = np_views(o["obs"], red)
            np.testing.assert_array_equal(o["view_actor"], wa, err_msg=f"t={t} actor views red={red}")
            np.testing.assert_array_equal(o["view_critic"], wc, err_msg=f"t={t} critic view red={red}")
            carried += int((wc[:, 1] > 1).sum())
    assert carried > 0                # own-cell values above 1 (carrying) were exercised


def body_layout_ids(make_env, N=40, T=90):
    """A table of L boards with an explicit per-env index (pm_env_desc.layout_id): env n plays layouts[ids[n]]."""
    from pacman_marl_2025_b200 import mazeGenerator
    L = 5
    mazes = [mazeGenerator.generateMaze(2000 + s).split("\n") for s in range(L)]
    ids = np.random.default_rng(1).integers(0, L, size=N)
    env = make_env(mazes, N, length=60, auto_reset=True, layout_ids=ids)
    ref = oracle.OracleBatch([mazes[i] for i in ids], N, length=60)
    obs, legal = env.reset()
    robs, rlegal = ref.reset()
    np.testing.assert_array_equal(legal, rlegal)
    np.testing.assert_array_equal(obs.transpose(1, 0, 2, 3, 4), robs)
    rng = np.random.default_rng(12)
    for t in range(T):
        a = biased_actions(rng, N, 0.6, t)
        o = env.step(a)
        r = ref.step(a, with_obs=True, nthreads=0, auto_reset=True)
        keep = ~r["done"].astype(bool)
        np.testing.assert_array_equal(o["reward"].view(np.uint64), r["reward"].T.copy().view(np.uint64), err_msg=f"t={t}")
        np.testing.assert_array_equal(o["done"], r["done"])
        np.testing.assert_array_equal(o["obs"].transpose(1, 0, 2, 3, 4)[keep], r["obs"][keep], err_msg=f"t={t} obs")


def body_bots_and_views_on_distinct_mazes(make_env, N=12, T=70):
    """Per-env boards (L = N) together with the baselineTeam look-ahead (one distance table per board) and the
    mirrored learner views (mirrored wall plane per board)."""
    from pacman_marl_2025_b200 import mazeGenerator
    mazes = [mazeGenerator.generateMaze(3000 + s).split("\n") for s in range(N)]
    env = make_env(mazes, N, length=50, auto_reset=True, bot_kind=[2, 0, 3, 0], bot_first_best=True)
    ref = oracle.OracleBatch(mazes, N, length=50)
    ref.set_bots([2, 0, 3, 0])
    env.set_views(True)
    env.reset(); ref.reset()
    rng = np.random.default_rng(41)
    for t in range(T):
        a = biased_actions(rng, N, 0.5, t)
        o = env.step(a)
        r = ref.step_bots(a, with_obs=True, auto_reset=True, nthreads=0)
        np.testing.assert_array_equal(o["actions_used"][:, [0, 2]], r["actions_used"][:, [0, 2]], err_msg=f"t={t} bot actions")
        np.testing.assert_array_equal(o["reward"].view(np.uint64), r["reward"].T.copy().view(np.uint64), err_msg=f"t={t}")
        keep = ~r["done"].astype(bool)
        np.testing.assert_array_equal(o["obs"].transpose(1, 0, 2, 3, 4)[keep], r["obs"][keep], err_msg=f"t={t} obs")
        wa, wc = np_views(o["obs"], True)
        np.testing.assert_array_equal(o["view_actor"], wa, err_msg=f"t={t} mirrored actor views")
        np.testing.assert_array_equal(o["view_critic"], wc, err_msg=f"t={t} mirrored critic view")


def body_step_in_several_launches(make_env, set_agents, rows, N=40, T=120, splits=((0, 1), (1, 2), (2, 3), (3, 4)), **kw):
    """pm_env_desc.agent_begin/agent_end: a step played as several launches (the hook for host-side opponents that
    decide at their own sub-move) gives exactly the outputs, observations and state of the one-launch step."""
    rng = np.random.default_rng(5)
    whole, parts = make_env(rows, N, length=50, auto_reset=True, **kw), make_env(rows, N, length=50, auto_reset=True, **kw)
    whole.reset(); parts.reset()
    for t in range(T):
        a = biased_actions(rng, N, 0.6, t)
        o = whole.step(a)
        obs = np.zeros_like(o["obs"])
        for b, e in splits:
            set_agents(parts, b, e)
            p = parts.step(a)
            obs[b:e] = p["obs"][b:e]
        set_agents(parts, 0, 0)
        for k in ("reward", "done", "legal", "score_change", "shaping", "shaped", "actions_used"):
            if k in o:
                np.testing.assert_array_equal(p[k], o[k], err_msg=f"t={t} {k}")
        keep = ~o["done"].astype(bool)     # a restarted env shows the new episode in all four observations only in the last launch
        np.testing.assert_array_equal(obs[:, keep], o["obs"][:, keep], err_msg=f"t={t} obs")
        np.testing.assert_array_equal(p["obs"][3:, ~keep], o["obs"][3:, ~keep])
    s1, s2 = whole.get_state(), parts.get_state()
    for k in s1:
        np.testing.assert_array_equal(s1[k], s2[k])


# ---- ABI 7 ---------------------------------------------------------------------------------------------------
def body_obs_bits(make_env, rows, N=40, T=130, **kw):
    """pm_step_out.obs_bits / pm_observe_bits: the packed observation (8*H*W-bit string + own-cell patch word + zero
    padding) decodes to exactly the fp32 observation of the same step."""
    from pacman_marl_2025_b200.gymPacMan import decode_obs_bits
    H, W = len(rows), len(rows[0])
    ncat = (8 * H * W + 31) // 32
    env = make_env(rows, N, length=100, auto_reset=True, **kw)
    obs, _ = env.reset()
    np.testing.assert_array_equal(decode_obs_bits(env.observe_bits(), H, W), obs)
    env.set_bits(True)
    rng = np.random.default_rng(3)
    patched = 0
    for t in range(T):
        o = env.step(biased_actions(rng, N, 0.6, t))
        bits = np.asarray(o["obs_bits"]).view(np.uint32)
        np.testing.assert_array_equal(decode_obs_bits(bits, H, W), o["obs"], err_msg=f"t={t}")
        assert not bits[..., ncat + 1:].any(), "padding words must be zero"
        patched += int((bits[..., ncat] != 0xFFFFFFFF).sum())
    assert patched > 0                # own-cell values above 1 (carrying) were exercised


M64 = (1 << 64) - 1


def py_random_u32(seed, n, episode, ctr):
    """pm_random_u32 (csrc/pm_env_core.cuh) restated."""
    z = (seed + 0x9E3779B97F4A7C15 * (n + 1) + 0xD1B54A32D192ED03 * ((episode << 32) + ctr + 1)) & M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z >> 32


def py_pool_layout(seed, n, episode, L):
    return py_random_u32(seed ^ 0x6C61796F75747321, n, episode, 0xFFFFFFF0) % L


def body_layout_pool(make_env, N=24, T=140, L=7, seed=77, base=0):
    """pm_env_desc.layout_pool: a new maze every episode (gymPacMan.py:98-100).  Env n plays board
    hash(seed, global env index, episode) % L; the oracle is re-initialised on that board whenever an env restarts."""
    from pacman_marl_2025_b200 import mazeGenerator
    mazes = [mazeGenerator.generateMaze(7000 + s).split("\n") for s in range(L)]
    env = make_env(mazes, N, length=30, auto_reset=True, layout_pool=True, random_seed=seed, env_index_base=base)
    obs, legal = env.reset()
    # the seed the kernels hash: the caller's seed, advanced once per FULL reset by the host class (so that successive
    # rollouts do not replay the same boards) - read it back from the descriptor
    seed = env.pool_seed()
    ids = np.array([py_pool_layout(seed, base + n, 0, L) for n in range(N)])
    episode = np.zeros(N, np.int64)
    ref = oracle.OracleBatch([mazes[i] for i in ids], N, length=30)
    robs, rlegal = ref.reset()
    np.testing.assert_array_equal(env.get_state()["layout_id"], ids)
    np.testing.assert_array_equal(legal, rlegal)
    np.testing.assert_array_equal(obs.transpose(1, 0, 2, 3, 4), robs)
    rng = np.random.default_rng(8)
    L_ = oracle.lib()
    moved = 0
    for t in range(T):
        a = biased_actions(rng, N, 0.6, t)
        o = env.step(a)
        r = ref.step(a, with_obs=True, nthreads=0)
        d = r["done"].astype(bool)
        np.testing.assert_array_equal(o["done"], r["done"], err_msg=f"t={t}")
        np.testing.assert_array_equal(o["reward"].view(np.uint64), r["reward"].T.copy().view(np.uint64), err_msg=f"t={t}")
        np.testing.assert_array_equal(o["score_change"], r["score_change"])
        want_obs, want_legal = r["obs"].copy(), r["legal"].copy()
        for n in np.nonzero(d)[0]:                  # restart on the next board of this env's sequence
            episode[n] += 1
            new = py_pool_layout(seed, base + int(n), int(episode[n]), L)
            moved += int(new != ids[n])
            ids[n] = new
            text, W, H = oracle._rows(mazes[new])
            assert L_.pmo_env_init(ref._env(int(n)), text, W, H, 30, 1, 1) == 0
            L_.pmo_reset(ref._env(int(n)))
        if d.any():
            idx = np.nonzero(d)[0]
            ro, rl = ref.reset_outputs_for(idx)
            want_obs[idx], want_legal[idx] = ro, rl
        np.testing.assert_array_equal(o["legal"], want_legal, err_msg=f"t={t} legal")
        np.testing.assert_array_equal(o["obs"].transpose(1, 0, 2, 3, 4), want_obs, err_msg=f"t={t} obs")
        s = env.get_state()
        np.testing.assert_array_equal(s["layout_id"], ids, err_msg=f"t={t} layout ids")
        np.testing.assert_array_equal(s["episode"], episode)
    assert moved > N                  # envs really changed boards


def body_elide_static_planes(make_env, layouts, N=40, T=120, **kw):
    """pm_step_out.obs_static_valid: with the env's persistent observation buffer the step may skip the elements that
    cannot have changed (walls; the capsule planes of capsule-free batches).  The tensors read back equal those of an
    env that writes everything - across restarts, and across moves to another board (layout pool)."""
    rng = np.random.default_rng(19)
    full, lean = make_env(layouts, N, length=35, auto_reset=True, **kw), make_env(layouts, N, length=35, auto_reset=True, **kw)
    lean.set_elide(True)
    o1, l1 = full.reset()
    o2, l2 = lean.reset()
    np.testing.assert_array_equal(o1, o2)
    for t in range(T):
        a = biased_actions(rng, N, 0.6, t)
        x, y = full.step(a), lean.step(a)
        np.testing.assert_array_equal(x["obs"], y["obs"], err_msg=f"t={t}")
        np.testing.assert_array_equal(x["reward"], y["reward"])


def body_make_observation(make_env, name):
    """GameState.makeObservation (capture.py:268-292) vs a fixture recorded from the live reference: visibility masks
    and - fed the recorded random.choice indices - the noisy sonar distances, after every step of the episodes."""
    z = np.load(os.path.join(helpers.GOLDEN, f"partial_{name}.npz"))
    rows = [str(r) for r in z["layout"]]
    N = 3
    env = make_env(rows, N, length=int(z["length"]))
    starts = list(z["ep_start"]) + [len(z["actions"])]
    hidden = 0
    for ep in range(len(z["ep_start"])):
        env.reset()
        for t in range(starts[ep], starts[ep + 1]):
            env.step(np.tile(z["actions"][t], (N, 1)))
            vis, noisy = env.make_observation(np.tile(z["draws"][t][None], (N, 1, 1)))
            np.testing.assert_array_equal(vis[1], z["visible"][t], err_msg=f"t={t} visibility")
            np.testing.assert_array_equal(noisy[1], z["noisy"][t], err_msg=f"t={t} noisy distances")
            hidden += int((z["visible"][t] != 15).sum())
            # without draws: hashed noise stays inside the sonar range of the true distance
            v2, n2 = env.make_observation()
            true = noisy[1] - (z["draws"][t].astype(np.int64) - 6)
            assert (np.abs(n2[1] - true) <= 6).all() and (v2 == vis).all()
    assert hidden > 100


def check_replay_move(z, k, st, n=0):
    """state after recorded move k == the state capture.replayGame reaches (fixture from the live reference)."""
    np.testing.assert_array_equal(st["agents"][n][:, 0:2], z["pos"][k], err_msg=f"move {k}: positions")
    np.testing.assert_array_equal(st["agents"][n][:, 2:7], z["misc"][k], err_msg=f"move {k}: dir/isPacman/scared/carrying/returned")
    W_H = st["food"].shape[1]
    np.testing.assert_array_equal(st["food"][n], helpers.unpack_cells(z["food"][k], W_H), err_msg=f"move {k}: food")
    assert int(st["score"][n]) == int(z["score"][k]), f"move {k}: score"
