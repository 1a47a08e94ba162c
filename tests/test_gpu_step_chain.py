"""pm_step_chain (ABI 8): T overlapping launches == T calls of pm_step.

The chain lets launch t+1 start before launch t has finished (programmatic dependent launch) and orders the two per
env tile with a release/acquire counter.  Whatever the overlap, every output of every step, the final state and the
final observations must be bit-identical to the sequential loop - at the full benchmark size (where launches really
overlap), at trainer size (8 envs per CTA), with the on-device opponents, with the learner views, on per-env mazes
with a layout pool, and when chains follow each other back to back."""
import numpy as np
import pytest
import torch

import helpers
from oracle import oracle
from parity_bodies import biased_actions, lay_rows

pytestmark = pytest.mark.gpu

SMALL = ("reward", "done", "legal", "score_change", "shaped_reward", "actions_used")


def make(layname, N, **kw):
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    kw.setdefault("self_play", True)
    rows = [str(r) for r in helpers.load_traj("capsSmall")["layout"]] if layname == "capsSmall" else lay_rows(layname)
    return gymPacMan_parallel_env(layouts=[rows], length=kw.pop("length", 60), num_envs=N, auto_reset=True, **kw)


def run_pair(layname, N, T, chunks, views=None, **kw):
    """The same action stream through step_batched (env a) and through step_chain in `chunks` calls (env b)."""
    a, b = make(layname, N, **kw), make(layname, N, **kw)
    bufs = {}
    for e, tag in ((a, "a"), (b, "b")):
        e.reset_batched()
        if views is not None:
            bufs[tag] = e.set_view_outputs(views, raw_obs=True)
    rng = np.random.default_rng(N + T)
    acts = torch.from_numpy(np.stack([biased_actions(rng, N, 0.6, t) for t in range(T)])).cuda()
    seq = {k: [] for k in SMALL}
    for t in range(T):
        o = a.step_batched(acts[t])
        for k in SMALL:
            seq[k].append(o[k].clone())
    t0 = 0
    for c in chunks:
        ob = b.step_chain(acts[t0:t0 + c], collect=True)
        for k in SMALL:
            want = torch.stack(seq[k][t0:t0 + c])
            assert torch.equal(ob["collected"][k], want), f"{k} differs in steps {t0}..{t0 + c}"
        t0 += c
    assert t0 == T
    torch.cuda.synchronize()
    assert torch.equal(a._obs, b._obs), "final observations"
    assert torch.equal(a._state, b._state), "final packed state"
    assert torch.equal(a._shaping, b._shaping)
    if views is not None:
        assert torch.equal(bufs["a"][0], bufs["b"][0]) and torch.equal(bufs["a"][1], bufs["b"][1]), "learner views"
    assert a.steps == b.steps
    return a, b


def test_chain_equals_sequential_steps_at_benchmark_size():
    # 65,536 envs x 4 agents on tinyCapture: 2,048 CTAs in two waves, consecutive launches really overlap
    a, b = run_pair("tinyCapture", 65536, 130, (1, 64, 2, 63), length=50)
    assert int(a._done.sum()) >= 0


def test_chain_against_the_oracle():
    """and not only against the product's own sequential path: final state and last-step outputs vs the C oracle"""
    N, T, length = 4096, 90, 40
    e = make("smallCapture", N, length=length)
    ref = oracle.OracleBatch(lay_rows("smallCapture"), N, length=length)
    e.reset_batched(); ref.reset(with_obs=False)
    rng = np.random.default_rng(7)
    acts = np.stack([biased_actions(rng, N, 0.6, t) for t in range(T)])
    o = e.step_chain(torch.from_numpy(acts).cuda(), collect=True)
    for t in range(T):
        r = ref.step(acts[t], with_obs=(t == T - 1), auto_reset=True, nthreads=0)
        idx = np.nonzero(r["done"])[0]
        if len(idx):                                     # restarted envs show the fresh episode's legal masks / observations
            ro, rl = ref.reset_outputs_for(idx)
            r["legal"][idx] = rl
            if r["obs"] is not None:
                r["obs"][idx] = ro
        np.testing.assert_array_equal(o["collected"]["reward"][t].cpu().numpy().view(np.uint64), r["reward"].T.copy().view(np.uint64), err_msg=f"t={t}")
        np.testing.assert_array_equal(o["collected"]["done"][t].cpu().numpy(), r["done"], err_msg=f"t={t}")
        np.testing.assert_array_equal(o["collected"]["legal"][t].cpu().numpy(), r["legal"], err_msg=f"t={t}")
    np.testing.assert_array_equal(o["obs"].cpu().numpy(), np.moveaxis(r["obs"], 1, 0))      # oracle: [N,4,...], env: [4,N,...]
    s = e.export_state()
    for n in range(0, N, 97):
        rs = ref.state(n)
        np.testing.assert_array_equal(s["agents"][n], rs["agents"])
        np.testing.assert_array_equal(s["food"][n], rs["food"])


@pytest.mark.parametrize("N", [1, 8, 1000, 1024])
def test_chain_small_batches(N):
    run_pair("smallCapture", N, 70, (35, 35), length=30)            # 8 envs per CTA, ragged last CTA


def test_chain_with_baseline_team_and_views():
    run_pair("tinyCapture", 16384, 60, (60,), views=False, self_play=False, enemieName="baselineTeam", opponent_seed=3, length=40)


def test_chain_with_random_team_views_only_red_learners():
    run_pair("capsSmall", 8192, 50, (25, 25), views=True, self_play=False, enemieName="randomTeam", opponent_seed=5, length=30)


def test_chain_with_layout_pool():
    from pacman_marl_2025_b200 import mazeGenerator
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    mazes = [mazeGenerator.generateMaze(900 + s).split("\n") for s in range(16)]
    envs = []
    for _ in range(2):
        e = gymPacMan_parallel_env(layouts=mazes, self_play=True, length=25, num_envs=4096, auto_reset=True, layout_pool=16,
                                   opponent_seed=11)
        e.reset_batched()
        envs.append(e)
    a, b = envs
    rng = np.random.default_rng(3)
    acts = torch.from_numpy(np.stack([biased_actions(rng, 4096, 0.6, t) for t in range(80)])).cuda()
    for t in range(80):
        a.step_batched(acts[t])
    b.step_chain(acts)
    assert torch.equal(a._obs, b._obs) and torch.equal(a._state, b._state) and torch.equal(a._layout_id, b._layout_id)


def test_chain_with_static_plane_elision_and_packed_bits():
    """the opt-in outputs ride along: obs_static_valid (skip planes that cannot have changed) inside a chain leaves the
    same observation tensor as plain sequential steps, and the packed observation bits of the last step agree"""
    N, T = 8192, 45
    a = make("smallCapture", N, length=20)
    b = make("smallCapture", N, length=20, elide_static_planes=True)
    for e in (a, b):
        e.reset_batched()
        e.set_bits_output(True)
    rng = np.random.default_rng(5)
    acts = torch.from_numpy(np.stack([biased_actions(rng, N, 0.6, t) for t in range(T)])).cuda()
    for t in range(T):
        oa = a.step_batched(acts[t])
    ob = b.step_chain(acts[:20])
    ob = b.step_chain(acts[20:])
    assert torch.equal(a._obs, b._obs) and torch.equal(a._state, b._state)
    assert torch.equal(oa["obs_bits"], ob["obs_bits"])


def test_chain_with_pinned_host_actions_and_records():
    """pm_step_chain_host_mapped: the per-step records in pinned host memory == step_host_mapped called T times"""
    N, T = 16384, 40
    a, b = make("tinyCapture", N, length=25), make("tinyCapture", N, length=25)
    a.reset_batched(); b.reset_batched()
    rng = np.random.default_rng(21)
    acts = torch.from_numpy(np.stack([biased_actions(rng, N, 0.6, t) for t in range(T)])).pin_memory()
    seq = []
    for t in range(T):
        r = a.step_host_mapped(acts[t])
        seq.append({k: v.clone() for k, v in r.items()})
    got = b.step_chain_host_mapped(acts)
    for k in ("reward", "shaped_reward", "score_change", "legal", "done"):
        want = torch.stack([s[k] for s in seq])
        assert torch.equal(got[k], want), k
    assert int(got["done"].sum()) >= N
    assert torch.equal(a._obs, b._obs) and torch.equal(a._state, b._state)
    with pytest.raises(ValueError):
        b.step_chain_host_mapped(torch.zeros((2, N, 4), dtype=torch.uint8))          # pageable


def test_chain_argument_errors():
    e = make("tinyCapture", 64)
    e.reset_batched()
    with pytest.raises(ValueError):
        e.step_chain(torch.zeros((3, 63, 4), dtype=torch.uint8, device="cuda"))
    import ctypes as C
    from pacman_marl_2025_b200 import _lib
    L = _lib.load()
    acts = torch.zeros((2, 64, 4), dtype=torch.uint8, device="cuda")
    outs = (_lib.PmStepOut * 1)()
    C.memmove(C.byref(outs[0]), C.byref(e._step_out), C.sizeof(_lib.PmStepOut))
    scratch = torch.zeros(64, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.pm_step_chain(C.byref(e._desc), acts.data_ptr(), 1, 256, 0, outs, 1, scratch.data_ptr(), st) == -1     # n_steps
    assert L.pm_step_chain(C.byref(e._desc), acts.data_ptr(), 1, 256, 2, outs, 1, None, st) == -1                  # scratch
    assert L.pm_step_chain(C.byref(e._desc), acts.data_ptr(), 4, 256, 2, outs, 1, scratch.data_ptr(), st) == -1    # itemsize
    e._desc.agent_begin, e._desc.agent_end = 0, 2
    assert L.pm_step_chain(C.byref(e._desc), acts.data_ptr(), 1, 256, 2, outs, 1, scratch.data_ptr(), st) == -1    # sub-moves
    e._desc.agent_begin, e._desc.agent_end = 0, 0
    assert L.pm_step_chain(C.byref(e._desc), acts.data_ptr(), 1, 256, 2, outs, 1, scratch.data_ptr(), st) == 0
    assert L.pm_chain_scratch_bytes(64) == 32 and L.pm_chain_scratch_bytes(65) == 36
