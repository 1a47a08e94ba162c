"""Pins the CPU oracle (oracle/pm_oracle.c) to the unmodified reference.

Every fixture under tests/golden/ was recorded by tests/golden/make_golden.py
from the live Python reference; here the oracle replays the same action
sequences / inputs and must agree on every field (bit-exact, incl. fp64 rewards).
"""
import hashlib
import struct

import numpy as np
import pytest

import helpers
from oracle import oracle


class OracleDriver:
    def __init__(self, fx):
        self.env = oracle.OracleBatch(fx["layout"], 1, length=int(fx["length"]),
                                      reward_forLegalAction=bool(fx["reward_forLegalAction"]),
                                      defenceReward=bool(fx["defenceReward"]))

    def reset(self):
        obs, legal = self.env.reset()
        return obs[0], legal[0]

    def step(self, a):
        o = self.env.step(a)
        return {k: (v[0] if v is not None else None) for k, v in o.items()}

    def state(self):
        return self.env.state(0)


@pytest.mark.parametrize("name", helpers.traj_names())
def test_trajectory_matches_reference(name):
    fx = helpers.load_traj(name)
    helpers.replay(fx, OracleDriver(fx), where=name)


@pytest.mark.parametrize("lay", ["tinyCapture", "smallCapture", "bloxCapture"])
def test_survey_a4_trajectory_digest(lay):
    """SURVEY A.4 recipe: 10 episodes of uniform random actions, SHA-256 over
    obs / rewards / legal masks / score / done."""
    import os
    rows = [ln.strip() for ln in open(os.path.join(os.path.dirname(helpers.GOLDEN), "..", "pacman_marl_2025_b200",
                                                   "layouts", lay + ".lay"))]
    env = oracle.OracleBatch(rows, 1, length=300)
    rng = np.random.default_rng(12345)
    h = hashlib.sha256()
    for _ in range(10):
        env.reset()
        done = False
        while not done:
            a = rng.integers(0, 5, size=4)
            o = env.step(a)
            score = env.state(0)["score"]
            for i in range(4):
                h.update(o["obs"][0, i].astype("<f4").tobytes())
                h.update(struct.pack("<d", o["reward"][0, i % 2]))
                h.update(bytes([int(o["legal"][0, i])]))
            done = bool(o["done"][0])
            h.update(struct.pack("<i?", score, done))
    assert h.hexdigest() == helpers.digests()["survey_traj_sha256"][lay]


@pytest.mark.parametrize("name", helpers.dist_names())
def test_distances_match_reference(name):
    rows, nodes, dist = helpers.load_dist(name)
    onodes, odist = oracle.distances(rows)
    np.testing.assert_array_equal(onodes, nodes)
    np.testing.assert_array_equal(odist, dist)
    d = helpers.digests()["dist"][name]
    sha = hashlib.sha256(np.where(odist < 0, np.iinfo(np.int32).max, odist).astype("<i4").tobytes()).hexdigest()
    assert sha == d["sha256"] and odist.max() == d["max"]


def test_gae_matches_reference_bit_exact():
    z = np.load(helpers.GOLDEN + "/gae.npz")
    for ci in range(int(z["n_cases"])):
        adv, ret = oracle.gae(z[f"c{ci}_rew"], z[f"c{ci}_val"], z[f"c{ci}_done"], z[f"c{ci}_last"],
                              float(z["gamma"]), float(z["lam"]))
        np.testing.assert_array_equal(adv.view(np.uint32), z[f"c{ci}_adv"].view(np.uint32), err_msg=f"case {ci} adv")
        np.testing.assert_array_equal(ret.view(np.uint32), z[f"c{ci}_ret"].view(np.uint32), err_msg=f"case {ci} ret")
    adv, ret = oracle.gae(z["kat_rew"][:, None], z["kat_val"][:, None], z["kat_done"][:, None], [0.37], 0.99, 0.95)
    np.testing.assert_array_equal(adv[:, 0].view(np.uint32), z["kat_adv"].view(np.uint32))
    np.testing.assert_array_equal(ret[:, 0].view(np.uint32), z["kat_ret"].view(np.uint32))
    # SURVEY A.4 printed values
    np.testing.assert_allclose(adv[:4, 0], [0.2268139124, 0.3359510303, -0.8172292113, -1.2273871899], rtol=1e-6)


def test_adv_norm_matches_reference():
    z = np.load(helpers.GOLDEN + "/gae.npz")
    out = oracle.adv_norm(z["norm_adv"], z["norm_inds"], int(z["norm_mb"]))
    np.testing.assert_allclose(out, z["norm_out"], rtol=1e-5, atol=1e-6)


class OracleBotDriver(OracleDriver):
    """Replays a fixture recorded with the reference's baselineTeam on the red side: the oracle's own bots
    must pick the recorded actions (ties -> first best, as pinned while recording) from the recorded best sets."""

    def __init__(self, fx):
        super().__init__(fx)
        self.fx, self.t = fx, 0
        self.env.set_bots(fx["bot_kind"])

    def step(self, a):
        kind = self.fx["bot_kind"]
        fed = np.where(kind >= 1, 7, np.asarray(a, np.int32))          # bot columns get garbage: must be ignored
        if int(self.fx.get("real_rng", 0)):
            self.env.set_draws(self.fx["draws"][self.t][None])           # the reference's own random.choice indices
        o = self.env.step_bots(fed)
        np.testing.assert_array_equal(o["actions_used"][0], a, err_msg=f"step {self.t}: bot actions")
        if (kind >= 2).any():
            np.testing.assert_array_equal(o["best_masks"][0][kind >= 2], self.fx["bot_best"][self.t][kind >= 2],
                                          err_msg=f"step {self.t}: best sets")
        self.t += 1
        return {k: (v[0] if v is not None else None) for k, v in o.items()}


@pytest.mark.parametrize("name", [n for n in helpers.traj_names() if n.startswith("bots_")])
def test_baseline_team_bots_match_reference(name):
    """agents/baselineTeam.py (Offensive + Defensive reflex agents on the red side, gymPacMan.py:157)."""
    fx = helpers.load_traj(name)
    assert list(fx["bot_kind"]) == [2, 0, 3, 0]
    helpers.replay(fx, OracleBotDriver(fx), where=name)


@pytest.mark.parametrize("name", [n for n in helpers.traj_names() if n.startswith("rng_")])
def test_bots_with_the_references_real_rng_stream(name):
    """BASELINE config 1 as the reference plays it: the bots' `random.choice` draws from the global MT19937
    (`random.seed(1)`; baselineTeam.py:92, randomTeam.py:100).  The fixture holds the index every call returned;
    fed those indices, the oracle's bots reproduce the reference's trajectory bit for bit."""
    fx = helpers.load_traj(name)
    assert int(fx["real_rng"]) == 1 and list(fx["bot_kind"]) in ([2, 0, 3, 0], [1, 0, 1, 0])
    helpers.replay(fx, OracleBotDriver(fx), where=name)
