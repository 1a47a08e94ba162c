"""Shared fixture plumbing for the parity tests (oracle and CUDA path alike)."""
import glob
import hashlib
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def traj_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "traj_*.npz")))


def dist_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "dist_*.npz")))


def load_traj(name):
    z = np.load(os.path.join(GOLDEN, f"traj_{name}.npz"))
    fx = {k: z[k] for k in z.files}
    fx["layout"] = [str(r) for r in fx["layout"]]
    fx["W"], fx["H"] = int(fx["WH"][0]), int(fx["WH"][1])
    return fx


def load_dist(name):
    z = np.load(os.path.join(GOLDEN, f"dist_{name}.npz"))
    return [str(r) for r in z["layout"]], z["nodes"], z["dist"]


def digests():
    return json.load(open(os.path.join(GOLDEN, "digests.json")))


def obs_digest(a):
    b = np.ascontiguousarray(a, dtype="<f4").tobytes()
    return np.frombuffer(hashlib.blake2b(b, digest_size=8).digest(), dtype="<u8")[0]


def unpack_cells(bits, HW):
    return np.unpackbits(np.asarray(bits, np.uint8), bitorder="little")[:HW]


def check_step(fx, t, out, state, where=""):
    """Compare one recorded step with a driver's outputs.

    out:   dict(obs[4,8,H,W] f32, reward[2] f64, done, legal[4], score_change, shaping[4] f64)
    state: dict(agents[4,7] (x,y,dir,isPacman,scared,carrying,returned), food[HW], caps[HW], score)
    """
    HW = fx["W"] * fx["H"]
    tag = f"{where} step {t}"
    ag = state["agents"]
    np.testing.assert_array_equal(ag[:, 0:2], fx["pos"][t], err_msg=f"{tag}: positions")
    np.testing.assert_array_equal(ag[:, 2:7], fx["misc"][t], err_msg=f"{tag}: dir/isPacman/scared/carrying/returned")
    np.testing.assert_array_equal(state["food"], unpack_cells(fx["food"][t], HW), err_msg=f"{tag}: food")
    np.testing.assert_array_equal(state["caps"], unpack_cells(fx["caps"][t], HW), err_msg=f"{tag}: capsules")
    assert state["score"] == int(fx["score"][t]), f"{tag}: score"
    assert int(out["score_change"]) == int(fx["score_change"][t]), f"{tag}: score_change"
    assert int(out["done"]) == int(fx["done"][t]), f"{tag}: done"
    np.testing.assert_array_equal(np.asarray(out["legal"], np.uint8), fx["legal"][t], err_msg=f"{tag}: legal masks")
    # bit-exact fp64 rewards and shaping
    np.testing.assert_array_equal(np.asarray(out["reward"], np.float64).view(np.uint64),
                                  fx["reward"][t].view(np.uint64), err_msg=f"{tag}: rewards")
    np.testing.assert_array_equal(np.asarray(out["shaping"], np.float64).view(np.uint64),
                                  fx["shaping"][t].view(np.uint64), err_msg=f"{tag}: shaping")
    if t < len(fx["full_obs"]):
        np.testing.assert_array_equal(out["obs"], fx["full_obs"][t], err_msg=f"{tag}: full obs")
    for i in range(4):
        assert obs_digest(out["obs"][i]) == fx["obs_dig"][t, i], f"{tag}: obs digest agent {i}"


def replay(fx, drv, where=""):
    """drv: object with reset() -> (obs[4,8,H,W], legal[4]); step(a[4]) -> out; state() -> dict."""
    ep_starts = list(fx["ep_start"]) + [len(fx["actions"])]
    for ep in range(len(fx["ep_start"])):
        obs, legal = drv.reset()
        np.testing.assert_array_equal(np.asarray(legal, np.uint8), fx["reset_legal"][ep], err_msg=f"{where} ep {ep}: reset legal")
        for i in range(4):
            assert obs_digest(obs[i]) == fx["reset_obs_dig"][ep, i], f"{where} ep {ep}: reset obs agent {i}"
        for t in range(ep_starts[ep], ep_starts[ep + 1]):
            out = drv.step(fx["actions"][t])
            check_step(fx, t, out, drv.state(), where)


class BotReplayDriver:
    """Wraps a single-env driver for fixtures recorded with the reference's bots on the red side: the env under
    test runs its OWN bots; the recorded red actions are withheld (garbage is fed in their place) and must come
    back in `actions_used`.  `bots_*` fixtures were recorded with `random.choice` pinned to the first best action
    (replay: tie-break = first); `rng_*` fixtures with the bots' REAL `random.choice` on the global MT19937 - the
    index each call returned is in `draws` and is fed back through `inner.set_draws` (pm_env_desc.bot_draws)."""

    def __init__(self, inner, fx):
        self.inner, self.fx, self.t = inner, fx, 0
        self.real_rng = bool(fx.get("real_rng", 0))

    def reset(self):
        return self.inner.reset()

    def step(self, a):
        bots = self.fx["bot_kind"] >= 1             # (non-bot columns report the EFFECTIVE action: illegal -> Stop)
        fed = np.where(bots, 9, np.asarray(a)).astype(np.uint8)
        if self.real_rng:
            self.inner.set_draws(self.fx["draws"][self.t])
        out = self.inner.step(fed)
        np.testing.assert_array_equal(np.asarray(out["actions_used"])[bots], np.asarray(a)[bots],
                                      err_msg=f"step {self.t}: bot actions")
        self.t += 1
        return out

    def state(self):
        return self.inner.state()
