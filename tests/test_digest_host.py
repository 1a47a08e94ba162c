"""The observation digest used by the full-size GPU parity tests (tests/test_gpu_full_size.py): the oracle's C
digest, a numpy restatement and the torch integer-op version the GPU side runs agree, and the digest separates
observations that differ by moves a position-linear checksum would cancel."""
import numpy as np
import torch

from oracle import oracle
from parity_bodies import biased_actions, lay_rows


def torch_digest(obs4, Kt):
    out = []
    for k in range(4):
        bits = obs4[k].reshape(obs4.shape[1], -1).view(torch.int32).to(torch.int64)
        out.append((bits * Kt).sum(dim=1))
    return torch.stack(out, dim=1)


def test_digest_three_ways():
    rows = lay_rows("smallCapture")
    N = 64
    a_env, b_env = oracle.OracleBatch(rows, N, length=40), oracle.OracleBatch(rows, N, length=40)
    a_env.reset(); b_env.reset()
    K = oracle.digest_weights(8 * 11 * 14)
    Kt = torch.from_numpy(K.view(np.int64))
    rng = np.random.default_rng(0)
    for t in range(60):
        act = biased_actions(rng, N, 0.6, t)
        d = a_env.step_digest(act, K, auto_reset=True)
        o = b_env.step(act, with_obs=True, auto_reset=True)
        obs = o["obs"]
        done = np.nonzero(o["done"])[0]
        if len(done):
            obs[done], _ = b_env.reset_outputs_for(done)
        want = (obs.reshape(N, 4, -1).view(np.uint32).astype(np.uint64) * K).sum(-1)
        np.testing.assert_array_equal(d["digest"], want)
        got = torch_digest(torch.from_numpy(obs.transpose(1, 0, 2, 3, 4).copy()), Kt).numpy().view(np.uint64)
        np.testing.assert_array_equal(got, want)
        np.testing.assert_array_equal(d["reward"], o["reward"])


def test_digest_weights_do_not_cancel_opposite_moves():
    K = oracle.digest_weights(1000)
    a = np.zeros(1000, np.float32); b = np.zeros(1000, np.float32)
    a[[100, 501]] = 1.0            # one marker moves +1, another moves -1: a position-linear checksum sees no change
    b[[101, 500]] = 1.0
    dig = lambda o: int((o.view(np.uint32).astype(np.uint64) * K).sum())
    assert dig(a) != dig(b)
