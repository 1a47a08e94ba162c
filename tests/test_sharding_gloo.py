"""N>1 host logic on CPU: world_size-2 gloo processes exercise the env partition,
the flat-bucket gradient all-reduce and the episode-stat reduction that the B200
run does over NCCL (pacman_marl_2025_b200/sharding.py).  The env kernels need no
collective, so what is checked for them is the partition itself and — through the
kernel-logic emulation — that two shards reproduce one batch (shard invariance)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_envs, q):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pacman_marl_2025_b200 import sharding
        torch.manual_seed(0)                                     # same init on every rank
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        if rank == 1:                                            # rank 1 starts perturbed: broadcast must fix it
            with torch.no_grad():
                for p in net.parameters():
                    p.add_(1.0)
        sharding.broadcast_module_(net, src=0)
        bucket = sharding.FlatGradBucket(net.parameters())
        g = torch.Generator().manual_seed(123)
        X, Y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
        half = slice(rank * 4, rank * 4 + 4)                     # each rank: its half of the global minibatch
        bucket.zero_()
        torch.nn.functional.mse_loss(net(X[half]), Y[half]).backward()
        local = bucket.flat.clone()
        bucket.all_reduce_mean_()
        # single-process truth: mean loss over the whole batch == mean of the two half-batch losses
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        ref.load_state_dict(net.state_dict())
        torch.nn.functional.mse_loss(ref(X), Y).backward()
        want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])          # views into the (padded) bucket
        ok_grad = torch.allclose(got, want, atol=1e-6) and bucket.flat.numel() % 64 == 0 and \
            abs(float(bucket.flat.sum()) - float(got.sum())) < 1e-5                # the padding holds zeros
        views_alias = all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
        norm_before = float(torch.linalg.vector_norm(bucket.flat))
        total = float(bucket.clip_grad_norm_(0.01))
        norm_after = float(torch.linalg.vector_norm(bucket.flat))
        # env partition + shard-invariant stepping through the kernel-logic emulation
        begin, count = sharding.shard_range(total_envs, rank, world)
        stats = torch.tensor([float(count), float(begin), 1.0], dtype=torch.float64)
        sharding.all_reduce_stats_(stats)
        from emu_driver import EmuEnv
        from parity_bodies import biased_actions, lay_rows
        env = EmuEnv(lay_rows("tinyCapture"), count, length=20, auto_reset=True, random_agent_mask=0b0101,
                     random_seed=5, env_index_base=begin, E=16, KB=1)
        env.reset()
        rng = np.random.default_rng(8)
        digest = []
        for t in range(30):
            a = biased_actions(rng, total_envs, 0.6, t)[begin:begin + count]
            o = env.step(a)
            digest.append((o["obs"].sum(axis=(0, 2, 3, 4)), o["reward"].copy(), o["done"].copy()))
        q.put((rank, ok_grad, views_alias, norm_before, total, norm_after, (begin, count), stats.tolist(),
               not torch.equal(local, bucket.flat), digest))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_bucket_stats_and_shard_invariance():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    total_envs = 80
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total_envs, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, ok_grad, alias, nb, total, na, rng_, stats, changed, _ in res:
        assert ok_grad and alias and changed
        assert abs(total - nb) < 1e-6 and na <= 0.01 + 1e-6
        assert stats == [float(total_envs), float(res[0][6][0] + res[1][6][0]), 2.0]
    (b0, c0), (b1, c1) = res[0][6], res[1][6]
    assert b0 == 0 and b1 == c0 and c0 + c1 == total_envs and b1 % 32 == 0
    # one process stepping the whole batch reproduces the concatenation of the two shards
    sys.path.insert(0, HERE)
    from emu_driver import EmuEnv
    from parity_bodies import biased_actions, lay_rows
    env = EmuEnv(lay_rows("tinyCapture"), total_envs, length=20, auto_reset=True, random_agent_mask=0b0101,
                 random_seed=5, E=16, KB=1)
    env.reset()
    rng = np.random.default_rng(8)
    for t in range(30):
        o = env.step(biased_actions(rng, total_envs, 0.6, t))
        s0, s1 = res[0][9][t], res[1][9][t]
        np.testing.assert_array_equal(o["obs"].sum(axis=(0, 2, 3, 4)), np.concatenate([s0[0], s1[0]]))
        np.testing.assert_array_equal(o["reward"], np.concatenate([s0[1], s1[1]], axis=1))
        np.testing.assert_array_equal(o["done"], np.concatenate([s0[2], s1[2]]))


@pytest.mark.parametrize("total,world", [(65536, 8), (100, 3), (31, 4), (0, 2), (4096, 1), (65, 2)])
def test_shard_ranges_tile_the_batch(total, world):
    from pacman_marl_2025_b200.sharding import shard_range
    pos = 0
    for r in range(world):
        b, c = shard_range(total, r, world)
        assert b == pos and c >= 0 and (b % 32 == 0 or c == 0)
        pos += c
    assert pos == total
