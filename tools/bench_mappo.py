#!/usr/bin/env python
"""MAPPO SPS (BASELINE config 5): learner-agent samples per second of the full rollout + GAE + PPO update
loop on smallCapture.  `python tools/bench_mappo.py [--envs N] [--steps T] [--updates U] [--amp 1]`;
under torchrun it runs one rank per GPU with the NCCL gradient all-reduce.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--updates", type=int, default=3)
    ap.add_argument("--amp", type=int, default=0)
    ap.add_argument("--layout", default="smallCapture")
    ap.add_argument("--pair", type=int, default=1)
    ap.add_argument("--views", type=int, default=1)
    ap.add_argument("--fused", type=int, default=1, help="row-f3 kernels (gather, fused loss, Adam+EMA) vs eager ops")
    ap.add_argument("--attn", default="bmm", help="critic attention: auto, kernel (pm_attn), bmm (written out) or module (nn.TransformerEncoder)")
    ap.add_argument("--graph", type=int, default=1, help="rollout steps replayed from CUDA graphs")
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    lay = os.path.join(os.path.dirname(M.__file__), "layouts", a.layout + ".lay")
    cfg = M.TrainConfig(layout_file=lay, num_envs=a.envs, num_steps=a.steps, amp_bf16=bool(a.amp), checkpoint_every=0,
                        pair_agents=bool(a.pair), use_views=bool(a.views), fused_pipeline=bool(a.fused), graph_rollout=bool(a.graph))
    M.MAPPOAgent.attn_impl = a.attn
    tr = M.MAPPOTrainer(cfg, device=f"cuda:{local}")
    tr.train_one(0)                                        # warm-up (cuDNN autotune, allocator)
    tr.train_one(0)                                        # second use of the opponent: the rollout steps are captured here
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    parts = dict(rollout=0.0, update=0.0)
    for u in range(1, a.updates + 1):
        sched = tr.schedule(u)
        ta = time.perf_counter()
        lv, stats = tr.collect(sched["mode"], sched["play_as_red"])
        torch.cuda.synchronize()
        tb = time.perf_counter()
        tr.update(lv, sched)
        torch.cuda.synchronize()
        parts["rollout"] += tb - ta
        parts["update"] += time.perf_counter() - tb
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    samples = 2 * a.envs * a.steps * a.updates * world
    if rank == 0:
        print(json.dumps({"metric": "MAPPO SPS (learner-agent samples/s, rollout + GAE + PPO update)", "value": samples / dt,
                          "unit": "samples/s", "n_gpus": world, "config": {"workload": f"{a.layout}, {a.envs} envs/GPU x {a.steps} steps, "
                          f"{cfg.update_epochs} epochs x {cfg.num_minibatches} minibatches, randomTeam opponent", "amp_bf16": bool(a.amp),
                          "pair_agents": bool(a.pair), "kernel_views": bool(a.views), "fused_pipeline": bool(a.fused),
                          "critic_attention": a.attn, "graph_rollout": bool(a.graph)},
                          "rollout_s": parts["rollout"], "update_s": parts["update"],
                          "env_agent_steps_per_s_in_rollout": 4 * a.envs * a.steps * a.updates * world / parts["rollout"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
