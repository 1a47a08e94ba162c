"""Dev / profiling tool: the step-kernel variants the trainer and the opponents really launch, each timed alone with
CUDA events (and, under ncu, captured by name).  Variants of pm_env_kernel<MODE=STEP, BOTS, VIEWS, E>:

  plain         <0,0,0,32>  self-play, fp32 observations               (the headline kernel)
  views         <0,0,1,32>  randomTeam in-kernel, learner views only (raw_obs off): what the trainer launches vs randomTeam
  views+obs     <0,0,1,32>  self-play, fp32 observations AND views:  what the trainer launches in self-play
  bots          <0,1,0,32>  baselineTeam in-kernel, fp32 observations
  bots+views    <0,1,1,32>  baselineTeam in-kernel, learner views only: what the trainer launches vs baselineTeam
  *-1024        the ,8> instantiations at a trainer-sized batch of 1 024 envs

    python tools/run_variants.py [--layout tinyCapture] [--envs 65536] [--reps 50] [--only plain,views]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacman_marl_2025_b200 import _lib  # noqa: E402
from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layout", default="tinyCapture")
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--warm", type=int, default=5)
ap.add_argument("--only", default="")
ap.add_argument("--small", type=int, default=1024)
ap.add_argument("--chain", type=int, default=0, help="> 0: also time the same steps as pm_step_chain runs of this length")
a = ap.parse_args()
lay = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pacman_marl_2025_b200", "layouts", a.layout + ".lay")
assert os.path.exists(lay), lay
VARIANTS = {
    "plain": dict(self_play=True, views=None, raw=True),
    "views": dict(self_play=False, team="randomTeam", views=False, raw=False),
    "views+obs": dict(self_play=True, views=False, raw=True),
    "bots": dict(self_play=False, team="baselineTeam", views=None, raw=True),
    "bots+views": dict(self_play=False, team="baselineTeam", views=False, raw=False),
    "plain-elide": dict(self_play=True, views=None, raw=True, elide=True),     # pm_step_out.obs_static_valid
}
def clocks():
    import subprocess
    try:
        return subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active",
                                        "--format=csv,noheader,nounits", "-i", "0"], text=True).strip()
    except Exception:                                   # noqa: BLE001
        return "n/a"


only = [v for v in a.only.split(",") if v]
print("lib", _lib.LIB_PATH, "| clocks before (sm MHz, mem MHz, W, reasons):", clocks(), flush=True)
out = {}
for N in (a.envs, a.small):
    if N <= 0:
        continue
    acts = torch.randint(0, 5, (8, N, 4), dtype=torch.uint8).cuda()
    for name, v in VARIANTS.items():
        if only and name not in only:
            continue
        env = gymPacMan_parallel_env(layout_file=lay, self_play=v["self_play"], enemieName=v.get("team", "randomTeam"),
                                     length=300, num_envs=N, auto_reset=True, elide_static_planes=v.get("elide", False))
        env.reset_batched()
        if v["views"] is not None or not v["raw"]:
            env.set_view_outputs(v["views"], raw_obs=v["raw"])
        for i in range(a.warm):
            env.step_batched(acts[i % 8])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(a.reps):
            env.step_batched(acts[i % 8])
        import time as _t
        t0 = _t.perf_counter()
        e1.record(); torch.cuda.synchronize()
        drain = (_t.perf_counter() - t0) * 1e6 / a.reps      # ~0 when the host, not the GPU, was the bottleneck
        us = e0.elapsed_time(e1) * 1000 / a.reps
        H, W = env.obs_shape[1:]
        b = (4 * 32 * H * W if v["raw"] else 0) + (3 * 8 * H * W if v["views"] is not None else 0) + 200
        out[f"{name}-{N}"] = {"us": us, "agent_steps_per_s": 4.0 * N / us * 1e6, "bytes_per_env_step": b,
                              "GBps": b * N / us / 1e3}
        print(f"{a.layout} {name:11s} N={N:6d}  {us:8.1f} us/launch  {4.0 * N / us:8.1f} M agent-steps/s  {b * N / us / 1e3:7.0f} GB/s  (drain {drain:.0f} us/launch, clocks {clocks()})",
              flush=True)
        if a.chain > 0:
            T = a.chain
            ca = torch.randint(0, 5, (T, N, 4), dtype=torch.uint8).cuda()
            env.step_chain(ca)
            torch.cuda.synchronize(); e0.record()
            reps = max(1, a.reps // T)
            for i in range(reps):
                env.step_chain(ca)
            e1.record(); torch.cuda.synchronize()
            usc = e0.elapsed_time(e1) * 1000 / (reps * T)
            out[f"{name}-{N}"]["chain_us"] = usc
            print(f"{a.layout} {name:11s} N={N:6d}  {usc:8.1f} us/step in chains of {T}  {4.0 * N / usc:8.1f} M agent-steps/s  {b * N / usc / 1e3:7.0f} GB/s",
                  flush=True)
        del env
print(json.dumps(out))
