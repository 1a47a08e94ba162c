"""Dev tool: where does the PPO update / rollout time go?  torch.profiler table of one update."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from pacman_marl_2025_b200 import pacman_mappo_resnet as M

amp = int(sys.argv[1]) if len(sys.argv) > 1 else 0
opts = sys.argv[2] if len(sys.argv) > 2 else ""
if "bench" in opts: torch.backends.cudnn.benchmark = True
if "tf32" in opts: torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
cfg = M.TrainConfig(num_envs=1024, num_steps=32, amp_bf16=bool(amp), checkpoint_every=0)
tr = M.MAPPOTrainer(cfg, device="cuda:0")
if "cl" in opts:
    tr.agent = tr.agent.to(memory_format=torch.channels_last)
tr.train_one(0)
sched = tr.schedule(1)
torch.cuda.synchronize(); t0 = time.perf_counter()
lv, _ = tr.collect(sched["mode"], sched["play_as_red"])
torch.cuda.synchronize(); t1 = time.perf_counter()
tr.update(lv, sched)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"amp={amp} opts={opts}: rollout {t1 - t0:.3f}s update {t2 - t1:.3f}s")
if "prof" in opts:
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        tr.update(lv, sched)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        tr.collect(sched["mode"], sched["play_as_red"])
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
