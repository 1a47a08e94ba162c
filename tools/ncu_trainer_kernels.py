"""Dev tool (run under ncu): one PPO minibatch step of the trainer's networks on a smallCapture minibatch, so that the
library's own kernels around the PyTorch layers (attention, LayerNorm, GroupNorm, gather, loss, Adam) can be captured:

    ncu --set full --clock-control none -k regex:pm_.*bwd -c 24 -f -o prof python tools/ncu_trainer_kernels.py

(the forward kernels also run in the warm-up update and the rollout, so select the launches of the last update with
-s, or the backward kernels by name as above; a full-set capture costs ~8 s of GPU time per launch).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pacman_marl_2025_b200 import pacman_mappo_resnet as M

torch.backends.cudnn.benchmark = True
torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
cfg = M.TrainConfig(num_envs=1024, num_steps=32, checkpoint_every=0, update_epochs=1, num_minibatches=8, graph_rollout=False)
tr = M.MAPPOTrainer(cfg, device="cuda:0")
tr.train_one(0)          # warm-up: autotune, allocator
torch.cuda.synchronize()
sched = tr.schedule(1)
lv, _ = tr.collect(sched["mode"], sched["play_as_red"])
tr.cfg.num_minibatches = 8
tr.update(lv, sched)     # profiled: 8 minibatch steps
torch.cuda.synchronize()
print("done")
