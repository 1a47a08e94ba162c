"""Dev tool: fwd+bwd time of the actor (conv backbone + head) on a PPO minibatch in NCHW (with / without the fused
GroupNorm kernels) and in channels-last with PyTorch's GroupNorm - does the layout remove cuDNN's transposes?"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacman_marl_2025_b200 import pacman_mappo_resnet as M
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
H, W, B = 11, 14, 8192
x0 = (torch.rand(B, 8, H, W, device="cuda") < 0.2).float()
for name, fused, cl in (("NCHW fused kernels", True, False), ("NCHW torch GroupNorm", False, False), ("channels_last torch GroupNorm", False, True)):
    torch.manual_seed(0)
    agent = M.MAPPOAgent((8, H, W), 5, 2).cuda()
    M.ResidualBlock.fused_norm = fused
    x = x0
    if cl:
        agent.actor_backbone = agent.actor_backbone.to(memory_format=torch.channels_last)
        x = x0.contiguous(memory_format=torch.channels_last)
    def it():
        agent.zero_grad(set_to_none=True)
        agent.logits(x).square().mean().backward()
    for _ in range(5):
        it()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        it()
    torch.cuda.synchronize()
    print(f"{name:32s} {(time.perf_counter() - t0) * 50:.2f} ms per fwd+bwd", flush=True)
M.ResidualBlock.fused_norm = True
