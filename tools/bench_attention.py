"""Dev tool: fwd+bwd time of the critic's two encoder layers on a PPO minibatch, nn.TransformerEncoder.forward vs the
written-out bmm formulation (MAPPOAgent.attn_impl), fp32+TF32 and bf16 autocast."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacman_marl_2025_b200 import pacman_mappo_resnet as M
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
for (H, W), B in (((11, 14), 4096), ((7, 20), 4096), ((20, 20), 1024)):
    agent = M.MAPPOAgent((8, H, W), 5, 2).cuda()
    x = (torch.rand(B, 8, H, W, device="cuda") < 0.2).float()
    for amp in (False, True):
        for impl in ("module", "bmm", "kernel"):
            agent.attn_impl = impl
            def it():
                agent.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    v = agent.get_value(x)
                v.float().square().mean().backward()
            for _ in range(3):
                it()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10):
                it()
            torch.cuda.synchronize()
            print(f"{H}x{W} B={B} {'bf16' if amp else 'fp32/tf32'} {impl:6s} {(time.perf_counter() - t0) * 100:.2f} ms per fwd+bwd  "
                  f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
