"""Dev tool: time pm_observe (encoder only: four observations per env from the current state)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacman_marl_2025_b200 import _lib
from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
for lay in ("tinyCapture", "smallCapture", "bloxCapture"):
    N = 65536
    env = gymPacMan_parallel_env(layout_file=os.path.join(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pacman_marl_2025_b200"), "layouts", lay + ".lay"),
                                 self_play=True, length=300, num_envs=N, auto_reset=True)
    acts = torch.randint(0, 5, (8, N, 4), dtype=torch.uint8).cuda()
    for i in range(30):
        env.step_batched(acts[i % 8])
    out = torch.empty_like(env._obs)
    for i in range(5):
        env.observe_batched(out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(50):
        env.observe_batched(out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 50
    print(f"{lay} observe-only {us:.1f} us/launch  {out.numel() * 4 / us / 1e6:.0f} GB/s")
    del env, out
    torch.cuda.empty_cache()
