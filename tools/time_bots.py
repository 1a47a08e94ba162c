"""Dev tool: time the step kernel with baselineTeam playing red (with / without the observation output)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacman_marl_2025_b200 import _lib
from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
N = 65536
for lay in ("tinyCapture", "bloxCapture"):
    env = gymPacMan_parallel_env(layout_file=os.path.join(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pacman_marl_2025_b200"), "layouts", lay + ".lay"),
                                 self_play=False, enemieName="baselineTeam", length=300, num_envs=N, auto_reset=True)
    acts = torch.randint(0, 5, (8, N, 4), dtype=torch.uint8).cuda()
    for raw in (True, False):
        env.set_view_outputs(None, raw_obs=raw)
        for i in range(10):
            env.step_batched(acts[i % 8])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for i in range(50):
            env.step_batched(acts[i % 8])
        e1.record(); torch.cuda.synchronize()
        print(f"{lay} baselineTeam {'with obs' if raw else 'rules only'} {e0.elapsed_time(e1) * 1000 / 50:.1f} us/launch")
    del env
