"""Dev tool: time pm_step with and without the observation output (rules-only)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pacman_marl_2025_b200 import _lib
from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
lay = sys.argv[1] if len(sys.argv) > 1 else "tinyCapture"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
env = gymPacMan_parallel_env(layout_file=os.path.join(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pacman_marl_2025_b200"), "layouts", lay + ".lay"),
                             self_play=True, length=300, num_envs=N)
acts = torch.randint(0, 5, (8, N, 4), dtype=torch.uint8).cuda()
L = _lib.load()
so = env._step_out
noobs = _lib.PmStepOut(None, 0, so.reward, so.done, so.legal, so.score_change, so.shaping, so.shaped_reward)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, out in (("with obs", so), ("rules only", noobs)):
    for i in range(5):
        L.pm_step(C.byref(env._desc), acts[i % 8].data_ptr(), 1, C.byref(out), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(50):
        L.pm_step(C.byref(env._desc), acts[i % 8].data_ptr(), 1, C.byref(out), st)
    e1.record(); torch.cuda.synchronize()
    print(lay, N, name, "%.1f us/launch" % (e0.elapsed_time(e1) * 1000 / 50))
