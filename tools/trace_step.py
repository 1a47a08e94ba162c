"""Dev tool: per-CTA timeline of one step launch (needs a library built with -DPM_TRACE).
Prints, for a few SMs, when each resident CTA started, finished its rules phase and finished its stores."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pacman_marl_2025_b200 import _lib
from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
lay = sys.argv[1] if len(sys.argv) > 1 else "tinyCapture"
N = 65536
env = gymPacMan_parallel_env(layout_file=os.path.join(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pacman_marl_2025_b200"), "layouts", lay + ".lay"),
                             self_play=True, length=300, num_envs=N, auto_reset=True)
acts = torch.randint(0, 5, (8, N, 4), dtype=torch.uint8).cuda()
L = _lib.load()
grid = N // 32
buf = torch.zeros((grid, 4), dtype=torch.int64, device="cuda")
for i in range(20):
    env.step_batched(acts[i % 8])
torch.cuda.synchronize()
L.pm_trace_set.argtypes = [C.c_void_p]
assert L.pm_trace_set(C.c_void_p(buf.data_ptr())) == 0
env.step_batched(acts[0])
torch.cuda.synchronize()
t = buf.cpu().numpy()
t0 = t[:, 1].min()
sm, st, ra, en = t[:, 0], t[:, 1] - t0, t[:, 2] - t0, t[:, 3] - t0
print(f"{lay}: kernel span {en.max() / 1e3:.1f} us; rules phase mean {np.mean(ra - st) / 1e3:.2f} us (p10 {np.percentile(ra - st, 10) / 1e3:.2f}, p90 {np.percentile(ra - st, 90) / 1e3:.2f}); "
      f"store phase mean {np.mean(en - ra) / 1e3:.1f} us")
first = np.sort(st)[:148 * 8]
print(f"first-wave start spread: {first.max() / 1e3:.2f} us; CTAs started after 5 us: {(st > 5000).sum()}")
# store-free time: union of [rules_done, end] intervals per SM
for s in (0, 1, 73, 147):
    idx = np.nonzero(sm == s)[0]
    order = idx[np.argsort(st[idx])]
    print(f"SM {s}: {len(idx)} CTAs")
    for b in order:
        print(f"   cta {b:5d}  start {st[b] / 1e3:7.1f}  rules_done {ra[b] / 1e3:7.1f}  end {en[b] / 1e3:7.1f}")
busy = 0.0
for s in np.unique(sm):
    idx = np.nonzero(sm == s)[0]
    iv = sorted((ra[b], en[b]) for b in idx)
    cur_s, cur_e, tot = iv[0][0], iv[0][1], 0
    for a, b in iv[1:]:
        if a > cur_e:
            tot += cur_e - cur_s; cur_s, cur_e = a, b
        else:
            cur_e = max(cur_e, b)
    tot += cur_e - cur_s
    busy += tot
print(f"mean time per SM with at least one CTA in its store phase: {busy / len(np.unique(sm)) / 1e3:.1f} us of {en.max() / 1e3:.1f}")
# how many CTAs are in the store phase at a given time (whole GPU)
for q in (5, 10, 15, 20, 50, 100, 110, 120, 130, 150, 190, 200):
    tt = q * 1000
    print(f"  t={q:4d} us: {int(((ra <= tt) & (en > tt)).sum()):5d} CTAs storing, {int(((st <= tt) & (ra > tt)).sum()):5d} in rules phase")
