"""Dev tool: pm_distances alone (tables packed and uploaded once): one board of each shipped layout, and batches of
mazeGenerator boards."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pacman_marl_2025_b200 import _lib, mazeGenerator
from pacman_marl_2025_b200.gymPacMan import pack_tables
from pacman_marl_2025_b200.layout import Layout, getLayout
L = _lib.load()
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pacman_marl_2025_b200", "layouts")


def run(name, lays, reps):
    W, H = lays[0].width, lays[0].height
    _, table, _ = pack_tables(lays)
    nmax = (max(int((~l.walls.array).sum()) if hasattr(l.walls, "array") else sum(1 for x in range(W) for y in range(H) if not l.walls[x][y]) for l in lays) + 7) & ~7
    tab = torch.from_numpy(table.view(np.int32)).cuda()
    nl = len(lays)
    n_open = torch.zeros(nl, dtype=torch.int32, device="cuda")
    node_of = torch.empty((nl, W * H), dtype=torch.int16, device="cuda")
    dist = torch.empty((nl, nmax, nmax), dtype=torch.int16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: L.pm_distances(W, H, tab.data_ptr(), nl, n_open.data_ptr(), node_of.data_ptr(), dist.data_ptr(), nmax, st)
    for _ in range(3):
        assert call() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name}: {us:.1f} us per launch, {us / nl:.2f} us per board, table {dist.numel() * 2 / 1e6:.1f} MB ({dist.numel() * 2 / us / 1e3:.0f} GB/s written)")


for n in ("tinyCapture", "smallCapture", "bloxCapture"):
    run(n, [getLayout(os.path.join(root, n + ".lay"))], 50)
for nl in (256, 1024):
    run(f"{nl} random mazes", [Layout(mazeGenerator.generateMaze(7000 + s).split("\n")) for s in range(nl)], 10)
