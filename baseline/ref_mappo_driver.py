"""Runs the UNMODIFIED reference trainer (baseline/_ref/pacman_mappo_resnet.py: `train()`, `:340-703`) for a bounded sample and
prints its samples/s.  Only module-level constants are changed (NUM_STEPS, TOTAL_UPDATES, EVAL_FREQ - the rollout length,
the number of updates and "never evaluate"), matplotlib (absent from this image) is stubbed, and the working directory is
a scratch directory that links the reference's layouts/ and agents/ (train() opens 'layouts/bloxCapture.lay' and writes
checkpoints relative to the cwd).  Used by bench.py (`also.mappo.reference`); nothing of this repo is on its path.

    python ref_mappo_driver.py <reference root> <NUM_STEPS> <TOTAL_UPDATES>      (cwd = scratch dir)
"""
import os, sys, time, json
from unittest.mock import MagicMock
sys.dont_write_bytecode = True
sys.modules["matplotlib"] = MagicMock(); sys.modules["matplotlib.pyplot"] = MagicMock()
ref, steps, updates = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sys.path.insert(0, ref)
import torch
torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
import pacman_mappo_resnet as R
R.NUM_STEPS, R.TOTAL_UPDATES, R.EVAL_FREQ = steps, updates, 10 ** 9
R.plt.subplots = lambda *a, **k: (MagicMock(), MagicMock())
t0 = time.perf_counter()
R.train()
dt = time.perf_counter() - t0
print(json.dumps({"seconds": dt, "updates": updates + 1, "num_steps": steps, "samples": 2 * steps * (updates + 1), "device": str(R.device)}))
