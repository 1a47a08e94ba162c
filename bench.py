#!/usr/bin/env python
"""bench.py — agent-steps/sec of the batched capture env (step + obs) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--layout tinyCapture] [--envs 65536]

One "step" = one pass of the hot path over the whole batch: ONE launch of the step
kernel (rules of 4 ordered sub-moves + four fp32 [8,H,W] observations per env, rewards,
done, legal masks, shaping) with pm_env_desc.auto_reset, i.e. envs that finish
are re-initialised inside the same launch, so episodes keep cycling.  The timed
loop issues the K launches through pm_step_chain (the actions of this workload do
not depend on the observations), so consecutive launches overlap per env tile;
the same K steps as separate pm_step launches are reported beside it.  Default workload = the configuration BASELINE.json's
metric is quoted on: tinyCapture, 65,536 envs per GPU, uniform random actions
(self-play, all four agents driven by the caller), length 300.  Envs shard per
GPU with no data-path collective ("weak" scaling: per-GPU batch fixed).

Printed JSON (rank 0, one line) follows the driver contract; see DESIGN.md §5.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes per agent-step = obs write 32*H*W + 1/4 of the per-env
# state read+write, actions and small outputs.
ALGO_BYTES_PER_AGENT_STEP = {"tinyCapture": 4529, "smallCapture": 4985, "bloxCapture": 12875}
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")       # the unmodified Python reference (copied there by build(); git-ignored,
                                                        # shipped to the GPU box with the snapshot)


def layout_path(name):
    return os.path.join(ROOT, "pacman_marl_2025_b200", "layouts", name + ".lay")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# The reference itself (unmodified Python: gymPacMan.py:143-193 + capture.py + agents/baselineTeam.py) timed on
# this box's host cores: one process per core, `torch.set_num_threads(1)`, a bounded number of seconds.
# ---------------------------------------------------------------------------------------------------------
def _pyref_worker(job):
    """One process = one env.  job = (rank, variant, layout, seconds).  -> (env_steps, seconds, episodes)"""
    rank, variant, layout, seconds = job
    import contextlib, io, random
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF_ROOT)
    os.chdir(REF_ROOT)                                  # loadAgents resolves agents/<team>.py from the checkout
    import numpy as np
    import torch
    torch.set_num_threads(1)
    quiet = contextlib.redirect_stdout(io.StringIO())
    with quiet:
        from gymPacMan import gymPacMan_parallel_env
        lay = os.path.join(REF_ROOT, "layouts", layout + ".lay")
        if variant == "config1":                        # BASELINE config 1: red = baselineTeam bots, blue = uniform over legal lists
            env = gymPacMan_parallel_env(layout_file=lay, enemieName="baselineTeam", self_play=False, length=299)
        else:                                           # env only: self-play, uniform random actions for all four agents
            env = gymPacMan_parallel_env(layout_file=lay, enemieName="randomTeam", self_play=True, length=299)
        random.seed(1 + rank)
        rng = np.random.default_rng(12345 + rank)
        obs, info = env.reset()
    steps = episodes = 0
    t0 = time.perf_counter()
    with quiet:
        while True:
            if variant == "config1":
                la = info["legal_actions"]
                acts = {env.agents[i]: (random.choice(la[env.agents[i]]) if i in (1, 3) else 0) for i in range(4)}
            else:
                a = rng.integers(0, 5, 4)
                acts = {i: int(a[i]) for i in range(4)}
            obs, rew, term, info = env.step(acts)
            steps += 1
            if any(term.values()):
                obs, info = env.reset()
                episodes += 1
            if (steps & 15) == 0 and time.perf_counter() - t0 >= seconds:
                break
    return steps, time.perf_counter() - t0, episodes


def python_reference_rate(layout, seconds=3.0, variants=("config1", "env_only")):
    """agent-steps/s of the unmodified Python reference on all host cores of this box, or a one-line reason why not.
    Workers are plain subprocesses of this script (`--pyref-worker`), one per core, each with a hard timeout."""
    if not os.path.exists(os.path.join(REF_ROOT, "gymPacMan.py")):
        return {"unavailable": "baseline/_ref/ holds no copy of the reference (build() copies it when /root/reference exists)"}
    cores = len(os.sched_getaffinity(0))
    out = {"cores": cores, "seconds_per_variant": seconds, "layout": layout,
           "how": "one process per host core, one env each, torch.set_num_threads(1), step loop incl. resets"}
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        for v in variants:
            procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--pyref-worker", f"{r},{v},{layout},{seconds}"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env, cwd=ROOT)
                     for r in range(cores)]
            res = []
            for pr in procs:
                try:
                    so, _ = pr.communicate(timeout=seconds + 120)
                    res.append(json.loads(so.strip().splitlines()[-1]))
                except Exception:                        # noqa: BLE001
                    pr.kill()
            if not res:
                out["unavailable"] = f"no worker of variant {v} finished"
                break
            rates = [4.0 * r["steps"] / r["seconds"] for r in res]
            out[v] = {"value": sum(rates), "per_core": sum(rates) / len(rates), "unit": "agent-steps/s",
                      "workers": len(res), "env_steps": int(sum(r["steps"] for r in res)),
                      "episodes": int(sum(r["episodes"] for r in res))}
    except Exception as e:                              # noqa: BLE001 - report, never fail the bench on the baseline leg
        out["unavailable"] = f"{type(e).__name__}: {e}"
    return out


def workload_config(args, world):
    """The `config` object of the JSON line - built the same way by both arms so the driver can match them."""
    rows = [ln.strip() for ln in open(layout_path(args.layout)) if ln.strip()]
    H, Wd = len(rows), len(rows[0])
    N = args.envs
    return {"workload": f"{args.layout}, {N} envs/GPU x 4 agents, self-play uniform random actions, length {args.length}, "
                        "finished envs restart inside the step (auto_reset)",
            "envs_per_gpu": N, "layout": args.layout, "n_gpus": world,
            "obs_bytes_per_step_per_gpu": 4 * N * 32 * H * Wd,
            "l2": "no flush: each step writes %.2f GB of observations per GPU (>> 126 MB L2)" % (4 * N * 32 * H * Wd / 1e9),
            "parallelism": f"env-sharded x{world}, no data-path collective"}


def oracle_rate(layout, n_envs, steps, warmup, length, nthreads=0):
    """agent-steps/s of the CPU port (oracle/pm_oracle.c, OpenMP over envs) incl. observations."""
    import numpy as np
    from oracle import oracle
    rows = [ln.strip() for ln in open(layout_path(layout))]
    env = oracle.OracleBatch(rows, n_envs, length=length)
    env.reset(with_obs=False)
    rng = np.random.default_rng(12345)
    acts = rng.integers(0, 5, size=(8, n_envs, 4)).astype(np.int32)
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would hide them)
    threads = len(os.sched_getaffinity(0)) if nthreads <= 0 else nthreads
    import ctypes as C
    L = oracle.lib()
    obs = np.empty((n_envs, 4, 8, env.H, env.W), np.float32)
    rew = np.zeros((n_envs, 2)); done = np.zeros(n_envs, np.uint8); legal = np.zeros((n_envs, 4), np.uint8)
    sc = np.zeros(n_envs, np.int32); sh = np.zeros((n_envs, 4))
    p = lambda a: a.ctypes.data_as(C.c_void_p)

    def one(i):
        L.pmo_step_batch(C.c_void_p(env.buf.ctypes.data), n_envs, p(acts[i % 8]), p(obs), p(rew), p(done), p(legal),
                         p(sc), p(sh), 1, threads)
    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    dt = time.perf_counter() - t0
    return 4.0 * n_envs * steps / dt, dt / steps, threads


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the path on the box's host cores; rank 0 only.
    `value` is the C port of capture.py / gymPacMan.py (oracle/pm_oracle.c, OpenMP over envs, all host threads) on
    one GPU's worth of the same workload - the FASTEST CPU form of the reference's algorithm, so the ratio the driver
    computes is conservative.  The unmodified Python reference is timed in the same run, on the same cores
    (`cpu_baseline.python_reference`): BASELINE config 1 and the env-only loop, one process per core."""
    if rank != 0:
        return
    n = args.envs
    v, sps, threads = oracle_rate(args.layout, n, args.steps, args.warmup, args.length)
    pyref = python_reference_rate(args.layout, seconds=3.0)
    line = {
        "impl": "reference", "metric": "agent-steps/sec (env step+obs)", "value": v, "unit": "agent-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitsets + f64 rewards -> f32 obs",
        "data": "synthetic (uniform random actions, seeded)",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of one GPU's batch ({n} envs) on {threads} OpenMP threads "
                                   "(C port of capture.py/gymPacMan.py incl. observations)",
                         "python_reference": pyref},
        "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_layout(torch, env_cls, layout, n_envs, steps, warmup, length, device, pool=8):  # noqa: ARG001 (steps)
    """-> dict(step_ms total loop, kernel_ms step kernel alone) for one layout on this rank."""
    env = env_cls(layout_file=layout_path(layout), self_play=True, length=length, num_envs=n_envs, device=device,
                  auto_reset=True)
    g = torch.Generator(device="cpu").manual_seed(1234)
    acts = torch.randint(0, 5, (pool, n_envs, 4), generator=g, dtype=torch.uint8).to(device)
    env.reset_batched()

    def one(i):
        env.step_batched(acts[i % pool])
    for i in range(warmup):
        one(i)
    return env, acts, one


def other_kernels(torch, device, N, peak):
    """BFS distance tables, GAE / advantage normalisation and the bot variant of the step kernel (informational)."""
    from pacman_marl_2025_b200 import mazeGenerator
    from pacman_marl_2025_b200.distanceCalculator import compute_distances
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    from pacman_marl_2025_b200.layout import Layout, getLayout
    from pacman_marl_2025_b200.mappo_ops import compute_gae, normalize_advantages
    out = {}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); ev0.record()
        for _ in range(reps):
            fn()
        ev1.record(); torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / reps
    # (3) all-pairs maze distances: one shipped board, and 256 distinct mazeGenerator boards in one launch (config 4)
    # `ms`: the call a user makes (host-side packing of the boards + upload + kernel); `kernel_ms`: pm_distances alone on
    # boards already packed on the device (what an env does with its own table)
    from pacman_marl_2025_b200.gymPacMan import pack_tables
    import numpy as _np

    def dev_table(lays):
        return torch.from_numpy(pack_tables(lays)[1].view(_np.int32)).to(device)
    blox = [getLayout(layout_path("bloxCapture"))]
    ms = timed(lambda: compute_distances(blox, device=device), 20)
    tb, ob = dev_table(blox), compute_distances(blox, device=device)
    kms = timed(lambda: compute_distances(blox, device=device, table=tb, out=ob), 50)
    out["bfs_bloxCapture"] = {"ms": ms, "kernel_ms": kms, "n_open": 190, "reference_python_ms_survey_container": 100.0}
    mazes = [Layout(mazeGenerator.generateMaze(1000 + s).split("\n")) for s in range(256)]
    tabs = compute_distances(mazes, device=device)
    ms = timed(lambda: compute_distances(mazes, device=device), 5)
    tm = dev_table(mazes)
    kms = timed(lambda: compute_distances(mazes, device=device, table=tm, out=tabs), 20)
    out["bfs_256_random_mazes"] = {"ms": ms, "ms_per_maze": ms / 256, "kernel_ms": kms, "kernel_us_per_maze": kms * 1e3 / 256,
                                   "table_bytes": int(tabs.dist.numel() * 2),
                                   "reference_python_ms_per_maze_survey_container": "190-335"}
    # (4) GAE over a [T, M] rollout: 20 algorithmic bytes per element (read r, V, done; write A, ret)
    T, M = 256, 2 * N
    r, v = torch.randn(T, M, device=device), torch.randn(T, M, device=device)
    d = (torch.rand(T, M, device=device) < 0.01).float()
    lv = torch.randn(M, device=device)
    ms = timed(lambda: compute_gae(r, v, d, lv), 20)
    gbs = 20.0 * T * M / (ms * 1e-3) / 1e9
    out["gae"] = {"ms": ms, "T": T, "columns": M, "achieved_GBps": gbs, "roofline_frac": gbs / peak}
    adv = torch.randn(T * M, device=device)
    idx = torch.randperm(T * M, device=device)
    ms = timed(lambda: normalize_advantages(adv, idx, 8192), 10)
    out["adv_norm"] = {"ms": ms, "samples": T * M, "minibatch": 8192}
    del r, v, d, adv, idx
    # (1b) step kernel with the on-device baselineTeam playing red (look-ahead over legal actions + maze distances)
    env = gymPacMan_parallel_env(layout_file=layout_path("tinyCapture"), self_play=False, enemieName="baselineTeam",
                                 length=300, num_envs=N, device=device, auto_reset=True)
    acts = torch.randint(0, 5, (4, N, 4), dtype=torch.uint8, device=device)
    env.reset_batched()
    k = [0]

    def one():
        env.step_batched(acts[k[0] % 4]); k[0] += 1
    ms = timed(one, 100)
    out["step_vs_baselineTeam_tinyCapture"] = {"ms": ms, "agent_steps_per_s": 4.0 * N / (ms * 1e-3),
                                               "roofline_frac": 4529 * 4.0 * N / (ms * 1e-3) / 1e9 / peak}
    # (1c) trainer mode: no fp32 observations, the step writes the learner views (2 actor + 1 critic, uint8) instead:
    #      3*8*H*W = 3 360 B per env-step on tiny (+ ~200 B of state and small outputs) instead of 4*32*H*W = 17 920 B
    for name, kw in (("trainer_views_vs_randomTeam_tinyCapture", dict(enemieName="randomTeam")),
                     ("trainer_views_vs_baselineTeam_tinyCapture", dict(enemieName="baselineTeam"))):
        env = gymPacMan_parallel_env(layout_file=layout_path("tinyCapture"), self_play=False, length=300, num_envs=N,
                                     device=device, auto_reset=True, **kw)
        env.reset_batched()
        env.set_view_outputs(False, raw_obs=False)

        def one_v():
            env.step_batched(acts[k[0] % 4]); k[0] += 1
        ms = timed(one_v, 100)
        vbytes = (3 * 8 * 140 + 200) * N
        out[name] = {"ms": ms, "agent_steps_per_s": 4.0 * N / (ms * 1e-3), "bytes_per_launch": vbytes,
                     "achieved_GBps": vbytes / (ms * 1e-3) / 1e9}
    # (1d) BASELINE config 4: per-env distinct mazeGenerator boards (4 096 different 20x20 mazes over the batch; walls,
    #      starts and the wall plane come from the layout table in HBM instead of shared memory)
    mz = [Layout(mazeGenerator.generateMaze(5000 + s).split("\n")) for s in range(1024)]
    envm = gymPacMan_parallel_env(layouts=mz, self_play=True, length=300, num_envs=N, device=device, auto_reset=True)
    envm.reset_batched()

    def one_m():
        envm.step_batched(acts[k[0] % 4]); k[0] += 1
    ms = timed(one_m, 40)
    out["step_1024_distinct_mazes_20x20"] = {"ms": ms, "agent_steps_per_s": 4.0 * N / (ms * 1e-3),
                                            "roofline_frac": 12875 * 4.0 * N / (ms * 1e-3) / 1e9 / peak}
    del envm
    torch.cuda.empty_cache()
    # (0) BASELINE config 1 through the reference-shaped API: ONE env, dict actions, Python scalars back, red side =
    #     baselineTeam (on-device), blue = uniform over the returned legal-action lists; wall clock incl. resets
    import random
    env1 = gymPacMan_parallel_env(layout_file=layout_path("tinyCapture"), self_play=False, enemieName="baselineTeam",
                                  length=299, num_envs=1, device=device)
    rnd = random.Random(1)
    obs, info = env1.reset()
    n_steps, t0 = 0, None
    for it in range(2300):
        if it == 300:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        acts = {i: rnd.choice(info["legal_actions"][i]) for i in (1, 3)}
        obs, rew, term, info = env1.step(acts)
        n_steps += it >= 300
        if term[1]:
            obs, info = env1.reset()
    dt = time.perf_counter() - t0
    out["single_env_reference_api_tinyCapture_vs_baselineTeam"] = {
        "env_steps_per_s": n_steps / dt, "agent_steps_per_s": 4.0 * n_steps / dt,
        "note": "one env, env.step({agent: action}) per step; the Python reference on the same config, same box: "
                                                               "cpu_baseline.python_reference.config1.per_core"}
    return out


def mappo_one(torch, device, layout, amp):
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    cfg = M.TrainConfig(layout_file=layout_path(layout), num_envs=1024, num_steps=32, amp_bf16=amp, checkpoint_every=0)
    tr = M.MAPPOTrainer(cfg, device=device)
    tr.train_one(0)                                     # eager warm-up (cuDNN autotune, allocator)
    tr.train_one(1)                                     # the rollout steps are captured into CUDA graphs here
    recs = [tr.train_one(2), tr.train_one(3)]
    best = max(recs, key=lambda r: r["sps"])
    del tr
    torch.cuda.empty_cache()
    return {"samples_per_s": best["sps"], "seconds_per_update": best["seconds"]}


def mappo_sps(torch, device, world=1):
    """BASELINE's second metric, MAPPO SPS (config 5): learner-agent samples per second of rollout + GAE + PPO update
    on smallCapture, whole job (every rank runs the same trainer on its env shard; one flat-bucket NCCL gradient
    all-reduce per optimizer step).  Short run: 2 warm-up updates (eager, then graph capture) + 2 timed.  Two arithmetic modes, labelled: the
    reference's nets are fp32; `fp32_tf32` keeps fp32 storage with TF32 tensor-core matmuls (narrower than fp32
    inside the matmul), `bf16_autocast` is narrower still."""
    out = {}
    for name, amp in (("fp32_tf32", False), ("bf16_autocast", True)):
        out[name] = mappo_one(torch, device, "smallCapture", amp)
    out["config"] = (f"smallCapture, 1024 envs/GPU x 32 steps x 2 learners per update, 3 epochs x 8 minibatches, randomTeam "
                     f"opponent, {world} GPU(s), NCCL gradient all-reduce per optimizer step" )
    out["reference"] = "timed at --gpus 1 (also.mappo.reference): the unmodified reference trainer on the same box"
    return out


def python_reference_mappo(num_steps=512, updates=1, timeout=240):
    """MAPPO SPS of the unmodified reference trainer on this box (its nets on this box's GPU, its env on one host core):
    baseline/ref_mappo_driver.py runs `pacman_mappo_resnet.train()` with NUM_STEPS / TOTAL_UPDATES shortened."""
    import shutil
    import tempfile
    if not os.path.exists(os.path.join(REF_ROOT, "pacman_mappo_resnet.py")):
        return {"unavailable": "baseline/_ref is not populated (run __graft_entry__.build() where /root/reference is mounted)"}
    tmp = tempfile.mkdtemp(prefix="refmappo_")
    try:
        for d in ("layouts", "agents"):
            os.symlink(os.path.join(REF_ROOT, d), os.path.join(tmp, d))
        env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
        for k in ("OMP_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
            env.pop(k, None)
        p = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_mappo_driver.py"), REF_ROOT, str(num_steps), str(updates)],
                           cwd=tmp, env=env, capture_output=True, text=True, timeout=timeout)
        line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
        if p.returncode != 0 or not line:
            return {"unavailable": (p.stderr or p.stdout).strip().splitlines()[-1][:200] if (p.stderr or p.stdout).strip() else f"rc {p.returncode}"}
        d = json.loads(line[-1])
        return {"samples_per_s": d["samples"] / d["seconds"], "seconds": d["seconds"], "samples": d["samples"],
                "nets_on": d["device"], "how": f"unmodified pacman_mappo_resnet.train(): 1 env on one host core (bloxCapture, hard-coded `:341-359`), "
                f"NUM_STEPS shortened to {num_steps}, {d['updates']} updates, no evaluation; includes process start-up of the two envs"}
    except Exception as e:                              # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


CHAIN = 50      # steps per pm_step_chain call in the timed loops


def run_chained(env, acts, steps):
    """`steps` consecutive env steps as pm_step_chain runs (acts [CHAIN, N, 4], reused cyclically)."""
    done = 0
    while done < steps:
        c = min(acts.shape[0], steps - done)
        env.step_chain(acts[:c])
        done += c


def time_chained(torch, env, acts, steps):
    """-> ms per step of `steps` chained steps (one CUDA-event pair around the whole run on the launching stream)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); run_chained(env, acts, steps); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def time_launches(torch, fn, reps):
    """-> (avg ms per call from one event pair per call, total ms of the back-to-back loop)"""
    ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for i in range(reps):
        ee[i][0].record(); fn(i); ee[i][1].record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ee) / reps, ee[0][0].elapsed_time(ee[-1][1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--layout", default="tinyCapture")
    ap.add_argument("--envs", type=int, default=65536, help="envs per GPU")
    ap.add_argument("--length", type=int, default=300)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / other layouts (profiling runs)")
    ap.add_argument("--pyref-worker", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.pyref_worker:                               # internal: one worker of python_reference_rate
        r, v, lay, sec = args.pyref_worker.split(",")
        st, dt, ep = _pyref_worker((int(r), v, lay, float(sec)))
        print(json.dumps({"steps": st, "seconds": dt, "episodes": ep}), flush=True)
        return
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from pacman_marl_2025_b200.gymPacMan import gymPacMan_parallel_env
    torch.cuda.set_device(local)
    device = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    N, K, W = args.envs, args.steps, args.warmup
    cfg = workload_config(args, world)
    H, Wd = gymPacMan_parallel_env(layout_file=layout_path(args.layout), self_play=True, num_envs=1,
                                   device=device).obs_shape[1:]
    env, acts, one = measure_layout(torch, gymPacMan_parallel_env, args.layout, N, K, W, args.length, device, pool=CHAIN)
    pool = acts.shape[0]
    NB = env.obs_bits_words()
    run_chained(env, acts, W)

    with ClockSampler(local) as clk:
        # ---- (1) whole-step throughput: K steps (auto_reset), actions resident in HBM.  The workload's actions do not
        #          depend on the observations (uniform random self-play), so the K launches go out as pm_step_chain runs
        #          of CHAIN steps: the same kernel, the same work per step, consecutive launches overlapping per env tile
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        run_chained(env, acts, K)
        ev1.record()
        barrier()
        loop_ms = max_over_ranks(ev0.elapsed_time(ev1))
        chain_ms = time_chained(torch, env, acts, min(K, 200))        # this rank, for the roofline

        # a write-only reference point measured in the same run: the copy peak in MEASURED_PEAKS.json is read + write
        # traffic, this kernel only writes, and write-only streams run higher on this part
        fill = torch.empty(1 << 29, dtype=torch.float32, device=device)          # 2 GiB
        memset_how = "cudaMemsetAsync"
        try:
            import ctypes as C
            rt = C.CDLL("libcudart.so.12")
            rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
            st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            do_fill = lambda: rt.cudaMemsetAsync(C.c_void_p(fill.data_ptr()), 0, fill.numel() * 4, st)
            assert do_fill() == 0
        except Exception:                               # noqa: BLE001
            memset_how = "torch fill kernel (float4 stores)"
            do_fill = lambda: fill.zero_()
        do_fill(); torch.cuda.synchronize()
        ev0.record()
        for _ in range(5):
            do_fill()
        ev1.record(); torch.cuda.synchronize()
        memset_gbps = 5 * fill.numel() * 4 / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
        del fill

        # ---- (1b) the same K steps as separate pm_step launches (what a loop with a policy between the steps sees) ----
        barrier()
        ev0.record()
        for i in range(K):
            one(i)
        ev1.record()
        barrier()
        sep_loop_ms = max_over_ranks(ev0.elapsed_time(ev1))

        # ---- (2) the dominant kernel alone (pm_env_kernel<STEP>): event pair around every separate launch ----
        KK = min(K, 200)
        kern_ms, _ = time_launches(torch, lambda i: env.step_batched(acts[i % pool]), KK)

        # ---- (3) end to end through the C ABI with HOST buffers: pinned actions in, results on the host when the call
        #          returns.  Four variants of WHAT comes back, each with its bytes per step:
        #            e2e        one 48-byte record per env (rewards, shaped rewards, score_change, legal, done); the fp32
        #                       observations stay in HBM for a policy net on the same GPU (the reference trainer's consumer)
        #            e2e_bits   + the PACKED observations (8*H*W-bit string + patch word per observation): everything a
        #                       HOST-side policy needs, 32x smaller than the fp32 planes
        #            e2e_views  + the uint8 learner views the batched trainer consumes (2 actor + 1 critic per env)
        #            with_obs   + the 4 fp32 [8,H,W] planes per env, i.e. the reference API's full return value
        acts_host = [acts[i].cpu().pin_memory() for i in range(pool)]
        chunks = int(os.environ.get("PM_BENCH_CHUNKS", "0"))      # 0 = zero-copy entry, n = memcpy pipeline with n ranges
        host_step = (lambda a: env.step_host_mapped(a)) if chunks == 0 else (lambda a: env.step_host_pipelined(a, n_chunks=chunks))

        def e2e_loop(fn, reps):
            for i in range(3):
                fn(acts_host[i % pool])
            barrier()
            t0 = time.perf_counter()
            ev0.record()
            for i in range(reps):
                fn(acts_host[i % pool])
            ev1.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3          # every call is host-synchronous: wall clock is the truth
            ms = max(ev0.elapsed_time(ev1), wall)
            barrier()
            return max_over_ranks(ms), wall

        KE = min(K, 200)
        e2e_ms, e2e_wall_ms = e2e_loop(host_step, KE)
        # the same traffic as chained runs: the actions of CHAIN steps are in pinned host memory before the call, the
        # per-step records are in pinned host memory when it returns (one synchronise per run instead of one per step)
        acts_host_chain = acts.cpu().pin_memory()
        env.step_chain_host_mapped(acts_host_chain)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, KE // CHAIN)):
            env.step_chain_host_mapped(acts_host_chain)
        e2e_chain_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e_chain_steps = max(1, KE // CHAIN) * CHAIN
        barrier()
        # the unpipelined entry (pm_step_host: one launch, six separate copies) for comparison
        plain_ms, _ = e2e_loop(lambda a: env.step_host(a), 50)
        e2e_plain = 4.0 * N * 50 * world / (plain_ms * 1e-3)
        # packed observations to the host: (a) zero-copy, the kernel writes them into pinned host memory itself;
        # (b) device buffer + cudaMemcpyAsync inside pm_step_host
        KB = min(K, 60)
        env.set_bits_output(True, raw_obs=False, host_mapped=True)
        bits_mapped_ms, _ = e2e_loop(lambda a: env.step_host_mapped(a), KB)
        env.set_bits_output(True, raw_obs=False, host_mapped=False)
        bits_copy_ms, _ = e2e_loop(lambda a: env.step_host(a, want_bits=True), KB)
        env.set_bits_output(False)
        bits_ms = min(bits_mapped_ms, bits_copy_ms)
        # uint8 learner views to the host (what the batched trainer reads, if its nets lived on the host)
        env.set_view_outputs(False, raw_obs=False)
        views_ms, _ = e2e_loop(lambda a: env.step_host(a, want_views=True), min(K, 20))
        env.set_view_outputs(None)
        # the reference API's full return value: fp32 observations to the host as well (PCIe-bound)
        obs_ms, _ = e2e_loop(lambda a: env.step_host(a, want_obs=True), 5)
    clocks = clk.summary()

    h2d = N * 4
    d2h = N * 48
    value = 4.0 * N * K * world / (loop_ms * 1e-3)
    peak, peak_src = peaks()
    algo = ALGO_BYTES_PER_AGENT_STEP.get(args.layout, 32 * H * Wd + 75)
    achieved = algo * 4.0 * N / (chain_ms * 1e-3) / 1e9
    achieved_sep = algo * 4.0 * N / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.layout}:{N}")
    rate = lambda ms, reps: 4.0 * N * reps * world / (ms * 1e-3)

    line = {
        "metric": "agent-steps/sec (env step+obs)", "value": value, "unit": "agent-steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": loop_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 bitsets + f64 rewards -> f32 obs", "data": "synthetic (uniform random actions, seeded)",
        "config": cfg,
        "roofline": {"bound": "hbm", "kernel": "pm_env_kernel<STEP>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo * 4 * N, "kernel_ms": chain_ms,
                     "how": f"average launch duration inside pm_step_chain runs of {CHAIN} steps: one CUDA-event pair around "
                            "the run on the launching stream / steps (consecutive launches overlap, so no event can sit "
                            "between them)",
                     "write_only_reference": {"memset_GBps": memset_gbps, "frac_of_memset": achieved / memset_gbps, "how": memset_how,
                                              "note": "peak is the measured COPY bandwidth (read + write); this kernel only "
                                                      "writes. A 2 GiB cudaMemset timed in this run is the write-only reference "
                                                      "point: a frac slightly above 1.0 of the copy peak is below it"},
                     "separate_launches": {"kernel_ms": kern_ms, "achieved": achieved_sep, "frac": achieved_sep / peak,
                                           "how": "one CUDA-event pair around every pm_step launch; this is the duration ncu "
                                                  "reports for a launch (profiles/: serialised, nothing to overlap with)"}},
        "separate_launches": {"value": 4.0 * N * K * world / (sep_loop_ms * 1e-3), "ms_per_step": sep_loop_ms / K,
                              "note": "the same K steps as K pm_step calls, each ordered after the previous one"},
        "e2e": {"value": rate(e2e_ms, KE), "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": KE, "wall_ms_rank0": e2e_wall_ms,
                "note": ("pm_step_host_mapped (zero-copy: the step kernel itself reads the pinned host actions and writes "
                         if chunks == 0 else f"pm_step_host_pipelined ({chunks} env ranges on two streams, cudaMemcpyAsync: ") +
                        "one 48-byte record per env (rewards, shaped rewards, score_change, legal, done) to pinned host "
                        "memory across PCIe every step); the call returns when the host holds the results; the fp32 "
                        "observations stay in HBM for a policy net on the same GPU.  e2e_bits / e2e_views / e2e_obs are the "
                        "same call with the observations ALSO delivered to the host in three formats",
                "unpipelined_pm_step_host_agent_steps_per_s": e2e_plain},
        "e2e_chain": {"value": rate(e2e_chain_ms, e2e_chain_steps), "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "steps": e2e_chain_steps,
                      "note": f"pm_step_chain_host_mapped: runs of {CHAIN} steps, every step reads its actions from and writes "
                              "its records to pinned host memory across PCIe; host wall clock around host-synchronous calls "
                              "(one synchronise per run).  For loops whose actions do not depend on the observations"},
        "e2e_bits": {"value": rate(bits_ms, KB), "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": d2h + 4 * N * NB * 4, "steps": KB,
                     "zero_copy_agent_steps_per_s": rate(bits_mapped_ms, KB), "memcpy_agent_steps_per_s": rate(bits_copy_ms, KB),
                     "note": f"records + packed observations ({NB * 4} B per observation instead of {32 * H * Wd}): "
                             "everything a host-side policy needs (gymPacMan.decode_obs_bits expands them)"},
        "e2e_views": {"value": rate(views_ms, min(K, 20)), "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": N * (16 + 1 + 4 + 4 + 16) + 3 * N * 8 * H * Wd, "steps": min(K, 20),
                      "note": "small outputs + the uint8 learner views (2 actor + 1 critic, 8*H*W B each) the batched trainer consumes"},
        "e2e_obs": {"value": rate(obs_ms, 5), "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": N * (16 + 1 + 4 + 4 + 16) + 4 * N * 32 * H * Wd, "steps": 5,
                    "note": "the reference API's full return value: four fp32 [8,H,W] observations per env to the host (PCIe-bound)"},
        "gpu_launches": K,
        "clocks": clocks,
    }

    # ---- BASELINE configs 3 and 5 at every N: bloxCapture step+obs (+ one distance table) and MAPPO SPS ----
    also = {}
    del env, acts, one
    torch.cuda.empty_cache()
    if not args.no_extras:
        for lay in ("smallCapture", "bloxCapture"):
            if lay == args.layout or (lay == "smallCapture" and world > 1):
                continue
            e2, a2, one2 = measure_layout(torch, gymPacMan_parallel_env, lay, N, 0, 5, args.length, device, pool=20)
            kk = 60
            run_chained(e2, a2, 20)
            barrier()
            cm = max_over_ranks(time_chained(torch, e2, a2, kk))
            barrier()
            km, _ = time_launches(torch, lambda i: e2.step_batched(a2[i % a2.shape[0]]), kk)
            ab = ALGO_BYTES_PER_AGENT_STEP[lay]
            also[lay] = {"value": 4.0 * N * world / (cm * 1e-3), "unit": "agent-steps/s", "n_gpus": world,
                         "envs_per_gpu": N, "kernel_ms": cm, "roofline_frac": ab * 4.0 * N / (cm * 1e-3) / 1e9 / peak,
                         "separate_launches": {"kernel_ms": km, "roofline_frac": ab * 4.0 * N / (km * 1e-3) / 1e9 / peak}}
            if lay == "bloxCapture":
                from pacman_marl_2025_b200.distanceCalculator import compute_distances
                dms, _ = time_launches(torch, lambda i: compute_distances([e2.layout], device=device), 10)
                also[lay]["distance_table_ms"] = dms
            del e2, a2, one2
            torch.cuda.empty_cache()
        also["mappo"] = mappo_sps(torch, device, world)

    if rank == 0 and world == 1 and not args.no_extras:
        # CPU side on this box's host cores: the C port (bounded sample of the same workload) and the Python reference
        ncpu = min(N, 16384)
        v, sps, threads = oracle_rate(args.layout, ncpu, 6, 2, args.length)
        line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                "sample": f"6 steps x {ncpu} envs of the same workload, {threads} OpenMP threads"}
        v1, _, _ = oracle_rate(args.layout, 4096, 4, 1, args.length, nthreads=1)
        line["cpu_baseline"]["one_core_value"] = v1
        line["cpu_baseline"]["python_reference"] = python_reference_rate(args.layout, seconds=3.0)
        if "mappo" in also:
            # the reference's own train() is hard-wired to bloxCapture: time it, and this trainer on the same board
            also["mappo"]["reference"] = python_reference_mappo()
            also["mappo"]["bloxCapture_fp32_tf32"] = mappo_one(torch, device, "bloxCapture", False)
        # the same layout at 4x the batch: the fixed cost of a launch (the store-free rules phase at its start) weighs less
        e4, a4, one4 = measure_layout(torch, gymPacMan_parallel_env, args.layout, 4 * N, 0, 5, args.length, device, pool=2)
        km, _ = time_launches(torch, lambda i: e4.step_batched(a4[i % 2]), 30)
        also[f"{args.layout}_x4_envs"] = {"envs": 4 * N, "value": 16.0 * N / (km * 1e-3), "kernel_ms": km,
                                          "roofline_frac": algo * 16.0 * N / (km * 1e-3) / 1e9 / peak}
        del e4, a4, one4
        torch.cuda.empty_cache()
        # BASELINE config 3 asks for bloxCapture over a range of batch sizes: the two ends beside the 65,536 above
        ab = ALGO_BYTES_PER_AGENT_STEP["bloxCapture"]
        sweep = {}
        for nb in (16384, 262144):
            eb, a_b, _ = measure_layout(torch, gymPacMan_parallel_env, "bloxCapture", nb, 0, 3, args.length, device, pool=10)
            run_chained(eb, a_b, 10)
            cm = time_chained(torch, eb, a_b, 20)
            km, _ = time_launches(torch, lambda i: eb.step_batched(a_b[i % 10]), 20)
            sweep[str(nb)] = {"value": 4.0 * nb / (cm * 1e-3), "kernel_ms": cm, "roofline_frac": ab * 4.0 * nb / (cm * 1e-3) / 1e9 / peak,
                              "separate_launches": {"kernel_ms": km, "roofline_frac": ab * 4.0 * nb / (km * 1e-3) / 1e9 / peak}}
            del eb, a_b
            torch.cuda.empty_cache()
        also["bloxCapture_envs_sweep"] = sweep
        # the other kernels of the path, each timed alone (CUDA events, this stream)
        also["kernels"] = other_kernels(torch, device, N, peak)
    if also:
        line["also"] = also

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
