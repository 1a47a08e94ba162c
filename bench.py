#!/usr/bin/env python
"""bench.py — agent-steps/sec of the batched capture env (step + obs) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--layout tinyCapture] [--envs 65536]

One "step" = one pass of the hot path over the whole batch: ONE pm_step launch
(rules of 4 ordered sub-moves + four fp32 [8,H,W] observations per env, rewards,
done, legal masks, shaping) with pm_env_desc.auto_reset, i.e. envs that finish
are re-initialised inside the same launch, so episodes keep cycling.  Default workload = the configuration BASELINE.json's
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
PY_REFERENCE_AGENT_STEPS_PER_CORE = {"tinyCapture": 1819, "smallCapture": 1554, "bloxCapture": 1031}


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
    """Reference arm: the reference's CPU implementation of the path on the box's host
    cores.  The reference is pure Python (cannot travel to the GPU box), so
    this times the oracle port with all host threads; rank 0 only."""
    if rank != 0:
        return
    n = args.envs
    v, sps, threads = oracle_rate(args.layout, n, args.steps, args.warmup, args.length)
    line = {
        "impl": "reference", "metric": "agent-steps/sec (env step+obs)", "value": v, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 bitsets + f64 rewards -> f32 obs",
        "data": "synthetic (uniform random actions, seeded)",
        "config": {"workload": f"{args.layout}, {n} envs x 4 agents, self-play random actions, length {args.length}, "
                               "auto-reset", "envs": n, "layout": args.layout},
        "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of the full {n}-env batch on {threads} OpenMP threads "
                                   f"(C port of capture.py/gymPacMan.py; the Python reference itself ran "
                                   f"{PY_REFERENCE_AGENT_STEPS_PER_CORE.get(args.layout)} agent-steps/s/core at survey time)"},
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
    blox = [getLayout(layout_path("bloxCapture"))]
    ms = timed(lambda: compute_distances(blox, device=device), 20)
    out["bfs_bloxCapture"] = {"ms": ms, "n_open": 190, "reference_python_ms": 100.0}
    mazes = [Layout(mazeGenerator.generateMaze(1000 + s).split("\n")) for s in range(256)]
    tabs = compute_distances(mazes, device=device)
    ms = timed(lambda: compute_distances(mazes, device=device), 5)
    out["bfs_256_random_mazes"] = {"ms": ms, "ms_per_maze": ms / 256, "table_bytes": int(tabs.dist.numel() * 2),
                                   "reference_python_ms_per_maze": "190-335"}
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
        "reference_python_agent_steps_per_s_1core": 1414, "note": "one env, env.step({agent: action}) per step"}
    return out


def mappo_sps(torch, device):
    """BASELINE's second metric, MAPPO SPS: learner-agent samples per second of rollout + GAE + PPO update on
    smallCapture (short run: 1 warm-up update + 1 timed; tools/bench_mappo.py is the long form)."""
    from pacman_marl_2025_b200 import pacman_mappo_resnet as M
    out = {}
    for name, amp in (("fp32_tf32", False), ("bf16_autocast", True)):
        cfg = M.TrainConfig(layout_file=layout_path("smallCapture"), num_envs=1024, num_steps=32, amp_bf16=amp,
                            checkpoint_every=0)
        tr = M.MAPPOTrainer(cfg, device=device)
        tr.train_one(0)
        rec = tr.train_one(1)
        out[name] = {"samples_per_s": rec["sps"], "seconds_per_update": rec["seconds"]}
        del tr
        torch.cuda.empty_cache()
    out["config"] = "smallCapture, 1024 envs x 32 steps x 2 learners per update, 3 epochs x 8 minibatches, randomTeam opponent, 1 GPU"
    return out


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
    args = ap.parse_args()
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

    N, K, W = args.envs, args.steps, args.warmup
    H, Wd = gymPacMan_parallel_env(layout_file=layout_path(args.layout), self_play=True, num_envs=1,
                                   device=device).obs_shape[1:]
    env, acts, one = measure_layout(torch, gymPacMan_parallel_env, args.layout, N, K, W, args.length, device)
    pool = acts.shape[0]

    with ClockSampler(local) as clk:
        # ---- (1) whole-step throughput: K x (pm_step + pm_reset(done)), inputs resident in HBM ----
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(K):
            one(i)
        ev1.record()
        barrier()
        loop_ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([loop_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            loop_ms = float(t.item())

        # ---- (2) the dominant kernel alone (pm_env_kernel<STEP>): event pair around every launch ----
        KK = min(K, 200)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KK)]
        torch.cuda.synchronize()
        for i in range(KK):
            evs[i][0].record()
            env.step_batched(acts[i % pool])
            evs[i][1].record()
        torch.cuda.synchronize()
        kern_ms = sum(a.elapsed_time(b) for a, b in evs) / KK

        # ---- (3) end to end through the C ABI with HOST buffers (pm_step_host_pipelined): pinned actions in,
        #          one 48-byte record per env (rewards, shaped rewards, score_change, legal, done) back to pinned
        #          host memory every step, call returns when the host has the results; obs stay in HBM for the policy
        acts_host = [acts[i].cpu().pin_memory() for i in range(pool)]
        chunks = int(os.environ.get("PM_BENCH_CHUNKS", "0"))      # 0 = zero-copy entry, n = memcpy pipeline with n ranges
        host_step = (lambda a: env.step_host_mapped(a)) if chunks == 0 else (lambda a: env.step_host_pipelined(a, n_chunks=chunks))
        for i in range(3):
            host_step(acts_host[i % pool])
        KE = min(K, 200)
        barrier()
        t0 = time.perf_counter()
        ev0.record()
        for i in range(KE):
            host_step(acts_host[i % pool])
        ev1.record()
        torch.cuda.synchronize()
        e2e_wall_ms = (time.perf_counter() - t0) * 1e3         # every call is host-synchronous: wall clock is the truth
        e2e_ms = max(ev0.elapsed_time(ev1), e2e_wall_ms)
        barrier()
        if world > 1:
            t = torch.tensor([e2e_ms], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        # the unpipelined entry (pm_step_host: one launch, six separate copies) for comparison
        for i in range(3):
            env.step_host(acts_host[i % pool])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(50):
            env.step_host(acts_host[i % pool])
        torch.cuda.synchronize()
        e2e_plain = 4.0 * N * 50 / (time.perf_counter() - t0)
        # same, but also hauling the 4 x [N,8,H,W] fp32 observations to the host (a CPU-side consumer)
        e2e_obs = None
        if rank == 0 and not args.no_extras:
            for i in range(2):
                env.step_host(acts_host[i % pool], want_obs=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(5):
                env.step_host(acts_host[i % pool], want_obs=True)
            torch.cuda.synchronize()
            e2e_obs = 4.0 * N * 5 / (time.perf_counter() - t0)
    clocks = clk.summary()

    h2d = N * 4
    d2h = N * 48
    value = 4.0 * N * K * world / (loop_ms * 1e-3)
    peak, peak_src = peaks()
    algo = ALGO_BYTES_PER_AGENT_STEP.get(args.layout, 32 * H * Wd + 75)
    achieved = algo * 4.0 * N / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{args.layout}:{N}")

    line = {
        "metric": "agent-steps/sec (env step+obs)", "value": value, "unit": "agent-steps/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": loop_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 bitsets + f64 rewards -> f32 obs", "data": "synthetic (uniform random actions, seeded)",
        "config": {"workload": f"{args.layout}, {N} envs/GPU x 4 agents, self-play random actions, length {args.length}, "
                               "finished envs restart inside the step launch (auto_reset)", "envs_per_gpu": N, "layout": args.layout,
                   "obs_bytes_per_step_per_gpu": 4 * N * 32 * H * Wd,
                   "l2": "no flush: each step writes %.2f GB of observations per GPU (>> 126 MB L2)" % (4 * N * 32 * H * Wd / 1e9),
                   "parallelism": f"env-sharded x{world}, no data-path collective"},
        "roofline": {"bound": "hbm", "kernel": "pm_env_kernel<STEP>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": algo * 4 * N, "kernel_ms": kern_ms},
        "e2e": {"value": 4.0 * N * KE * world / (e2e_ms * 1e-3), "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": KE, "wall_ms_rank0": e2e_wall_ms,
                "note": ("pm_step_host_mapped (zero-copy: the step kernel itself reads the pinned host actions and writes "
                         if chunks == 0 else f"pm_step_host_pipelined ({chunks} env ranges on two streams, cudaMemcpyAsync: ") +
                        "one 48-byte record per env (rewards, shaped rewards, score_change, legal, done) to pinned host "
                        "memory across PCIe every step); the call returns when the host holds the results; observations "
                        "stay in HBM for the policy net",
                "unpipelined_pm_step_host_agent_steps_per_s": e2e_plain,
                "with_obs_to_host_agent_steps_per_s": e2e_obs},
        "gpu_launches": K,
        "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_extras:
        # CPU port on this box's host cores: bounded sample of the same workload
        ncpu = min(N, 16384)
        v, sps, threads = oracle_rate(args.layout, ncpu, 6, 2, args.length)
        line["cpu_baseline"] = {"value": v, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                                "sample": f"6 steps x {ncpu} envs of the same workload, {threads} OpenMP threads; "
                                          f"Python reference at survey time: {PY_REFERENCE_AGENT_STEPS_PER_CORE.get(args.layout)} agent-steps/s/core"}
        v1, _, _ = oracle_rate(args.layout, 4096, 4, 1, args.length, nthreads=1)
        line["cpu_baseline"]["one_core_value"] = v1
        # the other shipped layouts, same batch, short loop (informational)
        also = {}
        del env, acts, one
        torch.cuda.empty_cache()
        for lay in ("smallCapture", "bloxCapture"):
            if lay == args.layout:
                continue
            e2, a2, one2 = measure_layout(torch, gymPacMan_parallel_env, lay, N, 0, 5, args.length, device)
            h2, w2 = e2.obs_shape[1:]
            kk = 60
            ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(kk)]
            torch.cuda.synchronize()
            ev0.record()
            for i in range(kk):
                ee[i][0].record(); e2.step_batched(a2[i % a2.shape[0]]); ee[i][1].record()
            ev1.record()
            torch.cuda.synchronize()
            km = sum(a.elapsed_time(b) for a, b in ee) / kk
            ab = ALGO_BYTES_PER_AGENT_STEP[lay]
            also[lay] = {"value": 4.0 * N * kk / (ev0.elapsed_time(ev1) * 1e-3), "kernel_ms": km,
                         "roofline_frac": ab * 4.0 * N / (km * 1e-3) / 1e9 / peak}
            del e2, a2, one2
            torch.cuda.empty_cache()
        # the same layout at 4x the batch: the fixed cost of a launch (the store-free rules phase at its start) weighs less
        e4, a4, one4 = measure_layout(torch, gymPacMan_parallel_env, args.layout, 4 * N, 0, 5, args.length, device, pool=2)
        ee = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
        torch.cuda.synchronize()
        for i in range(30):
            ee[i][0].record(); e4.step_batched(a4[i % 2]); ee[i][1].record()
        torch.cuda.synchronize()
        km = sum(a.elapsed_time(b) for a, b in ee) / 30
        also[f"{args.layout}_x4_envs"] = {"envs": 4 * N, "value": 16.0 * N / (km * 1e-3), "kernel_ms": km,
                                          "roofline_frac": algo * 16.0 * N / (km * 1e-3) / 1e9 / peak}
        del e4, a4, one4
        torch.cuda.empty_cache()
        # the other kernels of the path, each timed alone (CUDA events, this stream)
        also["kernels"] = other_kernels(torch, device, N, peak)
        also["mappo"] = mappo_sps(torch, device)
        line["also"] = also

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
