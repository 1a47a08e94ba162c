"""In-tree build of libpacman_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m pacman_marl_2025_b200.build [--force] [-v]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpacman_b200.so")
SOURCES = ["pm_env.cu", "pm_bfs.cu", "pm_gae.cu", "pm_ppo.cu", "pm_attn.cu", "pm_norm.cu", "pm_pack.cpp"]
HEADERS = ["pm_env_core.cuh", "pm_format.h", os.path.join("..", "..", "include", "pacman_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                      # rewards/GAE reproduce the reference's one-rounding-per-op arithmetic
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-shared", "-cudart", "static",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libpacman_b200.so")


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, out: str = OUT) -> str:
    if out == OUT and not force and not stale():
        return OUT
    extra = os.environ.get("PM_NVCC_EXTRA", "").split()      # tuning sweeps: -DPM_CONS_WARPS=.. -DPM_CTAS_PER_SM=..
    cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + \
          [os.path.join(CSRC, f) for f in SOURCES]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    outs = [a[6:] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else OUT))
