"""B200-native batched capture-the-flag Pac-Man environment.

Hot path: fused step + observation kernels, per-layout BFS distance
tables, shaping + GAE + advantage normalisation — hand-written sm_100a CUDA
behind the C ABI in ``include/pacman_b200.h`` (``libpacman_b200.so``), with this
package as the thin PyTorch host side that mirrors the reference's
``gymPacMan`` / ``layout`` / ``mazeGenerator`` / ``distanceCalculator`` modules.
"""
from . import layout, mazeGenerator  # noqa: F401  (pure host modules)

__all__ = ["layout", "mazeGenerator", "gymPacMan_parallel_env", "compute_distances", "Distancer",
           "compute_gae", "normalize_advantages"]


def __getattr__(name):  # torch / CUDA-dependent parts load lazily
    if name == "gymPacMan_parallel_env":
        from .gymPacMan import gymPacMan_parallel_env
        return gymPacMan_parallel_env
    if name in ("compute_distances", "Distancer", "DistanceTables"):
        from . import distanceCalculator
        return getattr(distanceCalculator, name)
    if name in ("compute_gae", "normalize_advantages"):
        from . import mappo_ops
        return getattr(mappo_ops, name)
    raise AttributeError(name)
