"""ctypes binding of libpacman_b200.so (C ABI: include/pacman_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or
``python -m pacman_marl_2025_b200.build``).  There is NO fallback: if the shared
object is missing, or a kernel launch fails, the calling op raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (PM_LIB_PATH: development builds of the same library with other launch-geometry macros, for B200 sweeps)
LIB_PATH = os.environ.get("PM_LIB_PATH") or os.path.join(_HERE, "libpacman_b200.so")

PM_ERRORS = {-1: "PM_EINVAL (bad argument)", -2: "PM_ELAYOUT (malformed layout text)",
             -3: "PM_ETOOBIG (board larger than 1024 cells or a side > 255)", -4: "PM_ENOGPU"}

ABI_VERSION = 8

EXPORTS = [
    "pm_abi_version", "pm_geom_words", "pm_layout_words", "pm_state_bytes", "pm_pack_geom", "pm_pack_layout", "pm_pack_layouts",
    "pm_reset", "pm_step", "pm_observe", "pm_step_host", "pm_get_state",
    "pm_set_state", "pm_distances", "pm_gae", "pm_adv_norm", "pm_pipe_create", "pm_pipe_destroy",
    "pm_step_host_pipelined", "pm_step_host_mapped", "pm_step_host_mapped_async", "pm_observe_views",
    "pm_obs_bits_words", "pm_observe_bits",
    "pm_chain_scratch_bytes", "pm_step_chain", "pm_step_chain_host_mapped", "pm_attn_fwd", "pm_attn_bwd",
    "pm_layernorm32_blocks", "pm_add_layernorm32_fwd", "pm_add_layernorm32_bwd", "pm_groupnorm_blocks", "pm_groupnorm_4x8_max_hw",
    "pm_groupnorm_4x8_fwd", "pm_groupnorm_4x8_bwd",
    "pm_make_observation", "pm_gather_rows", "pm_ppo_scratch_bytes", "pm_ppo_loss", "pm_adam_scratch_bytes", "pm_adam_ema",
]


class PmEnvDesc(C.Structure):
    """struct pm_env_desc"""
    _fields_ = [
        ("W", C.c_int32), ("H", C.c_int32), ("n_layouts", C.c_int32), ("has_caps", C.c_int32),
        ("n_envs", C.c_int64),
        ("geom", C.c_void_p), ("layouts", C.c_void_p), ("layout_id", C.c_void_p), ("state", C.c_void_p),
        ("length", C.c_int32), ("reward_for_legal_action", C.c_int32), ("defence_reward", C.c_int32),
        ("random_agent_mask", C.c_int32),
        ("shaping_scale", C.c_double),
        ("random_seed", C.c_uint64),
        # ABI 2
        ("auto_reset", C.c_int32), ("envs_per_cta", C.c_int32), ("env_index_base", C.c_int64),
        ("env_begin", C.c_int64), ("env_count", C.c_int64),
        # ABI 4
        ("bot_kind", C.c_int32), ("bot_first_best", C.c_int32), ("dist", C.c_void_p), ("node_of_cell", C.c_void_p),
        ("dist_nmax", C.c_int32), ("layout_pool", C.c_int32),
        # ABI 6
        ("agent_begin", C.c_int32), ("agent_end", C.c_int32),
        # ABI 7
        ("bot_draws", C.c_void_p),
    ]


class PmStepOut(C.Structure):
    """struct pm_step_out"""
    _fields_ = [
        ("obs", C.c_void_p), ("obs_agent_stride", C.c_int64), ("reward", C.c_void_p), ("done", C.c_void_p),
        ("legal", C.c_void_p), ("score_change", C.c_void_p), ("shaping", C.c_void_p), ("shaped_reward", C.c_void_p),
        ("packed", C.c_void_p),     # ABI 3
        ("actions_used", C.c_void_p),   # ABI 4
        # ABI 5: learner views
        ("view_actor", C.c_void_p), ("view_actor_stride", C.c_int64), ("view_critic", C.c_void_p),
        ("view_red", C.c_int32), ("obs_static_valid", C.c_int32),
        # ABI 7: packed observations
        ("obs_bits", C.c_void_p), ("obs_bits_agent_stride", C.c_int64),
    ]


class PmHostRecord(C.Structure):
    """struct pm_host_record (48 bytes)"""
    _fields_ = [("reward_red", C.c_double), ("reward_blue", C.c_double), ("shaped_reward", C.c_float * 4),
                ("score_change", C.c_int32), ("legal", C.c_uint8 * 4), ("done", C.c_uint8), ("pad", C.c_uint8 * 7)]


def declare_pack_api(L):
    """Prototypes of the host-only part of the ABI (also exported by the test emulation)."""
    L.pm_geom_words.restype = C.c_int
    L.pm_geom_words.argtypes = [C.c_int, C.c_int]
    L.pm_layout_words.restype = C.c_int
    L.pm_layout_words.argtypes = [C.c_int, C.c_int]
    L.pm_state_bytes.restype = C.c_size_t
    L.pm_state_bytes.argtypes = [C.c_int, C.c_int, C.c_int64]
    L.pm_obs_bits_words.restype = C.c_int
    L.pm_obs_bits_words.argtypes = [C.c_int, C.c_int]
    L.pm_pack_geom.restype = C.c_int
    L.pm_pack_geom.argtypes = [C.c_int, C.c_int, C.c_void_p]
    L.pm_pack_layout.restype = C.c_int
    L.pm_pack_layout.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
    if hasattr(L, "pm_pack_layouts"):
        L.pm_pack_layouts.restype = C.c_int
        L.pm_pack_layouts.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]
    return L


_lib = None


def load():
    """Open libpacman_b200.so; raise (never fall back) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root (needs nvcc). "
            "This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    declare_pack_api(L)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.pm_abi_version.restype = i32
    L.pm_reset.restype = i32
    L.pm_reset.argtypes = [C.POINTER(PmEnvDesc), vp, vp, i64, vp, vp]
    L.pm_step.restype = i32
    L.pm_step.argtypes = [C.POINTER(PmEnvDesc), vp, i32, C.POINTER(PmStepOut), vp]
    L.pm_chain_scratch_bytes.restype = C.c_size_t
    L.pm_chain_scratch_bytes.argtypes = [i64]
    L.pm_step_chain.restype = i32
    L.pm_step_chain.argtypes = [C.POINTER(PmEnvDesc), vp, i32, i64, i32, C.POINTER(PmStepOut), i32, vp, vp]
    L.pm_step_chain_host_mapped.restype = i32
    L.pm_step_chain_host_mapped.argtypes = [C.POINTER(PmEnvDesc), vp, i32, C.POINTER(PmStepOut), vp, vp, vp]
    L.pm_observe.restype = i32
    L.pm_observe.argtypes = [C.POINTER(PmEnvDesc), vp, i64, vp]
    L.pm_observe_bits.restype = i32
    L.pm_observe_bits.argtypes = [C.POINTER(PmEnvDesc), vp, i64, vp]
    L.pm_observe_views.restype = i32
    L.pm_observe_views.argtypes = [C.POINTER(PmEnvDesc), i32, vp, i64, vp, vp]
    L.pm_step_host.restype = i32
    L.pm_step_host.argtypes = [C.POINTER(PmEnvDesc), vp, vp, C.POINTER(PmStepOut), C.POINTER(PmStepOut), vp]
    L.pm_pipe_create.restype = i32
    L.pm_pipe_create.argtypes = [C.POINTER(vp), i32]
    L.pm_pipe_destroy.restype = i32
    L.pm_pipe_destroy.argtypes = [vp]
    L.pm_step_host_pipelined.restype = i32
    L.pm_step_host_pipelined.argtypes = [C.POINTER(PmEnvDesc), vp, vp, vp, C.POINTER(PmStepOut), vp, vp]
    L.pm_step_host_mapped.restype = i32
    L.pm_step_host_mapped.argtypes = [C.POINTER(PmEnvDesc), vp, C.POINTER(PmStepOut), vp, vp]
    L.pm_step_host_mapped_async.restype = i32
    L.pm_step_host_mapped_async.argtypes = [C.POINTER(PmEnvDesc), vp, C.POINTER(PmStepOut), vp, vp]
    L.pm_get_state.restype = i32
    L.pm_get_state.argtypes = [C.POINTER(PmEnvDesc), vp, vp, vp, vp, vp]
    L.pm_set_state.restype = i32
    L.pm_set_state.argtypes = [C.POINTER(PmEnvDesc), vp, vp, vp, vp, vp]
    L.pm_distances.restype = i32
    L.pm_distances.argtypes = [i32, i32, vp, i32, vp, vp, vp, i32, vp]
    L.pm_gae.restype = i32
    L.pm_gae.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, vp, vp, i32, i64, vp]
    L.pm_adv_norm.restype = i32
    L.pm_adv_norm.argtypes = [vp, vp, vp, i64, i32, vp]
    L.pm_make_observation.restype = i32
    L.pm_make_observation.argtypes = [C.POINTER(PmEnvDesc), vp, vp, vp, vp]
    L.pm_gather_rows.restype = i32
    L.pm_gather_rows.argtypes = [vp, vp, i64, i32, vp, i32, vp]
    L.pm_ppo_scratch_bytes.restype = C.c_size_t
    L.pm_ppo_loss.restype = i32
    L.pm_ppo_loss.argtypes = [vp, vp, vp, vp, vp, vp, i64, i64, i32, C.c_double, C.c_double, C.c_double, vp, vp, vp, vp, vp]
    L.pm_adam_scratch_bytes.restype = C.c_size_t
    L.pm_adam_ema.restype = i32
    L.pm_adam_ema.argtypes = [vp, vp, vp, vp, vp, i64, i64] + [C.c_double] * 7 + [vp, vp, vp]
    L.pm_attn_fwd.restype = i32
    L.pm_attn_fwd.argtypes = [vp, vp, vp, i32, i64, i32, vp]
    L.pm_attn_bwd.restype = i32
    L.pm_attn_bwd.argtypes = [vp, vp, vp, vp, vp, i32, i64, i32, vp]
    L.pm_layernorm32_blocks.restype = i32
    L.pm_layernorm32_blocks.argtypes = [i64]
    L.pm_add_layernorm32_fwd.restype = i32
    L.pm_add_layernorm32_fwd.argtypes = [vp, vp, vp, vp, vp, vp, i64, C.c_double, vp]
    L.pm_add_layernorm32_bwd.restype = i32
    L.pm_add_layernorm32_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, vp]
    L.pm_groupnorm_blocks.restype = i32
    L.pm_groupnorm_blocks.argtypes = [i64]
    L.pm_groupnorm_4x8_max_hw.restype = i32
    L.pm_groupnorm_4x8_fwd.restype = i32
    L.pm_groupnorm_4x8_fwd.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, C.c_double, i32, vp]
    L.pm_groupnorm_4x8_bwd.restype = i32
    L.pm_groupnorm_4x8_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, vp]
    if L.pm_abi_version() != ABI_VERSION:
        raise RuntimeError("libpacman_b200.so ABI version mismatch; rebuild")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"{what}: {PM_ERRORS.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc} (cudaError_t) from the kernel launch")
