"""Rollout post-processing ops used by the MAPPO trainer, on the GPU.

Reference: `/root/reference/pacman_mappo_resnet.py`
  * ``compute_gae``            `:277-290`  -> :func:`compute_gae` (kernel pm_gae)
  * minibatch adv. normalise   `:577`      -> :func:`normalize_advantages` (kernel pm_adv_norm)
  * ``compute_heuristic_shaping`` `:251-264` is fused into the step kernel
    (``step_batched()['shaping']`` / ``['shaped_reward']``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

GAMMA = 0.99        # pacman_mappo_resnet.py:19
GAE_LAMBDA = 0.95   # pacman_mappo_resnet.py:20


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def compute_gae(rewards, values, dones, last_value, gamma=GAMMA, gae_lambda=GAE_LAMBDA):
    """Reverse-scan GAE over M independent columns.

    rewards / values / dones: fp32 CUDA tensors [T] or [T, M] (any trailing shape is
    flattened to columns); last_value: Python float, or a tensor broadcastable to [M].
    Returns (advantages, returns) shaped like ``rewards``.  ``dones[t]`` gates the
    bootstrap at step t exactly as in the reference (quirk: not dones[t+1]).
    """
    L = _lib.load()
    if not rewards.is_cuda:
        raise RuntimeError("compute_gae expects CUDA tensors; there is no CPU fallback")
    dev = rewards.device
    shape = rewards.shape
    T = shape[0]
    r = rewards.reshape(T, -1).contiguous().float()
    v = values.reshape(T, -1).contiguous().float()
    d = dones.reshape(T, -1).contiguous().float()
    M = r.shape[1]
    lv = torch.as_tensor(last_value, dtype=torch.float32, device=dev).reshape(-1)
    lv = lv.expand(M).contiguous() if lv.numel() == 1 else lv.reshape(M).contiguous()
    adv = torch.empty_like(r)
    ret = torch.empty_like(r)
    with torch.cuda.device(dev):
        _lib.check(L.pm_gae(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), float(gamma), float(gae_lambda),
                            adv.data_ptr(), ret.data_ptr(), int(T), int(M), _stream(dev)), "pm_gae")
    return adv.reshape(shape), ret.reshape(shape)


def normalize_advantages(adv_flat, inds, minibatch):
    """out[s:s+mb] = (a - a.mean()) / (a.std() + 1e-8) with a = adv_flat[inds[s:s+mb]]
    for every consecutive minibatch of ``inds`` (unbiased std, like torch.std)."""
    L = _lib.load()
    if not adv_flat.is_cuda:
        raise RuntimeError("normalize_advantages expects CUDA tensors; there is no CPU fallback")
    dev = adv_flat.device
    a = adv_flat.contiguous().float()
    idx = inds.to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty(idx.numel(), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.pm_adv_norm(a.data_ptr(), idx.data_ptr(), out.data_ptr(), int(idx.numel()), int(minibatch),
                                 _stream(dev)), "pm_adv_norm")
    return out
