"""Env sharding and the trainer's two collectives.

The env path needs no exchange: rank r of R owns a contiguous range of the
global env indices and steps it with its own kernels.  The only collectives are
  (i)  one all-reduce of the flat gradient bucket per optimizer step, and
  (ii) one all-reduce of a small episode-statistics vector per update,
both through ``torch.distributed`` (NCCL over NVLink on the B200 box, gloo in
the CPU tests).  The reference has no distributed code at all
(`pacman_mappo_resnet.py` is single-process); the call sites these serve are its
``loss.backward()`` → ``optimizer.step()`` (`:587-590`) and its per-update log
(`:603-644`).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist

ENV_ALIGN = 32      # pm_env_desc.env_begin granularity (one warp of envs)


def shard_range(total_envs: int, rank: int, world: int, align: int = ENV_ALIGN) -> Tuple[int, int]:
    """Contiguous [begin, begin+count) of the global env indices owned by ``rank``.
    Every boundary is a multiple of ``align``; the last rank takes the remainder.  The
    ranges of all ranks tile [0, total_envs) exactly."""
    if not (0 <= rank < world) or total_envs < 0:
        raise ValueError("bad rank/world/total")
    blocks = (total_envs + align - 1) // align
    per, extra = divmod(blocks, world)
    b0 = rank * per + min(rank, extra)
    b1 = b0 + per + (1 if rank < extra else 0)
    begin, end = min(b0 * align, total_envs), min(b1 * align, total_envs)
    return begin, end - begin


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


FLAT_ALIGN = 64     # elements: every tensor starts on a 256-byte boundary inside a flat buffer (cuDNN / cuBLAS pick
                    # their vectorised / tensor-core kernels only for 16-byte aligned operands)


def flat_offsets(params, align: int = FLAT_ALIGN):
    """Start offset of every tensor in a flat buffer with aligned starts; returns (offsets, total elements)."""
    offs, o = [], 0
    for p in params:
        offs.append(o)
        o += (p.numel() + align - 1) // align * align
    return offs, o


class FlatGradBucket:
    """All gradients of a module in ONE contiguous buffer, so data-parallel
    averaging is a single all-reduce (9-25 MB for MAPPOAgent: latency-bound, so
    one bucket beats per-tensor calls).  ``p.grad`` of every parameter is made a
    view into the bucket once; backward then accumulates straight into it and no
    pack / unpack copy is needed around the collective.  Tensors start on 256-byte
    boundaries (zero padding in between, which every reduction over the buffer ignores)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        self.offsets, n = flat_offsets(self.params)
        self.flat = torch.zeros(n, dtype=dt, device=dev)
        for p, o in zip(self.params, self.offsets):
            if p.device != dev or p.dtype != dt:
                raise ValueError("FlatGradBucket needs all parameters on one device with one dtype")
            p.grad = self.flat[o:o + p.numel()].view_as(p)

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean_(self, group=None):
        """In-place average over the group (no-op for world size 1)."""
        rank, world = world_info(group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(world)
        return self.flat

    def all_reduce_sum_(self, group=None):
        """In-place SUM over the group (no-op for world size 1); the caller folds the 1/world into its next kernel."""
        rank, world = world_info(group)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        return self.flat

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """nn.utils.clip_grad_norm_ (pacman_mappo_resnet.py:589) on the flat view: one norm, one scale."""
        total = torch.linalg.vector_norm(self.flat)
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        self.flat.mul_(coef)
        return total


def all_reduce_stats_(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a small per-rank statistics vector (episode count, return sum, wins, ...) over ranks."""
    rank, world = world_info(group)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def broadcast_module_(module: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers."""
    rank, world = world_info(group)
    if world > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
