"""Maze-distance tables on the GPU — drop-in for ``distanceCalculator.Distancer``.

Reference: `/root/reference/distanceCalculator.py:26-65` (Distancer) and
`:111-150` (computeDistances: one priority-queue UCS per open cell, 9-335 ms per
board on a CPU core).  Here ``compute_distances`` builds the tables of L boards
in one launch of the bit-parallel BFS kernel (csrc/pm_bfs.cu) and keeps them in
HBM: ``dist[l, i, j]`` uint16 with nodes ordered like ``walls.asList(False)``
(x-major, y-minor) and ``node_of_cell[l, y*W+x]`` as the index map.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Sequence

import numpy as np
import torch

from . import _lib
from . import layout as _layout
from .gymPacMan import pack_tables

UNREACHABLE = 0xFFFF


class DistanceTables:
    """Device-resident result of :func:`compute_distances`."""

    def __init__(self, W, H, n_open, node_of_cell, dist):
        self.W, self.H = W, H
        self.n_open = n_open                # int32 [L]
        self.node_of_cell = node_of_cell    # int16 [L, H*W]
        self.dist = dist                    # int16 storage viewed as uint16 [L, nmax, nmax]

    def lookup(self, layout_index: torch.Tensor, cell_a: torch.Tensor, cell_b: torch.Tensor) -> torch.Tensor:
        """Batched maze distance between cells (c = y*W + x); int32, 65535 = unreachable / wall."""
        l = layout_index.long()
        i = self.node_of_cell[l, cell_a.long()].long()
        j = self.node_of_cell[l, cell_b.long()].long()
        ok = (i >= 0) & (j >= 0)
        d = self.dist[l, i.clamp(min=0), j.clamp(min=0)].to(torch.int32) & 0xFFFF
        return torch.where(ok, d, torch.full_like(d, UNREACHABLE))

    def matrix(self, l: int = 0) -> np.ndarray:
        """[n_open, n_open] int32 in reference node order, -1 = unreachable."""
        n = int(self.n_open[l].item())
        m = (self.dist[l, :n, :n].to(torch.int32) & 0xFFFF).cpu().numpy()
        return np.where(m == UNREACHABLE, -1, m).astype(np.int32)


def compute_distances(layouts: Sequence[_layout.Layout], device=None, nmax: int | None = None, table: torch.Tensor | None = None,
                      out: DistanceTables | None = None) -> DistanceTables:
    """distanceCalculator.computeDistances for every layout in ``layouts`` (same W x H).  ``table``: the layouts' packed
    records already on the device (``gymPacMan.pack_tables`` -> int32 [L, LS]), e.g. an env's own table - skips the
    host-side packing and the upload.  ``out``: a DistanceTables of a previous call over boards of the same shape whose
    buffers are written again (then the call is one kernel launch)."""
    L = _lib.load()
    if not torch.cuda.is_available():
        raise RuntimeError("compute_distances needs a CUDA device; there is no CPU fallback")
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    W, H = layouts[0].width, layouts[0].height
    nl = len(layouts)
    if table is None:
        _, packed, _, max_open = pack_tables(layouts, with_max_open=True)
        table = torch.from_numpy(packed.view(np.int32)).to(dev)
        if nmax is None:
            nmax = max_open
    elif table.device != dev or table.dtype != torch.int32 or table.shape[0] != nl or not table.is_contiguous():
        raise ValueError("table must be the contiguous int32 [L, LS] device table of these layouts")
    if out is not None:
        if (out.W, out.H) != (W, H) or out.dist.shape[0] != nl or out.dist.device != dev:
            raise ValueError("out does not match these layouts")
        n_open, node_of, dist = out.n_open, out.node_of_cell, out.dist
        nmax = int(dist.shape[1])
    else:
        if nmax is None:
            nmax = max(int((~lay.walls.array).sum()) for lay in layouts)
        nmax = (int(nmax) + 7) & ~7
        n_open = torch.zeros(nl, dtype=torch.int32, device=dev)
        node_of = torch.empty((nl, W * H), dtype=torch.int16, device=dev)
        dist = torch.empty((nl, nmax, nmax), dtype=torch.int16, device=dev)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(L.pm_distances(W, H, table.data_ptr(), nl, n_open.data_ptr(), node_of.data_ptr(), dist.data_ptr(),
                                  nmax, st), "pm_distances")
    return out if out is not None else DistanceTables(W, H, n_open, node_of, dist)


def manhattanDistance(x, y):
    return abs(x[0] - y[0]) + abs(x[1] - y[1])


class Distancer:
    """`distanceCalculator.Distancer` (`:26-65`): same constructor and
    ``getMazeDistances`` / ``getDistance`` calls; the table comes from the GPU
    kernel and is mirrored to the host once for scalar look-ups by CPU bots."""

    def __init__(self, layout, default=10000):
        self.layout = layout
        self.default = default
        self._distances = None

    def getMazeDistances(self):
        t = compute_distances([self.layout])
        n = int(t.n_open[0].item())
        self._matrix = (t.dist[0, :n, :n].to(torch.int32) & 0xFFFF).cpu().numpy()
        self._node = t.node_of_cell[0].cpu().numpy()
        self._distances = True
        self.tables = t

    def isReadyForMazeDistance(self):
        return self._distances is not None

    def getDistanceOnGrid(self, pos1, pos2):
        W = self.layout.width
        try:
            i = self._node[int(pos1[1]) * W + int(pos1[0])]
            j = self._node[int(pos2[1]) * W + int(pos2[0])]
        except IndexError:
            i = j = -1
        if i < 0 or j < 0:
            raise Exception("Positions not in grid: " + str((pos1, pos2)))
        d = int(self._matrix[i, j])
        return sys.maxsize if d == UNREACHABLE else d

    def getDistance(self, pos1, pos2):
        if self._distances is None:
            return manhattanDistance(pos1, pos2)
        if _is_int(pos1) and _is_int(pos2):
            return self.getDistanceOnGrid(pos1, pos2)
        best = self.default
        for p1, d1 in _grids2d(pos1):
            for p2, d2 in _grids2d(pos2):
                d = self.getDistanceOnGrid(p1, p2) + d1 + d2
                if best > d:
                    best = d
        return best


def _is_int(pos):
    return pos[0] == int(pos[0]) and pos[1] == int(pos[1])


def _grids1d(v):
    iv = int(v)
    return [(v, 0)] if v == iv else [(iv, v - iv), (iv + 1, iv + 1 - v)]


def _grids2d(pos):
    return [((x, y), dx + dy) for x, dx in _grids1d(pos[0]) for y, dy in _grids1d(pos[1])]
