"""Seeded random capture mazes — text-compatible with the reference generator.

The reference (`/root/reference/mazeGenerator.py:42-264`, entry point
`capture.randomLayout`, `capture.py:905-911`) builds a 18x9 half maze by
recursive wall splitting, mirrors it left-right, adds a border (-> 20x20), then
food, two capsules and the four starts.  Its output for a seed is defined by the
exact order of draws from CPython's global Mersenne Twister, so this module is
organised around that draw sequence:

  seed -> gauss(0.7, 0.2) -> randint(0,2), random()   [prison roll, result unused]
       -> per split: choice(range(1, n-1)), shuffle(slots)
       -> capsules: randint(row), randint(col) until 2 placed
       -> extra food: randint(row), randint(col) until 40 placed

Maze generation is a host-side layout source; the text it
returns feeds :class:`layout.Layout` and from there the packed device tables.
"""
from __future__ import annotations

import random
from typing import List

WALL, FOOD, CAPSULE, EMPTY = "%", ".", "o", " "

HALF_ROWS, HALF_COLS = 18, 9          # `mazeGenerator.py:259`
MAX_DIFFERENT_MAZES = 10000           # `mazeGenerator.py:253`


class _HalfMaze:
    """The left half under construction: one shared char grid plus a work list
    of rectangular rooms (top row, left col, rows, cols)."""

    def __init__(self, rows: int, cols: int):
        self.rows, self.cols = rows, cols
        self.grid: List[List[str]] = [[EMPTY] * cols for _ in range(rows)]

    # -- one wall through room (r0, c0, nr, nc) at offset `at` ---------------
    def split(self, r0, c0, nr, nc, at, gaps, vertical):
        """Returns the two sub-rooms, or None when the wall was not drawn
        (`mazeGenerator.py:85-118`)."""
        g = self.grid
        if vertical:
            gaps = min(nr, gaps)
            col = c0 + at
            span = list(range(r0, r0 + nr))
            # keep a doorway clear: drop an end slot that faces an open cell
            if 0 not in span:
                if g[span[0] - 1][col] == EMPTY:
                    span.pop(0)
                if len(span) <= gaps:
                    return None
            # (the reference tests the *column* count here; keep it)
            if (self.cols - 1) not in span and span[-1] + 1 < len(g):
                if g[span[-1] + 1][col] == EMPTY:
                    span.pop()
            if len(span) <= gaps:
                return None
            random.shuffle(span)
            for row in span[int(round(gaps)):]:
                g[row][col] = WALL
            return [(r0, c0, nr, at), (r0, c0 + at + 1, nr, nc - at - 1)]
        gaps = min(nc, gaps)
        row = r0 + at
        span = list(range(c0, c0 + nc))
        if 0 not in span:
            if g[row][span[0] - 1] == EMPTY:
                span.pop(0)
            if len(span) <= gaps:
                return None
        if (self.rows - 1) not in span and span[-1] + 1 < len(g[0]):
            if g[row][span[-1] + 1] == EMPTY:
                span.pop()
        if len(span) <= gaps:
            return None
        random.shuffle(span)
        for col in span[int(round(gaps)):]:
            g[row][col] = WALL
        return [(r0, c0, at, nc), (r0 + at + 1, c0, nr - at - 1, nc)]

    def carve(self, room, gaps, vertical, min_width, gapfactor):
        """Depth-first recursive splitting (`mazeGenerator.py:153-182`); the
        first child subtree is finished before the second starts, which fixes
        the order of RNG draws."""
        r0, c0, nr, nc = room
        if nr <= min_width and nc <= min_width:
            return
        extent = nc if vertical else nr
        if extent < min_width + 2:
            vertical = not vertical
            extent = nc if vertical else nr
        slots = range(1, extent - 1)
        if len(slots) == 0:
            return
        at = random.choice(slots)
        children = self.split(r0, c0, nr, nc, at, gaps, vertical)
        if children is None:
            return
        for child in children:
            self.carve(child, max(1, gaps * gapfactor), not vertical, min_width, gapfactor)

    def mirrored_with_border(self) -> List[List[str]]:
        """Left-right mirror + wall border (`mazeGenerator.py:57-75`)."""
        width = 2 * self.cols + 2
        full = [[WALL] * width]
        for row in self.grid:
            full.append([WALL] + row + row[::-1] + [WALL])
        full.append([WALL] * width)
        return full


def _decorate(grid: List[List[str]], max_food: int, max_capsules: int, toskip: int) -> None:
    """Dead-end food (two passes), starts, capsules, then random food up to
    ``max_food`` (`mazeGenerator.py:194-251`).  Mutates ``grid`` in place."""
    rows, cols = len(grid), len(grid[0])
    half = cols // 2
    total_food = 0
    for _depth in range(2):                      # max_depth = 2
        snapshot = [r[:] for r in grid]           # neighbours are read pre-pass
        added = 0
        for row in range(1, rows - 1):
            for col in range(1 + toskip, half - 1):
                if row > rows - 6 and col < 6:
                    continue
                if snapshot[row][col] != EMPTY:
                    continue
                open_nb = ((snapshot[row - 1][col] == EMPTY) + (snapshot[row][col - 1] == EMPTY)
                           + (snapshot[row + 1][col] == EMPTY) + (snapshot[row][col + 1] == EMPTY))
                if open_nb == 1:
                    grid[row][col] = FOOD
                    grid[row][cols - col - 1] = FOOD
                    added += 2
        total_food += added
        if added == 0:
            break

    grid[1][1], grid[2][1] = "3", "1"
    grid[1][cols - 2], grid[2][cols - 2] = "4", "2"

    placed = 0
    while placed < max_capsules:
        row = random.randint(1, rows - 1)
        col = random.randint(1 + toskip, half - 2)
        if row > rows - 6 and col < 6:
            continue
        if abs(col - cols / 2) < 3:
            continue
        if grid[row][col] == EMPTY:
            grid[row][col] = CAPSULE
            grid[row][cols - col - 1] = CAPSULE
            placed += 2

    while total_food < max_food:
        row = random.randint(1, rows - 1)
        col = random.randint(1 + toskip, half - 1)
        if row > rows - 6 and col < 6:
            continue
        if abs(col - cols // 2) < 3:
            continue
        if grid[row][col] == EMPTY:
            grid[row][col] = FOOD
            grid[row][cols - col - 1] = FOOD
            total_food += 2


def generateMaze(seed=None) -> str:
    """Layout text of the maze for ``seed`` (`mazeGenerator.py:255-264`)."""
    if not seed:
        seed = random.randint(1, MAX_DIFFERENT_MAZES)
    random.seed(seed)
    half = _HalfMaze(HALF_ROWS, HALF_COLS)
    gapfactor = random.gauss(0.7, 0.2)
    # "prison" roll: drawn, then overridden to 0 layers (`mazeGenerator.py:124-134`)
    random.randint(0, 2)
    random.random()
    half.carve((0, 0, HALF_ROWS, HALF_COLS), 5, True, 0, gapfactor)
    full = half.mirrored_with_border()
    rows, cols = len(full), len(full[0])
    _decorate(full, 2 * (rows * cols // 20), 2, 0)
    return "\n".join("".join(r) for r in full)


def randomLayout(seed=None) -> str:
    """`capture.randomLayout` (`capture.py:905-911`)."""
    if not seed:
        seed = random.randint(0, 99999999)
    return generateMaze(seed)


if __name__ == "__main__":
    import sys
    print(generateMaze(int(sys.argv[1]) if len(sys.argv) > 1 else None))
