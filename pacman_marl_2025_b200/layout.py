"""`.lay` layout loader — host side of the batched capture environment.

Keeps the text format and the attribute surface of the reference loader
(`/root/reference/layout.py:22-149`) so callers that read ``env.layout.width``,
``layout.walls[x][y]``, ``layout.walls.asList(False)``, ``layout.capsules`` or
``layout.agentPositions`` keep working, but stores the grids as numpy arrays so
they can be bit-packed for the device (see :mod:`packing`).

Format (reference `layout.py:95-130`):
  ``%`` wall, ``.`` food, ``o`` capsule, ``1``-``4`` agent starts (agent index =
  digit - 1 after sorting), ``P``/``G`` generic pacman/ghost starts, anything
  else empty.  Text row ``r`` is ``y = H-1-r`` (origin bottom-left), column is x.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import numpy as np

__all__ = ["BoolGrid", "Layout", "getLayout", "tryToLoad", "LAYOUT_DIR"]

LAYOUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "layouts")


class BoolGrid:
    """Boolean W x H grid addressed ``g[x][y]`` (reference `game.py:162-230`).

    Backed by ``self.array`` (numpy bool, shape [W, H]); ``self.data`` gives the
    list-of-lists view the reference's consumers index into.
    """

    def __init__(self, width: int, height: int, array: np.ndarray | None = None):
        self.width = int(width)
        self.height = int(height)
        if array is None:
            array = np.zeros((self.width, self.height), dtype=bool)
        self.array = np.asarray(array, dtype=bool)
        assert self.array.shape == (self.width, self.height)

    # --- reference-compatible surface -------------------------------------
    @property
    def data(self) -> List[List[bool]]:
        return self.array.tolist()

    def __getitem__(self, x):
        return self.array[x]

    def __eq__(self, other):
        return isinstance(other, BoolGrid) and np.array_equal(self.array, other.array)

    def __hash__(self):
        return hash(self.array.tobytes())

    def copy(self) -> "BoolGrid":
        return BoolGrid(self.width, self.height, self.array.copy())

    deepCopy = copy

    def count(self, item: bool = True) -> int:
        n = int(self.array.sum())
        return n if item else self.width * self.height - n

    def asList(self, key: bool = True) -> List[Tuple[int, int]]:
        """Cells equal to ``key`` in x-major, y-minor order (`game.py:225-230`)."""
        xs, ys = np.nonzero(self.array == bool(key))  # C order == x-major, y-minor
        return list(zip(xs.tolist(), ys.tolist()))

    def __str__(self):
        rows = []
        for y in range(self.height - 1, -1, -1):
            rows.append("".join("T" if self.array[x, y] else "F" for x in range(self.width)))
        return "\n".join(rows)


class Layout:
    """Static board description parsed from layout text (`layout.py:22-37`)."""

    def __init__(self, layoutText: Sequence[str]):
        layoutText = list(layoutText)
        self.width = len(layoutText[0])
        self.height = len(layoutText)
        self.walls = BoolGrid(self.width, self.height)
        self.food = BoolGrid(self.width, self.height)
        self.capsules: List[Tuple[int, int]] = []
        self.agentPositions: List[Tuple[bool, Tuple[int, int]]] = []
        self.numGhosts = 0
        self._parse(layoutText)
        self.layoutText = layoutText
        self.totalFood = self.food.count()

    def _parse(self, text: Sequence[str]) -> None:
        tagged = []
        top = self.height - 1
        # y outer / x inner, like the reference, so capsule order matches.
        for y in range(self.height):
            row = text[top - y]
            for x in range(self.width):
                ch = row[x]
                if ch == "%":
                    self.walls.array[x, y] = True
                elif ch == ".":
                    self.food.array[x, y] = True
                elif ch == "o":
                    self.capsules.append((x, y))
                elif ch == "P":
                    tagged.append((0, (x, y)))
                elif ch == "G":
                    tagged.append((1, (x, y)))
                    self.numGhosts += 1
                elif ch in "1234":
                    tagged.append((int(ch), (x, y)))
                    self.numGhosts += 1
        tagged.sort()
        self.agentPositions = [(tag == 0, pos) for tag, pos in tagged]

    # --- reference-compatible helpers --------------------------------------
    def getNumGhosts(self) -> int:
        return self.numGhosts

    def isWall(self, pos) -> bool:
        x, y = pos
        return bool(self.walls.array[x, y])

    def __str__(self):
        return "\n".join(self.layoutText)

    def deepCopy(self) -> "Layout":
        return Layout(self.layoutText[:])

    @property
    def starts(self) -> List[Tuple[int, int]]:
        return [pos for _, pos in self.agentPositions]


def tryToLoad(fullname: str):
    """Read one `.lay` file, stripping each line (`layout.py:145-149`)."""
    if not os.path.exists(fullname):
        return None
    with open(fullname) as f:
        return Layout([line.strip() for line in f])


def _candidates(name: str) -> Iterable[str]:
    if name.endswith(".lay"):
        yield os.path.join("layouts", name)
        yield name
    else:
        yield os.path.join("layouts", name + ".lay")
        yield name + ".lay"


def getLayout(name: str, back: int = 2):
    """Search order of the reference (`layout.py:131-143`): ``layouts/<name>``
    then ``<name>`` relative to the cwd and up to ``back`` parent directories;
    as a last resort the layouts bundled with this package."""
    here = os.path.abspath(".")
    for _ in range(back + 1):
        for cand in _candidates(name):
            lay = tryToLoad(os.path.join(here, cand) if not os.path.isabs(cand) else cand)
            if lay is not None:
                return lay
        here = os.path.dirname(here)
    base = os.path.basename(name)
    bundled = os.path.join(LAYOUT_DIR, base if base.endswith(".lay") else base + ".lay")
    return tryToLoad(bundled)
