"""Oracle: start / goal cell selection and the A* graph (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates, on the oracle's GridState,
  * `get_closest_grid_to_point` (utils.py:6-32): scan `grids` in LIST order (rows, then cells), skip empty
    cells, Euclidean distance from the point to the cell CENTRE (x + gs//2, y + gs//2; FrameProcessor.py:108,
    :148), strict `<` -> the FIRST minimum wins;
  * the start point of `_find_paths` (FrameProcessor.py:236): Coordinate(x=W // 2, y=H), and one goal per
    protrusion peak (:238-239);
  * `_create_graph` (FrameProcessor.py:184-207): for every non-empty cell of every LIST row, neighbours
    right, left, down, up (in that order) exist iff `grid_lookup.get((nx, ny))` is truthy - a pydantic
    model is always truthy, so the edge exists whenever the lookup HAS the key, also towards empty cells.

The distance is compared as np.sqrt of an exact Python int in the reference; sqrt is strictly increasing on
the integers that can occur (< 2**31), so comparing the squared integer distances selects the same cell.
"""
from __future__ import annotations

import numpy as np

from . import grid as ogrid


def closest_cell(st: ogrid.GridState, px: int, py: int):
    """utils.py:6-32 -> (list row k, column c) of the closest non-empty cell, or None."""
    best, best_d = None, None
    half = st.gs // 2
    for k, row in enumerate(st.grids):
        for c, g in enumerate(row):
            if g.empty:
                continue
            d = (px - (g.x + half)) ** 2 + (py - (g.y + half)) ** 2
            if best_d is None or d < best_d:
                best_d, best = d, (k, c)
    return best


def start_and_goals(st: ogrid.GridState, peaks) -> tuple:
    """FrameProcessor.py:236-239 -> ((k, c) of the start cell, [(k, c) per peak])."""
    start = closest_cell(st, st.W // 2, st.H)
    goals = [closest_cell(st, int(x), int(y)) for x, y in peaks]
    return start, goals


def neighbour_mask(st: ogrid.GridState) -> np.ndarray:
    """FrameProcessor.py:184-207 as a bit mask per LIST cell: bit0 right, bit1 left, bit2 down, bit3 up
    (0 for empty cells: they are not graph nodes)."""
    R, C = len(st.grids), len(st.grids[0]) if st.grids else 0
    out = np.zeros((R, C), np.uint8)
    gs = st.gs
    for k, row in enumerate(st.grids):
        for c, g in enumerate(row):
            if g.empty:
                continue
            m = 0
            for bit, (nx, ny) in enumerate(((g.x + gs, g.y), (g.x - gs, g.y), (g.x, g.y + gs), (g.x, g.y - gs))):
                if st.lookup.get((nx, ny)) is not None:
                    m |= 1 << bit
            out[k, c] = m
    return out


def lookup_rows(st: ogrid.GridState, orphan_y) -> np.ndarray:
    """For every pixel row y = ly * gs of the frame: index of the record row that owns grid_lookup at y
    (list position of the first list row holding that object, or R + j for the j-th orphan row), -1 if the
    lookup has no cell at y."""
    n = (st.H + st.gs - 1) // st.gs
    out = np.full(n, -1, np.int32)
    R = len(st.grids)
    first = {}
    for k, row in enumerate(st.grids):
        first.setdefault(id(row[0]), k)
    oy = [int(v) for v in orphan_y]
    for (x, y), g in st.lookup.items():
        if x != st.x0 or y % st.gs or not (0 <= y // st.gs < n):
            continue
        out[y // st.gs] = first[id(g)] if id(g) in first else R + oy.index(y)
    return out
