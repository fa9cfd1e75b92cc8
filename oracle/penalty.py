"""Oracle: PenaltyCalculator (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Literal restatement of PenaltyCalculator.py:26-142 and the driver FrameProcessor.py:173-182,
Python float64 arithmetic in the reference's exact operation order.
"""
from __future__ import annotations

import numpy as np

from .grid import GridState

# config.py:4-17 (BGR)
PENALTY_COLOUR_GRADIENT = {
    1.0000: (0, 0, 255), 0.9166: (0, 60, 255), 0.8333: (0, 88, 255), 0.7500: (0, 109, 255),
    0.6666: (0, 128, 255), 0.5833: (8, 145, 255), 0.5000: (0, 163, 249), 0.4166: (0, 183, 232),
    0.3333: (0, 202, 208), 0.1666: (0, 221, 176), 0.0833: (0, 239, 129), 0.0000: (0, 255, 15),
}


def pre_compute_easy_segments(np_grids: np.ndarray, grids: list):
    """PenaltyCalculator.py:26-55 -> (easy_rows, easy_cols): index -> ((x,y) first, (x,y) last)."""
    easy_rows, easy_cols = {}, {}
    if np_grids.ndim != 2:
        return easy_rows, easy_cols
    for row in range(np_grids.shape[0]):
        idx = np.where(np_grids[row, :] == 1)[0]
        if len(idx) > 0 and idx[-1] - idx[0] == len(idx) - 1:
            a, b = grids[row][idx[0]], grids[row][idx[-1]]
            easy_rows[row] = ((a.x, a.y), (b.x, b.y))
    for col in range(np_grids.shape[1]):
        idx = np.where(np_grids[:, col] == 1)[0]
        if len(idx) > 0 and idx[-1] - idx[0] == len(idx) - 1:
            a, b = grids[idx[0]][col], grids[idx[-1]][col]
            easy_cols[col] = ((a.x, a.y), (b.x, b.y))
    return easy_rows, easy_cols


def segment_penalty(cell, lookup: dict, direction: str, easy_rows: dict, easy_cols: dict, gs: int) -> float:
    """PenaltyCalculator.py:57-110."""
    sx, sy = cell.x, cell.y
    x, y = sx, sy
    if direction == "row" and cell.row in easy_rows:
        left, right = easy_rows[cell.row]
    elif direction == "col" and cell.col in easy_cols:
        left, right = easy_cols[cell.col]
    else:
        while True:
            nxt = (x - gs, y) if direction == "row" else (x, y - gs)
            if nxt not in lookup or lookup[nxt].empty:
                left = (x, y)
                break
            x, y = nxt
        x, y = sx, sy
        while True:
            nxt = (x + gs, y) if direction == "row" else (x, y + gs)
            if nxt not in lookup or lookup[nxt].empty:
                right = (x, y)
                break
            x, y = nxt
    den = right[0] - left[0] if direction == "row" else right[1] - left[1]
    if den == 0:
        ratio = 0.5
    else:
        ratio = (sx - left[0]) / den if direction == "row" else (sy - left[1]) / den
    return 2 * abs(ratio - 0.5)


def calculate_penalty(cell, lookup: dict, easy_rows: dict, easy_cols: dict, gs: int):
    """PenaltyCalculator.py:112-142."""
    if cell.empty:
        return 0
    rp = segment_penalty(cell, lookup, "row", easy_rows, easy_cols, gs)
    cp = segment_penalty(cell, lookup, "col", easy_rows, easy_cols, gs)
    if rp > 0.99 or cp > 0.99:
        return 1
    tot = rp + cp
    if tot == 0:
        return 0
    dom = abs(rp - cp) / tot
    rw = 0.5 + (0.25 * dom if rp > cp else -0.25 * dom)
    cw = 1 - rw
    return (rp * rw) + (cp * cw)


def calculate_penalties(st: GridState, use_easy: bool = True) -> np.ndarray:
    """FrameProcessor.py:173-182: score every non-empty cell of the LIST (duplicates included).

    Returns the penalty map [R, C] float64 in list order (NaN where the cell is empty).
    use_easy=False reproduces run_on_main.py's fixture path where np_grids is never set."""
    np_grids = st.np_grids if use_easy else np.empty((0, 0), dtype=np.uint8)
    easy_rows, easy_cols = pre_compute_easy_segments(np_grids, st.grids)
    R = len(st.grids)
    C = len(st.grids[0]) if R else 0
    out = np.full((R, C), np.nan, dtype=np.float64)
    for r, row in enumerate(st.grids):
        for c, cell in enumerate(row):
            if cell.empty:
                continue
            cell.penalty = calculate_penalty(cell, st.lookup, easy_rows, easy_cols, st.gs)
            out[r, c] = cell.penalty
    return out


def get_penalty_colour(p: float):
    """PenaltyCalculator.py:144-152 - nearest LUT key (first minimum in dict order)."""
    key = min(PENALTY_COLOUR_GRADIENT.keys(), key=lambda k: abs(k - p))
    return PENALTY_COLOUR_GRADIENT[key]
