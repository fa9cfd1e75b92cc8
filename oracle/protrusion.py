"""Oracle: ProtrusionDetector live path (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

`peaks_raster`  restates ProtrusionDetector.py:38-57 (`_create_binary_image`), :59-158
(`_find_peak`, whole-image branch) and :419-442/:535 (`__call__`) literally with OpenCV.
`peaks_closed_form` is the integer closed form the CUDA kernel uses; the tests
check raster == closed form == reference on the fixtures and on random grids.
"""
from __future__ import annotations

import cv2
import numpy as np


def create_binary_image(grids: list, H: int, W: int, gs: int) -> np.ndarray:
    """ProtrusionDetector.py:38-57 - every non-empty cell painted with INCLUSIVE corners."""
    binary = np.zeros((H, W), dtype=np.uint8)
    for row in grids:
        for g in row:
            if g.empty:
                continue
            x, y = g.x, g.y
            corners = np.array([[x, y], [x + gs, y], [x + gs, y + gs], [x, y + gs]], np.int32)
            cv2.fillPoly(binary, [corners], 255)
    return cv2.threshold(binary, 127, 255, cv2.THRESH_BINARY)[1]


def peaks_raster(grids: list, H: int, W: int, gs: int) -> list[tuple[int, int]]:
    """ProtrusionDetector.py:59-158 + :535 - centres of the top-most runs."""
    binary = create_binary_image(grids, H, W, gs)
    ys, xs = np.where(binary == 255)
    if not ys.size:
        return []
    min_y = np.min(ys)
    peak_x = np.sort(xs[ys == min_y])
    gaps = np.diff(peak_x)
    splits = np.where(gaps > (gs // 4))[0] + 1
    out = []
    for group in np.split(peak_x, splits):
        out.append((int(group[len(group) // 2]), int(min_y)))
    return out


def peaks_closed_form(rows_y, occ: np.ndarray, x0: int, W: int, gs: int) -> list[tuple[int, int]]:
    """Grid-level closed form.  rows_y[k] = pixel y of list row k, occ[k, c] = non-empty.

    y* = min y over non-empty cells; on pixel row y* only cells with y == y* are painted
    (a cell paints y..y+gs inclusive).  Each painted cell covers x..min(x+gs, W-1); adjacent
    cells overlap in one pixel column, non-adjacent cells are > gs//4 apart, so the runs are
    the maximal runs of adjacent occupied columns in the UNION of the list rows with y == y*.
    centre = sorted_xs[len//2] = xa + (xb - xa + 1)//2."""
    rows_y = np.asarray(rows_y)
    has = occ.any(axis=1) if occ.size else np.zeros(0, bool)
    if not has.any():
        return []
    ytop = int(rows_y[has].min())
    union = occ[(rows_y == ytop) & has].any(axis=0)
    out = []
    c, C = 0, union.shape[0]
    while c < C:
        if not union[c]:
            c += 1
            continue
        c1 = c
        while c1 + 1 < C and union[c1 + 1]:
            c1 += 1
        xa = x0 + c * gs
        xb = min(x0 + c1 * gs + gs, W - 1)
        out.append((xa + (xb - xa + 1) // 2, ytop))
        c = c1 + 1
    return out
