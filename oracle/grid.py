"""Oracle: mask polygon -> occupancy grid (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Literal restatement of `FrameProcessor._extract_grid_information` (FrameProcessor.py:50-171),
with slotted cells instead of pydantic models (models.py:17-36).  The list / dict operations of
the reference are replayed as they are written, so the accidental behaviours that bit-exact
parity depends on are reproduced by construction:

  * float -> int32 truncation of polygon points and cv2.boundingRect         (:75-76)
  * bbox snapping, w computed from the un-snapped width                        (:79-83)
  * occupancy = POINT SAMPLE of the filled polygon at the cell centre          (:88-97)
  * IndexError when a centre falls outside the frame                           (:97)
  * artificial bottom band from int(H*0.875) rounded up                        (:126-165)
  * `if row_idx < len(grids)-1: replace else: append` - duplicate row when
    row_idx == len-1, gap compression when the mask ends above the band, and Python
    negative-index replacement when the band starts above the mask bbox        (:162-165)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import cv2
import numpy as np


class Cell:
    """models.py:29-36 `Grid` (coords/centre flattened to ints)."""

    __slots__ = ("x", "y", "row", "col", "empty", "artificial", "penalty")

    def __init__(self, x, y, row, col, empty, artificial):
        self.x = int(x)
        self.y = int(y)
        self.row = int(row)
        self.col = int(col)
        self.empty = bool(empty)
        self.artificial = bool(artificial)
        self.penalty = None


@dataclass
class GridState:
    H: int
    W: int
    gs: int
    grids: list = field(default_factory=list)        # list[list[Cell]]   (FrameProcessor.grids)
    lookup: dict = field(default_factory=dict)       # (x, y) -> Cell     (FrameProcessor.grid_lookup)
    np_grids: np.ndarray = field(default_factory=lambda: np.empty((0, 0), dtype=np.uint8))
    x0: int = 0
    y0: int = 0
    n_mask_rows: int = 0
    bbox: tuple = (0, 0, 0, 0)                       # snapped x, y, w, h


def select_polygon(polys: list):
    """FrameProcessor.py:71-73 - the polygon with the largest cv2.contourArea (first max)."""
    return max(polys, key=lambda p: cv2.contourArea(p)) if len(polys) > 1 else polys[0]


def extract_grid_from_polygons(polys, H: int, W: int, gs: int) -> GridState:
    """FrameProcessor.py:50-171 for one result (`polys` = result.masks.xy or None)."""
    st = GridState(H=H, W=W, gs=gs)
    if polys is None or len(polys) == 0:          # result.masks is None -> continue (:68-69)
        return st
    mask = select_polygon(polys)
    points = np.int32([mask])                      # :75 truncation toward zero
    x, y, w, h = cv2.boundingRect(points)          # :76
    mask_img = np.zeros((H, W), dtype=np.uint8)
    cv2.fillPoly(mask_img, points, 1)              # :85-86
    return extract_grid_from_raster(mask_img, (x, y, w, h), st)


def extract_grid_from_raster(mask_img: np.ndarray, rect, st: GridState | None = None,
                             gs: int | None = None) -> GridState:
    """FrameProcessor.py:78-171 given the filled raster and the un-snapped bounding rect."""
    H, W = mask_img.shape
    if st is None:
        st = GridState(H=H, W=W, gs=gs)
    gs = st.gs
    x, y, w, h = (int(v) for v in rect)
    art_xs = set(range((W // 2) - (gs * 8), (W // 2) + (gs * (8 + 1)), gs))   # :60-65

    x = x - (x % gs)                                                           # :79
    y = y - (y % gs)                                                           # :80
    w = w + (gs - w % gs) if w % gs != 0 else w                                # :81
    w = W if w > W else w                                                      # :82
    h = h + (gs - h % gs) if h % gs != 0 else h                                # :83
    st.x0, st.y0, st.bbox = x, y, (x, y, w, h)

    j_vals = np.arange(x, x + w, gs)                                           # :88
    i_vals = np.arange(y, y + h, gs)                                           # :89
    rows, cols = len(i_vals), len(j_vals)
    in_mask = np.zeros((rows, cols), dtype=bool)
    for r, i in enumerate(i_vals):                                             # :94-97
        for c, j in enumerate(j_vals):
            cy, cx = int(i + gs // 2), int(j + gs // 2)
            if cy >= H or cx >= W:
                raise IndexError("cell centre outside the frame (FrameProcessor.py:97)")
            in_mask[r, c] = mask_img[cy, cx] > 0
    if not np.any(in_mask):                                                    # :99-101
        return st
    st.n_mask_rows = rows

    for row_idx, i in enumerate(i_vals):                                       # :104-124
        this_row = []
        for col_idx, j in enumerate(j_vals):
            g = Cell(j, i, row_idx, col_idx, not in_mask[row_idx, col_idx], False)
            this_row.append(g)
            st.lookup[(int(j), int(i))] = g
        st.grids.append(this_row)

    starting_y = int(H * 0.875)                                                # :126
    starting_y = starting_y + (gs - starting_y % gs) % gs                      # :127
    for i in np.arange(starting_y, H, gs):                                     # :130-165
        i = int(i)
        row_idx = (i - y) // gs
        this_row = []
        for col_idx, j in enumerate(j_vals):
            j = int(j)
            this_grid = st.lookup.get((j, i))
            previously_empty = this_grid.empty if this_grid else True
            is_art_col = j in art_xs
            if previously_empty:
                empty, artificial = (not is_art_col), is_art_col
            else:
                empty, artificial = False, False
            g = Cell(j, i, row_idx, col_idx, empty, artificial)
            st.lookup[(j, i)] = g
            this_row.append(g)
        if row_idx < len(st.grids) - 1:
            st.grids[row_idx] = this_row          # Python semantics: negative wraps, may raise IndexError
        else:
            st.grids.append(this_row)

    st.np_grids = np.array([[0 if g.empty else 1 for g in row] for row in st.grids], dtype=np.uint8)  # :168-171
    return st


def extract_grid_direct(mask: np.ndarray, gs: int) -> GridState:
    """What the CUDA path computes: no contour step, the binary mask itself is the raster and its
    pixel bounding box is the rect.  Identical to the polygon route whenever the mask is one
    hole-free 8-connected blob (SURVEY 7 'Hard parts'; checked in tests/test_oracle_mask.py)."""
    ys, xs = np.nonzero(mask)
    H, W = mask.shape
    if ys.size == 0:
        return GridState(H=H, W=W, gs=gs)
    rect = (int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1))
    return extract_grid_from_raster(mask, rect, gs=gs)


def grid_from_npy(grid_filled: np.ndarray, gs: int = 20) -> GridState:
    """utilities/generate_testing_grids/run_on_main.py:45-145 `convert_npy_to_grid_info`
    (the fixture loader: band rule 0.8375*H, in-place replacement, full-frame columns)."""
    H, W = grid_filled.shape[0] * gs, grid_filled.shape[1] * gs
    st = GridState(H=H, W=W, gs=gs)
    art_xs = set(range((W // 2) - (gs * 8), (W // 2) + (gs * (8 + 1)), gs))
    for row_idx in range(grid_filled.shape[0]):
        this_row = []
        for col_idx in range(grid_filled.shape[1]):
            g = Cell(col_idx * gs, row_idx * gs, row_idx, col_idx, not grid_filled[row_idx, col_idx], False)
            this_row.append(g)
            st.lookup[(g.x, g.y)] = g
        st.grids.append(this_row)
    starting_y = int(H * 0.8375) + (gs - int(H * 0.8375) % gs)
    for i in range(starting_y, H, gs):
        row_count = i // gs
        this_row = []
        for j in range(0, W, gs):
            old = st.lookup.get((j, i))
            previously_empty = old.empty if old else True
            if previously_empty:
                empty, artificial = (j not in art_xs), (j in art_xs)
            else:
                empty, artificial = False, False
            g = Cell(j, i, row_count, j // gs, empty, artificial)
            st.lookup[(j, i)] = g
            this_row.append(g)
        if row_count < len(st.grids):
            st.grids[row_count] = this_row
        else:
            st.grids.append(this_row)
    st.n_mask_rows = grid_filled.shape[0]
    st.bbox = (0, 0, W, H)
    # NOTE run_on_main.py never sets np_grids (FrameProcessor.np_grids stays (0,0)), so the
    # reference computes fixture penalties by pure traversal; callers choose.
    st.np_grids = np.array([[0 if g.empty else 1 for g in row] for row in st.grids], dtype=np.uint8)
    return st
