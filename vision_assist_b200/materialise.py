"""FrameRecord (arrays from the GPU) <-> the reference's object view (list[list[Grid]], grid_lookup)."""
from __future__ import annotations

import math

import numpy as np

from . import models
from .engine import FrameRecord


def record_to_objects(rec: FrameRecord, gs: int):
    """-> (grids, grid_lookup, np_grids) exactly as FrameProcessor._extract_grid_information +
    _calculate_penalties leave them (FrameProcessor.py:50-182): list order, Grid.row attribute,
    duplicate rows, and lookup entries that outlived their list row ("orphans")."""
    Coordinate, Grid, _ = models.classes()
    grids, lookup = [], {}
    half = gs // 2

    def make_row(y, attr, occ_row, pen_row):
        row = []
        for c in range(rec.C):
            x = rec.x0 + c * gs
            filled = bool(occ_row[c] & 1)
            p = None
            if filled and pen_row is not None and not math.isnan(pen_row[c]):
                p = float(pen_row[c])
            g = Grid(coords=Coordinate(x=x, y=int(y)), centre=Coordinate(x=x + half, y=int(y) + half), penalty=p,
                     row=int(attr), col=c, empty=not filled, artificial=bool(occ_row[c] & 2))
            row.append(g)
            lookup[(x, int(y))] = g
        return row

    for k in range(rec.n_orphans):                  # only reachable through grid_lookup
        make_row(rec.orphan_y[k], -1, rec.orphan_occ[k], None)
    for k in range(rec.R):                          # later list rows override earlier ones in the lookup
        grids.append(make_row(rec.rows_y[k], rec.rows_attr[k], rec.occ[k], rec.penalty[k]))
    np_grids = rec.np_grids if rec.R else np.empty((0, 0), dtype=np.uint8)
    return grids, lookup, np_grids


def objects_to_grid_input(grids, grid_lookup, gs: int, use_easy: bool = True) -> dict:
    """list[list[Grid]] (+ grid_lookup) -> the array form va_grid_to_penalty_peaks takes.

    Requires what every FrameProcessor-built structure satisfies: all rows share the same columns,
    x = x0 + c*gs, y a multiple of gs, lookup rows complete."""
    R = len(grids)
    if R == 0:
        return dict(x0=0, rows_y=np.zeros(0, np.int32), rows_attr=np.zeros(0, np.int32),
                    occ=np.zeros((0, 0), np.uint8), use_easy=int(use_easy))
    Cc = len(grids[0])
    x0 = grids[0][0].coords.x
    rows_y = np.array([row[0].coords.y for row in grids], np.int32)
    rows_attr = np.array([row[0].row for row in grids], np.int32)
    occ = np.zeros((R, Cc), np.uint8)
    for k, row in enumerate(grids):
        if len(row) != Cc:
            raise ValueError("ragged grid rows are not supported")
        for c, g in enumerate(row):
            if g.coords.x != x0 + c * gs or g.coords.y != rows_y[k] or g.coords.y % gs:
                raise ValueError("grid is not on a regular gs lattice")
            occ[k, c] = (0 if g.empty else 1) | (2 if g.artificial else 0)
    out = dict(x0=int(x0), rows_y=rows_y, rows_attr=rows_attr, occ=occ, use_easy=int(use_easy))
    if grid_lookup is not None:
        ys = sorted({y for (_, y) in grid_lookup})
        pocc = np.zeros((len(ys), Cc), np.uint8)
        for k, y in enumerate(ys):
            for c in range(Cc):
                g = grid_lookup.get((x0 + c * gs, y))
                if g is None:
                    raise ValueError("grid_lookup rows must cover every column")
                pocc[k, c] = 0 if g.empty else 1
            if y % gs:
                raise ValueError("grid_lookup is not on a regular gs lattice")
        if any((x - x0) % gs or not (0 <= (x - x0) // gs < Cc) for (x, _) in grid_lookup):
            raise ValueError("grid_lookup has columns outside the grid")
        out["plane_y"] = np.array(ys, np.int32)
        out["plane_occ"] = pocc
    return out
