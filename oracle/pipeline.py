"""Oracle: whole hot path for one frame (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

protos/coefs/boxes -> masks (oracle.mask_assembly) -> polygon or direct raster -> grid
(oracle.grid) -> penalties (oracle.penalty) -> peaks (oracle.protrusion), returned as a plain
dict of numpy arrays (`FrameResult`) that tests compare field by field with the decoded CUDA
record.  Two routes from masks to the grid:

  route="contour" : the reference's own route - masks2segments (cv2.findContours, contour with
                    most points) -> scale_coords -> contourArea-largest polygon -> fillPoly.
  route="lut"     : the same result computed the way the CUDA path does (oracle.contour): hole-filled top-level
                    components, table-driven point counts and doubled areas, no contour tracing.  Must equal
                    "contour" on every input (tests/test_contour_model.py).
  route="direct"  : round-1 approximation kept for comparison - instance with the largest PIXEL AREA (first max),
                    its pixel bounding box and the mask itself as raster.  Equal to "contour"
                    for hole-free single-blob masks only.
"""
from __future__ import annotations

import cv2
import numpy as np
import torch

from . import contour as ocontour
from . import graph as ograph
from . import grid as ogrid
from . import mask_assembly as oma
from . import penalty as open_
from . import protrusion as oprot

FLAG_EMPTY = 1            # no grid produced (reference returns [] at FrameProcessor.py:328-332)
FLAG_CENTRE_OOB = 2       # reference raises IndexError at FrameProcessor.py:97
FLAG_LIST_OOB = 4         # reference raises IndexError at FrameProcessor.py:163 (negative index)
FLAG_NON_SIMPLE = 8       # selected mask is not one hole-free blob
FLAG_NO_POLYGON = 32      # the selected polygon has no points: reference raises cv2.error in fillPoly (FrameProcessor.py:86)


def state_to_result(st: ogrid.GridState | None, flags: int = 0, sel: int = -1) -> dict:
    """Flatten a GridState (+penalties, peaks) into arrays in LIST order."""
    if st is None or not st.grids:
        return dict(flags=flags | FLAG_EMPTY, sel=sel, x0=0, y0=0, C=0, R=0,
                    rows_y=np.zeros(0, np.int32), rows_attr=np.zeros(0, np.int32),
                    occ=np.zeros((0, 0), np.uint8), penalty=np.zeros((0, 0), np.float64),
                    peaks=np.zeros((0, 2), np.int32), orphan_y=np.zeros(0, np.int32),
                    orphan_occ=np.zeros((0, 0), np.uint8), start=(-1, -1), goals=np.zeros((0, 2), np.int32),
                    nbr=np.zeros((0, 0), np.uint8), lookup_row=np.zeros(0, np.int32))
    pen = open_.calculate_penalties(st)
    R, C = len(st.grids), len(st.grids[0])
    rows_y = np.array([row[0].y for row in st.grids], np.int32)
    rows_attr = np.array([row[0].row for row in st.grids], np.int32)
    occ = np.array([[(0 if g.empty else 1) | (2 if g.artificial else 0) for g in row] for row in st.grids],
                   np.uint8)
    peaks = oprot.peaks_closed_form(rows_y, (occ & 1).astype(bool), st.x0, st.W, st.gs)
    # lookup entries whose object is no longer in the list ("orphans", negative-index quirk)
    in_list = {id(g) for row in st.grids for g in row}
    orphan_rows: dict[int, list] = {}
    for (x, y), g in st.lookup.items():
        if id(g) not in in_list:
            orphan_rows.setdefault(y, []).append(g)
    oy = sorted(orphan_rows)
    oocc = np.zeros((len(oy), C), np.uint8)
    for k, y in enumerate(oy):
        for g in orphan_rows[y]:
            oocc[k, (g.x - st.x0) // st.gs] = (0 if g.empty else 1) | (2 if g.artificial else 0)
    # start / goal cells of _find_paths and the _create_graph neighbourhood
    start, goals = ograph.start_and_goals(st, peaks)
    return dict(flags=flags, sel=sel, x0=st.x0, y0=st.y0, C=C, R=R, rows_y=rows_y, rows_attr=rows_attr,
                occ=occ, penalty=pen, peaks=np.array(peaks, np.int32).reshape(-1, 2),
                orphan_y=np.array(oy, np.int32), orphan_occ=oocc,
                start=start if start is not None else (-1, -1),
                goals=np.array([g if g is not None else (-1, -1) for g in goals], np.int32).reshape(-1, 2),
                nbr=ograph.neighbour_mask(st), lookup_row=ograph.lookup_rows(st, oy))


def euler_number_8(mask: np.ndarray) -> int:
    """#components(8-conn) - #holes via bit-quad counts (Gray 1971): (Q1 - Q3 - 2*QD)/4."""
    m = np.pad(mask.astype(np.int32) > 0, 1).astype(np.int32)
    a, b, c, d = m[:-1, :-1], m[:-1, 1:], m[1:, :-1], m[1:, 1:]
    s = a + b + c + d
    q1 = int((s == 1).sum())
    q3 = int((s == 3).sum())
    qd = int(((s == 2) & (a == d)).sum())
    return (q1 - q3 - 2 * qd) // 4


def frame_from_masks(masks: np.ndarray, gs: int, route: str = "contour", frame_shape=None) -> dict:
    """masks uint8 [n, H, W] -> FrameResult."""
    n, H, W = masks.shape
    frame_shape = frame_shape or (H, W)
    flags, sel, st = 0, -1, None
    try:
        if route == "contour":
            polys = oma.masks_to_polygons(masks, frame_shape) if n else None
            st = ogrid.extract_grid_from_polygons(polys, frame_shape[0], frame_shape[1], gs)
        elif route == "lut":
            sel, poly = ocontour.select_instance(masks) if n else (-1, None)
            if n and poly is None:
                return state_to_result(None, flags | FLAG_NO_POLYGON, sel)
            if poly is not None:
                x0, y0, x1, y1 = poly["bbox"]
                if not poly["simple"]:
                    flags |= FLAG_NON_SIMPLE
                st = ogrid.extract_grid_from_raster(poly["raster"].astype(np.uint8), (x0, y0, x1 - x0 + 1, y1 - y0 + 1),
                                                    gs=gs)
        else:
            if n:
                areas = masks.reshape(n, -1).sum(axis=1, dtype=np.int64)
                sel = int(np.argmax(areas))          # first maximum
                if areas[sel] > 0:
                    if euler_number_8(masks[sel]) != 1:
                        flags |= FLAG_NON_SIMPLE
                    st = ogrid.extract_grid_direct(masks[sel], gs)
    except IndexError as e:
        flags |= FLAG_CENTRE_OOB if "centre" in str(e) else FLAG_LIST_OOB
        return state_to_result(None, flags, sel)
    except cv2.error:
        return state_to_result(None, flags | FLAG_NO_POLYGON, sel)
    return state_to_result(st, flags, sel)


def frame_from_tensors(protos, coefs, boxes, shape, gs: int, route: str = "contour") -> dict:
    """protos [K,mh,mw], coefs [n,K], boxes [n,4] (torch CPU fp32) -> FrameResult (+ 'masks')."""
    protos = torch.as_tensor(protos, dtype=torch.float32)
    coefs = torch.as_tensor(coefs, dtype=torch.float32)
    boxes = torch.as_tensor(boxes, dtype=torch.float32)
    if coefs.shape[0] == 0:
        masks = np.zeros((0, shape[0], shape[1]), np.uint8)
    else:
        masks = oma.process_mask(protos, coefs, boxes, shape).numpy().astype(np.uint8)
    res = frame_from_masks(masks, gs, route)
    res["masks"] = masks
    return res
