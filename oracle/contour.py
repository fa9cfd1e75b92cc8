"""Oracle: the table-driven model of the reference's contour step (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

The reference turns every instance mask into ONE polygon and picks one polygon per frame:

    masks2segments   vendored ultralytics ops.py:837-859   cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE),
                                                           keep the contour with the MOST POINTS (np.argmax: first
                                                           maximum in OpenCV's output order)
    scale_coords     ops.py:784-816                        identity when the mask has the frame's shape
    selection        FrameProcessor.py:72-73               polygon with the largest cv2.contourArea (first maximum)
    raster           FrameProcessor.py:75-86               int32 truncation, cv2.boundingRect, cv2.fillPoly

`oracle.mask_assembly.masks2segments` + `oracle.grid.extract_grid_from_polygons` call OpenCV for all of that (route
"contour").  This module restates the same result WITHOUT tracing contours, the way the CUDA kernel computes it
(vision_assist_b200/csrc/va_contour_core.h, run by the tail kernel), so that every intermediate (points, doubled area, bbox, filled raster) can be
compared on the CPU, and so that the generated table is pinned against OpenCV itself (tests/test_contour_model.py):

  * G = complement of the 4-connected background region touching the frame; its 8-connected components are the
    top-level components with holes (and islands inside holes) filled - RETR_EXTERNAL returns exactly one contour
    per component of G, in REVERSE raster order of their first pixels;
  * points / doubled shoelace area of a component are sums of the 3x3 table of scripts/gen_contour_lut.py;
  * cv2.fillPoly of the kept contour == that component of G; cv2.boundingRect == its pixel bounding box.
"""
from __future__ import annotations

import os
import re

import numpy as np
from scipy import ndimage as ndi

_HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vision_assist_b200", "csrc",
                    "va_contour_lut.h")


def load_table() -> np.ndarray:
    """The generated table (product data, vision_assist_b200/csrc/va_contour_lut.h) as uint16[256]."""
    txt = open(_HDR).read()
    body = txt[txt.index("#define VA_CONTOUR_LUT_VALUES"):txt.index("static const uint16_t kContourLutHost")]
    vals = [int(v, 16) for v in re.findall(r"0x([0-9a-fA-F]+)", body)]
    assert len(vals) == 256
    return np.array(vals, np.uint16)


_T = None


def _tables():
    global _T
    if _T is None:
        t = load_table().astype(np.int64)
        _T = (t & 7, ((t >> 3) & 7) - 2, ((t >> 6) & 7) - 2)
    return _T


def filled_components(mask: np.ndarray):
    """-> (labels int32 [H,W] of G's 8-connected components (0 = outer background), count)."""
    m = np.pad(mask != 0, 1)
    bg, _ = ndi.label(~m)                                   # 4-connectivity
    G = bg != bg[0, 0]
    lab, n = ndi.label(G, structure=np.ones((3, 3), bool))
    return lab[1:-1, 1:-1].astype(np.int32), n


def component_sums(lab: np.ndarray, n: int):
    """Per component 1..n: (points, doubled signed area, first raster index, bbox x0,y0,x1,y1)."""
    H, W = lab.shape
    G = np.pad(lab > 0, 1).astype(np.int64)
    code = (G[:-2, :-2] | (G[:-2, 1:-1] << 1) | (G[:-2, 2:] << 2) | (G[1:-1, :-2] << 3) | (G[1:-1, 2:] << 4)
            | (G[2:, :-2] << 5) | (G[2:, 1:-1] << 6) | (G[2:, 2:] << 7))
    pts_t, dx_t, dy_t = _tables()
    ys, xs = np.nonzero(lab)
    l = lab[ys, xs]
    c = code[ys, xs]
    pts = np.bincount(l, weights=pts_t[c], minlength=n + 1).astype(np.int64)
    a2 = np.bincount(l, weights=xs * dy_t[c] - ys * dx_t[c], minlength=n + 1).astype(np.int64)
    first = np.full(n + 1, H * W, np.int64)
    np.minimum.at(first, l, ys * W + xs)
    out = []
    for k in range(1, n + 1):
        sel = l == k
        out.append(dict(points=int(pts[k]), area2=int(abs(a2[k])), first=int(first[k]),
                        bbox=(int(xs[sel].min()), int(ys[sel].min()), int(xs[sel].max()), int(ys[sel].max()))))
    return out


def instance_polygon(mask: np.ndarray):
    """What masks2segments keeps of one instance mask: None for an empty mask, else dict(points, area2, bbox,
    raster bool[H,W], n_components, simple) of the component whose contour has the most points (ties: the one OpenCV
    lists first = the LAST in raster order)."""
    lab, n = filled_components(mask)
    if n == 0:
        return None
    comps = component_sums(lab, n)
    best = max(range(n), key=lambda k: (comps[k]["points"], comps[k]["first"]))
    c = dict(comps[best])
    c["raster"] = lab == best + 1
    c["n_components"] = n
    c["simple"] = bool(n == 1 and np.array_equal(c["raster"], mask != 0))
    return c


def select_instance(masks: np.ndarray):
    """FrameProcessor.py:72-73 on the kept polygons: (sel, polygon dict or None).  `max(..., key=contourArea)` keeps
    the FIRST maximum; an empty polygon has area 0; with one instance no area is computed."""
    polys = [instance_polygon(m) for m in masks]
    if not polys:
        return -1, None
    areas = [p["area2"] if p else 0 for p in polys]
    sel = int(np.argmax(areas)) if len(polys) > 1 else 0
    return sel, polys[sel]
