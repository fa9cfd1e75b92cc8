"""The table-driven contour model (oracle/contour.py = what va_contour_core.h computes) against OpenCV itself.

Pins scripts/gen_contour_lut.py's table and the three structural facts the CUDA path relies on:
  (1) RETR_EXTERNAL contours <-> 8-connected components of the hole-filled image, in reverse raster order;
  (2) CHAIN_APPROX_SIMPLE point counts and cv2.contourArea are sums of the 3x3 table over a component's pixels;
  (3) cv2.fillPoly of the kept contour is that component, cv2.boundingRect its pixel bounding box.
Reference: masks2segments (vendored ultralytics ops.py:837-859), FrameProcessor.py:72-86.
"""
import os
import sys

import cv2
import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import contour as ocontour  # noqa: E402
from oracle import mask_assembly as oma  # noqa: E402
from oracle import pipeline as opl  # noqa: E402


def _rand_mask(rng, h, w, kind):
    if kind == 0:
        return (rng.random((h, w)) < rng.uniform(0.2, 0.85)).astype(np.uint8)
    z = rng.standard_normal((h // 4 + 2, w // 4 + 2)).astype(np.float32)
    z = cv2.resize(z, (w, h), interpolation=cv2.INTER_CUBIC)
    return (z > rng.uniform(-0.6, 0.6)).astype(np.uint8)


def test_table_matches_generator():
    from scripts import gen_contour_lut
    assert list(ocontour.load_table()) == gen_contour_lut.derive_table(n_images=1500)


def test_components_points_area_vs_opencv():
    rng = np.random.default_rng(7)
    for it in range(600):
        h, w = int(rng.integers(1, 48)), int(rng.integers(1, 48))
        m = _rand_mask(rng, h, w, it % 2)
        cs, _ = cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        lab, n = ocontour.filled_components(m)
        assert len(cs) == n
        comps = ocontour.component_sums(lab, n)
        # OpenCV lists the contours in reverse raster order of their first pixel
        order = sorted(range(n), key=lambda k: -comps[k]["first"])
        for c, k in zip(cs, order):
            pts = c.reshape(-1, 2)
            assert int(pts[0, 1]) * w + int(pts[0, 0]) == comps[k]["first"]
            assert len(pts) == comps[k]["points"]
            assert int(round(2 * cv2.contourArea(pts.astype(np.float32)))) == comps[k]["area2"]
            x, y, bw, bh = cv2.boundingRect(np.int32([pts.astype(np.float32)]))
            assert (x, y, x + bw - 1, y + bh - 1) == comps[k]["bbox"]
            fill = np.zeros((h, w), np.uint8)
            cv2.fillPoly(fill, np.int32([pts.astype(np.float32)]), 1)
            assert np.array_equal(fill.astype(bool), lab == k + 1)


def _results_equal(a, b):
    for key in ("flags", "x0", "y0", "C", "R"):
        va, vb = a[key], b[key]
        if key == "flags":
            va, vb = va & ~opl.FLAG_NON_SIMPLE, vb & ~opl.FLAG_NON_SIMPLE
        assert va == vb, key
    for key in ("rows_y", "rows_attr", "occ", "peaks", "orphan_y", "orphan_occ", "goals", "lookup_row"):
        assert np.array_equal(a[key], b[key]), key
    assert np.array_equal(np.isnan(a["penalty"]), np.isnan(b["penalty"]))
    assert np.array_equal(np.nan_to_num(a["penalty"]).view(np.uint64), np.nan_to_num(b["penalty"]).view(np.uint64))
    assert tuple(a["start"]) == tuple(b["start"])


@pytest.mark.parametrize("gs", [20, 8])
def test_lut_route_equals_contour_route(gs):
    """Whole frames: several random instances (blobs, noise, empty masks, single pixels, rings, near-ties)."""
    rng = np.random.default_rng(11 + gs)
    H = W = 160
    n_nonsimple = 0
    for it in range(120):
        n = int(rng.integers(1, 6))
        masks = np.zeros((n, H, W), np.uint8)
        for i in range(n):
            kind = int(rng.integers(0, 6))
            if kind == 0:
                pass                                             # empty mask: zeros((0,2)) polygon, area 0
            elif kind == 1:
                masks[i, rng.integers(0, H), rng.integers(0, W)] = 1     # single pixel
            elif kind == 2:                                      # ring (hole) with an island inside
                cv2.circle(masks[i], (int(rng.integers(40, 120)), int(rng.integers(40, 120))), int(rng.integers(15, 38)), 1, int(rng.integers(1, 6)))
                if rng.random() < 0.5:
                    c = np.argwhere(masks[i])
                    cy, cx = c.mean(0).astype(int)
                    masks[i, cy - 2:cy + 3, cx - 2:cx + 3] = 1
            elif kind == 3:                                      # two blobs of nearly equal size
                a = int(rng.integers(10, 30))
                masks[i, 20:20 + a, 10:10 + a] = 1
                masks[i, 90:90 + a + int(rng.integers(-1, 2)), 80:80 + a + int(rng.integers(-1, 2))] = 1
            else:
                masks[i] = _rand_mask(rng, H, W, 1)
        want = opl.frame_from_masks(masks, gs, "contour")
        got = opl.frame_from_masks(masks, gs, "lut")
        _results_equal(got, want)
        n_nonsimple += bool(got["flags"] & opl.FLAG_NON_SIMPLE)
    assert n_nonsimple > 20          # the inputs do exercise the non-simple branch


def test_lut_route_on_synthetic_noise_family():
    from vision_assist_b200 import synth
    H = W = 640
    for f in range(3):
        p, c, b = synth.make_frame(1000 + f, 4, H, W, 160, 160, family="noise")
        masks = oma.process_mask(p, c, b, (H, W)).numpy().astype(np.uint8)
        _results_equal(opl.frame_from_masks(masks, 20, "lut"), opl.frame_from_masks(masks, 20, "contour"))
