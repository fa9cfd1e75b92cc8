"""Deterministic polygon / grid generators shared by the oracle and parity tests."""
from __future__ import annotations

import numpy as np


def random_polygon(rng: np.random.Generator, H: int, W: int, kind: str | None = None) -> np.ndarray:
    """A simple (star-shaped) float32 polygon in frame pixels, like ultralytics masks.xy."""
    kind = kind or rng.choice(["blob", "blob", "tall", "bottom", "edge", "tiny"])
    if kind == "blob":
        cx, cy = rng.uniform(0.2 * W, 0.8 * W), rng.uniform(0.3 * H, 0.9 * H)
        rx, ry = rng.uniform(0.05 * W, 0.45 * W), rng.uniform(0.05 * H, 0.5 * H)
    elif kind == "tall":
        cx, cy = rng.uniform(0.3 * W, 0.7 * W), 0.55 * H
        rx, ry = rng.uniform(0.05 * W, 0.2 * W), 0.45 * H
    elif kind == "bottom":       # entirely inside / below the artificial band
        cx, cy = rng.uniform(0.2 * W, 0.8 * W), rng.uniform(0.9 * H, 0.97 * H)
        rx, ry = rng.uniform(0.05 * W, 0.3 * W), rng.uniform(0.01 * H, 0.06 * H)
    elif kind == "edge":         # touches the frame borders after clipping
        cx, cy = rng.choice([0.0, W - 1.0]), rng.uniform(0.2 * H, 0.9 * H)
        rx, ry = rng.uniform(0.1 * W, 0.4 * W), rng.uniform(0.1 * H, 0.5 * H)
    else:                        # tiny: a few pixels, often no cell centre inside
        cx, cy = rng.uniform(0.1 * W, 0.9 * W), rng.uniform(0.1 * H, 0.9 * H)
        rx, ry = rng.uniform(1, 14), rng.uniform(1, 14)
    k = int(rng.integers(5, 40))
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    rad = rng.uniform(0.55, 1.0, k)
    pts = np.stack([cx + rx * rad * np.cos(ang), cy + ry * rad * np.sin(ang)], 1)
    pts[:, 0] = pts[:, 0].clip(0, W - 1)
    pts[:, 1] = pts[:, 1].clip(0, H - 1)
    return pts.astype(np.float32)


def random_occupancy(rng: np.random.Generator, R: int, C: int, p: float | None = None) -> np.ndarray:
    """Random bool grid mixing blobs and salt noise (exercises easy and non-easy segments)."""
    p = rng.uniform(0.2, 0.9) if p is None else p
    g = rng.random((R, C)) < p
    if rng.random() < 0.5:       # carve a blob so that contiguous rows/cols exist too
        r0, r1 = sorted(rng.integers(0, R, 2)); c0, c1 = sorted(rng.integers(0, C, 2))
        g[r0:r1 + 1, c0:c1 + 1] = True
    return g
