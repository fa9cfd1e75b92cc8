"""Generate vision_assist_b200/csrc/va_contour_lut.h: the per-pixel table behind the GPU replacement of
cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + cv2.contourArea (reference: masks2segments, vendored
ultralytics ops.py:837-859, and FrameProcessor.py:72-73).

Fact used (established by this script, and pinned against OpenCV itself by tests/test_contour_model.py): let G be the
complement of the 4-connected background region that touches the image frame (= every top-level 8-connected
component with its holes filled).  OpenCV's border following (Suzuki & Abe, 8-connectivity) visits a pixel p of G
once per (incoming move, outgoing move) pair, and the SET of those pairs is a function of the 3x3 neighbourhood of p
in G alone.  So per top-level component

    number of CHAIN_APPROX_SIMPLE points = sum_p  #{visits of p with incoming != outgoing}     (1 for an isolated pixel)
    2 * contourArea (shoelace)           = | sum_p  x_p * sum(dy of outgoing moves) - y_p * sum(dx of outgoing moves) |

are sums of table look-ups over the component's pixels - no sequential tracing.  The script replays the border
following on random images (plain Python, no OpenCV), records the visit set of every pixel keyed by its 3x3 code,
asserts that the same code never yields two different sets and that all 256 codes occur, and writes the table.

Neighbourhood code bits: 0 NW, 1 N, 2 NE, 3 W, 4 E, 5 SW, 6 S, 7 SE.
Table entry: bits 0-2 points, bits 3-5 sum(dx)+2, bits 6-8 sum(dy)+2, bits 9-11 number of visits (= moves).
"""
from __future__ import annotations

import os
import sys

import numpy as np

DX = [1, 1, 0, -1, -1, -1, 0, 1]       # OpenCV chain codes: 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S, 7 SE (y down)
DY = [0, -1, -1, -1, 0, 1, 1, 1]
NB = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]


def outer_complement(img: np.ndarray) -> np.ndarray:
    """Zero-padded G: 1 where the pixel is NOT in the 4-connected background region touching the frame."""
    h, w = img.shape
    pad = np.zeros((h + 2, w + 2), np.uint8)
    pad[1:-1, 1:-1] = img != 0
    outer = np.zeros_like(pad)
    stack = [(0, 0)]
    outer[0, 0] = 1
    while stack:
        y, x = stack.pop()
        for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1)):
            yy, xx = y + dy, x + dx
            if 0 <= yy < h + 2 and 0 <= xx < w + 2 and not outer[yy, xx] and not pad[yy, xx]:
                outer[yy, xx] = 1
                stack.append((yy, xx))
    return (1 - outer).astype(np.uint8)


def trace(img: np.ndarray, x0: int, y0: int):
    """Outer border following from the raster-first pixel (x0, y0) of a component of the zero-padded image, as
    OpenCV's contour tracer does it: list of (x, y, incoming move, outgoing move); None for an isolated pixel."""
    s = s_end = 4
    while True:
        s = (s - 1) & 7
        if img[y0 + DY[s], x0 + DX[s]] or s == s_end:
            break
    if not img[y0 + DY[s], x0 + DX[s]]:
        return None
    x1, y1 = x0 + DX[s], y0 + DY[s]
    prev = s ^ 4
    x3, y3 = x0, y0
    out = []
    while True:
        while True:
            s = (s + 1) & 7
            if img[y3 + DY[s], x3 + DX[s]]:
                break
        out.append((x3, y3, prev, s))
        prev = s
        x4, y4 = x3 + DX[s], y3 + DY[s]
        if (x4, y4) == (x0, y0) and (x3, y3) == (x1, y1):
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return out


def derive_table(n_images: int = 4000, seed: int = 20260) -> list[int]:
    rng = np.random.default_rng(seed)
    seen: dict[int, tuple] = {}
    for it in range(n_images):
        h, w = int(rng.integers(1, 24)), int(rng.integers(1, 24))
        img = (rng.random((h, w)) < rng.uniform(0.25, 0.85)).astype(np.uint8)
        G = outer_complement(img)
        visits: dict[tuple, list] = {}
        done = np.zeros_like(G)
        for y in range(1, h + 1):
            for x in range(1, w + 1):
                if G[y, x] and not done[y, x]:
                    # raster-first pixel of a new component: trace it, then mark the component done (flood, 8-conn)
                    st = trace(G, x, y)
                    for (sx, sy, di, do) in st or []:
                        visits.setdefault((sx, sy), []).append((di, do))
                    stack = [(y, x)]
                    done[y, x] = 1
                    while stack:
                        cy, cx = stack.pop()
                        for dx, dy in NB:
                            if G[cy + dy, cx + dx] and not done[cy + dy, cx + dx]:
                                done[cy + dy, cx + dx] = 1
                                stack.append((cy + dy, cx + dx))
        for y in range(1, h + 1):
            for x in range(1, w + 1):
                if not G[y, x]:
                    continue
                code = sum(1 << k for k, (dx, dy) in enumerate(NB) if G[y + dy, x + dx])
                v = tuple(sorted(visits.get((x, y), [])))
                if seen.setdefault(code, v) != v:
                    raise AssertionError(f"3x3 code {code:#x} is not a function: {seen[code]} vs {v}")
    if len(seen) != 256:
        raise AssertionError(f"only {len(seen)} of 256 codes seen")
    table = []
    for code in range(256):
        v = seen[code]
        pts = sum(1 for di, do in v if di != do) if code else 1      # isolated pixel: one point, no move
        dxs = sum(DX[do] for _, do in v)
        dys = sum(DY[do] for _, do in v)
        assert 0 <= pts <= 4 and abs(dxs) <= 2 and abs(dys) <= 2 and len(v) <= 4
        table.append(pts | ((dxs + 2) << 3) | ((dys + 2) << 6) | (len(v) << 9))
    return table


def main() -> None:
    table = derive_table()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "vision_assist_b200", "csrc", "va_contour_lut.h")
    with open(path, "w") as f:
        f.write("// GENERATED by scripts/gen_contour_lut.py - do not edit.\n")
        f.write("// Per-pixel border-following table indexed by the 3x3 neighbourhood code of a pixel in the hole-filled\n")
        f.write("// image (bits: 0 NW, 1 N, 2 NE, 3 W, 4 E, 5 SW, 6 S, 7 SE).  Entry: bits 0-2 CHAIN_APPROX_SIMPLE points,\n")
        f.write("// bits 3-5 sum(dx of outgoing moves)+2, bits 6-8 sum(dy)+2, bits 9-11 moves.  See the script for the proof.\n")
        f.write("#pragma once\n#include <stdint.h>\n\n#define VA_CONTOUR_LUT_VALUES \\\n")
        for r in range(16):
            f.write("    " + ", ".join(f"0x{v:03x}" for v in table[16 * r:16 * r + 16]) + (", \\\n" if r < 15 else "\n"))
        f.write("\nstatic const uint16_t kContourLutHost[256] = {VA_CONTOUR_LUT_VALUES};\n")
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
