"""va_contour_core.h (the algorithm the CUDA contour kernels run) built for the host and compared with the oracle
model (oracle/contour.py, itself pinned against OpenCV in tests/test_contour_model.py): certificate path and general
path, u8 and bit-row inputs, several emulated thread counts.  No GPU needed; the kernel launch plumbing on top of
these phase functions is covered by the -m gpu tests."""
import ctypes
import os
import subprocess
import sys

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import contour as ocontour  # noqa: E402


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("native") / "libcontour_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(out),
                    os.path.join(ROOT, "tests", "native", "contour_host.cpp")], check=True)
    lib = ctypes.CDLL(str(out))
    lib.contour_host_instance.restype = ctypes.c_int
    return lib


def run(lib, mask, gs=20, fmt=0, force_general=0, nthreads=64, cap=None):
    H, W = mask.shape
    half = gs // 2
    lr, lc = -(-(H - half) // gs), -(-(W - half) // gs)
    lw = (lc + 31) // 32
    out = np.zeros(9, np.int32)
    lat = np.zeros((lr, lw), np.uint32)
    m = np.ascontiguousarray(mask, np.uint8)
    cap = cap if cap is not None else H * ((W + 1) // 2)
    lib.contour_host_instance(m.ctypes.data_as(ctypes.c_void_p), H, W, gs, fmt, force_general, nthreads, cap,
                              out.ctypes.data_as(ctypes.c_void_p), lat.ctypes.data_as(ctypes.c_void_p))
    bits = np.zeros((lr, lc), bool)
    for x in range(lc):
        bits[:, x] = (lat[:, x >> 5] >> np.uint32(x & 31)) & 1
    return dict(state=int(out[0]), area2=int(out[1]), bbox=tuple(int(v) for v in out[2:6]), points=int(out[6]),
                n_components=int(out[7]), general=bool(out[8]), light=int(out[8]) == 2, lattice=bits)


def check(lib, mask, gs, **kw):
    got = run(lib, mask, gs, **kw)
    want = ocontour.instance_polygon(mask)
    if want is None:
        assert got["state"] == 0
        return got
    assert got["state"] in (1, 2), got
    assert got["area2"] == want["area2"], (got, want["area2"])
    assert got["bbox"] == want["bbox"]
    half = gs // 2
    H, W = mask.shape
    assert np.array_equal(got["lattice"], want["raster"][half::gs, half::gs][:got["lattice"].shape[0], :got["lattice"].shape[1]])
    assert (got["state"] == 1) == want["simple"]
    if got["general"]:
        assert got["points"] == want["points"] and got["n_components"] == want["n_components"]
    return got


def _rand_mask(rng, h, w, kind):
    if kind == 0:
        return (rng.random((h, w)) < rng.uniform(0.15, 0.9)).astype(np.uint8)
    z = rng.standard_normal((h // 4 + 2, w // 4 + 2)).astype(np.float32)
    z = cv2.resize(z, (w, h), interpolation=cv2.INTER_CUBIC)
    return (z > rng.uniform(-0.6, 0.6)).astype(np.uint8)


def test_random_masks_general_and_certificate(lib):
    rng = np.random.default_rng(3)
    n_general = n_cert = 0
    for it in range(500):
        h, w = int(rng.integers(12, 90)), int(rng.integers(12, 300))
        m = _rand_mask(rng, h, w, it % 3 != 0)
        gs = int(rng.choice([4, 8, 20]))
        g = check(lib, m, gs, fmt=it % 2, nthreads=int(rng.choice([1, 7, 64, 256])))
        n_general += g["general"]
        n_cert += (not g["general"]) and g["state"] == 1
        check(lib, m, gs, fmt=(it + 1) % 2, force_general=1, nthreads=33)     # the light / general path must agree on simple masks too
        g2 = check(lib, m, gs, fmt=it % 2, force_general=2, nthreads=64)      # ... and the full path where the light one applies
        assert not g2["light"]
    assert n_general > 100


def test_notched_masks_take_the_light_path(lib):
    """Row-convex blobs with notches one or several rows deep, stubs beside the outline, bumps and slits: one component
    without holes decided from the run ends (phase_light_check follows bands of adjacent multi-run rows while they stay
    parallel).  The same masks with a covered gap (a hole), an island, a run that hangs in the air or notches of
    different depths side by side must fall through to the full path - every result is checked against the OpenCV
    model."""
    rng = np.random.default_rng(8)
    n_light = n_full = 0
    for it in range(700):
        H, W = int(rng.integers(16, 70)), int(rng.integers(40, 300))
        m = np.zeros((H, W), np.uint8)
        y0 = int(rng.integers(0, H // 3)); y1 = int(rng.integers(2 * H // 3, H))
        a, b = sorted(int(v) for v in rng.integers(0, W, 2))
        b = max(b, min(W - 1, a + 12))
        spans = {}
        for y in range(y0, y1 + 1):
            a = int(np.clip(a + rng.integers(-3, 4), 0, W - 13)); b = int(np.clip(b + rng.integers(-3, 4), a + 12, W - 1))
            m[y, a:b + 1] = 1
            spans[y] = (a, b)
        kind = it % 5
        # notches in the top / bottom row (never holes: one neighbour row is missing)
        for y in ([y0] if kind in (0, 3) else [y1] if kind == 1 else [y0, y1]):
            a, b = spans[y]
            step = 1 if y == y0 else -1
            same_depth = int(rng.integers(1, 7))
            for _ in range(int(rng.choice([1, 1, 1, 2, 3]))):
                c0 = int(rng.integers(a + 1, b)); c1 = min(b - 1, c0 + int(rng.integers(0, 9)))
                depth = same_depth if rng.random() < 0.5 else int(rng.integers(1, 7))                 # a notch several rows deep, narrowing or drifting as it goes in
                for k in range(depth):
                    yy = y + step * k
                    if not (y0 <= yy <= y1) or c0 > c1:
                        break
                    m[yy, c0:c1 + 1] = 0
                    c0 += int(rng.integers(-1, 2)); c1 -= int(rng.integers(0, 2))
        # interior rows: a stub beside the main run (a notch when a neighbour row leaves the gap open, a hole when both
        # cover it, an island when the stub touches neither neighbour), sometimes a slit inside the run (a hole)
        rows = [y for y in range(y0 + 2, y1 - 1, int(rng.integers(2, 9)))][:int(rng.integers(0, 4))]
        for y in rows:
            a, b = spans[y]
            if rng.random() < 0.75:
                gap = int(rng.integers(1, 4)); ln = int(rng.integers(1, 5))
                if rng.random() < 0.5 and a - gap - ln >= 0:
                    m[y, a - gap - ln:a - gap] = 1
                elif b + gap + ln < W:
                    m[y, b + gap + 1:b + gap + 1 + ln] = 1
            else:
                c0 = int(rng.integers(a + 1, b))
                m[y, c0:min(b - 1, c0 + 2) + 1] = 0
        if kind == 3 and rows:                              # two adjacent rows with several runs: never light
            a, b = spans[rows[0] + 1]
            m[rows[0] + 1, (a + b) // 2] = 0
        g = check(lib, m, int(rng.choice([4, 8, 20])), fmt=it % 2, force_general=1, nthreads=int(rng.choice([1, 32, 64, 256])))
        n_light += g["light"]; n_full += g["general"] and not g["light"]
        check(lib, m, 20, fmt=it % 2, force_general=2, nthreads=64)
    assert n_light > 100 and n_full > 100, (n_light, n_full)


def test_row_convex_shapes_take_the_certificate(lib):
    rng = np.random.default_rng(4)
    for it in range(300):
        H, W = int(rng.integers(12, 60)), int(rng.integers(12, 280))
        m = np.zeros((H, W), np.uint8)
        y0 = int(rng.integers(0, H - 1)); y1 = int(rng.integers(y0, H))
        a, b = sorted(int(v) for v in rng.integers(0, W, 2))
        for y in range(y0, y1 + 1):
            m[y, a:b + 1] = 1
            for _ in range(50):
                a2, b2 = sorted(int(v) for v in rng.integers(0, W, 2))
                if a2 <= b + 1 and b2 >= a - 1:
                    a, b = a2, b2
                    break
        g = check(lib, m, 10)
        assert not g["general"] and g["state"] == 1


def test_shapes(lib):
    H, W = 200, 300
    cases = []
    m = np.zeros((H, W), np.uint8); cases.append(m.copy())                                   # empty
    m = np.zeros((H, W), np.uint8); m[50, 60] = 1; cases.append(m.copy())                    # single pixel
    m = np.zeros((H, W), np.uint8); cv2.circle(m, (150, 100), 60, 1, 5); cases.append(m.copy())   # ring
    cv2.circle(m, (150, 100), 20, 1, -1); cases.append(m.copy())                             # ring + island
    cv2.circle(m, (150, 100), 8, 0, -1); cases.append(m.copy())                              # ring + ring island
    m = np.zeros((H, W), np.uint8); m[20:60, 20:60] = 1; m[100:140, 200:240] = 1; cases.append(m.copy())   # equal blobs (tie)
    m[100:141, 200:240] = 1; cases.append(m.copy())
    m = np.ones((H, W), np.uint8); cases.append(m.copy())                                    # full frame
    m[1:-1, 1:-1] = 0; cases.append(m.copy())                                                # frame-hugging ring
    m = np.zeros((H, W), np.uint8); m[::2, ::2] = 1; cases.append(m.copy())                  # isolated pixels everywhere
    m = (np.indices((H, W)).sum(0) % 2).astype(np.uint8); cases.append(m.copy())             # checkerboard: one 8-connected net
    m = np.zeros((H, W), np.uint8)
    for k in range(0, 90, 4):                                                                # spiral-ish nested squares with gaps
        cv2.rectangle(m, (10 + k, 10 + k), (289 - k, 189 - k), 1, 1)
        m[100, 10 + k] = 0
    cases.append(m.copy())
    for m in cases:
        for fmt in (0, 1):
            check(lib, m, 20, fmt=fmt)
            check(lib, m, 4, fmt=fmt, force_general=1, nthreads=128)


def test_run_capacity_overflow_is_reported(lib):
    m = np.zeros((64, 256), np.uint8); m[::2, ::2] = 1
    assert run(lib, m, 20, cap=100)["state"] == 4
