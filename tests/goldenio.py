"""Readers for tests/golden/*.npz (written by tests/golden/make_golden.py from the reference)."""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name: str):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def polygon_cases():
    z = load("polygons.npz")
    off = z["poly_off"]
    pi = 0
    for i in range(len(z["H"])):
        k = int(z["case_npoly"][i])
        polys = [z["poly_pts"][off[pi + j]:off[pi + j + 1]] for j in range(k)]
        pi += k
        R, C = int(z["R"][i]), int(z["C"][i])
        yield dict(idx=i, H=int(z["H"][i]), W=int(z["W"][i]), gs=int(z["gs"][i]), err=int(z["err"][i]),
                   R=R, C=C, x0=int(z["x0"][i]), polys=polys,
                   rows_y=z["rows_y"][i][:R], rows_attr=z["rows_attr"][i][:R], occ=z["occ"][i][:R, :C],
                   pen=z["pen"][i][:R, :C], peaks=z["peaks"][i][:int(z["npk"][i])],
                   start=z["start"][i], goals=z["goals"][i][:int(z["npk"][i])], nbr=z["nbr"][i][:R, :C])


def assert_result_matches(res: dict, case: dict, what=""):
    """Compare an oracle / decoded-CUDA FrameResult with a golden case (bit-exact)."""
    R, C = case["R"], case["C"]
    assert res["R"] == R and (R == 0 or res["C"] == C), (what, res["R"], res["C"], R, C)
    if R == 0:
        return
    assert res["x0"] == case["x0"], what
    assert np.array_equal(res["rows_y"], case["rows_y"]), what
    assert np.array_equal(res["rows_attr"], case["rows_attr"]), what
    assert np.array_equal(res["occ"], case["occ"]), what
    a, b = np.asarray(res["penalty"], np.float64), case["pen"]
    assert np.array_equal(np.isnan(a), np.isnan(b)), what
    assert np.array_equal(a[~np.isnan(a)].view(np.uint64), b[~np.isnan(b)].view(np.uint64)), \
        (what, np.nanmax(np.abs(a - b)))
    assert np.array_equal(np.asarray(res["peaks"]).reshape(-1, 2), case["peaks"].reshape(-1, 2)), \
        (what, res["peaks"], case["peaks"])
    if "start" in case and "start" in res:      # start / goal cells, graph neighbourhood
        assert tuple(int(v) for v in res["start"]) == tuple(int(v) for v in case["start"]), (what, res["start"], case["start"])
        assert np.array_equal(np.asarray(res["goals"]).reshape(-1, 2), case["goals"].reshape(-1, 2)), what
        if "nbr" in res and "nbr" in case:
            assert np.array_equal(res["nbr"], case["nbr"]), what
