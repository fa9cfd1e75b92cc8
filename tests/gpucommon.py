"""Helpers shared by the GPU parity tests: compare decoded CUDA records with oracle results."""
from __future__ import annotations

import numpy as np
import torch

from oracle import pipeline as opl


def assert_record_equals_oracle(rec, res: dict, what=""):
    """rec: vision_assist_b200.engine.FrameRecord, res: oracle FrameResult dict.  Bit-exact."""
    err_mask = opl.FLAG_EMPTY | opl.FLAG_CENTRE_OOB | opl.FLAG_LIST_OOB
    assert (rec.flags & err_mask) == (res["flags"] & err_mask), (what, rec.flags, res["flags"])
    assert rec.R == res["R"], (what, rec.R, res["R"])
    if rec.R == 0:
        assert rec.peaks.shape[0] == 0
        return
    assert (rec.C, rec.x0, rec.y0) == (res["C"], res["x0"], res["y0"]), what
    assert np.array_equal(rec.rows_y, res["rows_y"]), what
    assert np.array_equal(rec.rows_attr, res["rows_attr"]), what
    assert np.array_equal(rec.occ, res["occ"]), what
    a, b = rec.penalty, np.asarray(res["penalty"], np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), what
    assert np.array_equal(a[~np.isnan(a)].view(np.uint64), b[~np.isnan(b)].view(np.uint64)), \
        (what, float(np.nanmax(np.abs(a - b))))
    assert np.array_equal(rec.peaks, res["peaks"].reshape(-1, 2)), (what, rec.peaks, res["peaks"])
    assert np.array_equal(rec.orphan_y, res["orphan_y"]), what
    assert np.array_equal(rec.orphan_occ, res["orphan_occ"]), what
    # path start / end cells, grid_lookup row table, implicit _create_graph neighbourhood
    assert tuple(rec.start) == tuple(int(v) for v in res["start"]), (what, rec.start, res["start"])
    assert np.array_equal(rec.goals, np.asarray(res["goals"]).reshape(-1, 2)), (what, rec.goals, res["goals"])
    n = len(res["lookup_row"])
    assert np.array_equal(rec.lookup_row[:n], res["lookup_row"]) and (rec.lookup_row[n:] == -1).all(), \
        (what, rec.lookup_row, res["lookup_row"])
    assert np.array_equal(rec.neighbour_mask(), res["nbr"]), what


def to_dev(*ts):
    return tuple(t.cuda() for t in ts)


def band_mismatch_report(gpu_masks: np.ndarray, up_logits: np.ndarray, band: float = 1e-4):
    """Binary-mask parity rule of the north star: masks must be bit-exact except pixels whose
    reference upsampled logit is within `band` of the threshold (and not exactly 0).  Returns (n_diff, n_diff_outside_band)."""
    ref = (up_logits > 0).astype(np.uint8)
    diff = gpu_masks != ref
    # a reference logit of exactly 0 is a pixel crop_mask (or an all-zero blend) produced: no tolerance there
    outside = diff & ((np.abs(up_logits) >= band) | (up_logits == 0))
    return int(diff.sum()), int(outside.sum())
