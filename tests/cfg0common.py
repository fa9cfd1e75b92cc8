"""Shared by the CPU and GPU halves of the BASELINE configs[0] test: golden access and path comparison."""
import numpy as np

import goldenio


def golden_case(z, k):
    R, C, npk = int(z[f"{k}/R"]), int(z[f"{k}/C"]), int(z[f"{k}/npk"])
    return dict(R=R, C=C, x0=int(z[f"{k}/x0"]), rows_y=z[f"{k}/rows_y"][:R], rows_attr=z[f"{k}/rows_attr"][:R],
                occ=z[f"{k}/occ"][:R, :C], pen=z[f"{k}/pen"][:R, :C], peaks=z[f"{k}/peaks"][:npk],
                start=z[f"{k}/start"], goals=z[f"{k}/goals"][:npk])


def assert_paths_match(z, k, paths, what=""):
    """paths: [(cells [(x, y), ...], cost)] after the similarity filter, in the reference's order."""
    want_len = z[f"{k}/path_len"]
    cells = z[f"{k}/path_cells"]
    off = np.concatenate([[0], np.cumsum(want_len)])
    assert len(paths) == int(z[f"{k}/n_paths"]), what
    for j, (got_cells, got_cost) in enumerate(paths):
        assert [tuple(int(v) for v in c) for c in got_cells] == [tuple(int(v) for v in q) for q in cells[off[j]:off[j + 1]]], (what, j)
        assert np.float64(got_cost).view(np.uint64) == z[f"{k}/path_cost"][j].view(np.uint64), (what, j)


def similarity_filter(found):
    """FrameProcessor.py:255-269 on (cells, cost) tuples."""
    paths = sorted((p for p in found if p is not None), key=lambda pc: len(pc[0]), reverse=True)
    unique = []
    for cells, cost in paths:
        a = set(cells)
        ok = True
        for other, _ in unique:
            b = set(other)
            inter = len(a & b)
            sim = 0.0 if not a or not b else 1.0 if inter in (len(a), len(b)) else inter / len(a | b)
            if sim >= 0.90:
                ok = False
                break
        if ok:
            unique.append((cells, cost))
    return unique
