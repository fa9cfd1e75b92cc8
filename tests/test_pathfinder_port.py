"""the array-based A* (vision_assist_b200/PathFinder.py) against the reference's own PathFinder
(PathFinder.py:119-186) on the reference's own graph (FrameProcessor._create_graph) - identical cell sequences and
bit-identical costs, with the reference's `angle_cache` carried over from search to search on both sides (it stores
radians and serves them as degrees the second time, so the order of the searches matters).

Here (build container) the reference is imported; on the GPU box the same comparison runs against the committed golden
paths of tests/golden/cfg0.npz (tests/test_gpu_cfg0.py)."""
import numpy as np
import pytest

import polygen
import refharness
from oracle import grid as og
from oracle import pipeline as opl
from test_dropin_cpu import record_from_oracle
from vision_assist_b200.PathFinder import ArrayPathFinder


@pytest.mark.reference
def test_array_astar_equals_reference_astar():
    ref = refharness.load()
    rng = np.random.default_rng(21)
    H = W = 640
    ref.PathFinder.path_finder.angle_cache.clear()          # one shared cache per side for the whole sequence
    mine = ArrayPathFinder()
    done = n_paths = 0
    for it in range(120):
        polys = [polygen.random_polygon(rng, H, W, kind="blob") for _ in range(int(rng.integers(1, 3)))]
        fp = refharness.new_frame_processor(ref)
        fp.frame = np.zeros((H, W, 3), np.uint8)
        try:
            fp._extract_grid_information([refharness.FakeResult(polys)])
        except IndexError:
            continue
        if not fp.grids:
            continue
        fp._calculate_penalties()
        graph = fp._create_graph()
        peaks = fp.protrusion_detector(fp.frame, fp.grids, fp.grid_lookup)
        start = ref.utils.get_closest_grid_to_point(ref.models.Coordinate(x=W // 2, y=H), fp.grids)
        want = []
        for peak in peaks:
            end = ref.utils.get_closest_grid_to_point(peak, fp.grids)
            grids, cost = ref.PathFinder.path_finder.find_path(graph, start, end, fp.grid_lookup)
            want.append(([(g.coords.x, g.coords.y) for g in grids], cost) if grids else None)

        rec = record_from_oracle(opl.state_to_result(og.extract_grid_from_polygons(polys, H, W, 20)))
        got = mine.find_paths(rec, 20)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert (a is None) == (b is None)
            if a is not None:
                assert a[0] == b[0]
                assert np.float64(a[1]).view(np.uint64) == np.float64(b[1]).view(np.uint64)
                n_paths += 1
        done += 1
    assert done >= 60 and n_paths >= 50
    # both caches saw the same windows in the same order
    theirs = ref.PathFinder.path_finder.angle_cache
    assert set(mine.angle_cache) == set(theirs)
    assert all(np.float64(mine.angle_cache[k]).view(np.uint64) == np.float64(theirs[k]).view(np.uint64) for k in theirs)
