"""CPU: the host-side mirror of the reference interface (materialise.py, models.py) without a GPU.

A FrameRecord is built from the ORACLE's result (the CUDA record is bit-identical to it, see
tests/test_gpu_parity.py) and turned into the reference's object view; in the build container the
objects are then pushed through the UNMODIFIED reference host stages (graph, A*, PathAnalyser) and must
give the same paths and the same final answer as the reference FrameProcessor working on its own objects.
"""
import numpy as np
import pytest

import polygen
import refharness
from oracle import grid as og
from oracle import pipeline as opl
from vision_assist_b200 import config as vconfig
from vision_assist_b200 import models as vmodels
from vision_assist_b200.engine import FrameRecord
from vision_assist_b200.FrameProcessor import FrameProcessor as VFP
from vision_assist_b200.materialise import objects_to_grid_input, record_to_objects


def record_from_oracle(res: dict) -> FrameRecord:
    R = res["R"]
    return FrameRecord(flags=res["flags"], sel=res["sel"], x0=res["x0"], y0=res["y0"], C=res["C"], R=R,
                       n_orphans=len(res["orphan_y"]), area=0, bbox=(0, 0, 0, 0), contour_area2=0,
                       rows_y=res["rows_y"], rows_attr=res["rows_attr"], occ=res["occ"], penalty=res["penalty"],
                       peaks=res["peaks"], orphan_y=res["orphan_y"], orphan_occ=res["orphan_occ"],
                       start=tuple(res["start"]), goals=res["goals"], lookup_row=res["lookup_row"])


def test_record_to_objects_roundtrip():
    rng = np.random.default_rng(5)
    n = 0
    for it in range(80):
        polys = [polygen.random_polygon(rng, 640, 640)]
        try:
            st = og.extract_grid_from_polygons(polys, 640, 640, 20)
        except IndexError:
            continue
        if not st.grids:
            continue
        res = opl.state_to_result(st)
        grids, lookup, np_grids = record_to_objects(record_from_oracle(res), 20)
        assert np.array_equal(np_grids, st.np_grids)
        assert len(grids) == len(st.grids)
        for rr, ro in zip(grids, st.grids):
            for g, o in zip(rr, ro):
                assert (g.coords.x, g.coords.y, g.centre.x, g.centre.y, g.row, g.col, g.empty, g.artificial) == \
                       (o.x, o.y, o.x + 10, o.y + 10, o.row, o.col, o.empty, o.artificial)
                assert (g.penalty is None) == (o.penalty is None)
                if o.penalty is not None:
                    assert float(g.penalty) == float(o.penalty)
        assert set(lookup) == set(st.lookup)
        for key, o in st.lookup.items():
            g = lookup[key]
            assert (g.empty, g.artificial) == (o.empty, o.artificial)
        # the lookup points at the LAST list row with that y (duplicate-row quirk)
        for row in grids:
            for g in row:
                assert lookup[(g.coords.x, g.coords.y)].coords == g.coords
        gi = objects_to_grid_input(grids, lookup, 20)
        assert np.array_equal(gi["occ"], res["occ"]) and np.array_equal(gi["rows_y"], res["rows_y"])
        assert np.array_equal(gi["rows_attr"], res["rows_attr"])
        n += 1
    assert n > 40


def test_models_match_reference_fields():
    c = vmodels.Coordinate(x=3, y=4)
    assert c.to_tuple() == (3, 4) and c.midpoint == (3 + vconfig.grid_size // 2, 4 + vconfig.grid_size // 2)
    g = vmodels.Grid(coords=c, centre=c, penalty=None, row=1, col=2, empty=True, artificial=False)
    assert set(type(g).model_fields) == {"coords", "centre", "penalty", "row", "col", "empty", "artificial"}
    if refharness.available():
        ref = refharness.load()
        for name in ("Coordinate", "Grid", "Peak"):
            assert set(getattr(ref.models, name).model_fields) == set(getattr(vmodels, name).model_fields), name
        assert ref.config.grid_size == vconfig.grid_size
        assert dict(ref.config.penalty_colour_gradient) == dict(vconfig.penalty_colour_gradient)


@pytest.mark.reference
def test_reference_host_stages_on_materialised_objects():
    """A* paths and the final instruction are identical when the reference's host stages consume the
    objects materialised from the record instead of the reference's own objects."""
    ref = refharness.load()
    rng = np.random.default_rng(11)
    H = W = 640
    vmodels.bind_models(ref.models)          # produce the reference's own pydantic classes
    try:
        done = 0
        for it in range(40):
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
            graph_ref = fp._create_graph()
            graph_ref_snapshot = {k: list(v) for k, v in graph_ref.items()}   # A* adds empty entries to the defaultdict
            peaks_ref = fp.protrusion_detector(fp.frame, fp.grids, fp.grid_lookup)
            ref.PathFinder.path_finder.angle_cache.clear()
            paths_ref = fp._find_paths(peaks_ref, graph_ref)
            want = [[(g.coords.x, g.coords.y) for g in p.grids] for p in paths_ref]
            want_cost = [p.total_cost for p in paths_ref]

            st = og.extract_grid_from_polygons(polys, H, W, 20)
            res_full = opl.state_to_result(st)
            res_nbr = res_full["nbr"]
            rec = record_from_oracle(res_full)
            fp2 = refharness.new_frame_processor(ref)
            fp2.frame = fp.frame
            fp2.grids, fp2.grid_lookup, fp2.np_grids = record_to_objects(rec, 20)
            graph2 = fp2._create_graph()
            assert {k: list(v) for k, v in graph2.items()} == graph_ref_snapshot
            peaks2 = [ref.models.Coordinate(x=int(x), y=int(y)) for x, y in rec.peaks]
            assert [(p.x, p.y) for p in peaks2] == [(p.x, p.y) for p in peaks_ref]
            ref.PathFinder.path_finder.angle_cache.clear()
            paths2 = fp2._find_paths(peaks2, graph2)
            assert [[(g.coords.x, g.coords.y) for g in p.grids] for p in paths2] == want
            assert [p.total_cost for p in paths2] == want_cost
            # the drop-in FrameProcessor._find_paths with the start / end cells taken from the record
            # instead of utils.get_closest_grid_to_point; A* itself is the reference's
            VFP._instance, VFP._initialized = None, False
            vfp = VFP(model=None)
            vfp.bind_host_stages(path_finder=ref.PathFinder.path_finder, Path=ref.models.Path)
            vfp.frame, vfp.frame_record = fp.frame, rec
            vfp.grids, vfp.grid_lookup, vfp.np_grids = fp2.grids, fp2.grid_lookup, fp2.np_grids
            ref.PathFinder.path_finder.angle_cache.clear()
            paths3 = vfp._find_paths(peaks2, fp2._create_graph())
            assert [[(g.coords.x, g.coords.y) for g in p.grids] for p in paths3] == want
            assert [p.total_cost for p in paths3] == want_cost
            assert np.array_equal(rec.neighbour_mask(), res_nbr)
            done += 1
        assert done >= 20
    finally:
        vmodels.bind_models(vmodels)
        VFP._instance, VFP._initialized = None, False


def test_lazy_object_view_of_the_dropin():
    """FrameProcessor.grids / grid_lookup are built from the pending record on first access only, and the protrusion
    detector's `.grids` follows them (no GPU: the record comes from the oracle)."""
    rng = np.random.default_rng(9)
    rec = None
    while rec is None:
        try:
            st = og.extract_grid_from_polygons([polygen.random_polygon(rng, 640, 640, kind="blob")], 640, 640, 20)
        except IndexError:
            continue
        if st.grids:
            rec = record_from_oracle(opl.state_to_result(st))
    VFP._instance, VFP._initialized = None, False
    try:
        fp = VFP(model=None)
        fp.frame = np.zeros((640, 640, 3), np.uint8)
        assert fp.grids == [] and fp.grid_lookup == {} and not fp._has_grid()
        fp.frame_record, fp.np_grids, fp._pending_record = rec, rec.np_grids, rec          # what _extract_grid_information leaves
        assert fp._has_grid() and fp._grids == []                                          # nothing built yet
        peaks = fp.protrusion_detector.from_record(fp.frame, rec.peaks, lambda: fp.grids)
        assert [(p.x, p.y) for p in peaks] == [tuple(int(v) for v in q) for q in rec.peaks]
        assert fp._pending_record is rec                                                   # still nothing built
        want_grids, want_lookup, _ = record_to_objects(rec, 20)
        got = fp.grids                                                                     # first access builds the objects
        assert fp._pending_record is None and len(got) == rec.R
        assert [[(g.coords.x, g.coords.y, g.empty, g.artificial, g.penalty) for g in row] for row in got] == \
               [[(g.coords.x, g.coords.y, g.empty, g.artificial, g.penalty) for g in row] for row in want_grids]
        assert set(fp.grid_lookup) == set(want_lookup)
        assert fp.protrusion_detector.grids is got                                         # the detector's lazy view resolves to the same list
        fp.grids = []                                                                      # assigning drops any pending record
        fp._pending_record = rec
        fp.grid_lookup = {}
        assert fp._pending_record is None and fp.grids == []
    finally:
        VFP._instance, VFP._initialized = None, False
