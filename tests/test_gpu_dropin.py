"""GPU: the FrameProcessor / PenaltyCalculator / ProtrusionDetector drop-ins (reference call signatures)
against the oracle.  BASELINE config 0 shape: a single 640x640 frame through FrameProcessor with a
random-init head-output model (no weights exist in the reference repository)."""
import numpy as np
import pytest
import torch

import polygen
from oracle import grid as og
from oracle import penalty as open_
from oracle import pipeline as opl

pytestmark = pytest.mark.gpu


def _fresh():
    from vision_assist_b200 import FrameProcessor as FPmod, PenaltyCalculator as PCmod, ProtrusionDetector as PDmod
    FPmod.FrameProcessor._instance = None
    FPmod.FrameProcessor._initialized = False
    PDmod.ProtrusionDetector._instance = None
    PDmod.ProtrusionDetector._initialized = False
    return FPmod, PCmod, PDmod


def test_cfg0_single_frame_head_output_model():
    from vision_assist_b200 import synth
    FPmod, _, _ = _fresh()
    H = W = 640
    for idx in (0, 1, 2, 3):
        protos, coefs, boxes = synth.make_frame(idx, 8, H, W, 160, 160)
        model = FPmod.HeadOutputModel(lambda frame: (protos.cuda(), coefs.cuda(), boxes.cuda()))
        FPmod.FrameProcessor._instance = None
        FPmod.FrameProcessor._initialized = False
        fp = FPmod.FrameProcessor(model, verbose=False, debug=False)          # main.py:44 omits imshow
        frame = np.zeros((H, W, 3), np.uint8)
        peaks = fp(frame)
        want = opl.frame_from_tensors(protos, coefs, boxes, (H, W), 20, "contour")
        assert [(p.x, p.y) for p in peaks] == [tuple(p) for p in want["peaks"].tolist()]
        assert np.array_equal(fp.np_grids, (want["occ"] & 1))
        assert len(fp.grids) == want["R"]
        for k, row in enumerate(fp.grids):
            for c, g in enumerate(row):
                assert (g.coords.x, g.coords.y, g.row, g.col) == (want["x0"] + 20 * c, int(want["rows_y"][k]), int(want["rows_attr"][k]), c)
                assert g.empty == (not (want["occ"][k, c] & 1)) and g.artificial == bool(want["occ"][k, c] & 2)
                if g.empty:
                    assert g.penalty is None
                else:
                    assert g.penalty == want["penalty"][k, c]
        graph = fp._create_graph()
        assert len(graph) > 0
        # the same call with no detections returns [] like the reference (FrameProcessor.py:328-332)
        empty = FPmod.HeadOutputModel(lambda frame: (protos.cuda(), coefs.cuda()[:0], boxes.cuda()[:0]))
        fp.model = empty
        assert fp(frame) == []


def test_raw_head_output_route():
    """Raw head output -> va_nms -> fused path inside the drop-in FrameProcessor equals the route that starts
    from the oracle's non_max_suppression rows."""
    from oracle import nms as onms
    from vision_assist_b200 import synth
    FPmod, _, _ = _fresh()
    H = W = 640
    pred = synth.make_head_output(77, 1, A=8400, nc=1, n_objects=5)[0]
    protos = synth.make_frame(5, 8, H, W, 160, 160)[0]
    rows = onms.nms_image(pred.numpy(), conf_thres=0.5, iou_thres=0.7, nc=1, max_det=32)
    assert rows.shape[0] >= 2

    class RawModel:
        def predict(self, frame, conf=0.5, verbose=False):
            return [FPmod.RawHeadResult(protos.cuda(), pred.cuda(), nc=1, conf=conf)]

    FPmod.FrameProcessor._instance = None
    FPmod.FrameProcessor._initialized = False
    fp = FPmod.FrameProcessor(RawModel(), verbose=False, debug=False)
    frame = np.zeros((H, W, 3), np.uint8)
    peaks = fp(frame)
    rec_raw = fp.frame_record
    import torch
    model = FPmod.HeadOutputModel(lambda f: (protos.cuda(), torch.from_numpy(rows[:, 6:]).cuda(), torch.from_numpy(rows[:, :4]).cuda()))
    FPmod.FrameProcessor._instance = None
    FPmod.FrameProcessor._initialized = False
    fp2 = FPmod.FrameProcessor(model, verbose=False, debug=False)
    peaks2 = fp2(frame)
    rec = fp2.frame_record
    if rec is None:
        assert rec_raw is None and peaks == [] and peaks2 == []
        return
    assert [(p.x, p.y) for p in peaks] == [(p.x, p.y) for p in peaks2]
    assert np.array_equal(rec_raw.occ, rec.occ) and np.array_equal(rec_raw.rows_y, rec.rows_y)
    assert np.array_equal(np.nan_to_num(rec_raw.penalty, nan=-1.0), np.nan_to_num(rec.penalty, nan=-1.0))
    assert rec_raw.start == rec.start and np.array_equal(rec_raw.goals, rec.goals)


def test_polygon_model_route_and_reference_errors():
    FPmod, _, _ = _fresh()

    class Masks:
        def __init__(self, xy): self.xy = xy

    class Result:
        def __init__(self, xy): self.masks = Masks(xy) if xy is not None else None

    class PolyModel:
        def __init__(self): self.xy = None
        def predict(self, frame, conf=0.5, verbose=False): return [Result(self.xy)]

    rng = np.random.default_rng(3)
    model = PolyModel()
    fp = FPmod.FrameProcessor(model, False, False, False)
    checked = 0
    for H, W in ((640, 640), (650, 650)):
        frame = np.zeros((H, W, 3), np.uint8)
        for it in range(25):
            model.xy = [polygen.random_polygon(rng, H, W) for _ in range(int(rng.integers(1, 3)))]
            fp.frame = frame
            err = None
            try:
                st = og.extract_grid_from_polygons(model.xy, H, W, 20)
            except IndexError as e:
                err = e
            if err is not None:
                with pytest.raises(IndexError):
                    fp._extract_grid_information(model.predict(frame))
                continue
            fp._extract_grid_information(model.predict(frame))
            assert np.array_equal(fp.np_grids, st.np_grids)
            if st.grids:
                fp._calculate_penalties()
                pen = open_.calculate_penalties(st)
                for k, row in enumerate(fp.grids):
                    for c, g in enumerate(row):
                        assert (g.penalty is None and np.isnan(pen[k, c])) or g.penalty == pen[k, c]
                checked += 1
    model.xy = None
    assert fp(np.zeros((640, 640, 3), np.uint8)) == []
    assert checked > 20


def test_penalty_calculator_and_protrusion_detector_dropins():
    """Reference call pattern: _pre_compute_easy_segments(np_grids, grids) then calculate_penalty(grid, lookup)
    per cell (FrameProcessor.py:173-182), ProtrusionDetector()(frame, grids, lookup) (:341)."""
    FPmod, PCmod, PDmod = _fresh()
    from vision_assist_b200.materialise import record_to_objects
    from test_dropin_cpu import record_from_oracle
    rng = np.random.default_rng(9)
    pc = PCmod.PenaltyCalculator()
    assert pc is PCmod.penalty_calculator
    pd = PDmod.ProtrusionDetector(debug=False, imshow=False)
    frame = np.zeros((640, 640, 3), np.uint8)
    n = 0
    for it in range(12):
        polys = [polygen.random_polygon(rng, 640, 640, kind="blob")]
        st = og.extract_grid_from_polygons(polys, 640, 640, 20)
        if not st.grids:
            continue
        res = opl.state_to_result(st)
        rec = record_from_oracle(res)
        grids, lookup, np_grids = record_to_objects(rec, 20)
        for row in grids:
            for g in row:
                g.penalty = None
        pc._pre_compute_easy_segments(np_grids, grids)
        for k, row in enumerate(grids):
            for c, g in enumerate(row):
                p = pc.calculate_penalty(g, lookup)
                if g.empty:
                    assert p == 0
                else:
                    assert p == res["penalty"][k, c]
        peaks = pd(frame, grids, lookup)
        assert [(p.x, p.y) for p in peaks] == [tuple(p) for p in res["peaks"].tolist()]
        n += 1
    assert n >= 8
    assert pc.get_penalty_colour(0.49) == open_.get_penalty_colour(0.49)
