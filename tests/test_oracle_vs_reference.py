"""CPU, build container only: the oracle against the UNMODIFIED reference imported from
/root/reference (skipped on the GPU box, where only the golden vectors travel)."""
import numpy as np
import pytest
import torch

import polygen
import refharness
from oracle import graph as ograph
from oracle import grid as og
from oracle import mask_assembly as oma
from oracle import nms as onms
from oracle import penalty as open_
from oracle import pipeline as opl
from oracle import protrusion as oprot

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    return refharness.load()


def _set_gs(ref, gs):
    for m in ("FrameProcessor", "PenaltyCalculator", "ProtrusionDetector", "models"):
        setattr(getattr(ref, m), "grid_size", gs)


@pytest.mark.parametrize("H,W,gs,n", [(640, 640, 20, 220), (720, 1280, 20, 40), (640, 640, 8, 12),
                                      (650, 650, 20, 30), (1080, 1920, 20, 6)])
def test_grid_penalty_peaks_match_reference(ref, H, W, gs, n):
    rng = np.random.default_rng(H * 7 + gs)
    _set_gs(ref, gs)
    try:
        for it in range(n):
            polys = [polygen.random_polygon(rng, H, W) for _ in range(int(rng.integers(1, 4)))]
            fp = refharness.new_frame_processor(ref)
            fp.frame = np.zeros((H, W, 3), np.uint8)
            rerr = oerr = None
            try:
                fp._extract_grid_information([refharness.FakeResult(polys)])
            except IndexError as e:
                rerr = "centre" if "out of bounds" in str(e) else "list"
            try:
                st = og.extract_grid_from_polygons(polys, H, W, gs)
            except IndexError as e:
                oerr = "centre" if "centre" in str(e) else "list"
            assert rerr == oerr
            if rerr:
                continue
            assert len(fp.grids) == len(st.grids)
            if not fp.grids:
                continue
            for rr, ro in zip(fp.grids, st.grids):
                for gr, go in zip(rr, ro):
                    assert (gr.coords.x, gr.coords.y, gr.centre.x, gr.centre.y, gr.row, gr.col, gr.empty,
                            gr.artificial) == (go.x, go.y, go.x + gs // 2, go.y + gs // 2, go.row, go.col,
                                               go.empty, go.artificial)
            assert set(fp.grid_lookup) == set(st.lookup)
            for key, g in fp.grid_lookup.items():
                o = st.lookup[key]
                assert (g.empty, g.artificial, g.row) == (o.empty, o.artificial, o.row)
            assert np.array_equal(fp.np_grids, st.np_grids)
            fp._calculate_penalties()
            pen = open_.calculate_penalties(st)
            for r, rr in enumerate(fp.grids):
                for c, g in enumerate(rr):
                    if g.empty:
                        assert g.penalty is None and np.isnan(pen[r, c])
                    else:
                        assert float(g.penalty) == pen[r, c]
            want = [(p.x, p.y) for p in fp.protrusion_detector(fp.frame, fp.grids, fp.grid_lookup)]
            res = opl.state_to_result(st)
            assert want == oprot.peaks_raster(st.grids, H, W, gs) == [tuple(p) for p in res["peaks"].tolist()]
            # start / goal cells (utils.py:6-32 as called at FrameProcessor.py:236-239) and _create_graph (:184-207)
            Coordinate = ref.models.Coordinate

            def where(obj):
                return next(((k, c) for k, rr in enumerate(fp.grids) for c, g in enumerate(rr) if g is obj), (-1, -1))

            sg = ref.utils.get_closest_grid_to_point(Coordinate(x=W // 2, y=H), fp.grids)
            assert (where(sg) if sg is not None else (-1, -1)) == tuple(res["start"])
            for (px, py), goal in zip(want, res["goals"].tolist()):
                eg = ref.utils.get_closest_grid_to_point(Coordinate(x=px, y=py), fp.grids)
                assert where(eg) == tuple(goal)
            graph = fp._create_graph()
            nbr = ograph.neighbour_mask(st)
            for r, rr in enumerate(fp.grids):
                for c, g in enumerate(rr):
                    x, y = g.coords.x, g.coords.y
                    have = {pos for pos, _ in graph.get((x, y), [])} if not g.empty else set()
                    bits = sum(1 << b for b, pos in enumerate(((x + gs, y), (x - gs, y), (x, y + gs), (x, y - gs)))
                               if pos in have)
                    assert bits == nbr[r, c]
    finally:
        _set_gs(ref, 20)


def test_penalty_colour_lut(ref):
    pc = ref.PenaltyCalculator.penalty_calculator
    for p in np.linspace(0, 1.2, 241):
        assert pc.get_penalty_colour(float(p)) == open_.get_penalty_colour(float(p))


def test_process_mask_matches_vendored_ops(ref):
    torch.manual_seed(3)
    for (mh, mw, ih, iw, n) in [(160, 160, 640, 640, 5), (40, 40, 270, 480, 3), (96, 160, 384, 640, 2)]:
        protos = torch.randn(32, mh, mw)
        coefs = torch.randn(n, 32)
        x1 = torch.rand(n) * iw * 0.5
        y1 = torch.rand(n) * ih * 0.5
        boxes = torch.stack([x1, y1, x1 + torch.rand(n) * iw * 0.5, y1 + torch.rand(n) * ih * 0.5], 1)
        want = ref.ops.process_mask(protos, coefs, boxes, (ih, iw), upsample=True)
        assert torch.equal(want, oma.process_mask(protos, coefs, boxes, (ih, iw)))
        got_np = oma.process_mask_np(protos.numpy(), coefs.numpy(), boxes.numpy(), (ih, iw))
        # numpy matmul may sum in a different order than torch: allow flips only at |logit| ~ 0
        assert (got_np != want.numpy().astype(np.uint8)).mean() < 1e-5
        segs_ref = ref.ops.masks2segments(want)
        segs = oma.masks2segments(want.numpy())
        assert all(np.array_equal(a, b) for a, b in zip(segs_ref, segs))
        for s in segs:
            a = ref.ops.scale_coords((ih, iw), s.copy(), (720, 1280))
            assert np.array_equal(a, oma.scale_coords((ih, iw), s, (720, 1280)))


def test_nms_matches_vendored_ops(ref):
    """oracle.nms against ops.non_max_suppression (vendored ultralytics) with torchvision.ops.nms: random head
    outputs, 1 and 3 classes, score ties, sparse and dense overlaps - bit-exact rows."""
    for seed in range(24):
        g = torch.Generator().manual_seed(seed)
        nc = 1 if seed % 3 else 3
        A = [300, 2100, 8400][seed % 3]
        pred = torch.rand(2, 4 + nc + 32, A, generator=g)
        pred[:, :2] *= 600
        pred[:, 2:4] = pred[:, 2:4] * (60 if seed % 2 else 300) + 5
        pred[:, 4:4 + nc] = pred[:, 4:4 + nc] ** (3 if seed % 2 else 1)
        if seed % 5 == 0:
            pred[:, 4] = torch.round(pred[:, 4] * 20) / 20
        want = ref.ops.non_max_suppression(pred.clone(), conf_thres=0.5, iou_thres=0.7, nc=nc, max_det=300)
        got = onms.nms_batch(pred.numpy(), conf_thres=0.5, iou_thres=0.7, nc=nc, max_det=300)
        for w, m in zip(want, got):
            assert tuple(w.shape) == m.shape, seed
            assert np.array_equal(w.numpy().view(np.uint32), m.view(np.uint32)), seed


def test_scale_boxes_matches_vendored_ops(ref):
    """oracle.nms.scale_boxes against ops.scale_boxes + clip_boxes (ops.py:139-174): letterboxed 640x640 input back to
    frames of several shapes (gain < 1, > 1, pads that round either way), boxes partly outside the frame - bit-exact."""
    g = torch.Generator().manual_seed(5)
    for (h1, w1), (h0, w0) in [((640, 640), (720, 1280)), ((640, 640), (1080, 1920)), ((640, 640), (480, 640)), ((384, 640), (1080, 1920)),
                               ((640, 640), (640, 640)), ((640, 640), (333, 517)), ((640, 480), (1000, 601)), ((320, 320), (97, 131))]:
        boxes = torch.rand(64, 4, generator=g) * torch.tensor([w1, h1, w1, h1]) * 1.2 - 0.1 * max(h1, w1)
        want = ref.ops.scale_boxes((h1, w1), boxes.clone(), (h0, w0)).numpy()
        got = onms.scale_boxes((h1, w1), boxes.numpy(), (h0, w0))
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), ((h1, w1), (h0, w0))
