"""CPU: the oracle against the committed golden vectors (generated from the unmodified reference
by tests/golden/make_golden.py).  Runs everywhere, including the GPU box."""
import numpy as np
import pytest
import torch

import goldenio
from oracle import graph as ograph
from oracle import grid as og
from oracle import mask_assembly as oma
from oracle import nms as onms
from oracle import penalty as open_
from oracle import pipeline as opl
from oracle import protrusion as oprot
from vision_assist_b200 import synth


def test_polygon_cases_bit_exact():
    n = 0
    for case in goldenio.polygon_cases():
        H, W, gs = case["H"], case["W"], case["gs"]
        try:
            st = og.extract_grid_from_polygons(case["polys"], H, W, gs)
            err = 0
        except IndexError as e:
            err = 1 if "centre" in str(e) else 2
            st = None
        assert err == case["err"], case["idx"]
        res = opl.state_to_result(st)
        goldenio.assert_result_matches(res, case, f"polygon case {case['idx']}")
        if st is not None and st.grids:
            assert oprot.peaks_raster(st.grids, H, W, gs) == [tuple(p) for p in case["peaks"].tolist()]
        n += 1
    assert n >= 150


def test_reference_fixtures_penalties_and_peaks():
    z = goldenio.load("fixtures.npz")
    for nm in z["names"]:
        st = og.grid_from_npy(z[f"{nm}/grid"].astype(bool))
        assert np.array_equal(np.array([r[0].y for r in st.grids]), z[f"{nm}/rows_y"])
        occ = np.array([[(0 if g.empty else 1) | (2 if g.artificial else 0) for g in r] for r in st.grids], np.uint8)
        assert np.array_equal(occ, z[f"{nm}/occ"])
        for use_easy, key in ((False, "pen_traversal"), (True, "pen_easy")):
            pen = open_.calculate_penalties(st, use_easy=use_easy)
            gold = z[f"{nm}/{key}"]
            assert np.array_equal(np.isnan(pen), np.isnan(gold))
            assert np.array_equal(pen[~np.isnan(pen)].view(np.uint64), gold[~np.isnan(gold)].view(np.uint64)), nm
        rows_y = np.array([r[0].y for r in st.grids])
        pk = oprot.peaks_closed_form(rows_y, (occ & 1).astype(bool), 0, st.W, 20)
        assert pk == [tuple(p) for p in z[f"{nm}/peaks"].tolist()], nm
        assert oprot.peaks_raster(st.grids, st.H, st.W, 20) == pk
        # start / goal cells and graph neighbourhood as the reference's own functions give them
        start, goals = ograph.start_and_goals(st, pk)
        assert tuple(start) == tuple(int(v) for v in z[f"{nm}/start"]), nm
        assert [tuple(g) for g in goals] == [tuple(g) for g in z[f"{nm}/goals"].tolist()], nm
        assert np.array_equal(ograph.neighbour_mask(st), z[f"{nm}/nbr"]), nm


def test_reference_png_known_answers():
    """The 5 live `*_processed.png` renders of the reference: per-cell LUT colour."""
    z = goldenio.load("fixtures.npz")
    keys = list(open_.PENALTY_COLOUR_GRADIENT.keys())
    checked = 0
    for nm in z["live_png"]:
        st = og.grid_from_npy(z[f"{nm}/grid"].astype(bool))
        pen = open_.calculate_penalties(st, use_easy=False)
        col = z[f"{nm}/png_colour_idx"]
        for r in range(pen.shape[0]):
            for c in range(pen.shape[1]):
                if not np.isnan(pen[r, c]):
                    want = keys[int(col[r, c])]
                    assert open_.get_penalty_colour(pen[r, c]) == open_.PENALTY_COLOUR_GRADIENT[want], (nm, r, c)
                    checked += 1
    assert checked > 2000


def test_mask_assembly_golden():
    z = goldenio.load("mask_assembly.npz")
    for i, (f, n, H, W, mh, mw) in enumerate(z["cases"].tolist()):
        fam = str(z["families"][i])
        p, c, b = synth.make_frame(f, n, H, W, mh, mw, 32, fam)
        gold = np.unpackbits(z[f"{i}/masks_packed"], axis=-1)[..., :W]
        m = oma.process_mask(p, c, b, (H, W)).numpy().astype(np.uint8)
        assert np.array_equal(m, gold), i
        cl = oma.cropped_logits(p, c, b, (H, W)).numpy()
        g = z[f"{i}/cropped_logits"]
        assert np.array_equal(cl == 0, g == 0)
        assert np.allclose(cl, g, rtol=1e-5, atol=1e-5)
        # explicit numpy specification: identical upsampling arithmetic given the same logits
        up_np = oma.bilinear_upsample_np(g, (H, W))
        up_t = torch.nn.functional.interpolate(torch.from_numpy(g)[None], (H, W), mode="bilinear",
                                               align_corners=False)[0].numpy()
        assert np.array_equal(up_np, up_t), i
        assert np.array_equal((up_np > 0).astype(np.uint8), gold), i


def test_frames_golden_contour_and_direct_routes():
    z = goldenio.load("frames.npz")
    n_direct_diff = 0
    for ci, (H, W, mh, mw, n, gs, first, count) in enumerate(z["cases"].tolist()):
        fam = str(z["families"][ci])
        for k in range(count):
            p, c, b = synth.make_frame(first + k, n, H, W, mh, mw, 32, fam)
            res = opl.frame_from_tensors(p, c, b, (H, W), gs, "contour")
            R, C = int(z[f"{ci}/R"][k]), int(z[f"{ci}/C"][k])
            case = dict(R=R, C=C, x0=int(z[f"{ci}/x0"][k]), rows_y=z[f"{ci}/rows_y"][k][:R],
                        rows_attr=z[f"{ci}/rows_attr"][k][:R], occ=z[f"{ci}/occ"][k][:R, :C],
                        pen=z[f"{ci}/pen"][k][:R, :C], peaks=z[f"{ci}/peaks"][k][:int(z[f"{ci}/npk"][k])],
                        start=z[f"{ci}/start"][k], goals=z[f"{ci}/goals"][k][:int(z[f"{ci}/npk"][k])],
                        nbr=z[f"{ci}/nbr"][k][:R, :C])
            goldenio.assert_result_matches(res, case, f"frame case {ci}/{k}")
            assert np.array_equal(res["masks"].reshape(n, -1).sum(1), z[f"{ci}/areas"][k])
            rd = opl.frame_from_masks(res["masks"], gs, "direct")
            try:
                goldenio.assert_result_matches(rd, case)
            except AssertionError:
                n_direct_diff += 1
                assert fam == "noise" or (rd["flags"] & opl.FLAG_NON_SIMPLE), (ci, k)
    # the sidewalk family is hole-free single blobs: direct == contour there
    assert n_direct_diff <= 6


def test_euler_number():
    m = np.zeros((12, 12), np.uint8)
    m[2:9, 2:9] = 1
    assert opl.euler_number_8(m) == 1
    m[4:6, 4:6] = 0
    assert opl.euler_number_8(m) == 0
    m[10, 10] = 1
    assert opl.euler_number_8(m) == 1      # 2 components, 1 hole
    m[9, 9] = 1                            # diagonal touch joins under 8-connectivity
    assert opl.euler_number_8(m) == 0


def test_nms_golden():
    """the oracle's non_max_suppression against the vendored ultralytics function (+ torchvision nms)."""
    z = goldenio.load("nms.npz")
    rows = 0
    for ci, (first, B, A, nc, nobj, ties, md) in enumerate(z["cases"].tolist()):
        ct, it = (float(v) for v in z["thres"][ci])
        pred = synth.make_head_output(first, B, A=A, nc=nc, n_objects=nobj, ties=bool(ties)).numpy()
        got = onms.nms_batch(pred, conf_thres=ct, iou_thres=it, nc=nc, max_det=md)
        for b in range(B):
            want = z[f"{ci}/{b}"]
            assert got[b].shape == want.shape, (ci, b, got[b].shape, want.shape)
            assert np.array_equal(got[b].view(np.uint32), want.view(np.uint32)), (ci, b)
            rows += want.shape[0]
    assert rows > 100
