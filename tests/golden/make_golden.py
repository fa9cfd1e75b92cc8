"""Generate the golden vectors in tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

Everything written here is produced by executing reference code (FrameProcessor,
PenaltyCalculator, ProtrusionDetector, the vendored ultralytics ops.py) or by reading reference
data files (the 13 `*_grids.npy` inputs and the `*_processed.png` known-answer renders) - the
oracle in oracle/ is NOT used, so the vectors pin the oracle as well as the CUDA path.

Files:
  fixtures.npz      13 reference occupancy fixtures (utilities/generate_testing_grids/examples)
                    + reference penalties / peaks for them + per-cell LUT colour index read from
                    the 5 live `*_processed.png` renders.
  polygons.npz      seeded random polygons through reference FrameProcessor._extract_grid_information
                    -> _calculate_penalties -> ProtrusionDetector (list rows, attrs, flags, penalties, peaks).
  mask_assembly.npz seeded synthetic head outputs through the vendored ops.process_mask
                    (cropped logits + packed binary masks).
  frames.npz        seeded synthetic frames through the whole reference chain
                    process_mask -> masks2segments -> scale_coords -> FrameProcessor -> penalties -> peaks.
  nms.npz           seeded synthetic raw head outputs through the vendored ops.non_max_suppression.
"""
from __future__ import annotations

import importlib.util
import inspect
import os
import sys
import textwrap

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import polygen  # noqa: E402
import refharness  # noqa: E402
from vision_assist_b200 import synth  # noqa: E402

RMAX, CMAX, PMAX = 80, 100, 48
LUT_KEYS = [1.0, 0.9166, 0.8333, 0.75, 0.6666, 0.5833, 0.5, 0.4166, 0.3333, 0.1666, 0.0833, 0.0]


def ref_state_arrays(fp, ref, H, W):
    """Dump a reference FrameProcessor's grid state (+penalties, peaks) into fixed-size arrays."""
    R = len(fp.grids)
    C = len(fp.grids[0]) if R else 0
    assert R <= RMAX and C <= CMAX
    rows_y = np.full(RMAX, -1, np.int32)
    rows_attr = np.full(RMAX, -1, np.int32)
    occ = np.zeros((RMAX, CMAX), np.uint8)
    pen = np.full((RMAX, CMAX), np.nan)
    peaks = np.full((PMAX, 2), -1, np.int32)
    x0 = fp.grids[0][0].coords.x if R else 0
    npk = 0
    if R:
        fp._calculate_penalties()
        for r, row in enumerate(fp.grids):
            rows_y[r] = row[0].coords.y
            rows_attr[r] = row[0].row
            for c, g in enumerate(row):
                occ[r, c] = (0 if g.empty else 1) | (2 if g.artificial else 0)
                if not g.empty:
                    pen[r, c] = float(g.penalty)
        pk = fp.protrusion_detector(fp.frame, fp.grids, fp.grid_lookup)
        npk = len(pk)
        for k, p in enumerate(pk):
            peaks[k] = (p.x, p.y)
    # the reference's own start / goal cell choice (utils.get_closest_grid_to_point as used by
    # FrameProcessor._find_paths :236-239) and _create_graph (:184-207) as a neighbour bit mask per list cell
    start = np.full(2, -1, np.int32)
    goals = np.full((PMAX, 2), -1, np.int32)
    nbr = np.zeros((RMAX, CMAX), np.uint8)
    if R:
        gs = ref.FrameProcessor.grid_size

        def where(obj):
            for k, row in enumerate(fp.grids):
                for c, g in enumerate(row):
                    if g is obj:
                        return k, c
            return -1, -1

        Coordinate = ref.models.Coordinate
        sg = ref.utils.get_closest_grid_to_point(Coordinate(x=W // 2, y=H), fp.grids)
        if sg is not None:
            start[:] = where(sg)
        for k in range(npk):
            eg = ref.utils.get_closest_grid_to_point(Coordinate(x=int(peaks[k, 0]), y=int(peaks[k, 1])), fp.grids)
            if eg is not None:
                goals[k] = where(eg)
        graph = fp._create_graph()
        for r, row in enumerate(fp.grids):
            for c, g in enumerate(row):
                if g.empty:
                    continue
                x, y = g.coords.x, g.coords.y
                have = {pos for pos, _ in graph.get((x, y), [])}
                for bit, pos in enumerate(((x + gs, y), (x - gs, y), (x, y + gs), (x, y - gs))):
                    if pos in have:
                        nbr[r, c] |= 1 << bit
    return dict(R=R, C=C, x0=x0, rows_y=rows_y, rows_attr=rows_attr, occ=occ, pen=pen, peaks=peaks, npk=npk,
                start=start, goals=goals, nbr=nbr)


def gen_fixtures(ref):
    exdir = os.path.join(refharness.REFERENCE_ROOT, "utilities/generate_testing_grids/examples")
    names = sorted(f[:-len("_grids.npy")] for f in os.listdir(exdir) if f.endswith("_grids.npy"))
    live_png = ["insane_case2", "obstacle_on_path", "outrageous_case", "right_turn_on_path", "right_turn"]
    # the fixture loader, executed from the reference source with the reference's own models
    src_path = os.path.join(refharness.REFERENCE_ROOT, "utilities/generate_testing_grids/run_on_main.py")
    src = open(src_path).read()
    start = src.index("def convert_npy_to_grid_info")
    end = src.index("class SingleSavedFrameFrameProcessor")
    ns = {"np": np, "Coordinate": ref.models.Coordinate, "Grid": ref.models.Grid, "print": lambda *a, **k: None}
    exec(compile(src[start:end], src_path, "exec"), ns)
    convert = ns["convert_npy_to_grid_info"]
    out = {"names": np.array(names), "live_png": np.array(live_png)}
    for nm in names:
        path = os.path.join(exdir, nm + "_grids.npy")
        g = np.load(path)
        out[f"{nm}/grid"] = g.astype(np.uint8)
        fp = refharness.new_frame_processor(ref)
        H, W = g.shape[0] * 20, g.shape[1] * 20
        fp.frame = np.zeros((H, W, 3), np.uint8)
        fp.grids, fp.grid_lookup = convert(path)
        # run_on_main never sets np_grids -> pure traversal (use_easy=False in the oracle)
        a = ref_state_arrays(fp, ref, H, W)
        out[f"{nm}/pen_traversal"] = a["pen"][:a["R"], :a["C"]]
        out[f"{nm}/occ"] = a["occ"][:a["R"], :a["C"]]
        out[f"{nm}/rows_y"] = a["rows_y"][:a["R"]]
        out[f"{nm}/rows_attr"] = a["rows_attr"][:a["R"]]
        out[f"{nm}/peaks"] = a["peaks"][:a["npk"]]
        out[f"{nm}/start"] = a["start"]
        out[f"{nm}/goals"] = a["goals"][:a["npk"]]
        out[f"{nm}/nbr"] = a["nbr"][:a["R"], :a["C"]]
        # also with easy segments (what FrameProcessor.__call__ would do with np_grids set)
        fp2 = refharness.new_frame_processor(ref)
        fp2.frame = fp.frame
        fp2.grids, fp2.grid_lookup = convert(path)
        fp2.np_grids = np.array([[0 if q.empty else 1 for q in row] for row in fp2.grids], dtype=np.uint8)
        a2 = ref_state_arrays(fp2, ref, H, W)
        out[f"{nm}/pen_easy"] = a2["pen"][:a2["R"], :a2["C"]]
        if nm in live_png:
            img = cv2.imread(os.path.join(exdir, "outputs", nm + "_processed.png"))
            assert img is not None and img.shape[:2] == (H, W), (nm, None if img is None else img.shape)
            lut = ref.config.penalty_colour_gradient
            keys = list(lut.keys())
            assert [round(k, 4) for k in keys] == LUT_KEYS
            col = np.full(g.shape, -1, np.int8)
            for r in range(a["R"]):
                for c in range(a["C"]):
                    if a["occ"][r, c] & 1:
                        y, x = a["rows_y"][r] + 10, a["x0"] + c * 20 + 10
                        bgr = tuple(int(v) for v in img[y, x])
                        hits = [i for i, k in enumerate(keys) if lut[k] == bgr]
                        assert len(hits) == 1, (nm, r, c, bgr)
                        col[r, c] = hits[0]
            out[f"{nm}/png_colour_idx"] = col
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **out)
    print("fixtures.npz:", len(names), "fixtures,", len(live_png), "live PNG answers")


def gen_polygons(ref, n_cases=172):
    rng = np.random.default_rng(20261018)
    cfgs = [(640, 640, 20)] * 110 + [(720, 1280, 20)] * 20 + [(640, 640, 16)] * 10 + \
           [(640, 640, 32)] * 10 + [(384, 640, 8)] * 10 + [(650, 650, 20)] * 12
    out = {}
    keep = dict(H=[], W=[], gs=[], err=[], R=[], C=[], x0=[], npk=[])
    arrs = dict(rows_y=[], rows_attr=[], occ=[], pen=[], peaks=[], start=[], goals=[], nbr=[])
    polys_all, poly_off = [], [0]
    case_poly = []
    for H, W, gs in cfgs[:n_cases]:
        for modname in ("FrameProcessor", "PenaltyCalculator", "ProtrusionDetector", "models"):
            setattr(getattr(ref, modname), "grid_size", gs)        # config.grid_size is imported by value
        k = int(rng.integers(1, 4))
        polys = [polygen.random_polygon(rng, H, W) for _ in range(k)]
        fp = refharness.new_frame_processor(ref)
        fp.frame = np.zeros((H, W, 3), np.uint8)
        err = 0
        try:
            fp._extract_grid_information([refharness.FakeResult(polys)])
            a = ref_state_arrays(fp, ref, H, W)
        except IndexError as e:
            err = 1 if "out of bounds" in str(e) else 2
            fp.grids = []
            a = ref_state_arrays(fp, ref, H, W)
        case_poly.append(k)
        for p in polys:
            polys_all.append(p)
            poly_off.append(poly_off[-1] + len(p))
        for key, v in dict(H=H, W=W, gs=gs, err=err, R=a["R"], C=a["C"], x0=a["x0"], npk=a["npk"]).items():
            keep[key].append(v)
        for key in arrs:
            arrs[key].append(a[key if key != "pen" else "pen"])
    for modname in ("FrameProcessor", "PenaltyCalculator", "ProtrusionDetector", "models"):
        setattr(getattr(ref, modname), "grid_size", 20)
    out.update({k: np.array(v, np.int32) for k, v in keep.items()})
    out.update({k: np.stack(v) for k, v in arrs.items()})
    out["poly_pts"] = np.concatenate(polys_all).astype(np.float32)
    out["poly_off"] = np.array(poly_off, np.int64)
    out["case_npoly"] = np.array(case_poly, np.int32)
    np.savez_compressed(os.path.join(HERE, "polygons.npz"), **out)
    print("polygons.npz:", len(keep["H"]), "cases; errors:", int((out["err"] > 0).sum()),
          "empty:", int((out["R"] == 0).sum()))


MASK_CASES = [  # (frame_idx, n, H, W, mh, mw, family)
    (0, 4, 160, 160, 40, 40, "sidewalk"),
    (1, 3, 160, 160, 40, 40, "noise"),
    (2, 2, 640, 640, 160, 160, "sidewalk"),
    (3, 2, 640, 640, 160, 160, "noise"),
    (4, 3, 270, 480, 40, 40, "noise"),          # anisotropic non-integer scale (6.75 x 12)
    (5, 2, 384, 640, 96, 160, "sidewalk"),
]


def gen_mask_assembly(ref):
    out = {"cases": np.array([c[:6] for c in MASK_CASES], np.int32),
           "families": np.array([c[6] for c in MASK_CASES])}
    for i, (f, n, H, W, mh, mw, fam) in enumerate(MASK_CASES):
        p, c, b = synth.make_frame(f, n, H, W, mh, mw, 32, fam)
        m = ref.ops.process_mask(p, c, b, (H, W), upsample=True)
        lo = ref.ops.process_mask(p, c, b, (H, W), upsample=False)          # sign of cropped logits
        cl = ref.ops.crop_mask((c @ p.view(32, -1)).view(-1, mh, mw),
                               b * torch.tensor([mw / W, mh / H, mw / W, mh / H]))
        out[f"{i}/masks_packed"] = np.packbits(m.numpy().astype(np.uint8), axis=-1)
        out[f"{i}/cropped_logits"] = cl.numpy()
        out[f"{i}/lowres_packed"] = np.packbits(lo.numpy().astype(np.uint8), axis=-1)
    np.savez_compressed(os.path.join(HERE, "mask_assembly.npz"), **out)
    print("mask_assembly.npz:", len(MASK_CASES), "cases")


FRAME_CASES = [(640, 640, 160, 160, 8, 20, "sidewalk", 100, 24),
               (640, 640, 160, 160, 3, 20, "noise", 200, 6),
               (1080, 1920, 160, 160, 4, 20, "sidewalk", 300, 2)]


def gen_frames(ref):
    out = {"cases": np.array([c[:6] + c[7:] for c in FRAME_CASES], np.int32),
           "families": np.array([c[6] for c in FRAME_CASES])}
    for ci, (H, W, mh, mw, n, gs, fam, first, count) in enumerate(FRAME_CASES):
        acc = dict(R=[], C=[], x0=[], npk=[], rows_y=[], rows_attr=[], occ=[], pen=[], peaks=[], start=[], goals=[],
                   nbr=[], areas=[])
        for f in range(first, first + count):
            p, c, b = synth.make_frame(f, n, H, W, mh, mw, 32, fam)
            masks = ref.ops.process_mask(p, c, b, (H, W), upsample=True)
            segs = ref.ops.masks2segments(masks)
            xy = [ref.ops.scale_coords((H, W), s, (H, W), normalize=False) for s in segs]
            fp = refharness.new_frame_processor(ref)
            fp.frame = np.zeros((H, W, 3), np.uint8)
            fp._extract_grid_information([refharness.FakeResult(xy)])
            a = ref_state_arrays(fp, ref, H, W)
            for k in ("R", "C", "x0", "npk", "rows_y", "rows_attr", "occ", "pen", "peaks", "start", "goals", "nbr"):
                acc[k].append(a[k])
            acc["areas"].append(masks.reshape(n, -1).sum(1).numpy().astype(np.int64))
        for k, v in acc.items():
            out[f"{ci}/{k}"] = np.array(v) if np.ndim(v[0]) == 0 else np.stack(v)
    np.savez_compressed(os.path.join(HERE, "frames.npz"), **out)
    print("frames.npz:", sum(c[-1] for c in FRAME_CASES), "frames")


NMS_CASES = [  # (first_idx, B, A, nc, n_objects, ties, conf_thres, iou_thres, max_det)
    (0, 4, 8400, 1, 6, False, 0.5, 0.7, 300),
    (10, 3, 8400, 1, 12, True, 0.5, 0.7, 300),
    (20, 3, 2100, 3, 8, False, 0.5, 0.45, 300),
    (30, 2, 8400, 2, 10, True, 0.5, 0.6, 8),
    (40, 2, 300, 1, 0, False, 0.5, 0.7, 300),       # nothing above the threshold
]


def gen_nms(ref):
    """the vendored ops.non_max_suppression (with torchvision.ops.nms) on synthetic head outputs."""
    out = {"cases": np.array([[c[0], c[1], c[2], c[3], c[4], int(c[5]), c[8]] for c in NMS_CASES], np.int32),
           "thres": np.array([[c[6], c[7]] for c in NMS_CASES], np.float32)}
    total = 0
    for ci, (first, B, A, nc, nobj, ties, ct, it, md) in enumerate(NMS_CASES):
        pred = synth.make_head_output(first, B, A=A, nc=nc, n_objects=nobj, ties=ties)
        res = ref.ops.non_max_suppression(pred.clone(), conf_thres=ct, iou_thres=it, nc=nc, max_det=md)
        for b, r in enumerate(res):
            out[f"{ci}/{b}"] = r.numpy().astype(np.float32)
            total += r.shape[0]
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **out)
    print("nms.npz:", len(NMS_CASES), "cases,", total, "kept rows")


CFG0_FRAMES, CFG0_EVERY, CFG0_SEED = 30, 5, 9000


def cfg0_clip_frame(i: int, H: int = 640, W: int = 640) -> np.ndarray:
    """Deterministic synthetic BGR frame i of the MockCamera clip (a moving gradient with the index stamped in)."""
    ys, xs = np.mgrid[0:H, 0:W]
    f = np.stack([(xs + 7 * i) % 256, (ys + 3 * i) % 256, (xs + ys) // 5 % 256], -1).astype(np.uint8)
    cv2.putText(f, f"frame {i}", (40, 80), cv2.FONT_HERSHEY_SIMPLEX, 2.0, (255, 255, 255), 3)
    return f


def write_cfg0_clip(path: str, H: int = 640, W: int = 640) -> None:
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30.0, (W, H))
    assert vw.isOpened()
    for i in range(CFG0_FRAMES):
        vw.write(cfg0_clip_frame(i, H, W))
    vw.release()


def gen_cfg0(ref):
    """BASELINE configs[0]: a 30-frame 640x640 MJPG clip read through the reference's MockCamera; every 5th frame goes
    through the UNMODIFIED reference FrameProcessor.__call__ (model = polygons of the reference's own mask assembly of
    synthetic head tensors, seed 9000 + frame index).  Stored per processed frame: a checksum of the decoded frame, the
    final answer, the grid state and the A* paths (cells + costs, after the similarity filter) - what the drop-in must
    reproduce on the GPU box from the same clip and the same head tensors."""
    import tempfile
    import zlib
    H = W = 640
    d = tempfile.mkdtemp(prefix="va_cfg0_")
    clip = os.path.join(d, "clip.avi")
    write_cfg0_clip(clip, H, W)
    mc_spec = importlib.util.spec_from_file_location("va_ref_mockcamera", os.path.join(refharness.REFERENCE_ROOT, "MockCamera.py"))
    mc = importlib.util.module_from_spec(mc_spec)
    mc_spec.loader.exec_module(mc)
    cam = mc.MockCamera(clip, target_fps=10000)
    assert cam.isOpened() and (cam.frame_width, cam.frame_height) == (W, H)
    ref.PathFinder.path_finder.angle_cache.clear()
    out = {"meta": np.array([CFG0_FRAMES, CFG0_EVERY, CFG0_SEED, H, W], np.int32)}
    idx, k = 0, 0
    answers = []
    while True:
        ret, frame = cam.read()
        if not ret:
            break
        if idx % CFG0_EVERY == 0:
            p, c, b = synth.make_frame(CFG0_SEED + idx, 8, H, W, 160, 160)
            masks = ref.ops.process_mask(p, c, b, (H, W), upsample=True)
            xy = [ref.ops.scale_coords((H, W), s_, (H, W), normalize=False) for s_ in ref.ops.masks2segments(masks)]
            fp = refharness.new_frame_processor(ref, refharness.FakeModel(xy))
            fp.frame = frame
            fp._extract_grid_information(fp.model.predict(frame))
            a = ref_state_arrays(fp, ref, H, W)
            graph = fp._create_graph()
            peaks = fp.protrusion_detector(frame, fp.grids, fp.grid_lookup)
            paths = fp._find_paths(peaks, graph)
            answer = ref.PathAnalyser.path_analyser(H, W, paths)
            answers.append(str(answer))
            cells = [np.array([(g.coords.x, g.coords.y) for g in pth.grids], np.int32) for pth in paths]
            out[f"{k}/frame_index"] = np.int32(idx)
            out[f"{k}/frame_crc"] = np.uint32(zlib.crc32(frame.tobytes()))
            out[f"{k}/n_paths"] = np.int32(len(paths))
            out[f"{k}/path_len"] = np.array([len(c_) for c_ in cells], np.int32)
            out[f"{k}/path_cells"] = np.concatenate(cells, 0) if cells else np.zeros((0, 2), np.int32)
            out[f"{k}/path_cost"] = np.array([pth.total_cost for pth in paths], np.float64)
            for key in ("R", "C", "x0", "npk", "rows_y", "rows_attr", "occ", "pen", "peaks", "start", "goals"):
                out[f"{k}/{key}"] = a[key]
            k += 1
        idx += 1
    cam.release()
    out["n"] = np.int32(k)
    out["answers"] = np.array(answers)
    np.savez_compressed(os.path.join(HERE, "cfg0.npz"), **out)
    print("cfg0.npz:", k, "frames processed of", idx, "answers:", answers)


if __name__ == "__main__":
    ref = refharness.load()
    torch.set_num_threads(1)
    gen_fixtures(ref)
    gen_polygons(ref)
    gen_mask_assembly(ref)
    gen_frames(ref)
    gen_nms(ref)
    gen_cfg0(ref)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
