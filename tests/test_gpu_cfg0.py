"""BASELINE configs[0] on the GPU box: a 30-frame 640x640 MJPG clip -> frame source with MockCamera's `read()` contract
-> drop-in FrameProcessor (model shim = synthetic YOLOv8-seg head tensors, one launch sequence per frame) -> array A*.

Expected values (tests/golden/cfg0.npz) were produced in the build container by tests/golden/make_golden.py::gen_cfg0:
the SAME clip read through the reference's own MockCamera (MockCamera.py:32-54) and pushed through the UNMODIFIED
reference FrameProcessor (FrameProcessor.py:301-360) with the polygons of the reference's own mask assembly of the same
head tensors: grid, penalties, peaks, start / end cells, the A* paths that survive the similarity filter and their
costs.  /root/reference does not exist on the GPU box, hence golden vectors; tests/test_cfg0_cpu.py is the CPU half."""
import os
import sys
import zlib

import cv2
import numpy as np
import pytest

import goldenio
from cfg0common import assert_paths_match, golden_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))

pytestmark = pytest.mark.gpu

from vision_assist_b200 import synth  # noqa: E402
from vision_assist_b200.FrameProcessor import FrameProcessor, HeadOutputModel  # noqa: E402
from vision_assist_b200.PathFinder import path_finder as array_path_finder  # noqa: E402


class ClipCamera:
    """The reference MockCamera's contract (MockCamera.py:32-54: `read() -> (ret, frame)`, `isOpened()`, `release()`)
    without its real-time throttle."""

    def __init__(self, path):
        self.cap = cv2.VideoCapture(str(path))

    def isOpened(self):
        return self.cap.isOpened()

    def read(self):
        return self.cap.read()

    def release(self):
        self.cap.release()


def test_cfg0_mockcamera_clip_through_dropin(tmp_path):
    import make_golden as mg          # only the clip writer is used (pure OpenCV / numpy); the reference is not imported
    z = goldenio.load("cfg0.npz")
    n_frames, every, seed, H, W = (int(v) for v in z["meta"])
    clip = tmp_path / "clip.avi"
    mg.write_cfg0_clip(str(clip), H, W)
    cam = ClipCamera(clip)
    assert cam.isOpened()
    state = {"idx": 0}

    def head_tensors(frame):
        p, c, b = synth.make_frame(seed + state["idx"], 8, H, W, 160, 160)
        return p.cuda(), c.cuda(), b.cuda()

    FrameProcessor._instance, FrameProcessor._initialized = None, False
    fp = FrameProcessor(HeadOutputModel(head_tensors), verbose=False, debug=False)
    array_path_finder.angle_cache.clear()        # the golden sequence started from an empty cache
    k = idx = 0
    try:
        while True:
            ret, frame = cam.read()
            if not ret:
                break
            if idx % every == 0:
                assert int(z[f"{k}/frame_index"]) == idx
                if zlib.crc32(frame.tobytes()) != int(z[f"{k}/frame_crc"]):
                    pytest.skip("this OpenCV build decodes the MJPG clip differently from the build container's")
                state["idx"] = idx
                peaks = fp(frame)
                rec = fp.frame_record
                case = golden_case(z, k)
                goldenio.assert_result_matches(rec.as_dict(), case, f"cfg0 frame {idx}")
                assert [(p.x, p.y) for p in peaks] == [tuple(int(v) for v in q) for q in case["peaks"]]
                # the object view the reference's host stages consume is built on first access only
                assert fp._pending_record is rec and np.array_equal(fp.np_grids, case["occ"] & 1)
                assert len(fp.grids) == case["R"] and fp._pending_record is None
                assert fp.protrusion_detector.grids is fp.grids
                assert all(fp.grid_lookup[(g.coords.x, g.coords.y)].col == g.col for g in fp.grids[-1])
                # A* paths and costs: array port on the GPU record vs the reference's PathFinder + similarity filter
                assert_paths_match(z, k, fp.array_paths, f"cfg0 frame {idx}")
                k += 1
            idx += 1
    finally:
        cam.release()
        FrameProcessor._instance, FrameProcessor._initialized = None, False
    assert idx == n_frames and k == int(z["n"])
