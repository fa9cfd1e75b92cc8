"""BASELINE configs[0], CPU half: the oracle's record of every processed clip frame + the array A* port reproduce the
golden grid state, paths and costs that the UNMODIFIED reference FrameProcessor produced behind its own MockCamera
(tests/golden/make_golden.py::gen_cfg0).  The GPU half (tests/test_gpu_cfg0.py) runs the drop-in on the same clip."""
import goldenio
from cfg0common import assert_paths_match, golden_case, similarity_filter
from oracle import pipeline as opl
from test_dropin_cpu import record_from_oracle
from vision_assist_b200 import synth
from vision_assist_b200.PathFinder import ArrayPathFinder


def test_cfg0_oracle_and_array_astar_vs_reference_golden():
    z = goldenio.load("cfg0.npz")
    n_frames, every, seed, H, W = (int(v) for v in z["meta"])
    finder = ArrayPathFinder()                 # the golden sequence started from an empty angle cache
    assert int(z["n"]) == (n_frames + every - 1) // every
    assert set(z["answers"].tolist()) <= {"move_left", "move_right", "continue_forward"}
    for k in range(int(z["n"])):
        idx = int(z[f"{k}/frame_index"])
        p, c, b = synth.make_frame(seed + idx, 8, H, W, 160, 160)
        res = opl.frame_from_tensors(p, c, b, (H, W), 20, "contour")
        goldenio.assert_result_matches(res, golden_case(z, k), f"cfg0 frame {idx}")
        paths = similarity_filter(finder.find_paths(record_from_oracle(res), 20))
        assert_paths_match(z, k, paths, f"cfg0 frame {idx}")
