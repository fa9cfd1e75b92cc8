"""cProfile of the drop-in FrameProcessor.__call__ on one synthetic frame (GPU box): where the host time goes."""
import cProfile, io, os, pstats, sys, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_assist_b200 import synth
from vision_assist_b200.FrameProcessor import FrameProcessor, HeadOutputModel

H = W = 640
p, c, b = synth.make_frame(424242, 8, H, W, 160, 160)
p1, c1, b1 = p.cuda(), c.cuda(), b.cuda()
fp = FrameProcessor(HeadOutputModel(lambda frame: (p1, c1, b1)))
frame = np.zeros((H, W, 3), np.uint8)
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(20):
        fp(frame)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        fp(frame)
    pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
