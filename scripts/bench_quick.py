"""Quick device-timing loop for tuning (GPU box): prints ms per 256-frame step for the fused path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine
B, n = 256, 8
eng = MaskGridEngine(H=640, W=640, mh=160, mw=160, max_n=n, gs=20, max_batch=B)
hp, hc, hb, hn = synth.make_batch(0, 64, n, 640, 640, 160, 160, max_n=n)
protos = hp.repeat(4, 1, 1, 1).cuda(); coefs = hc.repeat(4, 1, 1).cuda(); boxes = hb.repeat(4, 1, 1).cuda(); counts = hn.repeat(4).cuda()
masks = torch.empty((B, n, 640, 640), dtype=torch.uint8, device="cuda")
rec = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
for _ in range(5): eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=rec)
torch.cuda.synchronize()
if os.environ.get('VA_PROFILE', '1') == '1': eng.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 100
for _ in range(K): eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=rec)
e1.record(); torch.cuda.synchronize()
a, t, c = eng.profile_read() if os.environ.get('VA_PROFILE', '1') == '1' else (0.0, 0.0, 1)
print(f"{os.environ.get('TAG','')}: step {e0.elapsed_time(e1)/K:.4f} ms  fused {a/c:.4f} ms  tail {t/c:.4f} ms  -> {B*K/e0.elapsed_time(e1)*1000:.0f} fps")
