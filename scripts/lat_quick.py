"""B=1 latency breakdown (GPU box): device time of the assembly and tail kernels + host-visible latency."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine
eng = MaskGridEngine(H=640, W=640, mh=160, mw=160, max_n=8, gs=20, max_batch=1)
hp, hc, hb, hn = synth.make_batch(424242, 1, 8, 640, 640, 160, 160, max_n=8)
protos, coefs, boxes, counts = hp.cuda(), hc.cuda(), hb.cuda(), hn.cuda()
rec = torch.empty((1, eng.record_bytes), dtype=torch.uint8, device="cuda")
hrec = torch.empty((1, eng.record_bytes), dtype=torch.uint8, pin_memory=True)
for _ in range(20):
    eng.run(protos, coefs, boxes, counts, records_out=rec, write_masks=False)
torch.cuda.synchronize()
prof = os.environ.get('VA_PROFILE', '1') == '1'
if prof: eng.profile(True)
ts = []
for _ in range(300):
    t0 = time.perf_counter()
    eng.run(protos, coefs, boxes, counts, records_out=rec, write_masks=False)
    hrec.copy_(rec, non_blocking=True)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e6)
a, t, c = eng.profile_read() if prof else (0.0, 0.0, 1)
ts.sort()
print(f"{os.environ.get('TAG','')}: p50 {ts[150]:.1f} us  p99 {ts[296]:.1f} us | device: assemble {1e3*a/c:.1f} us  tail {1e3*t/c:.1f} us")
