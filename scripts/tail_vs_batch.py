"""Tail-stage time against the batch size (GPU box): shows how the tail kernel's CTAs are scheduled over the SMs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine

n, M, S = 8, 160, 640
hp, hc, hb, hn = synth.make_batch(0, 32, n, S, S, M, M, max_n=n)
for B in (32, 64, 128, 148, 160, 192, 256, 296, 320):
    eng = MaskGridEngine(H=S, W=S, mh=M, mw=M, max_n=n, gs=20, max_batch=B)
    reps = (B + 31) // 32
    t = [x.repeat(reps, *([1] * (x.dim() - 1)))[:B].contiguous().cuda() for x in (hp, hc, hb, hn)]
    masks = torch.empty((B, n, S, S), dtype=torch.uint8, device="cuda")
    rec = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
    for _ in range(5):
        eng.run(*t, masks_out=masks, records_out=rec)
    torch.cuda.synchronize()
    eng.profile(True)
    for _ in range(40):
        eng.run(*t, masks_out=masks, records_out=rec)
    a, tl, c = eng.profile_read()
    print(f"B={B:4d}: mask kernel {1e3 * a / c:7.1f} us   tail {1e3 * tl / c:6.1f} us", flush=True)
    del eng, masks, rec
