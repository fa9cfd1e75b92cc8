"""Small driver for ncu captures (GPU box): a few va_run_fused calls on the cfg1 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine

B = int(os.environ.get("VA_B", "256")); n = int(os.environ.get("VA_N", "8"))
M = int(os.environ.get("VA_M", "160")); S = 4 * M
HH, WW = (int(v) for v in os.environ.get("VA_HW", f"{S}x{S}").split("x"))
tc = os.environ.get("VA_TC", "1") == "1"
eng = MaskGridEngine(H=HH, W=WW, mh=M, mw=M, max_n=n, gs=int(os.environ.get("VA_GS", "20")), max_batch=B, tensor_core=tc)
hp, hc, hb, hn = synth.make_batch(0, 32, n, HH, WW, M, M, max_n=n)
reps = B // 32
protos = hp.repeat(reps, 1, 1, 1).cuda(); coefs = hc.repeat(reps, 1, 1).cuda(); boxes = hb.repeat(reps, 1, 1).cuda(); counts = hn.repeat(reps).cuda()
masks = torch.empty((B, n, HH, WW), dtype=torch.uint8, device="cuda")
rec = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
for _ in range(int(os.environ.get("VA_ITERS", "4"))):
    eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=rec)
torch.cuda.synchronize()
print("ok tc=", eng.uses_tensor_core)
