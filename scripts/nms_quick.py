"""Device time of va_nms and of the chain va_nms -> va_run_fused (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine
B = 256
eng = MaskGridEngine(H=640, W=640, mh=160, mw=160, max_n=8, gs=20, max_batch=B)
pred = synth.make_head_output(0, 32, A=8400, nc=1, n_objects=6).repeat(B // 32, 1, 1).contiguous().cuda()
protos = synth.make_batch(0, 32, 8, 640, 640, 160, 160, max_n=8)[0].repeat(B // 32, 1, 1, 1).contiguous().cuda()
masks = torch.empty((B, 8, 640, 640), dtype=torch.uint8, device="cuda")
rec = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
for _ in range(5):
    coefs, boxes, conf, cls, counts = eng.nms(pred)
    eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=rec)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
K = 50
e[0].record()
for _ in range(K):
    coefs, boxes, conf, cls, counts = eng.nms(pred)
e[1].record()
for _ in range(K):
    coefs, boxes, conf, cls, counts = eng.nms(pred)
    eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=rec)
e[2].record()
torch.cuda.synchronize()
nms_ms = e[0].elapsed_time(e[1]) / K
chain_ms = e[1].elapsed_time(e[2]) / K
inb = pred.numel() * 4 / 1e6
print(f"va_nms: {nms_ms:.4f} ms per {B} images ({inb:.0f} MB of head output -> {inb / nms_ms:.0f} GB/s); "
      f"nms + run: {chain_ms:.4f} ms -> {B / chain_ms * 1e3:.0f} frames/s; mean kept {float(counts.float().mean()):.1f}")
