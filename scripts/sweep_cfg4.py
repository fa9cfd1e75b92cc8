#!/usr/bin/env python
"""SURVEY 8(d) cfg4 sweeps, run on the GPU box: frames/s and HBM fraction of va_run_fused for
gs in {4,8,16,20,32} at 640^2/160^2 and proto in {160..320}^2 (H=W=4*mh) x n in {8,32}, plus cfg2.

  python scripts/sweep_cfg4.py > gpurun_out/sweep_cfg4.md

Device-timed with CUDA events over `iters` back-to-back calls after warm-up; inputs + masks of every
point exceed the 126 MB L2.  Prints a markdown table (copied to profiles/ by hand).
"""
import json, os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine

peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
PEAK = 6553.3
try:
    PEAK = float(json.load(open(peaks)).get("hbm_gbs", PEAK))
except Exception:
    pass


def point(H, W, mh, mw, n, gs, B, iters=30, uniq=16):
    eng = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=n, gs=gs, max_batch=B)
    hp, hc, hb, hn = synth.make_batch(0, uniq, n, H, W, mh, mw, max_n=n)
    r = B // uniq
    protos = hp.repeat(r, 1, 1, 1).cuda(); coefs = hc.repeat(r, 1, 1).cuda()
    boxes = hb.repeat(r, 1, 1).cuda(); counts = hn.repeat(r).cuda()
    masks = torch.empty((B, n, H, W), dtype=torch.uint8, device="cuda")
    rec = torch.empty((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=rec)
    torch.cuda.synchronize()
    eng.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        eng.run(protos, coefs, boxes, counts, masks_out=masks, records_out=rec)
    e1.record(); torch.cuda.synchronize()
    a, t, c = eng.profile_read()
    ms = e0.elapsed_time(e1) / iters
    by = eng.algorithmic_bytes_per_frame(n)
    fps = B / ms * 1e3
    path = "tcgen05" if eng.uses_tensor_core else "cuda-core"
    print(f"| {H}x{W} | {mh}x{mw} | {n} | {gs} | {B} | {path} | {ms:.3f} | {a / c:.3f} | {t / c:.3f} | {fps:,.0f} | "
          f"{by / 1e6:.2f} | {100 * fps * by / (PEAK * 1e9):.1f}% |", flush=True)
    del eng, masks, rec, protos
    torch.cuda.empty_cache()


print("| frame | protos | n | gs | B | path | step ms | assemble ms | tail ms | frames/s | MB/frame | HBM frac (step) |")
print("|---|---|---:|---:|---:|---|---:|---:|---:|---:|---:|---:|")
for gs in (4, 8, 16, 20, 32):
    point(640, 640, 160, 160, 8, gs, 256)
for m in (160, 192, 224, 256, 320):
    for n in (8, 16, 32):
        B = 256 if n == 8 else 128 if n == 16 else 64
        if m >= 256:
            B //= 2
        point(4 * m, 4 * m, m, m, n, 20, B)
point(1080, 1920, 160, 160, 32, 20, 32)
