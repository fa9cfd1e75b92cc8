"""Which record bytes does a call leave unwritten?  (GPU box)  Runs the same batch into a 0x00- and a 0xAB-filled buffer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine

for (H, W, n, B) in ((1080, 1920, 32, 8), (640, 640, 8, 32)):
    eng = MaskGridEngine(H=H, W=W, mh=160, mw=160, max_n=n, gs=20, max_batch=B)
    hp, hc, hb, hn = synth.make_batch(4242, B, n, H, W, 160, 160, max_n=n)
    dev = [t.contiguous().cuda() for t in (hp, hc, hb, hn)]
    a = torch.zeros((B, eng.record_bytes), dtype=torch.uint8, device="cuda")
    b = torch.full((B, eng.record_bytes), 0xAB, dtype=torch.uint8, device="cuda")
    for wm in (True, False):
        eng.run(*dev, records_out=a, write_masks=wm)
        eng.run(*dev, records_out=b, write_masks=wm)
        d = (a != b).cpu().numpy()
        L = eng.layout
        print(f"{H}x{W} n={n} write_masks={wm}: unwritten bytes per frame {d.sum(1).tolist()}")
        if d.any():
            fr = int(np.nonzero(d.any(1))[0][0])
            off = np.nonzero(d[fr])[0]
            names = [(k, getattr(L, k)) for k in ("off_header", "off_row_y", "off_row_attr", "off_penalty", "off_peaks", "off_occ", "off_goals", "off_lookup")]
            print("  frame", fr, "first offsets", off[:12].tolist(), "last", off[-3:].tolist(), "sections", names, "record_bytes", eng.record_bytes)
            r = eng.decode(a[fr:fr + 1])[0]
            print("  R", r.R, "C", r.C, "flags", r.flags, "n_peaks", len(r.peaks))
