"""Repeatability soak with a diagnosis (GPU box): which bytes of which frame's record differ between identical calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_assist_b200 import synth
from vision_assist_b200.engine import MaskGridEngine

H, W, mh, mw, n, B = 1080, 1920, 160, 160, 32, 8
eng = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=n, gs=20, max_batch=B)
hp, hc, hb, hn = synth.make_batch(4242, B, n, H, W, mh, mw, max_n=n)
dev = [t.contiguous().cuda() for t in (hp, hc, hb, hn)]
rec0, masks0 = eng.run(*dev)
rec0, masks0 = rec0.clone(), masks0.clone()
rec = torch.empty_like(rec0); masks = torch.empty_like(masks0)
bad = 0
for it in range(int(os.environ.get("VA_SOAK", "400"))):
    wm = it % 3 != 2
    eng.run(*dev, masks_out=masks, records_out=rec, write_masks=wm)
    if not torch.equal(rec, rec0):
        bad += 1
        d = (rec != rec0).cpu().numpy()
        for b in np.nonzero(d.any(1))[0]:
            off = np.nonzero(d[b])[0]
            a, g = eng.decode(rec0[b:b + 1])[0], eng.decode(rec[b:b + 1])[0]
            print(f"iter {it} write_masks={wm} frame {b}: {len(off)} bytes differ, first offsets {off[:6].tolist()}; "
                  f"sel {a.sel}->{g.sel} flags {a.flags}->{g.flags} area2 {a.contour_area2}->{g.contour_area2} bbox {a.bbox}->{g.bbox} R {a.R}->{g.R}", flush=True)
        if bad >= 3:
            break
    if wm and it % 20 == 0 and not torch.equal(masks, masks0):
        print(f"iter {it}: masks differ", flush=True)
print("done, bad iterations:", bad)
