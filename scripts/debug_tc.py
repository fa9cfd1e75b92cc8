"""GPU debug helper (not a test): structured inputs through the tcgen05 path vs the CUDA-core path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vision_assist_b200.engine import MaskGridEngine

H = W = 640; mh = mw = 160; n = 8; B = 2
tc = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=8, gs=20, max_batch=B, tensor_core=True)
cc = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=8, gs=20, max_batch=B, tensor_core=False)
print("uses tc:", tc.uses_tensor_core, tc.lib.va_last_error(tc._ctx))
P = mh * mw
boxes = torch.tensor([0., 0., 640., 640.]).repeat(B, 8, 1).cuda()
counts = torch.full((B,), n, dtype=torch.int32).cuda()

def run(name, protos, coefs):
    protos = protos.contiguous().cuda(); coefs = coefs.contiguous().cuda()
    _, lt = tc.assemble_masks(protos, coefs, boxes, counts, want_logits=True)
    _, lc = cc.assemble_masks(protos, coefs, boxes, counts, want_logits=True)
    torch.cuda.synchronize()
    lt = lt.cpu().numpy(); lc = lc.cpu().numpy()
    d = np.abs(lt - lc)
    print(f"== {name}: tc nonzero frac {np.mean(lt != 0):.4f}  cc nonzero frac {np.mean(lc != 0):.4f}  max|diff| {d.max():.6g}  rel {d.max() / max(np.abs(lc).max(), 1e-30):.3g}")
    print("   tc[0,0,0,:8]  ", lt[0, 0, 0, :8])
    print("   cc[0,0,0,:8]  ", lc[0, 0, 0, :8])
    print("   tc[0,:,0,0]   ", lt[0, :, 0, 0])
    print("   cc[0,:,0,0]   ", lc[0, :, 0, 0])
    print("   tc[0,1,1,30:36]", lt[0, 1, 1, 30:36], " cc", lc[0, 1, 1, 30:36])
    print("   tc[1,2,80,100:104]", lt[1, 2, 80, 100:104], " cc", lc[1, 2, 80, 100:104])
    bad = np.argwhere(d > 1e-3 * max(np.abs(lc).max(), 1e-30))
    print("   #bad", len(bad), "first bad", bad[:5].tolist())
    return lt, lc

ones = torch.ones(B, 32, mh, mw)
c_inst = torch.arange(1, 9).float()[None, :, None].repeat(B, 1, 32)
run("A: protos=1, coef=i+1 (expect 32*(i+1))", ones, c_inst)
pr = torch.zeros(B, 32, P); pr[:, 0] = (torch.arange(P) % 128).float()
c0 = torch.zeros(B, 8, 32); c0[:, :, 0] = 1
run("B: protos[0][p]=p%128, coef[:,0]=1 (expect p%128)", pr.view(B, 32, mh, mw), c0)
pk = torch.arange(32).float()[None, :, None].repeat(B, 1, P)
ck = torch.zeros(B, 8, 32)
for i in range(8): ck[:, i, i * 4] = 1
run("C: protos[k][p]=k, coef[i][4i]=1 (expect 4i)", pk.view(B, 32, mh, mw), ck)
torch.manual_seed(0)
run("D: random", torch.randn(B, 32, mh, mw), torch.randn(B, 8, 32))
