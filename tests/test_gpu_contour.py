"""GPU parity of the contour step: the CUDA path against the reference's own OpenCV route
(masks2segments -> contourArea selection -> boundingRect / fillPoly, vendored ops.py:837-859 and FrameProcessor.py:72-86)
on masks that are NOT single hole-free blobs: several components, rings, islands inside holes, area near-ties where
pixel area and contourArea order the instances differently, empty and single-pixel masks.  Both entry points:
caller-provided masks (va_mask_to_records) and the fused path from prototypes (va_run_fused, u8 and grid-only)."""
import cv2
import numpy as np
import pytest
import torch

from gpucommon import assert_record_equals_oracle, to_dev
from oracle import contour as ocontour
from oracle import pipeline as opl

pytestmark = pytest.mark.gpu

from vision_assist_b200.engine import MaskGridEngine  # noqa: E402


def _rand_mask(rng, h, w):
    z = rng.standard_normal((h // 8 + 2, w // 8 + 2)).astype(np.float32)
    z = cv2.resize(z, (w, h), interpolation=cv2.INTER_CUBIC)
    return (z > rng.uniform(-0.5, 0.5)).astype(np.uint8)


def _adversarial_frames(H, W, n):
    """-> list of uint8 [n, H, W] frames."""
    rng = np.random.default_rng(99)
    frames = []

    def blank():
        return np.zeros((n, H, W), np.uint8)

    f = blank(); cv2.circle(f[0], (W // 2, H // 2), H // 4, 1, 9); frames.append(f)                       # ring
    f = blank(); cv2.circle(f[0], (W // 2, H // 2), H // 4, 1, 9); cv2.circle(f[0], (W // 2, H // 2), 20, 1, -1); frames.append(f)   # + island
    f = blank(); f[0, 40:140, 40:140] = 1; f[0, 300:420, 300:420] = 1; frames.append(f)                   # two blobs, one instance
    f = blank(); f[0, 40:140, 40:140] = 1; f[0, 300:400, 300:400] = 1; frames.append(f)                   # equal blobs: tie on points
    # near-tie between instances: a ring (fewer pixels, larger contourArea) against a solid blob
    f = blank(); cv2.circle(f[0], (200, 200), 100, 1, 12); cv2.circle(f[1], (400, 400), 95, 1, -1); frames.append(f)
    f = blank(); cv2.circle(f[1], (200, 200), 100, 1, 12); cv2.circle(f[0], (400, 400), 95, 1, -1); frames.append(f)
    f = blank(); f[0, 100:200, 100:200] = 1; f[1, 300:400, 300:400] = 1; frames.append(f)                 # exact tie: first instance wins
    f = blank(); f[1, 100:200, 100:200] = 1; frames.append(f)                                             # instance 0 empty, a later one not
    f = blank(); frames.append(f)                                                                         # all empty: cv2.error in the reference
    f = blank(); f[0, 333, 222] = 1; frames.append(f)                                                     # single pixel
    f = blank(); f[0, 100, 50:300] = 1; f[1, 200:205, 200:205] = 1; frames.append(f)                      # zero-area line vs small blob
    f = blank(); f[0] = 1; f[0, 1:-1, 1:-1] = 0; frames.append(f)                                         # ring hugging the frame
    f = blank(); f[0] = 1; frames.append(f)                                                               # full frame
    f = blank(); f[0, ::2, ::4] = 1; frames.append(f)                 # isolated pixels: H*W/8 components = exactly the run capacity
    f = blank(); f[0] = ((np.indices((H, W)) // 4).sum(0) % 2).astype(np.uint8); frames.append(f)         # 4x4 checkerboard: one 8-connected net, H*W/8 runs
    for _ in range(12):
        f = blank()
        for i in range(n):
            if rng.random() < 0.8:
                f[i] = _rand_mask(rng, H, W)
        frames.append(f)
    # row-convex blobs with a few rows of several runs: notches in the top / bottom row and stubs beside the outline
    # (light path: one component without holes decided from the run ends), slits inside the blob (holes), stubs that
    # touch nothing (islands) and adjacent notched rows (full path) - and a tall one (more rows than threads)
    for k in range(10):
        f = blank()
        for i in range(min(n, 3)):
            y0 = int(rng.integers(5, H // 4)); y1 = int(rng.integers(H // 2, H - 5)) if k != 9 else H - 3
            a, b_ = int(rng.integers(20, W // 3)), int(rng.integers(2 * W // 3, W - 20))
            spans = {}
            for y in range(y0, y1 + 1):
                a = int(np.clip(a + rng.integers(-3, 4), 8, W // 2 - 10)); b_ = int(np.clip(b_ + rng.integers(-3, 4), W // 2 + 10, W - 9))
                f[i, y, a:b_ + 1] = 1
                spans[y] = (a, b_)
            for y in ([y0], [y1], [y0, y1])[k % 3]:
                a, b_ = spans[y]
                step = 1 if y == y0 else -1
                depth = int(rng.integers(1, 7))                 # notches several rows deep (bands of adjacent multi-run rows)
                for _ in range(int(rng.choice([1, 1, 2, 3]))):
                    c0 = int(rng.integers(a + 1, b_)); c1 = min(b_ - 1, c0 + int(rng.integers(0, 9)))
                    dd = depth if rng.random() < 0.6 else int(rng.integers(1, 7))
                    for q in range(dd):
                        if c0 > c1:
                            break
                        f[i, y + step * q, c0:c1 + 1] = 0
                        c0 += int(rng.integers(-1, 2)); c1 -= int(rng.integers(0, 2))
            for y in range(y0 + 3, y1 - 3, int(rng.integers(17, 60))):
                a, b_ = spans[y]
                r_ = rng.random()
                if r_ < 0.6:
                    gap, ln = int(rng.integers(1, 4)), int(rng.integers(1, 5))
                    f[i, y, a - gap - ln:a - gap] = 1
                elif r_ < 0.8:
                    f[i, y, (a + b_) // 2:(a + b_) // 2 + 3] = 0                      # a hole
                else:
                    f[i, y:y + 2, b_ - 4] = 0                                        # adjacent rows with several runs
        frames.append(f)
    return frames


@pytest.mark.parametrize("gs", [20, 8])
def test_mask_to_records_adversarial(gs):
    H = W = 640
    n = 3
    frames = _adversarial_frames(H, W, n)
    B = len(frames)
    eng = MaskGridEngine(H=H, W=W, mh=160, mw=160, max_n=n, gs=gs, max_batch=B)
    masks = torch.from_numpy(np.stack(frames)).cuda()
    counts = torch.full((B,), n, dtype=torch.int32).cuda()
    for rep in range(2):                      # twice: the scratch (work list, summaries) must be left re-armed
        recs = eng.decode(eng.masks_to_records(masks, counts))
        for b, f in enumerate(frames):
            want = opl.frame_from_masks(f, gs, "contour")
            assert_record_equals_oracle(recs[b], want, f"adversarial frame {b} rep {rep}")
            lut = opl.frame_from_masks(f, gs, "lut")
            assert bool(recs[b].flags & opl.FLAG_NO_POLYGON) == bool(lut["flags"] & opl.FLAG_NO_POLYGON), b
            assert bool(recs[b].flags & opl.FLAG_NON_SIMPLE) == bool(lut["flags"] & opl.FLAG_NON_SIMPLE), b
            assert recs[b].sel == lut["sel"], (b, recs[b].sel, lut["sel"])
            sel, poly = ocontour.select_instance(f)
            if poly is not None:
                assert recs[b].contour_area2 == poly["area2"], b
                assert recs[b].bbox == poly["bbox"], b


def test_run_capacity_overflow_is_flagged():
    """More than H*W/8 pixel runs in one mask (impossible for 4x-upsampled masks) is reported, not mis-computed."""
    H = W = 640
    eng = MaskGridEngine(H=H, W=W, mh=160, mw=160, max_n=1, gs=20, max_batch=2)
    m = np.zeros((2, 1, H, W), np.uint8)
    m[0, 0, ::2, ::2] = 1
    m[1, 0, 100:300, 100:300] = 1
    recs = eng.decode(eng.masks_to_records(torch.from_numpy(m).cuda(), torch.ones(2, dtype=torch.int32).cuda()))
    assert recs[0].flags & 16 and recs[0].R == 0
    assert_record_equals_oracle(recs[1], opl.frame_from_masks(m[1], 20, "contour"), "frame after an overflow")


def _field_protos(mh, mw, K):
    """Prototype channels whose positive regions are: 0 a ring, 1 two discs, 2 a solid disc slightly smaller than the
    ring's outline, 3 a disc with a thin slit (one component, no hole, not row-convex)."""
    ys, xs = np.mgrid[0:mh, 0:mw].astype(np.float32)
    p = np.full((K, mh, mw), -1.0, np.float32)
    r = np.hypot(xs - 0.35 * mw, ys - 0.4 * mh)
    p[0] = 0.05 * mw - np.abs(r - 0.2 * mw)
    d1 = 0.1 * mw - np.hypot(xs - 0.25 * mw, ys - 0.25 * mh)
    d2 = 0.13 * mw - np.hypot(xs - 0.7 * mw, ys - 0.7 * mh)
    p[1] = np.maximum(d1, d2)
    p[2] = 0.235 * mw - np.hypot(xs - 0.68 * mw, ys - 0.62 * mh)
    p[3] = np.minimum(0.3 * mw - np.hypot(xs - 0.5 * mw, ys - 0.5 * mh), np.maximum(np.abs(xs - 0.5 * mw) - 1.5, ys - 0.45 * mh))
    return p


@pytest.mark.parametrize("tc", [pytest.param(False, id="cuda-core"), pytest.param(True, id="tcgen05")])
def test_fused_path_rings_blobs_and_near_ties(tc):
    H = W = 640
    mh = mw = 160
    K, n = 32, 4
    eng = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=n, gs=20, max_batch=8, tensor_core=tc)
    if tc:
        assert eng.uses_tensor_core
    base = torch.from_numpy(_field_protos(mh, mw, K))
    combos = [[0], [1], [0, 2], [2, 0], [3], [1, 3, 0, 2], [3, 1], [2]]
    B = len(combos)
    protos = base[None].repeat(B, 1, 1, 1).contiguous()
    coefs = torch.zeros(B, n, K)
    boxes = torch.zeros(B, n, 4)
    counts = torch.zeros(B, dtype=torch.int32)
    for b, chans in enumerate(combos):
        counts[b] = len(chans)
        for i, ch in enumerate(chans):
            coefs[b, i, ch] = 1.0
            boxes[b, i] = torch.tensor([0.0, 0.0, W - 1.0, H - 1.0])
    dev = to_dev(protos, coefs, boxes, counts)
    records, masks = eng.run(*dev)
    recs = eng.decode(records)
    masks = masks.cpu().numpy()
    n_general = 0
    for b in range(B):
        nb = int(counts[b])
        want = opl.frame_from_masks(masks[b, :nb], 20, "contour")
        assert want["R"] > 0
        assert_record_equals_oracle(recs[b], want, f"field frame {b}")
        full = opl.frame_from_tensors(protos[b], coefs[b, :nb], boxes[b, :nb], (H, W), 20, "contour")
        assert np.array_equal(masks[b, :nb], full["masks"])
        lut = opl.frame_from_masks(masks[b, :nb], 20, "lut")
        assert recs[b].sel == lut["sel"]
        n_general += bool(recs[b].flags & opl.FLAG_NON_SIMPLE)
    assert n_general >= 4
    # the ring (instance 0) has fewer pixels than the disc (instance 1) but the larger contourArea: selected
    assert recs[2].sel == 0 and recs[3].sel == 1
    # grid-only mode (bit-packed masks in context scratch) gives the same records
    rec2, none = eng.run(*dev, write_masks=False)
    assert none is None and torch.equal(rec2, records)
    rec3, _ = eng.run(*dev)
    assert torch.equal(rec3, records)
