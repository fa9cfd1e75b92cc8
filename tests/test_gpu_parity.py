"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed
golden vectors.  Bit-exact for grids / penalties / peaks; binary masks bit-exact except pixels
whose reference logit is within 1e-4 of the threshold (counted); logits within 1e-3 relative."""
import numpy as np
import pytest
import torch

import goldenio
import polygen
from gpucommon import assert_record_equals_oracle, band_mismatch_report, to_dev
from oracle import grid as og
from oracle import mask_assembly as oma
from oracle import nms as onms
from oracle import penalty as open_
from oracle import pipeline as opl
from oracle import protrusion as oprot

pytestmark = pytest.mark.gpu

from vision_assist_b200 import synth  # noqa: E402
from vision_assist_b200.engine import MaskGridEngine  # noqa: E402

PATHS = [pytest.param(False, id="cuda-core"), pytest.param(True, id="tcgen05")]


def make_engine(tc, **kw):
    e = MaskGridEngine(tensor_core=tc, **kw)
    if tc and not e.uses_tensor_core:
        why = e.lib.va_last_error(e._ctx).decode()
        # documented limits of the tcgen05 kernel: up-sampling scales with W % 16 == 0, mw >= 52 and at
        # most 8 dst rows per prototype row; other configurations run on the CUDA-core contraction (the other id)
        if any(k in why for k in ("needs H=4*mh", "max_n <=", "mw >= 52", "vertical scales", "owns no dst row", "W % 16", "mw % 4")):
            pytest.skip(why)
        pytest.fail("tcgen05 plan unavailable: " + why)
    return e


@pytest.mark.parametrize("tc", PATHS)
@pytest.mark.parametrize("H,W,mh,mw,n,family", [(640, 640, 160, 160, 8, "sidewalk"), (640, 640, 160, 160, 5, "noise"),
                                                (160, 160, 40, 40, 3, "noise"), (384, 640, 96, 160, 2, "sidewalk")])
def test_logits_and_masks_vs_oracle(tc, H, W, mh, mw, n, family):
    B = 3
    eng = make_engine(tc, H=H, W=W, mh=mh, mw=mw, max_n=8, gs=20, max_batch=B)
    protos, coefs, boxes, counts = synth.make_batch(1000, B, n, H, W, mh, mw, family=family, max_n=8)
    masks, logits = eng.assemble_masks(*to_dev(protos, coefs, boxes, counts), want_logits=True)
    torch.cuda.synchronize()
    masks, logits = masks.cpu().numpy(), logits.cpu().numpy()
    total_diff = 0
    for b in range(B):
        cl = oma.cropped_logits(protos[b], coefs[b, :n], boxes[b, :n], (H, W)).numpy()
        got = logits[b, :n]
        assert np.array_equal(cl == 0, got == 0), "crop pattern"
        scale = np.abs(cl).max()
        assert np.abs(got - cl).max() <= 1e-3 * scale, "logits within 1e-3 relative"
        assert np.abs(got - cl).max() <= 2e-5 * scale, "fp32-class accuracy expected"
        up = oma.upsampled_logits(protos[b], coefs[b, :n], boxes[b, :n], (H, W)).numpy()
        nd, nout = band_mismatch_report(masks[b, :n], up)
        assert nout == 0, f"{nout} mask pixels differ outside the 1e-4 band"
        total_diff += nd
        # identical logits in -> bit-identical upsample arithmetic
        spec = (oma.bilinear_upsample_np(got, (H, W)) > 0).astype(np.uint8)
        assert np.array_equal(spec, masks[b, :n]), "upsample arithmetic differs from the specification"
    print(f"[parity] {family} {H}x{W}: {total_diff} mask pixels inside the 1e-4 band differ")


@pytest.mark.parametrize("tc", PATHS)
def test_mask_assembly_golden(tc):
    z = goldenio.load("mask_assembly.npz")
    for i, (f, n, H, W, mh, mw) in enumerate(z["cases"].tolist()):
        fam = str(z["families"][i])
        if H < 8 or W < 16:
            continue
        eng = make_engine(tc, H=H, W=W, mh=mh, mw=mw, max_n=8, gs=20 if min(H, W) >= 80 else 8, max_batch=1)
        protos, coefs, boxes, counts = synth.make_batch(f, 1, n, H, W, mh, mw, family=fam, max_n=8)
        masks = eng.assemble_masks(*to_dev(protos, coefs, boxes, counts)).cpu().numpy()[0, :n]
        gold = np.unpackbits(z[f"{i}/masks_packed"], axis=-1)[..., :W]
        up = oma.bilinear_upsample_np(z[f"{i}/cropped_logits"], (H, W))
        nd, nout = band_mismatch_report(masks, up)
        assert nout == 0 and nd <= 8, (i, nd, nout)
        assert (masks != gold).sum() == nd


@pytest.mark.parametrize("tc", PATHS)
def test_frames_golden(tc):
    z = goldenio.load("frames.npz")
    for ci, (H, W, mh, mw, n, gs, first, count) in enumerate(z["cases"].tolist()):
        fam = str(z["families"][ci])
        eng = make_engine(tc, H=H, W=W, mh=mh, mw=mw, max_n=8, gs=gs, max_batch=count, check_simple=True)
        protos, coefs, boxes, counts = synth.make_batch(first, count, n, H, W, mh, mw, family=fam, max_n=8)
        records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
        recs = eng.decode(records)
        for k, rec in enumerate(recs):
            if fam == "sidewalk":
                assert rec.area == int(z[f"{ci}/areas"][k].max())
            R, C = int(z[f"{ci}/R"][k]), int(z[f"{ci}/C"][k])
            case = dict(R=R, C=C, x0=int(z[f"{ci}/x0"][k]), rows_y=z[f"{ci}/rows_y"][k][:R],
                        rows_attr=z[f"{ci}/rows_attr"][k][:R], occ=z[f"{ci}/occ"][k][:R, :C],
                        pen=z[f"{ci}/pen"][k][:R, :C], peaks=z[f"{ci}/peaks"][k][:int(z[f"{ci}/npk"][k])])
            # golden = the reference's own contour route (masks2segments -> contourArea -> fillPoly): every family,
            # multi-component masks with holes included
            if fam == "sidewalk":
                assert not (rec.flags & opl.FLAG_NON_SIMPLE)
            goldenio.assert_result_matches(rec.as_dict(), case, f"golden frame {ci}/{k}")


@pytest.mark.parametrize("tc", PATHS)
@pytest.mark.parametrize("family", ["sidewalk", "noise"])
def test_run_fused_vs_oracle(tc, family):
    H = W = 640
    B, n = 24, 8
    eng = make_engine(tc, H=H, W=W, mh=160, mw=160, max_n=8, gs=20, max_batch=B, check_simple=True)
    protos, coefs, boxes, counts = synth.make_batch(5000, B, n, H, W, 160, 160, family=family, max_n=8)
    counts[3] = 0          # a frame without detections
    counts[5] = 1
    counts[7] = 3
    records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
    recs = eng.decode(records)
    masks = masks.cpu().numpy()
    n_nonsimple = n_band_frames = 0
    for b in range(B):
        nb = int(counts[b])
        # the reference's contour route (OpenCV) on the SAME masks: bit-exact on every family - components, holes,
        # contour point counts, contourArea selection
        res = opl.frame_from_masks(masks[b, :nb], 20, "contour")
        assert_record_equals_oracle(recs[b], res, f"{family} frame {b}")
        lut = opl.frame_from_masks(masks[b, :nb], 20, "lut")
        assert bool(recs[b].flags & opl.FLAG_NON_SIMPLE) == bool(lut["flags"] & opl.FLAG_NON_SIMPLE), (family, b)
        assert bool(recs[b].flags & opl.FLAG_NO_POLYGON) == bool(lut["flags"] & opl.FLAG_NO_POLYGON), (family, b)
        n_nonsimple += bool(recs[b].flags & opl.FLAG_NON_SIMPLE)
        if nb:
            assert recs[b].sel == lut["sel"], (family, b, recs[b].sel, lut["sel"])
        # and from the tensors: identical whenever no mask pixel sits inside the 1e-4 band
        full = opl.frame_from_tensors(protos[b], coefs[b, :nb], boxes[b, :nb], (H, W), 20, "contour")
        if np.array_equal(masks[b, :nb], full["masks"]):
            assert_record_equals_oracle(recs[b], full, f"{family} frame {b} (from tensors)")
        else:
            n_band_frames += 1
            up = oma.upsampled_logits(protos[b], coefs[b, :nb], boxes[b, :nb], (H, W)).numpy()
            assert band_mismatch_report(masks[b, :nb], up)[1] == 0
    assert n_band_frames <= 2
    if family == "noise":
        assert n_nonsimple >= B // 4          # the family does exercise the general path
    print(f"[parity] {family}: {B}/{B} records bit-exact vs the reference's contour route; {n_nonsimple} non-simple frames, "
          f"{n_band_frames} frames with in-band mask pixels")


@pytest.mark.parametrize("tc", PATHS)
def test_cfg1_full_batch_256(tc):
    """BASELINE config 1: 256 synthetic 640x640 frames, bit-exact grid / penalty / peaks check."""
    H = W = 640
    B, n = 256, 8
    eng = make_engine(tc, H=H, W=W, mh=160, mw=160, max_n=8, gs=20, max_batch=B)
    protos, coefs, boxes, counts = synth.make_batch(0, B, n, H, W, 160, 160, max_n=8)
    dev = to_dev(protos, coefs, boxes, counts)
    records, masks = eng.run(*dev)
    recs = eng.decode(records)
    band_pixels = n_general = 0
    masks_h = masks.cpu().numpy()
    for b in range(B):
        full = opl.frame_from_tensors(protos[b], coefs[b], boxes[b], (H, W), 20, "contour")
        assert_record_equals_oracle(recs[b], full, f"cfg1 frame {b}")
        # binary-mask rule of the north star on EVERY frame: bit-exact except pixels within 1e-4 of the threshold
        if not np.array_equal(masks_h[b], full["masks"]):
            up = oma.upsampled_logits(protos[b], coefs[b], boxes[b], (H, W)).numpy()
            nd, nout = band_mismatch_report(masks_h[b], up)
            assert nout == 0, (b, nd, nout)
            band_pixels += nd
        n_general += bool(recs[b].flags & opl.FLAG_NON_SIMPLE)
    print(f"[parity] cfg1: 256/256 records bit-exact; {band_pixels} in-band mask pixels over all 256 frames; "
          f"{n_general} non-simple frames")
    # size-independent properties: idempotence (scratch reset), batch-split invariance, grid-only mode
    records2, _ = eng.run(*dev)
    assert torch.equal(records, records2)
    half = [t[:128].contiguous() for t in dev]
    r_half, _ = eng.run(*half)
    assert torch.equal(r_half, records[:128])
    r_nomask, none = eng.run(*dev, write_masks=False)
    assert none is None and torch.equal(r_nomask, records)


@pytest.mark.parametrize("tc", PATHS)
def test_sixteen_instance_accumulator_path(tc):
    """max_n in (8, 16]: the tcgen05 kernel reads 16 accumulator columns per pass."""
    H = W = 640
    B, n = 6, 12
    eng = make_engine(tc, H=H, W=W, mh=160, mw=160, max_n=16, gs=20, max_batch=B)
    protos, coefs, boxes, counts = synth.make_batch(6000, B, n, H, W, 160, 160, max_n=16)
    counts[1] = 16 if False else 12
    counts[2] = 9
    records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
    recs = eng.decode(records)
    masks = masks.cpu().numpy()
    for b in range(B):
        nb = int(counts[b])
        up = oma.upsampled_logits(protos[b], coefs[b, :nb], boxes[b, :nb], (H, W)).numpy()
        nd, nout = band_mismatch_report(masks[b, :nb], up)
        assert nout == 0, (nd, nout)
        assert_record_equals_oracle(recs[b], opl.frame_from_masks(masks[b, :nb], 20, "contour"), f"n=12 frame {b}")


@pytest.mark.parametrize("tc", PATHS)
def test_instance_groups_ragged_counts_and_dead_bands(tc):
    """max_n = 32 (two instance groups of 16 on the tcgen05 path), ragged counts that leave a group empty or
    partly filled, and boxes chosen so that whole bands lie outside the hull of a group's boxes: thin boxes at the
    top / bottom edge, boxes outside the frame, empty and inverted boxes."""
    H = W = 640
    B, n = 8, 32
    eng = make_engine(tc, H=H, W=W, mh=160, mw=160, max_n=32, gs=20, max_batch=B)
    protos, coefs, boxes, counts = synth.make_batch(7000, B, n, H, W, 160, 160, max_n=32)
    for b, c in enumerate([32, 17, 16, 5, 1, 0, 24, 32]):
        counts[b] = c
    g = torch.Generator().manual_seed(7)
    boxes[0, 16:] = torch.tensor([40., 600., 600., 639.])            # group 1 lives in the bottom band only
    boxes[1, 16] = torch.tensor([100., 0., 500., 9.])                # one thin instance at the very top
    boxes[2, :16, 1] = 300.; boxes[2, :16, 3] = 340.                 # all boxes inside one band
    boxes[3, 0] = torch.tensor([-50., -80., -10., -5.])              # outside the frame (negative)
    boxes[3, 1] = torch.tensor([700., 700., 900., 900.])             # outside the frame (beyond)
    boxes[3, 2] = torch.tensor([200., 300., 200., 300.])             # empty
    boxes[3, 3] = torch.tensor([400., 500., 300., 100.])             # inverted
    boxes[6, :24] = torch.rand(24, 4, generator=g) * 640             # arbitrary, mostly inverted / degenerate
    boxes[7, :, 1] = 0.; boxes[7, :, 3] = 640.                       # full height: nothing is dead
    records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
    recs = eng.decode(records)
    masks = masks.cpu().numpy()
    for b in range(B):
        nb = int(counts[b])
        if nb:
            up = oma.upsampled_logits(protos[b], coefs[b, :nb], boxes[b, :nb], (H, W)).numpy()
            nd, nout = band_mismatch_report(masks[b, :nb], up)
            assert nout == 0, (b, nd, nout)
        assert_record_equals_oracle(recs[b], opl.frame_from_masks(masks[b, :nb], 20, "contour"), f"groups frame {b}")
    r_nomask, _ = eng.run(*to_dev(protos, coefs, boxes, counts), write_masks=False)
    assert torch.equal(r_nomask, records)


@pytest.mark.parametrize("tc", PATHS)
@pytest.mark.parametrize("max_n,seed", [(8, 1), (16, 2), (32, 3)])
def test_box_fuzz_masks_and_records(tc, max_n, seed):
    """Adversarial boxes against every exactness shortcut of the mask kernels (hull of the group's boxes, live chunk
    ranges, epilogue skips, zero fills): edges on / next to integer proto coordinates, sub-pixel and inverted boxes,
    boxes outside the frame, +-inf and NaN coordinates, ragged counts.  Masks must equal the oracle's process_mask
    except inside the 1e-4 logit band; records must be bit-exact for the masks produced."""
    H = W = 640
    B = 24
    eng = make_engine(tc, H=H, W=W, mh=160, mw=160, max_n=max_n, gs=20, max_batch=B)
    g = torch.Generator().manual_seed(1000 + seed)
    protos, coefs, boxes, counts = synth.make_batch(9000 + 100 * seed, B, max_n, H, W, 160, 160, family="noise", max_n=max_n)
    protos[B // 2:] = synth.make_batch(9500 + 100 * seed, B - B // 2, max_n, H, W, 160, 160, family="sidewalk", max_n=max_n)[0]
    counts[:] = torch.randint(0, max_n + 1, (B,), generator=g, dtype=torch.int32)
    counts[0], counts[1] = max_n, 1
    for b in range(B):
        for i in range(max_n):
            kind = int(torch.randint(0, 10, (1,), generator=g))
            r = torch.rand(4, generator=g)
            if kind == 0:      # edges exactly on proto pixel boundaries (multiples of 4 frame pixels)
                x1, y1 = 4 * int(r[0] * 100), 4 * int(r[1] * 100)
                bx = [x1, y1, x1 + 4 * int(1 + r[2] * 60), y1 + 4 * int(1 + r[3] * 60)]
            elif kind == 1:    # just off the boundaries
                x1, y1 = 4 * int(r[0] * 100) + 1e-3, 4 * int(r[1] * 100) - 1e-3
                bx = [x1, y1, x1 + 4 * int(1 + r[2] * 60) - 2e-3, y1 + 4 * int(1 + r[3] * 60) + 2e-3]
            elif kind == 2:    # thin horizontal strip (one or two proto rows)
                y1 = float(r[1] * 630)
                bx = [float(r[0] * 300), y1, float(300 + r[2] * 340), y1 + float(r[3] * 8)]
            elif kind == 3:    # thin vertical strip
                x1 = float(r[0] * 630)
                bx = [x1, float(r[1] * 300), x1 + float(r[2] * 8), float(300 + r[3] * 340)]
            elif kind == 4:    # inverted / empty
                bx = [float(r[0] * 640), float(r[1] * 640), float(r[0] * 640 - r[2] * 50), float(r[1] * 640 - r[3] * 50)]
            elif kind == 5:    # outside the frame
                bx = [-200 + float(r[0] * 100), 700 + float(r[1] * 100), -50 + float(r[2] * 40), 900 + float(r[3] * 40)]
            elif kind == 6:    # covers everything, infinite bounds
                bx = [float("-inf"), -5.0, float("inf"), 1e9]
            elif kind == 7:    # NaN coordinate: every comparison is false -> the mask is empty
                bx = [float(r[0] * 300), float("nan"), float(300 + r[2] * 300), float(300 + r[3] * 300)]
            else:              # ordinary random box
                x1, y1 = float(r[0] * 500), float(r[1] * 500)
                bx = [x1, y1, x1 + float(r[2] * 400), y1 + float(r[3] * 400)]
            boxes[b, i] = torch.tensor(bx)
    records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
    recs = eng.decode(records)
    masks = masks.cpu().numpy()
    n_band = 0
    for b in range(B):
        nb = int(counts[b])
        if nb:
            up = oma.upsampled_logits(protos[b], coefs[b, :nb], boxes[b, :nb], (H, W)).numpy()
            nd, nout = band_mismatch_report(masks[b, :nb], np.nan_to_num(up, nan=0.0))
            assert nout == 0, (b, nd, nout)
            # a reference logit of exactly 0 comes from crop_mask: no tolerance there
            assert int(masks[b, :nb][up == 0].sum()) == 0, (b, int(masks[b, :nb][up == 0].sum()))
            n_band += nd
        assert_record_equals_oracle(recs[b], opl.frame_from_masks(masks[b, :nb], 20, "contour"), f"fuzz frame {b}")
    r_nomask, _ = eng.run(*to_dev(protos, coefs, boxes, counts), write_masks=False)
    assert torch.equal(r_nomask, records)
    print(f"[parity] box fuzz max_n={max_n}: {n_band} mask pixels inside the 1e-4 band")


def test_band_count_invariance():
    """The tensor-core kernel cuts a frame into 1 .. 20 bands depending on the batch size (work items per SM);
    masks and records must not depend on that.  The same 32 frames at B = 1, 3, 32, 96, 640 and 1280 (one band per
    frame) against the B = 32 result, which test_cfg1_full_batch_256-style parity pins to the oracle."""
    H = W = 640
    n = 8
    protos, coefs, boxes, counts = synth.make_batch(300, 32, n, H, W, 160, 160, max_n=n)
    counts[5] = 0
    counts[9] = 3
    base = None
    for B in (32, 1, 3, 96, 640, 1280):
        eng = make_engine(True, H=H, W=W, mh=160, mw=160, max_n=n, gs=20, max_batch=B)
        reps = (B + 31) // 32
        dev = [t.repeat(reps, *([1] * (t.dim() - 1)))[:B].contiguous().cuda() for t in (protos, coefs, boxes, counts)]
        records, masks = eng.run(*dev)
        if base is None:
            base = (records.clone(), masks.clone())
            continue
        idx = torch.arange(B, device="cuda") % 32
        assert torch.equal(records, base[0][idx]), B
        live = (torch.arange(n)[None, :] < counts[:, None]).cuda()        # instances >= count are never written
        for b0 in range(0, B, 32):                       # masks block-wise to bound memory
            hi = min(b0 + 32, B)
            assert torch.equal(masks[b0:hi][live[: hi - b0]], base[1][: hi - b0][live[: hi - b0]]), (B, b0)
        del eng, records, masks, dev
        torch.cuda.empty_cache()


@pytest.mark.parametrize("H,W,mh,mw,gs", [(270, 480, 40, 40, 10), (333, 500, 48, 64, 20), (200, 304, 100, 152, 8),
                                           (96, 128, 96, 128, 8), (80, 96, 160, 192, 4), (1000, 1000, 36, 52, 20),
                                           (641, 643, 160, 160, 20), (540, 960, 80, 160, 20), (405, 720, 60, 120, 10),
                                           (300, 1024, 100, 64, 8)])
@pytest.mark.parametrize("tc", PATHS)
def test_generic_geometry_fuzz(tc, H, W, mh, mw, gs):
    """Generic-scale paths (tcgen05 kernel with the row table where its limits allow, CUDA-core contraction + generic
    upsample everywhere): non-integer, anisotropic, unit and down-sampling scales, widths that are not a multiple of
    16 (ragged right edge) or of the cell size."""
    B, n = 4, 5
    eng = make_engine(tc, H=H, W=W, mh=mh, mw=mw, max_n=8, gs=gs, max_batch=B)
    for fam, first in (("sidewalk", 1200), ("noise", 1300)):
        protos, coefs, boxes, counts = synth.make_batch(first, B, n, H, W, mh, mw, family=fam, max_n=8)
        counts[1] = 2
        records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
        recs = eng.decode(records)
        masks = masks.cpu().numpy()
        for b in range(B):
            nb = int(counts[b])
            up = oma.upsampled_logits(protos[b], coefs[b, :nb], boxes[b, :nb], (H, W)).numpy()
            nd, nout = band_mismatch_report(masks[b, :nb], up)
            assert nout == 0, (fam, b, nd, nout)
            assert_record_equals_oracle(recs[b], opl.frame_from_masks(masks[b, :nb], gs, "contour"), f"{fam} {H}x{W} frame {b}")


@pytest.mark.parametrize("tc", PATHS)
def test_mixed_call_sequence_keeps_scratch_clean(tc):
    """The context's reduction scratch (areas, bounding boxes, lattice bits) is consumed and reset by whichever call
    used it: any interleaving of the entry points must leave the next call's result unchanged."""
    H = W = 640
    B, n = 12, 8
    eng = make_engine(tc, H=H, W=W, mh=160, mw=160, max_n=n, gs=20, max_batch=B)
    protos, coefs, boxes, counts = synth.make_batch(8100, B, n, H, W, 160, 160, max_n=n)
    counts[2] = 0
    counts[7] = 5
    dev = to_dev(protos, coefs, boxes, counts)
    rec0, masks0 = eng.run(*dev)
    masks_only = eng.assemble_masks(*dev)
    live = (torch.arange(n)[None, :] < counts[:, None]).cuda()
    assert torch.equal(masks_only[live], masks0[live])
    m = torch.where(live[:, :, None, None], masks0, torch.zeros_like(masks0))
    rec1 = eng.masks_to_records(m, dev[3])
    assert torch.equal(rec1, rec0)                      # same masks through the mask-driven entry point
    rec2, _ = eng.run(*dev, write_masks=False)
    assert torch.equal(rec2, rec0)
    half = [t[:5].contiguous() for t in dev]
    eng.assemble_masks(*half)
    rec3, masks3 = eng.run(*dev)
    assert torch.equal(rec3, rec0) and torch.equal(masks3[live], masks0[live])
    h = [t.cpu().pin_memory() for t in (protos, coefs, boxes, counts)]
    rec4 = eng.run_host(*h)
    assert torch.equal(rec4.cuda(), rec0)
    rec5, _ = eng.run(*dev)
    assert torch.equal(rec5, rec0)


def test_nms_golden_and_oracle():
    """va_nms against the golden rows of the vendored non_max_suppression and against the oracle on
    further seeds; then head output -> va_nms -> va_run_fused must equal va_run_fused on the oracle's rows."""
    z = goldenio.load("nms.npz")
    eng = MaskGridEngine(H=640, W=640, mh=160, mw=160, max_n=32, gs=20, max_batch=8)

    def check(pred, want_rows, ct, it, nc, md):
        coefs, boxes, conf, cls, counts = eng.nms(pred.cuda(), conf_thres=ct, iou_thres=it, nc=nc, max_det=min(md, 32))
        for b, want in enumerate(want_rows):
            k = int(counts[b])
            assert k == min(want.shape[0], md, 32), (b, k, want.shape)
            got = np.concatenate([boxes[b, :k].cpu().numpy(), conf[b, :k, None].cpu().numpy(),
                                  cls[b, :k, None].cpu().numpy().astype(np.float32), coefs[b, :k].cpu().numpy()], 1)
            assert np.array_equal(got.view(np.uint32), want[:k].view(np.uint32)), b
            assert float(boxes[b, k:].abs().sum()) == 0.0 and float(coefs[b, k:].abs().sum()) == 0.0
        return coefs, boxes, counts

    for ci, (first, B, A, nc, nobj, ties, md) in enumerate(z["cases"].tolist()):
        ct, it = (float(v) for v in z["thres"][ci])
        pred = synth.make_head_output(first, B, A=A, nc=nc, n_objects=nobj, ties=bool(ties))
        check(pred, [z[f"{ci}/{b}"] for b in range(B)], ct, it, nc, md)
    for seed in range(6):                                   # more seeds against the oracle
        nc = 1 + seed % 3
        pred = synth.make_head_output(500 + 10 * seed, 4, A=8400, nc=nc, n_objects=4 + 2 * seed, ties=bool(seed % 2))
        want = onms.nms_batch(pred.numpy(), conf_thres=0.5, iou_thres=0.7, nc=nc, max_det=32)
        coefs, boxes, counts = check(pred, want, 0.5, 0.7, nc, 32)
    # the front end feeds the hot path without a host round trip
    protos = synth.make_batch(40, 4, 8, 640, 640, 160, 160, max_n=32)[0].cuda()
    rec_a, masks_a = eng.run(protos, coefs, boxes, counts)
    c2 = torch.zeros_like(coefs); b2 = torch.zeros_like(boxes)
    for b, rows in enumerate(want):
        k = rows.shape[0]
        b2[b, :k] = torch.from_numpy(rows[:, :4]).cuda(); c2[b, :k] = torch.from_numpy(rows[:, 6:]).cuda()
    rec_b, masks_b = eng.run(protos, c2, b2, torch.tensor([r.shape[0] for r in want], dtype=torch.int32).cuda())
    assert torch.equal(rec_a, rec_b)
    # more candidates than one 512-candidate tile: tiles in order of (score, anchor order), each thinned by the survivors
    # of the earlier ones - the reference's result for any number of candidates
    g = torch.Generator().manual_seed(99)
    dense = torch.rand(5, 37, 4000, generator=g)
    dense[:, :2] *= 600
    dense[:, 2:4] = 8 + dense[:, 2:4] * 60
    dense[0, 4] = 0.9                                        # all tied: anchor order decides
    dense[1, 4] = 0.5 + 0.5 * torch.rand(4000, generator=g)  # distinct scores
    dense[2, 4] = torch.round((0.5 + 0.5 * torch.rand(4000, generator=g)) * 64) / 64   # many ties on the tile boundaries
    dense[3, :4] = torch.tensor([300., 300., 100., 100.])[:, None]                      # every box identical: one survivor,
    dense[3, 4] = 0.9                                                                   # found after visiting every tile
    dense[4, :2] = 300 + 40 * torch.rand(2, 4000, generator=g)                          # one crowded spot: few survivors,
    dense[4, 2:4] = 80 + 40 * torch.rand(2, 4000, generator=g)                          # spread over many tiles
    dense[4, 4] = torch.round((0.5 + 0.5 * torch.rand(4000, generator=g)) * 16) / 16
    want = onms.nms_batch(dense.numpy(), conf_thres=0.5, iou_thres=0.7, nc=1, max_det=32)
    assert want[3].shape[0] == 1 and 1 < want[4].shape[0] < 32
    check(dense, want, 0.5, 0.7, 1, 32)
    check(dense, onms.nms_batch(dense.numpy(), conf_thres=0.5, iou_thres=0.3, nc=1, max_det=32), 0.5, 0.3, 1, 32)
    check(dense, onms.nms_batch(dense.numpy(), conf_thres=0.5, iou_thres=0.7, nc=1, max_det=5), 0.5, 0.7, 1, 5)
    # scale_boxes: the kept boxes back in original-frame pixels (ops.py:139-174)
    coefs, boxes, conf, cls, counts = eng.nms(dense.cuda(), conf_thres=0.5, iou_thres=0.7, nc=1)
    for img1, img0 in [((640, 640), (720, 1280)), ((640, 640), (480, 640)), ((640, 640), (333, 517)), ((384, 640), (1080, 1920))]:
        got = eng.scale_boxes(boxes, counts, img1, img0).cpu().numpy()
        for b in range(dense.shape[0]):
            k = int(counts[b])
            want_b = onms.scale_boxes(img1, boxes[b, :k].cpu().numpy(), img0)
            assert np.array_equal(got[b, :k].view(np.uint32), want_b.view(np.uint32)), (img1, img0, b)
            assert not got[b, k:].any()


@pytest.mark.parametrize("tc", PATHS)
def test_cfg2_1080p_generic_scale(tc):
    H, W, B, n = 1080, 1920, 2, 32
    eng = make_engine(tc, H=H, W=W, mh=160, mw=160, max_n=32, gs=20, max_batch=B)
    assert eng.uses_tensor_core == tc          # BASELINE configs[2] runs on the tcgen05 kernel (generic scale 6.75 x 12)
    protos, coefs, boxes, counts = synth.make_batch(7000, B, n, H, W, 160, 160, max_n=32)
    records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
    recs = eng.decode(records)
    masks = masks.cpu().numpy()
    for b in range(B):
        up = oma.upsampled_logits(protos[b], coefs[b], boxes[b], (H, W)).numpy()
        nd, nout = band_mismatch_report(masks[b], up)
        assert nout == 0 and nd < 50, (nd, nout)
        assert_record_equals_oracle(recs[b], opl.frame_from_masks(masks[b], 20, "contour"), f"1080p frame {b}")


@pytest.mark.parametrize("tc", PATHS)
@pytest.mark.parametrize("gs", [4, 8, 16, 32])
def test_cfg4_cell_size_sweep(tc, gs):
    H = W = 640
    B, n = 4, 8
    eng = make_engine(tc, H=H, W=W, mh=160, mw=160, max_n=8, gs=gs, max_batch=B)
    protos, coefs, boxes, counts = synth.make_batch(8000, B, n, H, W, 160, 160, max_n=8)
    records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
    recs = eng.decode(records)
    masks = masks.cpu().numpy()
    for b in range(B):
        assert_record_equals_oracle(recs[b], opl.frame_from_masks(masks[b], gs, "contour"), f"gs={gs} frame {b}")


@pytest.mark.parametrize("tc", PATHS)
@pytest.mark.parametrize("m,n", [(192, 8), (224, 8), (256, 32), (320, 8)])
def test_cfg4_proto_sweep(tc, m, n):
    H = W = 4 * m
    B = 2
    eng = make_engine(tc, H=H, W=W, mh=m, mw=m, max_n=n, gs=20 if H % 20 == 0 else 16, max_batch=B)
    protos, coefs, boxes, counts = synth.make_batch(9000, B, n, H, W, m, m, max_n=n)
    records, masks = eng.run(*to_dev(protos, coefs, boxes, counts))
    recs = eng.decode(records)
    masks = masks.cpu().numpy()
    for b in range(B):
        up = oma.upsampled_logits(protos[b], coefs[b], boxes[b], (H, W)).numpy()
        nd, nout = band_mismatch_report(masks[b], up)
        assert nout == 0, (m, nd, nout)
        assert_record_equals_oracle(recs[b], opl.frame_from_masks(masks[b], eng.gs, "contour"), f"proto {m} frame {b}")


def test_host_buffer_path_matches_device_path():
    H = W = 640
    B, n = 70, 8                   # > 2 host chunks
    eng = MaskGridEngine(H=H, W=W, mh=160, mw=160, max_n=8, gs=20, max_batch=B)
    protos, coefs, boxes, counts = synth.make_batch(100, B, n, H, W, 160, 160, max_n=8, pin=True)
    rec_host = eng.run_host(protos, coefs, boxes, counts)
    rec_dev, _ = eng.run(*to_dev(protos, coefs, boxes, counts), write_masks=False)
    assert np.array_equal(rec_host.numpy(), rec_dev.cpu().numpy())
    masks_h = torch.empty((B, 8, H, W), dtype=torch.uint8, pin_memory=True)
    rec_host2 = eng.run_host(protos, coefs, boxes, counts, masks_out=masks_h)
    _, masks_d = eng.run(*to_dev(protos, coefs, boxes, counts))
    assert np.array_equal(rec_host2.numpy(), rec_dev.cpu().numpy())
    assert np.array_equal(masks_h.numpy(), masks_d.cpu().numpy())
    # fp16 prototypes in host memory (model run with half=True): widened exactly on the device, i.e. the fp32 path on
    # protos.half().float() - what the reference computes (protos.float(), ops.py:724) - with half the PCIe bytes
    ph = protos.half().pin_memory()
    rec_f16 = eng.run_host(ph, coefs, boxes, counts)
    rec_ref, _ = eng.run(*to_dev(ph.float(), coefs, boxes, counts), write_masks=False)
    assert np.array_equal(rec_f16.numpy(), rec_ref.cpu().numpy())
    for b_ in (0, 33, 69):                      # and against the oracle's process_mask on the widened prototypes
        m_ = oma.process_mask(ph[b_].float(), coefs[b_], boxes[b_], (H, W)).numpy().astype(np.uint8)
        up_ = oma.upsampled_logits(ph[b_].float(), coefs[b_], boxes[b_], (H, W)).numpy()
        masks_b = eng.run(*to_dev(ph[b_:b_ + 1].float(), coefs[b_:b_ + 1], boxes[b_:b_ + 1], counts[b_:b_ + 1]))[1][0].cpu().numpy()
        nd, nout = band_mismatch_report(masks_b, up_)
        assert nout == 0
        assert_record_equals_oracle(eng.decode(rec_f16[b_:b_ + 1])[0], opl.frame_from_masks(masks_b, 20, "contour"), f"f16 frame {b_}")


def test_polygon_route_golden():
    """ultralytics-style masks.xy polygons: host fillPoly + boundingRect, device grid/penalty/peaks."""
    import cv2
    engines = {}
    n = 0
    for case in goldenio.polygon_cases():
        H, W, gs = case["H"], case["W"], case["gs"]
        key = (H, W, gs)
        if key not in engines:
            engines[key] = MaskGridEngine(H=H, W=W, mh=H // 4 // 4 * 4 or 4, mw=max(4, W // 4 // 4 * 4), max_n=1, gs=gs, max_batch=1)
        eng = engines[key]
        poly = og.select_polygon(case["polys"])
        pts = np.int32([poly])
        rect = cv2.boundingRect(pts)
        raster = np.zeros((H, W), np.uint8)
        cv2.fillPoly(raster, pts, 1)
        m = torch.from_numpy(raster)[None, None].cuda()
        rec = eng.decode(eng.masks_to_records(m, torch.tensor([1], dtype=torch.int32).cuda(),
                                              rects=torch.tensor([list(rect)], dtype=torch.int32).cuda(),
                                              sel=torch.tensor([0], dtype=torch.int32).cuda()))[0]
        err = 1 if rec.flags & opl.FLAG_CENTRE_OOB else 2 if rec.flags & opl.FLAG_LIST_OOB else 0
        assert err == case["err"], case["idx"]
        goldenio.assert_result_matches(rec.as_dict(), case, f"polygon case {case['idx']}")
        n += 1
    assert n >= 150


def test_reference_fixtures_grid_mode():
    """The reference's 13 *_grids.npy fixtures and 5 live PNG known answers through va_grid_to_penalty_peaks."""
    z = goldenio.load("fixtures.npz")
    eng = MaskGridEngine(H=1280, W=720, mh=320, mw=180, max_n=1, gs=20, max_batch=32)   # fixtures are 64 rows x 36 cols
    keys = list(open_.PENALTY_COLOUR_GRADIENT.keys())
    for use_easy, key in ((0, "pen_traversal"), (1, "pen_easy")):
        inputs = [dict(x0=0, rows_y=z[f"{nm}/rows_y"], rows_attr=z[f"{nm}/rows_attr"], occ=z[f"{nm}/occ"], use_easy=use_easy)
                  for nm in z["names"]]
        recs = eng.decode(eng.grids_to_records(inputs))
        for nm, rec in zip(z["names"], recs):
            gold = z[f"{nm}/{key}"]
            assert np.array_equal(np.isnan(rec.penalty), np.isnan(gold)), nm
            assert np.array_equal(rec.penalty[~np.isnan(gold)].view(np.uint64), gold[~np.isnan(gold)].view(np.uint64)), nm
            assert np.array_equal(rec.peaks, z[f"{nm}/peaks"]), nm
            if nm in z["live_png"] and not use_easy:
                col = z[f"{nm}/png_colour_idx"]
                for r in range(rec.R):
                    for c in range(rec.C):
                        if rec.occ[r, c] & 1:
                            want = open_.PENALTY_COLOUR_GRADIENT[keys[int(col[r, c])]]
                            assert open_.get_penalty_colour(rec.penalty[r, c]) == want, (nm, r, c)


def test_random_grids_grid_mode_vs_oracle():
    rng = np.random.default_rng(77)
    gs = 20
    eng = MaskGridEngine(H=720, W=1280, mh=180, mw=320, max_n=1, gs=gs, max_batch=64)
    done = 0
    for it in range(12):
        inputs, want = [], []
        for k in range(64):
            R, C = int(rng.integers(1, 37)), int(rng.integers(1, 65))
            g = polygen.random_occupancy(rng, R, C)
            st = og.grid_from_npy(np.pad(g, ((0, 36 - R), (0, 64 - C))))
            use_easy = bool(rng.integers(0, 2))
            pen = open_.calculate_penalties(st, use_easy=use_easy)
            rows_y = np.array([r[0].y for r in st.grids], np.int32)
            occ = np.array([[(0 if q.empty else 1) | (2 if q.artificial else 0) for q in r] for r in st.grids], np.uint8)
            inputs.append(dict(x0=0, rows_y=rows_y, rows_attr=np.array([r[0].row for r in st.grids], np.int32), occ=occ,
                               use_easy=int(use_easy)))
            want.append((pen, oprot.peaks_closed_form(rows_y, (occ & 1).astype(bool), 0, st.W, gs)))
        recs = eng.decode(eng.grids_to_records(inputs))
        for rec, (pen, pk) in zip(recs, want):
            assert np.array_equal(np.isnan(rec.penalty), np.isnan(pen))
            assert np.array_equal(rec.penalty[~np.isnan(pen)].view(np.uint64), pen[~np.isnan(pen)].view(np.uint64))
            assert [tuple(p) for p in rec.peaks.tolist()] == pk
            done += 1
    assert done == 768


def test_capacity_and_argument_errors():
    from vision_assist_b200 import _lib
    eng = MaskGridEngine(H=640, W=640, mh=160, mw=160, max_n=8, gs=20, max_batch=2)
    protos, coefs, boxes, counts = synth.make_batch(0, 3, 2, 640, 640, 160, 160, max_n=8)
    with pytest.raises(ValueError):
        eng.run(*to_dev(protos, coefs, boxes, counts))
    with pytest.raises(ValueError):
        eng.run(protos, coefs, boxes, counts)          # CPU tensors on the device API
    with pytest.raises(_lib.VaError):
        MaskGridEngine(H=640, W=640, mh=160, mw=160, max_n=64)
    r0, _ = eng.run(*[t[:0] for t in to_dev(protos, coefs, boxes, counts)])
    assert r0.shape[0] == 0


# last in the file on purpose: the longest-running test of the suite
@pytest.mark.parametrize("H,W,mh,mw,n,B", [(640, 640, 160, 160, 8, 256), (640, 640, 160, 160, 32, 48), (1080, 1920, 160, 160, 32, 8)])
def test_repeatability_soak(H, W, mh, mw, n, B):
    """Work stealing, atomics into the reduction scratch and the self-resetting work counter must not make the
    result depend on scheduling: 60 back-to-back calls on the same inputs give bit-identical records and masks."""
    eng = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=n, gs=20, max_batch=B)
    uniq = min(B, 16)
    hp, hc, hb, hn = synth.make_batch(4242, uniq, n, H, W, mh, mw, max_n=n)
    r = (B + uniq - 1) // uniq
    dev = [t.repeat(r, *([1] * (t.dim() - 1)))[:B].contiguous().cuda() for t in (hp, hc, hb, hn)]
    rec0, masks0 = eng.run(*dev)
    rec0, masks0 = rec0.clone(), masks0.clone()
    rec = torch.empty_like(rec0)
    masks = torch.empty_like(masks0)
    for it in range(60):
        eng.run(*dev, masks_out=masks, records_out=rec, write_masks=(it % 3 != 2))
        if not torch.equal(rec, rec0):              # say what moved: frame, byte offsets, the decoded fields
            d = (rec != rec0).cpu().numpy()
            fr = int(np.nonzero(d.any(1))[0][0])
            a_, g_ = eng.decode(rec0[fr:fr + 1])[0], eng.decode(rec[fr:fr + 1])[0]
            raise AssertionError(f"call {it}: record of frame {fr} differs at byte offsets {np.nonzero(d[fr])[0][:8].tolist()} "
                                 f"({int(d[fr].sum())} bytes): sel {a_.sel}->{g_.sel} flags {a_.flags}->{g_.flags} "
                                 f"area2 {a_.contour_area2}->{g_.contour_area2} bbox {a_.bbox}->{g_.bbox} R {a_.R}->{g_.R}")
        if it % 3 != 2 and it % 10 == 0:
            assert torch.equal(masks, masks0), it
