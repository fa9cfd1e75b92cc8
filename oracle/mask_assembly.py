"""Oracle: YOLOv8-seg mask assembly (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates `ultralytics.utils.ops.process_mask` as vendored in the reference at
`testing/old/segmenting_using_tflite/ops.py`:
    crop_mask        ops.py:688-704
    process_mask     ops.py:707-737
    masks2segments   ops.py:837-859
    scale_coords     ops.py:784-816   (+ clip_coords ops.py:388-405)

Two restatements are kept:
  * `process_mask` uses the same torch CPU operators the reference calls (fp32 matmul,
    F.interpolate bilinear align_corners=False) - this is the parity oracle;
  * `process_mask_np` spells the arithmetic out in numpy fp32 (explicit source-index /
    lambda computation and the 4-tap blend) so the CUDA kernel has a bit-level specification
    that does not depend on torch internals.  tests/test_oracle_mask.py checks they agree.
"""
from __future__ import annotations

import cv2
import numpy as np
import torch
import torch.nn.functional as F


def crop_mask(masks: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """ops.py:688-704 - zero everything outside the (float, half-open) box."""
    _, h, w = masks.shape
    x1, y1, x2, y2 = torch.chunk(boxes[:, :, None], 4, 1)
    r = torch.arange(w, dtype=x1.dtype)[None, None, :]
    c = torch.arange(h, dtype=x1.dtype)[None, :, None]
    return masks * ((r >= x1) * (r < x2) * (c >= y1) * (c < y2))


def cropped_logits(protos: torch.Tensor, masks_in: torch.Tensor, bboxes: torch.Tensor, shape):
    """ops.py:722-734 - the proto-resolution logits after crop (before upsample/threshold)."""
    c, mh, mw = protos.shape
    ih, iw = shape
    masks = (masks_in @ protos.float().view(c, -1)).view(-1, mh, mw)
    width_ratio = mw / iw
    height_ratio = mh / ih
    b = bboxes.clone()
    b[:, 0] *= width_ratio
    b[:, 2] *= width_ratio
    b[:, 3] *= height_ratio
    b[:, 1] *= height_ratio
    return crop_mask(masks, b)


def upsampled_logits(protos, masks_in, bboxes, shape) -> torch.Tensor:
    """ops.py:736 - float map whose sign is the mask."""
    m = cropped_logits(protos, masks_in, bboxes, shape)
    return F.interpolate(m[None], shape, mode="bilinear", align_corners=False)[0]


def process_mask(protos, masks_in, bboxes, shape, upsample: bool = True) -> torch.Tensor:
    """ops.py:707-737 - returns float {0,1} masks [n, ih, iw]."""
    m = cropped_logits(protos, masks_in, bboxes, shape)
    if upsample:
        m = F.interpolate(m[None], shape, mode="bilinear", align_corners=False)[0]
    return m.gt_(0.0)


# ---------------------------------------------------------------------------------------------
# explicit numpy specification
# ---------------------------------------------------------------------------------------------
def source_index_np(out_size: int, in_size: int):
    """PyTorch align_corners=False source index, fp32 exactly as ATen (torch 2.11 CPU) computes it.

    scale = float(in)/out ; src = max(fma(scale, dst+0.5, -0.5), 0) ; i0 = int(src) ;
    i1 = i0 + (i0 < in-1) ; l1 = src - i0 ; l0 = 1 - l1       (all float32; the multiply-subtract
    is contracted into ONE fma by the compiler - probed bit-for-bit on non-integer scales,
    tests/test_oracle_golden.py::test_mask_assembly_golden case 4).
    """
    scale = np.float32(in_size) / np.float32(out_size)
    dst = np.arange(out_size, dtype=np.float32)
    src = (np.float64(scale) * (dst + np.float32(0.5)).astype(np.float64) - 0.5).astype(np.float32)
    src = np.maximum(src, np.float32(0.0)).astype(np.float32)
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    l1 = np.clip(src - i0.astype(np.float32), np.float32(0), np.float32(1)).astype(np.float32)
    l0 = (np.float32(1.0) - l1).astype(np.float32)
    return i0, i1, l0, l1


def _fma32(p, q, r):
    """fl32(p*q + r) with a single rounding (the product of two fp32 is exact in fp64)."""
    return (p.astype(np.float64) * q.astype(np.float64) + r.astype(np.float64)).astype(np.float32)


def bilinear_upsample_np(m: np.ndarray, shape, contract: bool = True) -> np.ndarray:
    """Bit-level specification of the 4-tap blend (fp32).

    contract=True  : top = fma(a, l0x, fl(b*l1x)) ; out = fma(top, l0y, fl(bot*l1y))
                     - this is what torch 2.11 CPU (ATen UpSampleKernel Interpolate<2>, compiled
                     with FMA contraction) produces bit-for-bit (probe in tests/test_oracle_mask.py)
                     and what the CUDA kernel implements with FMUL + FFMA.
    contract=False : all four products and both sums individually rounded.
    """
    ih, iw = shape
    n, mh, mw = m.shape
    y0, y1, ly0, ly1 = source_index_np(ih, mh)
    x0, x1, lx0, lx1 = source_index_np(iw, mw)
    m = m.astype(np.float32)
    a = m[:, y0][:, :, x0]
    b = m[:, y0][:, :, x1]
    c = m[:, y1][:, :, x0]
    d = m[:, y1][:, :, x1]
    LX0 = np.broadcast_to(lx0, a.shape)
    LX1 = np.broadcast_to(lx1, a.shape)
    LY0 = np.broadcast_to(ly0[None, :, None], a.shape)
    LY1 = np.broadcast_to(ly1[None, :, None], a.shape)
    if not contract:
        top = a * LX0 + b * LX1
        bot = c * LX0 + d * LX1
        return (top * LY0 + bot * LY1).astype(np.float32)
    top = _fma32(a, LX0, (b * LX1).astype(np.float32))
    bot = _fma32(c, LX0, (d * LX1).astype(np.float32))
    return _fma32(top, LY0, (bot * LY1).astype(np.float32))


def cropped_logits_np(protos: np.ndarray, coefs: np.ndarray, boxes: np.ndarray, shape) -> np.ndarray:
    c, mh, mw = protos.shape
    ih, iw = shape
    logits = (coefs.astype(np.float32) @ protos.reshape(c, -1).astype(np.float32)).reshape(-1, mh, mw)
    wr = np.float32(mw / iw)
    hr = np.float32(mh / ih)
    b = boxes.astype(np.float32)
    x1 = (b[:, 0] * wr)[:, None, None]
    x2 = (b[:, 2] * wr)[:, None, None]
    y1 = (b[:, 1] * hr)[:, None, None]
    y2 = (b[:, 3] * hr)[:, None, None]
    r = np.arange(mw, dtype=np.float32)[None, None, :]
    cc = np.arange(mh, dtype=np.float32)[None, :, None]
    keep = (r >= x1) & (r < x2) & (cc >= y1) & (cc < y2)
    return np.where(keep, logits, np.float32(0.0)).astype(np.float32)


def process_mask_np(protos, coefs, boxes, shape) -> np.ndarray:
    """uint8 {0,1} masks [n, ih, iw] from the explicit numpy arithmetic."""
    up = bilinear_upsample_np(cropped_logits_np(protos, coefs, boxes, shape), shape)
    return (up > 0).astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# masks -> polygons (host side of ultralytics Results.masks.xy)
# ---------------------------------------------------------------------------------------------
def masks2segments(masks: np.ndarray) -> list[np.ndarray]:
    """ops.py:837-859, strategy='largest' (contour with the most points)."""
    segments = []
    for x in masks.astype("uint8"):
        c = cv2.findContours(x, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[0]
        if c:
            c = np.array(c[np.array([len(q) for q in c]).argmax()]).reshape(-1, 2)
        else:
            c = np.zeros((0, 2))
        segments.append(c.astype("float32"))
    return segments


def scale_coords(img1_shape, coords: np.ndarray, img0_shape) -> np.ndarray:
    """ops.py:784-816 with ratio_pad=None, padding=True, normalize=False; clip ops.py:388-405."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (img1_shape[1] - img0_shape[1] * gain) / 2, (img1_shape[0] - img0_shape[0] * gain) / 2
    coords = coords.copy()
    coords[..., 0] -= pad[0]
    coords[..., 1] -= pad[1]
    coords[..., 0] /= gain
    coords[..., 1] /= gain
    coords[..., 0] = coords[..., 0].clip(0, img0_shape[1])
    coords[..., 1] = coords[..., 1].clip(0, img0_shape[0])
    return coords


def masks_to_polygons(masks: np.ndarray, frame_shape) -> list[np.ndarray]:
    """`Results.masks.xy`: masks2segments then scale_coords to the original frame."""
    ih, iw = masks.shape[1:]
    return [scale_coords((ih, iw), s, frame_shape) for s in masks2segments(masks)]
