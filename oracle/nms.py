"""Oracle: the front of the path - candidate filter + NMS on raw segmentation-head output
(TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).  SURVEY 8(f3).

Restates `non_max_suppression` of the vendored ultralytics ops (testing/old/segmenting_using_tflite/ops.py:
214-363) for the configuration FrameProcessor uses through `model.predict(frame, conf=0.5)`
(FrameProcessor.py:322): best class only (multi_label=False), no class filter, no apriori labels, not rotated,
and torchvision.ops.nms (CPU kernel, torchvision/csrc/ops/cpu/nms_kernel.cpp of the 0.2x series) which it calls:

  * candidates: anchors whose best class confidence is > conf_thres, in ANCHOR ORDER           (:277, :307-308)
  * xywh -> xyxy in fp32: xy - wh/2, xy + wh/2                                                  (:472-480)
  * more than max_nms candidates: the max_nms best by confidence                               (:316-317)
  * boxes offset by class * max_wh (fp32) before the IoU test unless agnostic                   (:319-325)
  * greedy NMS in order of descending score, STABLE (ties keep anchor order); box j is suppressed by a kept
    box i iff  inter / (area_i + area_j - inter) > iou_thres  in fp32, with
    inter = max(0, min(x2) - max(x1)) * max(0, min(y2) - max(y1)),  area = (x2 - x1) * (y2 - y1)
  * the first max_det survivors, rows (x1, y1, x2, y2, conf, class, mask coefficients...)        (:327-330)
"""
from __future__ import annotations

import numpy as np

F = np.float32


def nms_image(pred: np.ndarray, conf_thres: float, iou_thres: float, nc: int, max_det: int = 300,
              agnostic: bool = False, max_wh: int = 7680, max_nms: int = 30000) -> np.ndarray:
    """pred: [4 + nc + nm, A] fp32 (one image, as the head emits it) -> [n, 6 + nm] fp32."""
    pred = np.asarray(pred, F)
    nm = pred.shape[0] - 4 - nc
    cls = pred[4:4 + nc]                                   # [nc, A]
    conf = cls.max(axis=0)
    j = cls.argmax(axis=0)                                 # first maximum, like torch.max on CPU
    keep = np.nonzero(conf > F(conf_thres))[0]
    if keep.size == 0:
        return np.zeros((0, 6 + nm), F)
    x, y, w, h = (pred[k, keep] for k in range(4))
    hw, hh = w / F(2), h / F(2)
    box = np.stack([x - hw, y - hh, x + hw, y + hh], 1).astype(F)
    rows = np.concatenate([box, conf[keep, None], j[keep, None].astype(F), pred[4 + nc:, keep].T], 1).astype(F)
    if rows.shape[0] > max_nms:
        rows = rows[np.argsort(-rows[:, 4], kind="stable")[:max_nms]]
    c = rows[:, 5:6] * F(0 if agnostic else max_wh)
    b = (rows[:, :4] + c).astype(F)
    scores = rows[:, 4]
    order = np.argsort(-scores, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = ((x2 - x1).astype(F) * (y2 - y1).astype(F)).astype(F)
    suppressed = np.zeros(rows.shape[0], bool)
    kept = []
    thr = F(iou_thres)
    for _i, i in enumerate(order):
        if suppressed[i]:
            continue
        kept.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(x1[i], x1[rest]); yy1 = np.maximum(y1[i], y1[rest])
        xx2 = np.minimum(x2[i], x2[rest]); yy2 = np.minimum(y2[i], y2[rest])
        ww = np.maximum(F(0), (xx2 - xx1).astype(F)); hh2 = np.maximum(F(0), (yy2 - yy1).astype(F))
        inter = (ww * hh2).astype(F)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = (inter / ((areas[i] + areas[rest]).astype(F) - inter).astype(F)).astype(F)
        suppressed[rest[ovr > thr]] = True
    return rows[np.array(kept[:max_det], np.int64)]


def nms_batch(pred: np.ndarray, **kw) -> list:
    return [nms_image(p, **kw) for p in pred]


def scale_boxes(img1_shape, boxes: np.ndarray, img0_shape) -> np.ndarray:
    """ops.scale_boxes (ops.py:139-174; ratio_pad None, padding True, xyxy) + clip_boxes (:367-385) on fp32 rows [n, 4]:
    gain / pad in Python arithmetic, the tensor arithmetic in fp32 (torch CPU: boxes -= pad; boxes /= fp32(gain); clamp)."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad = (round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1), round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1))
    out = np.array(boxes, F, copy=True)
    out[..., 0] -= F(pad[0]); out[..., 2] -= F(pad[0])
    out[..., 1] -= F(pad[1]); out[..., 3] -= F(pad[1])
    out[..., :4] /= F(gain)
    out[..., 0] = out[..., 0].clip(0, img0_shape[1]); out[..., 2] = out[..., 2].clip(0, img0_shape[1])
    out[..., 1] = out[..., 1].clip(0, img0_shape[0]); out[..., 3] = out[..., 3].clip(0, img0_shape[0])
    return out
