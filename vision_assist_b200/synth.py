"""Deterministic synthetic inputs for the mask -> grid -> penalty -> protrusion path.

There are no weights in the reference repository (`model/.MISSING_LARGE_BLOBS`) and no network,
so every test / bench frame is synthesised: YOLOv8-seg head outputs (prototype maps, per-instance
mask coefficients, xyxy boxes) whose assembled masks look like sidewalk segmentations.

Per frame the seed is 0xB2000000 + frame_idx, generated with a CPU torch.Generator so
the build container, the GPU box and the committed golden vectors all see identical tensors.

  family "sidewalk" : instance 0 is a sheared trapezoid that widens towards the bottom of the frame
                      (channel 0 is its signed-distance-like field), instances 1.. are small discs
                      (channels 1..3), channels 4..K-1 are smooth noise that roughens the outlines.
                      Masks are (almost always) single hole-free blobs, instance 0 has the largest area.
  family "noise"    : every channel is smooth noise, coefficients N(0,1): masks with many components
                      and holes - used for logit / binary-mask parity and the non-simple-mask flag.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

SEED_BASE = 0xB2000000
N_DISC_CH = 3


def _smooth_noise(g: torch.Generator, ch: int, mh: int, mw: int) -> torch.Tensor:
    z = torch.randn(1, ch, 6, 6, generator=g)
    return F.interpolate(z, (mh, mw), mode="bicubic", align_corners=False)[0]


def make_frame(frame_idx: int, n: int, H: int, W: int, mh: int, mw: int, K: int = 32,
               family: str = "sidewalk"):
    """-> protos [K,mh,mw] f32, coefs [n,K] f32, boxes [n,4] f32 (xyxy, frame pixels)."""
    g = torch.Generator().manual_seed(SEED_BASE + int(frame_idx))
    sx, sy = W / mw, H / mh
    if family == "noise":
        protos = _smooth_noise(g, K, mh, mw)
        coefs = torch.randn(n, K, generator=g)
        c = torch.rand(n, 2, generator=g) * torch.tensor([W * 0.6, H * 0.6])
        wh = (0.2 + 0.6 * torch.rand(n, 2, generator=g)) * torch.tensor([float(W), float(H)])
        boxes = torch.cat([c, torch.minimum(c + wh, torch.tensor([W - 1.0, H - 1.0]))], 1)
        return protos.contiguous(), coefs.contiguous(), boxes.contiguous()

    u = torch.rand(16, generator=g)
    nrm = torch.randn(8, generator=g)
    ys = torch.arange(mh, dtype=torch.float32)[:, None]
    xs = torch.arange(mw, dtype=torch.float32)[None, :]
    cx0 = mw / 2 + nrm[0] * mw / 8
    y_top = mh * (0.125 + 0.375 * u[0])
    hw_top = mw * (1 / 16 + u[1] / 8)
    hw_bot = mw * (0.25 + 0.25 * u[2])
    shear = 0.3 * nrm[1]
    t = ((ys - y_top) / max(mh - 1 - float(y_top), 1.0)).clamp(0, 1)
    halfw = hw_top + (hw_bot - hw_top) * t
    cx = cx0 + shear * (ys - y_top)
    field = torch.minimum(halfw - (xs - cx).abs(), ys - y_top + 0.5)

    protos = torch.zeros(K, mh, mw)
    protos[0] = field
    disc = []
    for d in range(N_DISC_CH):
        dcx = mw * (0.1 + 0.8 * u[3 + 2 * d])
        dcy = mh * (0.1 + 0.8 * u[4 + 2 * d])
        rad = mw * (1 / 16 + u[9 + d] / 16)
        protos[1 + d] = rad - ((xs - dcx) ** 2 + (ys - dcy) ** 2).sqrt()
        disc.append((float(dcx), float(dcy), float(rad)))
    n_noise = K - 1 - N_DISC_CH
    protos[1 + N_DISC_CH:] = 0.3 * _smooth_noise(g, n_noise, mh, mw)

    coefs = torch.zeros(n, K)
    boxes = torch.zeros(n, 4)
    jit = torch.rand(n, 4, generator=g) * 8.0
    for i in range(n):
        coefs[i, 1 + N_DISC_CH:] = 0.3 * torch.randn(n_noise, generator=g)
        if i == 0:
            coefs[i, 0] = 1.0
            pos = (field > -1.0).nonzero()
            if pos.numel() == 0:
                x1 = y1 = 0.0
                x2, y2 = W - 1.0, H - 1.0
            else:
                y1, x1 = (pos.min(0).values.float() * torch.tensor([sy, sx])).tolist()
                y2, x2 = ((pos.max(0).values.float() + 1) * torch.tensor([sy, sx])).tolist()
        else:
            d = (i - 1) % N_DISC_CH
            coefs[i, 1 + d] = 0.75 + 0.5 * float(torch.rand(1, generator=g))
            dcx, dcy, rad = disc[d]
            shrink = 1.0 / (1 + (i - 1) // N_DISC_CH)       # later instances: smaller boxes
            x1, x2 = (dcx - rad * shrink) * sx, (dcx + rad * shrink) * sx
            y1, y2 = (dcy - rad * shrink) * sy, (dcy + rad * shrink) * sy
        b = torch.tensor([x1 - jit[i, 0], y1 - jit[i, 1], x2 + jit[i, 2], y2 + jit[i, 3]])
        b[0::2] = b[0::2].clamp(0, W - 1.0)
        b[1::2] = b[1::2].clamp(0, H - 1.0)
        boxes[i] = b
    return protos.contiguous(), coefs.contiguous(), boxes.contiguous()


def make_batch(first_idx: int, B: int, n: int, H: int, W: int, mh: int, mw: int, K: int = 32,
               family: str = "sidewalk", max_n: int | None = None, pin: bool = False):
    """-> protos [B,K,mh,mw], coefs [B,max_n,K], boxes [B,max_n,4], counts [B] int32 (CPU tensors)."""
    max_n = max_n or n
    protos = torch.empty(B, K, mh, mw, pin_memory=pin)
    coefs = torch.zeros(B, max_n, K, pin_memory=pin)
    boxes = torch.zeros(B, max_n, 4, pin_memory=pin)
    counts = torch.full((B,), n, dtype=torch.int32)
    for b in range(B):
        p, c, bx = make_frame(first_idx + b, n, H, W, mh, mw, K, family)
        protos[b] = p
        coefs[b, :n] = c
        boxes[b, :n] = bx
    return protos, coefs, boxes, counts


def make_head_output(first_idx: int, B: int, A: int = 8400, nc: int = 1, K: int = 32, H: int = 640, W: int = 640,
                     n_objects: int = 6, ties: bool = False):
    """Synthetic raw segmentation-head output [B, 4 + nc + K, A] (cx, cy, w, h, class confidences, mask
    coefficients): a few objects, each answered by a cluster of overlapping anchors with high confidence (what
    NMS has to thin out), over a background of low-confidence anchors.  Seeded per image like make_frame."""
    out = torch.empty(B, 4 + nc + K, A)
    for b in range(B):
        g = torch.Generator().manual_seed(0xB3000000 + first_idx + b)
        p = torch.rand(4 + nc + K, A, generator=g)
        p[0] *= W; p[1] *= H
        p[2] = 8 + p[2] * 60; p[3] = 8 + p[3] * 60
        p[4:4 + nc] *= 0.45                                   # background below conf = 0.5
        p[4 + nc:] = (p[4 + nc:] - 0.5) * 2
        for o in range(n_objects):
            cx, cy = float(torch.rand(1, generator=g)) * W, float(torch.rand(1, generator=g)) * H
            bw, bh = 40 + float(torch.rand(1, generator=g)) * 300, 40 + float(torch.rand(1, generator=g)) * 300
            cls = int(torch.randint(0, nc, (1,), generator=g))
            m = int(torch.randint(3, 40, (1,), generator=g))
            idx = torch.randint(0, A, (m,), generator=g)
            jit = torch.randn(4, m, generator=g)
            p[0, idx] = cx + jit[0] * 6; p[1, idx] = cy + jit[1] * 6
            p[2, idx] = bw * (1 + 0.08 * jit[2]); p[3, idx] = bh * (1 + 0.08 * jit[3])
            conf = 0.5 + 0.5 * torch.rand(m, generator=g)
            if ties:
                conf = torch.round(conf * 16) / 16
            p[4:4 + nc, idx] = 0.1
            p[4 + cls, idx] = conf
        out[b] = p
    return out


# ---------------------------------------------------------------------------------------------
# on-device generator for frame STREAMS (BASELINE configs[3]: 65,536 frames do not fit in HBM at once and must not
# cross PCIe in the timed region): the same "sidewalk" family, every random number a counter-based hash of
# (SEED_BASE + frame index, draw index) - a frame's tensors depend on its index only, not on the chunking, the shard
# or the GPU that generates it.  (Not bit-identical to make_frame's CPU torch.Generator stream; parity tests use
# make_frame, throughput streams use this.)
# ---------------------------------------------------------------------------------------------
def _hash_u32(x: torch.Tensor) -> torch.Tensor:
    """splitmix-style avalanche on int64 lanes holding 32-bit values."""
    m = 0xFFFFFFFF
    x = (x ^ (x >> 16)) * 0x7FEB352D & m
    x = (x ^ (x >> 15)) * 0x846CA68B & m
    return (x ^ (x >> 16)) & m


def _uniform(seed: torch.Tensor, k: int, draws: int) -> torch.Tensor:
    """seed int64 [B] -> float32 [B, draws] in (0, 1), stream k."""
    idx = torch.arange(draws, device=seed.device, dtype=torch.int64)[None, :]
    h = _hash_u32(_hash_u32(seed[:, None] & 0xFFFFFFFF) ^ (idx * 0x9E3779B1 + k * 0x85EBCA6B & 0xFFFFFFFF))
    return ((h.to(torch.float64) + 0.5) / 4294967296.0).to(torch.float32)


def _normal(seed: torch.Tensor, k: int, draws: int) -> torch.Tensor:
    u1, u2 = _uniform(seed, 2 * k + 100, draws), _uniform(seed, 2 * k + 101, draws)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)


def make_batch_device(first_idx: int, B: int, n: int, H: int, W: int, mh: int, mw: int, K: int = 32, device="cuda"):
    """-> protos [B,K,mh,mw], coefs [B,n,K], boxes [B,n,4], counts [B] on `device` ("sidewalk" family)."""
    dev = torch.device(device)
    seed = (SEED_BASE + first_idx + torch.arange(B, device=dev, dtype=torch.int64))
    u = _uniform(seed, 0, 16)
    nrm = _normal(seed, 1, 8)
    sx, sy = W / mw, H / mh
    ys = torch.arange(mh, device=dev, dtype=torch.float32)[None, :, None]
    xs = torch.arange(mw, device=dev, dtype=torch.float32)[None, None, :]
    col = lambda t: t[:, None, None]
    cx0 = col(mw / 2 + nrm[:, 0] * mw / 8)
    y_top = col(mh * (0.125 + 0.375 * u[:, 0]))
    hw_top = col(mw * (1 / 16 + u[:, 1] / 8))
    hw_bot = col(mw * (0.25 + 0.25 * u[:, 2]))
    shear = col(0.3 * nrm[:, 1])
    t = ((ys - y_top) / (mh - 1 - y_top).clamp(min=1.0)).clamp(0, 1)
    halfw = hw_top + (hw_bot - hw_top) * t
    cx = cx0 + shear * (ys - y_top)
    field = torch.minimum(halfw - (xs - cx).abs(), ys - y_top + 0.5)            # [B, mh, mw]
    protos = torch.empty(B, K, mh, mw, device=dev)
    protos[:, 0] = field
    discs = []
    for d in range(N_DISC_CH):
        dcx, dcy = col(mw * (0.1 + 0.8 * u[:, 3 + 2 * d])), col(mh * (0.1 + 0.8 * u[:, 4 + 2 * d]))
        rad = col(mw * (1 / 16 + u[:, 9 + d] / 16))
        protos[:, 1 + d] = rad - ((xs - dcx) ** 2 + (ys - dcy) ** 2).sqrt()
        discs.append((dcx[:, 0, 0], dcy[:, 0, 0], rad[:, 0, 0]))
    n_noise = K - 1 - N_DISC_CH
    z = _normal(seed, 2, n_noise * 36).view(B, n_noise, 6, 6)
    protos[:, 1 + N_DISC_CH:] = 0.3 * F.interpolate(z, (mh, mw), mode="bicubic", align_corners=False)
    coefs = torch.zeros(B, n, K, device=dev)
    coefs[:, :, 1 + N_DISC_CH:] = 0.3 * _normal(seed, 3, n * n_noise).view(B, n, n_noise)
    boxes = torch.zeros(B, n, 4, device=dev)
    jit = _uniform(seed, 4, n * 4).view(B, n, 4) * 8.0
    amp = 0.75 + 0.5 * _uniform(seed, 5, n)
    pos = field > -1.0
    anyy, anyx = pos.any(2), pos.any(1)                                         # [B, mh], [B, mw]
    ar_y, ar_x = torch.arange(mh, device=dev), torch.arange(mw, device=dev)
    big = 1 << 20
    y1 = torch.where(anyy, ar_y, big).min(1).values.float() * sy
    y2 = (torch.where(anyy, ar_y, -1).max(1).values.float() + 1) * sy
    x1 = torch.where(anyx, ar_x, big).min(1).values.float() * sx
    x2 = (torch.where(anyx, ar_x, -1).max(1).values.float() + 1) * sx
    none = ~anyy.any(1)
    for i in range(n):
        if i == 0:
            coefs[:, 0, 0] = 1.0
            b = torch.stack([torch.where(none, 0.0, x1), torch.where(none, 0.0, y1),
                             torch.where(none, W - 1.0, x2), torch.where(none, H - 1.0, y2)], 1)
        else:
            d = (i - 1) % N_DISC_CH
            coefs[:, i, 1 + d] = amp[:, i]
            dcx, dcy, rad = discs[d]
            shrink = 1.0 / (1 + (i - 1) // N_DISC_CH)
            b = torch.stack([(dcx - rad * shrink) * sx, (dcy - rad * shrink) * sy, (dcx + rad * shrink) * sx,
                             (dcy + rad * shrink) * sy], 1)
        b = b + jit[:, i] * torch.tensor([-1.0, -1.0, 1.0, 1.0], device=dev)
        b[:, 0::2] = b[:, 0::2].clamp(0, W - 1.0)
        b[:, 1::2] = b[:, 1::2].clamp(0, H - 1.0)
        boxes[:, i] = b
    counts = torch.full((B,), n, dtype=torch.int32, device=dev)
    return protos.contiguous(), coefs.contiguous(), boxes.contiguous(), counts
