"""Batched host API above the C ABI: device buffers in, per-frame records out.

`MaskGridEngine` owns one `va_ctx` (one CUDA device, one geometry).  torch is used only for
device memory, streams and pinned host buffers; all compute is in libva_sm100.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


@dataclass
class FrameRecord:
    """Decoded per-frame record (numpy views into the record blob, list order; see the C header)."""
    flags: int
    sel: int
    x0: int
    y0: int
    C: int
    R: int
    n_orphans: int
    area: int
    bbox: tuple
    contour_area2: int      # 2 * cv2.contourArea of the kept polygon (FrameProcessor.py:72-73)
    rows_y: np.ndarray      # int32 [R]
    rows_attr: np.ndarray   # int32 [R]
    occ: np.ndarray         # uint8 [R, C]   bit0 = non-empty, bit1 = artificial
    penalty: np.ndarray     # float64 [R, C] NaN where empty
    peaks: np.ndarray       # int32 [n_peaks, 2]
    orphan_y: np.ndarray    # int32 [n_orphans]
    orphan_occ: np.ndarray  # uint8 [n_orphans, C]
    _gs: int = 20
    start: tuple = (-1, -1)             # (list row, column) of the path start cell (FrameProcessor.py:236), (-1, -1) if none
    goals: np.ndarray | None = None     # int32 [n_peaks, 2]: (list row, column) of each peak's end cell (:238-239)
    lookup_row: np.ndarray | None = None   # int32 [lookup_rows]: record row owning grid_lookup at y = ly * gs, -1 if none

    @property
    def np_grids(self) -> np.ndarray:
        """FrameProcessor.np_grids (FrameProcessor.py:168-171)."""
        return (self.occ & 1).astype(np.uint8)

    def as_dict(self) -> dict:
        return dict(flags=self.flags, sel=self.sel, x0=self.x0, y0=self.y0, C=self.C, R=self.R,
                    rows_y=self.rows_y, rows_attr=self.rows_attr, occ=self.occ, penalty=self.penalty,
                    peaks=self.peaks, orphan_y=self.orphan_y, orphan_occ=self.orphan_occ, start=self.start,
                    goals=self.goals, lookup_row=self.lookup_row)

    def neighbour_mask(self) -> np.ndarray:
        """FrameProcessor._create_graph (FrameProcessor.py:184-207) from the record's implicit form: uint8 [R, C],
        bit0 right, bit1 left, bit2 down, bit3 up; 0 for empty cells (they are not graph nodes)."""
        R, C = self.R, self.C
        out = np.zeros((R, C), np.uint8)
        if R == 0:
            return out
        node = (self.occ & 1).astype(bool)
        ly = self.rows_y // self._gs
        n = len(self.lookup_row)
        down = np.array([(l + 1 < n) and self.lookup_row[l + 1] >= 0 for l in ly])
        up = np.array([(l - 1 >= 0) and self.lookup_row[l - 1] >= 0 for l in ly])
        cols = np.arange(C)
        out |= (node & (cols + 1 < C)[None, :]).astype(np.uint8)
        out |= (node & (cols - 1 >= 0)[None, :]).astype(np.uint8) << 1
        out |= (node & down[:, None]).astype(np.uint8) << 2
        out |= (node & up[:, None]).astype(np.uint8) << 3
        return out

    def raise_reference_errors(self) -> None:
        """Re-raise what the reference raises for this frame."""
        if self.flags & _lib.VA_FLAG_CENTRE_OOB:
            raise IndexError("index out of bounds: cell centre outside the frame (FrameProcessor.py:97)")
        if self.flags & _lib.VA_FLAG_LIST_OOB:
            raise IndexError("list assignment index out of range (FrameProcessor.py:163)")
        if self.flags & _lib.VA_FLAG_NO_POLYGON:
            import cv2
            raise cv2.error("fillPoly: the selected polygon has no points (FrameProcessor.py:86)")
        if self.flags & _lib.VA_FLAG_OVERFLOW:
            raise RuntimeError("record capacity exceeded")


class MaskGridEngine:
    def __init__(self, H: int, W: int, mh: int, mw: int, max_n: int = 8, gs: int = 20, max_batch: int = 256,
                 device: int | None = None, K: int = 32, check_simple: bool = False, tensor_core: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("vision_assist_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        flags = (_lib.VA_CFG_CHECK_SIMPLE if check_simple else 0) | (0 if tensor_core else _lib.VA_CFG_NO_TENSOR_CORE)
        self.cfg = _lib.VaConfig(self.device, H, W, mh, mw, K, max_n, gs, max_batch, flags)
        ctx = C.c_void_p()
        torch.cuda.init()
        rc = self.lib.va_create(C.byref(ctx), C.byref(self.cfg))
        if rc != 0:
            raise _lib.VaError(rc, self.lib.va_last_error(None).decode())
        self._ctx = ctx
        self.layout = _lib.VaLayout()
        self._check(self.lib.va_get_layout(self._ctx, C.byref(self.layout)))
        self.H, self.W, self.mh, self.mw, self.K, self.max_n, self.gs, self.max_batch = H, W, mh, mw, K, max_n, gs, max_batch

    # -- plumbing ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self.lib.va_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise _lib.VaError(rc, self.lib.va_last_error(self._ctx).decode())

    @property
    def record_bytes(self) -> int:
        return self.layout.record_bytes

    @property
    def uses_tensor_core(self) -> bool:
        return bool(self.lib.va_uses_tensor_core(self._ctx))

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.va_last_launch_count(self._ctx))

    def profile(self, on: bool | int = True) -> None:
        """Record CUDA events around the assembly and tail kernels of every `on`-th run() (True = every call;
        C ABI va_profile_enable)."""
        self._check(self.lib.va_profile_enable(self._ctx, int(on)))

    def profile_read(self):
        """-> (assemble_ms_total, tail_ms_total, calls) since the last read; synchronises."""
        a, t, n = C.c_float(), C.c_float(), C.c_int32()
        self._check(self.lib.va_profile_read(self._ctx, C.byref(a), C.byref(t), C.byref(n)))
        return a.value, t.value, n.value

    def algorithmic_bytes_per_frame(self, n: int, write_masks: bool = True) -> int:
        """protos in + coefs + boxes + u8 masks out + grid/penalty record + header."""
        L = self.layout
        return (4 * self.K * self.mh * self.mw + 4 * n * self.K + 16 * n + (n * self.H * self.W if write_masks else 0)
                + L.rmax * L.cmax * 9 + 64)

    def _dev(self, t: torch.Tensor, dtype, shape_tail):
        if not (t.is_cuda and t.device.index == self.device and t.dtype == dtype and t.is_contiguous()):
            raise ValueError(f"expected contiguous {dtype} CUDA tensor on device {self.device}, got {t.dtype} {t.device}")
        if tuple(t.shape[1:]) != tuple(shape_tail):
            raise ValueError(f"expected trailing shape {tuple(shape_tail)}, got {tuple(t.shape)}")
        return C.c_void_p(t.data_ptr())

    def _inputs(self, protos, coefs, boxes, counts):
        B = protos.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        return (B, self._dev(protos, torch.float32, (self.K, self.mh, self.mw)),
                self._dev(coefs, torch.float32, (self.max_n, self.K)), self._dev(boxes, torch.float32, (self.max_n, 4)),
                self._dev(counts, torch.int32, ()))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- device-buffer API ---------------------------------------------------------------------
    def assemble_masks(self, protos, coefs, boxes, counts, want_logits: bool = False):
        """ops.process_mask for a batch -> masks u8 [B,max_n,H,W] (+ cropped logits [B,max_n,mh,mw])."""
        B, p, c, b, n = self._inputs(protos, coefs, boxes, counts)
        masks = torch.empty((B, self.max_n, self.H, self.W), dtype=torch.uint8, device=protos.device)
        # zeros: the kernels only write the rows of live instances (slots >= counts[b] stay 0)
        logits = torch.zeros((B, self.max_n, self.mh, self.mw), dtype=torch.float32, device=protos.device) if want_logits else None
        self._check(self.lib.va_assemble_masks(self._ctx, p, c, b, n, B, C.c_void_p(masks.data_ptr()),
                                               C.c_void_p(logits.data_ptr()) if want_logits else None, self._stream()))
        return (masks, logits) if want_logits else masks

    def nms(self, pred: torch.Tensor, conf_thres: float = 0.5, iou_thres: float = 0.7, nc: int = 1,
            max_det: int | None = None, agnostic: bool = False, max_wh: int = 7680):
        """ops.non_max_suppression (+ torchvision nms) on the raw head output `pred` [B, 4 + nc + K, A] ->
        (coefs [B,max_n,K], boxes [B,max_n,4] xyxy, conf [B,max_n], cls [B,max_n] i32, counts [B] i32): the first
        three in exactly the layout `run()` takes.  Any number of candidates (tiles of 512 in score order)."""
        if not (pred.is_cuda and pred.dtype == torch.float32 and pred.is_contiguous() and pred.dim() == 3):
            raise ValueError("expected a contiguous float32 CUDA tensor [B, 4 + nc + K, A]")
        B, Cc, A = pred.shape
        if Cc != 4 + nc + self.K:
            raise ValueError(f"expected {4 + nc + self.K} rows (4 + nc + K), got {Cc}")
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        dev = pred.device
        coefs = torch.empty((B, self.max_n, self.K), dtype=torch.float32, device=dev)
        boxes = torch.empty((B, self.max_n, 4), dtype=torch.float32, device=dev)
        conf = torch.empty((B, self.max_n), dtype=torch.float32, device=dev)
        cls = torch.empty((B, self.max_n), dtype=torch.int32, device=dev)
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        prm = _lib.VaNmsParams(conf_thres, iou_thres, nc, self.max_n if max_det is None else max_det, int(agnostic), max_wh)
        self._check(self.lib.va_nms(self._ctx, C.c_void_p(pred.data_ptr()), A, C.byref(prm), B, C.c_void_p(coefs.data_ptr()),
                                    C.c_void_p(boxes.data_ptr()), C.c_void_p(conf.data_ptr()), C.c_void_p(cls.data_ptr()),
                                    C.c_void_p(counts.data_ptr()), self._stream()))
        return coefs, boxes, conf, cls, counts

    def scale_boxes(self, boxes: torch.Tensor, counts: torch.Tensor, img1_shape, img0_shape) -> torch.Tensor:
        """ops.scale_boxes(img1_shape, boxes, img0_shape) + clip_boxes (ops.py:139-174) on `nms()`'s boxes [B,max_n,4]:
        letterboxed model input (h, w) -> original frame (h, w).  Returns a new tensor; slots >= counts are zero."""
        if not (boxes.is_cuda and boxes.dtype == torch.float32 and boxes.is_contiguous() and boxes.shape[1:] == (self.max_n, 4)):
            raise ValueError(f"expected a contiguous float32 CUDA tensor [B, {self.max_n}, 4]")
        B = boxes.shape[0]
        if counts.shape != (B,) or counts.dtype != torch.int32 or not counts.is_cuda:
            raise ValueError("counts must be an int32 CUDA tensor [B]")
        out = torch.empty_like(boxes)
        self._check(self.lib.va_scale_boxes(self._ctx, C.c_void_p(boxes.data_ptr()), C.c_void_p(counts.data_ptr()), B,
                                            int(img1_shape[0]), int(img1_shape[1]), int(img0_shape[0]), int(img0_shape[1]),
                                            C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def run(self, protos, coefs, boxes, counts, masks_out: torch.Tensor | None = None,
            records_out: torch.Tensor | None = None, write_masks: bool = True, records_ptr: int | None = None):
        """Whole path; returns (records u8 [B, record_bytes], masks or None), all on the device.
        `records_ptr`: raw device address for the records instead of a tensor - e.g. a slot of another GPU's buffer
        mapped with `peer_open` (the tail kernel then stores the records over NVLink); returns (None, masks)."""
        B, p, c, b, n = self._inputs(protos, coefs, boxes, counts)
        if records_ptr is None and records_out is None:
            records_out = torch.empty((B, self.record_bytes), dtype=torch.uint8, device=protos.device)
        if write_masks and masks_out is None:
            masks_out = torch.empty((B, self.max_n, self.H, self.W), dtype=torch.uint8, device=protos.device)
        mptr = C.c_void_p(masks_out.data_ptr()) if (write_masks and masks_out is not None) else None
        rptr = C.c_void_p(records_ptr if records_ptr is not None else records_out.data_ptr())
        self._check(self.lib.va_run_fused(self._ctx, p, c, b, n, B, mptr, rptr, self._stream()))
        return (records_out if records_ptr is None else None), (masks_out if write_masks else None)

    # -- multi-GPU record sink (peer memory over NVLink, C ABI va_peer_* / va_signal / va_wait_flags) ---------
    def peer_alloc(self, nbytes: int) -> tuple[int, bytes]:
        """-> (device address, 64-byte IPC handle) of a zeroed buffer other processes can map with peer_open."""
        ptr, h = C.c_void_p(), C.create_string_buffer(_lib.VA_IPC_HANDLE_BYTES)
        self._check(self.lib.va_peer_alloc(self._ctx, int(nbytes), C.byref(ptr), h))
        return int(ptr.value), bytes(h.raw)

    def peer_open(self, handle: bytes) -> int:
        ptr = C.c_void_p()
        self._check(self.lib.va_peer_open(self._ctx, C.create_string_buffer(handle, _lib.VA_IPC_HANDLE_BYTES), C.byref(ptr)))
        return int(ptr.value)

    def peer_close(self, ptr: int) -> None:
        self._check(self.lib.va_peer_close(self._ctx, C.c_void_p(ptr)))

    def peer_free(self, ptr: int) -> None:
        self._check(self.lib.va_peer_free(self._ctx, C.c_void_p(ptr)))

    def peer_put(self, dst_ptr: int, src_ptr: int, nbytes: int) -> None:
        """Enqueue a device-to-device copy (copy engine) into a local or peer-mapped buffer on the current stream."""
        self._check(self.lib.va_peer_put(self._ctx, C.c_void_p(dst_ptr), C.c_void_p(src_ptr), int(nbytes), self._stream()))

    def signal(self, flag_ptr: int, value: int) -> None:
        """Enqueue `*flag = value` (system scope) behind everything already queued on the current stream."""
        self._check(self.lib.va_signal(self._ctx, C.c_void_p(flag_ptr), int(value), self._stream()))

    def wait_flags(self, flags_ptr: int, n: int, value: int) -> None:
        """Enqueue a wait on the current stream until flags[i] >= value for i < n."""
        self._check(self.lib.va_wait_flags(self._ctx, C.c_void_p(flags_ptr), int(n), int(value), self._stream()))

    def masks_to_records(self, masks, counts, rects: torch.Tensor | None = None, sel: torch.Tensor | None = None):
        B = masks.shape[0]
        m = self._dev(masks, torch.uint8, (self.max_n, self.H, self.W))
        n = self._dev(counts, torch.int32, ())
        rec = torch.empty((B, self.record_bytes), dtype=torch.uint8, device=masks.device)
        r = self._dev(rects, torch.int32, (4,)) if rects is not None else None
        s = self._dev(sel, torch.int32, ()) if sel is not None else None
        self._check(self.lib.va_mask_to_records(self._ctx, m, n, B, r, s, C.c_void_p(rec.data_ptr()), self._stream()))
        return rec

    def grids_to_records(self, grids: list[dict]):
        """grids: [{x0, rows_y, rows_attr, occ[R,C] u8, use_easy, plane_y?, plane_occ?}] -> records (device)."""
        B = len(grids)
        L = self.layout
        hdr = (_lib.VaGridInput * B)()
        row_y = np.zeros((B, L.rmax), np.int32)
        row_attr = np.zeros((B, L.rmax), np.int32)
        occ = np.zeros((B, L.rmax, L.cmax), np.uint8)
        plane_y = np.zeros((B, L.rmax), np.int32)
        plane_occ = np.zeros((B, L.rmax, L.cmax), np.uint8)
        any_plane = False
        for i, g in enumerate(grids):
            o = np.asarray(g["occ"], np.uint8)
            R, Cc = o.shape
            if R > L.rmax or Cc > L.cmax:
                raise ValueError(f"grid {R}x{Cc} exceeds engine capacity {L.rmax}x{L.cmax}")
            npl = 0
            if g.get("plane_y") is not None:
                po = np.asarray(g["plane_occ"], np.uint8)
                npl = po.shape[0]
                if npl > L.rmax:
                    raise ValueError("too many lookup rows")
                plane_y[i, :npl] = g["plane_y"]
                plane_occ[i, :npl, :Cc] = po
                any_plane = True
            hdr[i] = _lib.VaGridInput(int(g.get("x0", 0)), Cc, R, npl, int(g.get("use_easy", 1)))
            row_y[i, :R] = g["rows_y"]
            row_attr[i, :R] = g["rows_attr"]
            occ[i, :R, :Cc] = o
        dev = torch.device("cuda", self.device)
        d_hdr = torch.frombuffer(bytearray(bytes(hdr)), dtype=torch.uint8).to(dev)
        d_y, d_a, d_o = (torch.from_numpy(a).to(dev) for a in (row_y, row_attr, occ))
        d_py, d_po = (torch.from_numpy(a).to(dev) for a in (plane_y, plane_occ)) if any_plane else (None, None)
        rec = torch.empty((B, self.record_bytes), dtype=torch.uint8, device=dev)
        self._check(self.lib.va_grid_to_penalty_peaks(
            self._ctx, C.c_void_p(d_hdr.data_ptr()), C.c_void_p(d_y.data_ptr()), C.c_void_p(d_a.data_ptr()),
            C.c_void_p(d_o.data_ptr()), C.c_void_p(d_py.data_ptr()) if any_plane else None,
            C.c_void_p(d_po.data_ptr()) if any_plane else None, B, C.c_void_p(rec.data_ptr()), self._stream()))
        return rec

    # -- host-buffer API -----------------------------------------------------------------------
    def run_host(self, protos, coefs, boxes, counts, records_out: torch.Tensor | None = None,
                 masks_out: torch.Tensor | None = None):
        """Host tensors in (pinned recommended), records (and optionally masks) back in host memory.  `protos` may be
        float16 (a model run with half=True; the reference computes on protos.float(), ops.py:724): half the PCIe bytes,
        widened exactly on the device."""
        f16 = protos.dtype == torch.float16
        for t, dt in ((protos, torch.float16 if f16 else torch.float32), (coefs, torch.float32), (boxes, torch.float32),
                      (counts, torch.int32)):
            if t.is_cuda or t.dtype != dt or not t.is_contiguous():
                raise ValueError("run_host expects contiguous CPU tensors (f32 or f16 prototypes, f32, f32, i32)")
        B = protos.shape[0]
        if B > self.max_batch:
            raise ValueError(f"batch {B} > max_batch {self.max_batch}")
        if records_out is None:
            records_out = torch.empty((B, self.record_bytes), dtype=torch.uint8, pin_memory=True)
        fn = self.lib.va_run_fused_host_f16 if f16 else self.lib.va_run_fused_host
        self._check(fn(
            self._ctx, C.c_void_p(protos.data_ptr()), C.c_void_p(coefs.data_ptr()), C.c_void_p(boxes.data_ptr()),
            C.c_void_p(counts.data_ptr()), B, C.c_void_p(masks_out.data_ptr()) if masks_out is not None else None,
            C.c_void_p(records_out.data_ptr())))
        return records_out

    # -- decoding ------------------------------------------------------------------------------
    def decode(self, records) -> list[FrameRecord]:
        """records: u8 [B, record_bytes] torch (any device) or numpy -> list of FrameRecord."""
        if isinstance(records, torch.Tensor):
            records = records.cpu().numpy()
        return [decode_record(records[i], self.layout, self.gs) for i in range(records.shape[0])]


def decode_record(blob: np.ndarray, L, gs: int = 20) -> FrameRecord:
    blob = np.ascontiguousarray(blob)
    h = blob[:64].view(np.int32)
    flags, sel, x0, y0, Cc, R, norph, npk, area, rm, minx, miny, maxx, maxy, area2, start = (int(v) for v in h[:16])
    ry = blob[L.off_row_y:L.off_row_y + 4 * L.rmax].view(np.int32)
    ra = blob[L.off_row_attr:L.off_row_attr + 4 * L.rmax].view(np.int32)
    pen = blob[L.off_penalty:L.off_penalty + 8 * L.rmax * L.cmax].view(np.float64).reshape(L.rmax, L.cmax)
    pk = blob[L.off_peaks:L.off_peaks + 8 * L.pmax].view(np.int32).reshape(L.pmax, 2)
    occ = blob[L.off_occ:L.off_occ + L.rmax * L.cmax].reshape(L.rmax, L.cmax)
    return FrameRecord(flags=flags, sel=sel, x0=x0, y0=y0, C=Cc, R=R, n_orphans=norph, area=area,
                       bbox=(minx, miny, maxx, maxy), contour_area2=area2, rows_y=ry[:R].copy(), rows_attr=ra[:R].copy(),
                       occ=occ[:R, :Cc].copy(), penalty=pen[:R, :Cc].copy(), peaks=pk[:npk].copy(),
                       orphan_y=ry[R:R + norph].copy(), orphan_occ=occ[R:R + norph, :Cc].copy(), _gs=gs,
                       start=(start >> 16, start & 0xffff) if start >= 0 else (-1, -1),
                       goals=blob[L.off_goals:L.off_goals + 8 * L.pmax].view(np.int32).reshape(L.pmax, 2)[:npk].copy(),
                       lookup_row=blob[L.off_lookup:L.off_lookup + 4 * L.lookup_rows].view(np.int32).copy())
