"""Drop-in for the reference's FrameProcessor (FrameProcessor.py:17-360) with the per-frame
data-parallel stage on the GPU.

Kept: the singleton, constructor `(model, verbose, debug, imshow)` (imshow defaults to False because
the reference's own main.py:44 omits it), the attributes other modules read (`frame`, `grids`,
`grid_lookup`, `np_grids`, `protrusion_detector`, `model`), the de-facto public privates
(`_extract_grid_information`, `_calculate_penalties`, `_create_graph`, `_find_paths`) and
`__call__`'s return conventions (`[]` / `(frame, [])` when no grid).

Two model duck types are accepted by `model.predict(frame, conf=0.5, verbose=...)`:
  * head tensors: a result with `.protos [K,mh,mw]`, `.coefs [n,K]`, `.boxes [n,4]` CUDA tensors
    (see `HeadOutputModel`) -> one `va_run_fused` launch sequence does mask assembly, grid,
    penalties and peaks;
  * raw head output: a result with `.pred [4+nc+K, A]` and `.protos` (see `RawHeadResult`) -> `va_nms` (confidence
    filter + NMS of ops.non_max_suppression) and the same fused path, still without a host round trip;
  * ultralytics-style `result.masks.xy` polygons (FrameProcessor.py:67-73) -> host cv2.fillPoly of
    the contourArea-largest polygon, then `va_mask_to_records`.
A* path finding / path analysis stay host code: pass the reference's `path_finder`,
`path_analyser` (and `path_visualiser`, `get_closest_grid_to_point`) via `bind_host_stages`.
"""
from __future__ import annotations

from collections import defaultdict
from typing import ClassVar, Optional

import cv2
import numpy as np
import torch

from . import config, models
from .engine import FrameRecord, MaskGridEngine
from .materialise import record_to_objects
from .PenaltyCalculator import penalty_calculator
from .ProtrusionDetector import ProtrusionDetector


class HeadOutputResult:
    """One frame of YOLOv8-seg head outputs after NMS (what ops.process_mask consumes)."""

    def __init__(self, protos: torch.Tensor, coefs: torch.Tensor, boxes: torch.Tensor):
        self.protos, self.coefs, self.boxes = protos, coefs, boxes
        self.masks = None


class RawHeadResult:
    """One frame of RAW segmentation-head output: `pred` [4 + nc + K, A] (cx, cy, w, h, class confidences, mask
    coefficients per anchor) and `protos` [K, mh, mw]; the confidence filter and NMS of
    ops.non_max_suppression run on the GPU (va_nms) in front of the mask path."""

    def __init__(self, protos: torch.Tensor, pred: torch.Tensor, nc: int = 1, conf: float = 0.5, iou: float = 0.7):
        self.protos, self.pred, self.nc, self.conf, self.iou = protos, pred, nc, conf, iou
        self.masks = None


class HeadOutputModel:
    """Model shim: predict() returns pre-computed head tensors (random-init stand-in for the absent
    weights, or a wrapper around a real network's (protos, coefs, boxes))."""

    def __init__(self, fn):
        self.fn = fn

    def predict(self, frame, conf=0.5, verbose=False):
        protos, coefs, boxes = self.fn(frame)
        return [HeadOutputResult(protos, coefs, boxes)]


class FrameProcessor:
    _instance: ClassVar[Optional["FrameProcessor"]] = None
    _initialized: bool = False

    def __new__(cls, model, verbose: bool = False, debug: bool = False, imshow: bool = False) -> "FrameProcessor":
        if cls._instance is None:
            cls._instance = super().__new__(cls)
        return cls._instance

    def __init__(self, model, verbose: bool = False, debug: bool = False, imshow: bool = False) -> None:
        if not self._initialized:
            self._initialized = True
            self.model = model
            self.verbose = verbose
            self.debug = debug
            self.imshow = imshow
            self.frame: Optional[np.ndarray] = None
            self._pending_record: FrameRecord | None = None   # record whose Grid objects have not been built yet
            self._grids: list = []
            self._grid_lookup: dict = {}
            self.np_grids: np.ndarray = np.empty((0, 0), dtype=np.uint8)
            self.protrusion_detector = ProtrusionDetector(debug=debug, imshow=imshow)
            self.frame_record: FrameRecord | None = None
            self._engine: MaskGridEngine | None = None
            self._penalties_ready = False
            self.path_finder = None
            self.path_analyser = None
            self.path_visualiser = None
            self.get_closest_grid_to_point = None
            self.Path = None
            self.array_paths: list = []          # [(cells [(x, y), ...], cost)] of the last frame (array A*, no host stages)

    # -- grids / grid_lookup: the reference's object view of the last frame, built from the GPU record on first access
    #    (about a thousand validated Grid / Coordinate objects per frame cost milliseconds, the record 60 us) ---------
    def _materialise(self) -> None:
        rec = self._pending_record
        if rec is not None:
            self._pending_record = None
            self._grids, self._grid_lookup, _ = record_to_objects(rec, config.grid_size)
            self.protrusion_detector.set_precomputed(self._grids, rec.peaks)

    @property
    def grids(self) -> list:
        self._materialise()
        return self._grids

    @grids.setter
    def grids(self, value: list) -> None:
        self._pending_record = None
        self._grids = value

    @property
    def grid_lookup(self) -> dict:
        self._materialise()
        return self._grid_lookup

    @grid_lookup.setter
    def grid_lookup(self, value: dict) -> None:
        self._pending_record = None
        self._grid_lookup = value

    def _has_grid(self) -> bool:
        return self._pending_record is not None or bool(self._grids)

    def bind_host_stages(self, path_finder=None, path_analyser=None, path_visualiser=None,
                         get_closest_grid_to_point=None, Path=None) -> None:
        """Attach the reference's host-side stages (PathFinder.py, PathAnalyser.py, utils.py, models.Path)."""
        self.path_finder, self.path_analyser, self.path_visualiser = path_finder, path_analyser, path_visualiser
        self.get_closest_grid_to_point, self.Path = get_closest_grid_to_point, Path

    # -- engine management ---------------------------------------------------------------------
    def _engine_for(self, H, W, mh, mw, n) -> MaskGridEngine:
        e = self._engine
        gs = config.grid_size
        if e is None or (e.H, e.W, e.mh, e.mw, e.gs) != (H, W, mh, mw, gs) or e.max_n < n:
            e = MaskGridEngine(H=H, W=W, mh=mh, mw=mw, max_n=max(8, n), gs=gs, max_batch=1)
            self._engine = e
            penalty_calculator.bind_engine(e)
            self.protrusion_detector.bind_engine(e)
        return e

    # -- FrameProcessor.py:50-171 ----------------------------------------------------------------
    def _extract_grid_information(self, results) -> None:
        self.grids = []
        self.grid_lookup = {}
        self.np_grids = np.empty((0, 0), dtype=np.uint8)
        self.frame_record = None
        self._penalties_ready = False
        H, W = self.frame.shape[:2]
        gs = config.grid_size
        for result in results:
            rec = None
            if getattr(result, "pred", None) is not None:
                # raw head output: confidence filter + NMS (va_nms), then the fused mask -> grid path
                K, mh, mw = result.protos.shape
                eng = self._engine_for(H, W, mh, mw, 32)
                coefs, boxes, _, _, counts = eng.nms(result.pred.contiguous()[None], conf_thres=result.conf,
                                                     iou_thres=result.iou, nc=result.nc)
                n = int(counts[0])
                if n == 0:
                    continue
                records, _ = eng.run(result.protos.contiguous()[None], coefs, boxes, counts, write_masks=False)
                rec = eng.decode(records)[0]
            elif getattr(result, "protos", None) is not None:
                n = int(result.coefs.shape[0])
                if n == 0:
                    continue
                if n > 32:
                    raise ValueError(f"{n} instances in one frame: the GPU path carries at most 32 per frame "
                                     "(DESIGN.md section 6); pass the raw head output (RawHeadResult) to get the 32 best "
                                     "survivors of NMS, or keep the 32 highest-confidence rows")
                K, mh, mw = result.protos.shape
                eng = self._engine_for(H, W, mh, mw, n)
                dev = result.protos.device
                coefs = torch.zeros((1, eng.max_n, K), dtype=torch.float32, device=dev)
                boxes = torch.zeros((1, eng.max_n, 4), dtype=torch.float32, device=dev)
                coefs[0, :n] = result.coefs
                boxes[0, :n] = result.boxes
                counts = torch.tensor([n], dtype=torch.int32, device=dev)
                records, _ = eng.run(result.protos.contiguous()[None], coefs, boxes, counts, write_masks=False)
                rec = eng.decode(records)[0]
            else:
                if result.masks is None:
                    continue
                xy = result.masks.xy
                mask = max(xy, key=lambda p: cv2.contourArea(p)) if len(xy) > 1 else xy[0]   # :72-73
                points = np.int32([mask])                                                    # :75
                rect = cv2.boundingRect(points)                                              # :76
                raster = np.zeros((H, W), dtype=np.uint8)
                cv2.fillPoly(raster, points, 1)                                              # :85-86
                eng = self._engine_for(H, W, max(2, H // 4), max(4, (W // 4) // 4 * 4), 1)
                dev = torch.device("cuda", eng.device)
                m = torch.zeros((1, eng.max_n, H, W), dtype=torch.uint8, device=dev)
                m[0, 0] = torch.from_numpy(raster).to(dev)
                rec = eng.decode(eng.masks_to_records(
                    m, torch.tensor([1], dtype=torch.int32, device=dev),
                    rects=torch.tensor([list(rect)], dtype=torch.int32, device=dev),
                    sel=torch.tensor([0], dtype=torch.int32, device=dev)))[0]
            rec.raise_reference_errors()
            if rec.R == 0:
                return                                                                       # :99-101
            self.frame_record = rec
            self.np_grids = rec.np_grids
            self._pending_record = rec                  # grids / grid_lookup are built on first access
            self._penalties_ready = True

    # -- FrameProcessor.py:173-182 ---------------------------------------------------------------
    def _calculate_penalties(self) -> None:
        if self._penalties_ready:
            return                      # the fused launch already filled Grid.penalty
        penalty_calculator._pre_compute_easy_segments(self.np_grids, self.grids)
        for grid_row in self.grids:
            for grid in grid_row:
                if grid.empty:
                    continue
                grid.penalty = penalty_calculator.calculate_penalty(grid, self.grid_lookup)

    # -- FrameProcessor.py:184-207 (host, unchanged semantics) ------------------------------------
    def _create_graph(self) -> defaultdict:
        gs = config.grid_size
        graph = defaultdict(list)
        for grid_row in self.grids:
            for grid in grid_row:
                if grid.empty:
                    continue
                x, y = grid.coords.x, grid.coords.y
                for nx, ny in ((x + gs, y), (x - gs, y), (x, y + gs), (x, y - gs)):
                    if self.grid_lookup.get((nx, ny)):
                        graph[(x, y)].append(((nx, ny), np.sqrt((x - nx) ** 2 + (y - ny) ** 2)))
        return graph

    # -- FrameProcessor.py:230-271 (host; needs the reference's path_finder / Path) ---------------
    def _find_paths(self, protrusion_peaks, graph):
        if self.path_finder is None or self.Path is None:
            raise RuntimeError("bind_host_stages(path_finder=..., Path=...) first: "
                               "A* stays host code from the reference (PathFinder.py)")
        Coordinate, _, _ = models.classes()
        all_paths = []
        if not self.grids:
            return all_paths
        # start / end cells (utils.get_closest_grid_to_point, FrameProcessor.py:236-239): chosen on the GPU with the
        # record when the peaks are the record's own; otherwise the bound host function
        rec = self.frame_record
        device_cells = (rec is not None and rec.goals is not None and rec.start[0] >= 0
                        and len(rec.goals) == len(protrusion_peaks)
                        and all((p.x, p.y) == (int(q[0]), int(q[1])) for p, q in zip(protrusion_peaks, rec.peaks)))
        if device_cells:
            start = self.grids[rec.start[0]][rec.start[1]]
            ends = [self.grids[int(k)][int(c)] for k, c in rec.goals]
        else:
            if self.get_closest_grid_to_point is None:
                raise RuntimeError("bind_host_stages(get_closest_grid_to_point=...) is needed for peaks that did not "
                                   "come from the GPU record")
            start = self.get_closest_grid_to_point(Coordinate(x=self.frame.shape[1] // 2, y=self.frame.shape[0]), self.grids)
            ends = [self.get_closest_grid_to_point(peak, self.grids) for peak in protrusion_peaks]
        for end in ends:
            grid_path, total_cost = self.path_finder.find_path(graph, start, end, self.grid_lookup)
            if grid_path:
                all_paths.append(self.Path(grids=grid_path, total_cost=total_cost, path_type="path"))
            else:
                print("No path found.")
        unique, out = [], sorted(all_paths, key=lambda p: len(p.grids), reverse=True)
        for path in out:
            a = {(g.coords.x, g.coords.y) for g in path.grids}
            ok = True
            for other in unique:
                b = {(g.coords.x, g.coords.y) for g in other.grids}
                inter = len(a & b)
                sim = 0.0 if not a or not b else 1.0 if inter in (len(a), len(b)) else inter / len(a | b)
                if sim >= 0.90:
                    ok = False
                    break
            if ok:
                unique.append(path)
        return unique

    # -- FrameProcessor.py:230-271 on the record's arrays -----------------------------
    def _find_paths_arrays(self) -> list:
        """A* from the record's start cell to every peak's end cell with the array-based port of the reference's
        PathFinder (vision_assist_b200/PathFinder.py), then the reference's similarity filter (:255-269): longest paths
        first, a path is dropped when it shares >= 90 % of its cells (Jaccard; 1.0 for a subset) with a kept one.
        -> [(cells [(x, y), ...], total cost)]; no Grid object is created."""
        from .PathFinder import path_finder as array_path_finder
        rec = self.frame_record
        if rec is None or rec.R == 0:
            return []
        found = [p for p in array_path_finder.find_paths(rec, config.grid_size)]
        for p in found:
            if p is None:
                print("No path found.")
        paths = sorted((p for p in found if p is not None), key=lambda pc: len(pc[0]), reverse=True)   # stable, as list.sort
        unique = []
        for cells, cost in paths:
            a = set(cells)
            ok = True
            for other, _ in unique:
                b = set(other)
                inter = len(a & b)
                sim = 0.0 if not a or not b else 1.0 if inter in (len(a), len(b)) else inter / len(a | b)
                if sim >= 0.90:
                    ok = False
                    break
            if ok:
                unique.append((cells, cost))
        return unique

    # -- FrameProcessor.py:272-299 (debug drawing; de-facto public: utilities/generate_testing_grids/run_on_main.py:196) --
    def _draw_grid(self, grid, color) -> None:
        """Fill one cell on self.frame (corners inclusive, as cv2.fillPoly draws them)."""
        gs = config.grid_size
        x, y = grid.coords.x, grid.coords.y
        corners = np.array([[x, y], [x + gs, y], [x + gs, y + gs], [x, y + gs]], np.int32)
        cv2.fillPoly(self.frame, [corners], color)

    def _draw_non_path_grids(self) -> None:
        """Every non-empty cell in its penalty colour (PenaltyCalculator.get_penalty_colour, config.py:4-17)."""
        for grid_row in self.grids:
            for grid in grid_row:
                if grid.empty:
                    continue
                self._draw_grid(grid, penalty_calculator.get_penalty_colour(grid.penalty or 0))

    # -- FrameProcessor.py:301-360 ---------------------------------------------------------------
    def __call__(self, frame: np.ndarray):
        self.frame = frame
        results = self.model.predict(frame, conf=0.5, verbose=self.verbose)
        self._extract_grid_information(results)
        if not self._has_grid():
            return (self.frame, []) if self.debug else []
        self._calculate_penalties()
        if self.path_finder is None or self.path_analyser is None:
            # no host stages bound: the peaks of the record go back to the caller, the paths come from the array A*
            # (kept in self.array_paths) - no Grid object is built unless somebody reads .grids / .grid_lookup
            rec = self.frame_record
            if self._pending_record is not None and rec is not None:
                protrusion_peaks = self.protrusion_detector.from_record(frame, rec.peaks, lambda: self.grids)
            else:
                protrusion_peaks = self.protrusion_detector(frame, self.grids, self.grid_lookup)
            if not protrusion_peaks:
                print("No protrusions detected.")
            self.array_paths = self._find_paths_arrays()
            return (self.frame, protrusion_peaks) if self.debug else protrusion_peaks
        graph = self._create_graph()
        protrusion_peaks = self.protrusion_detector(frame, self.grids, self.grid_lookup)
        if not protrusion_peaks:
            print("No protrusions detected.")
        paths = self._find_paths(protrusion_peaks, graph)
        final_answer = self.path_analyser(frame.shape[0], frame.shape[1], paths)
        if self.debug:
            self._draw_non_path_grids()                                                      # :352
            if self.path_visualiser is not None:
                self.frame = self.path_visualiser(self.frame, paths)                         # :355
            return self.frame, final_answer
        return final_answer
