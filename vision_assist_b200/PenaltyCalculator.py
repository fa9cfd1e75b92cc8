"""Drop-in for the reference's PenaltyCalculator (PenaltyCalculator.py:8-156), computed on the GPU.

Same singleton, same three public methods and argument meaning.  The reference scores one cell per
call by walking `grid_lookup`; here the first `calculate_penalty` after a `_pre_compute_easy_segments`
scores the WHOLE grid with one `va_grid_to_penalty_peaks` launch and later calls are lookups into
that penalty map.  No CPU fallback: without the CUDA library the calls raise.
"""
from __future__ import annotations

from typing import ClassVar, Optional

import numpy as np

from . import config
from .engine import MaskGridEngine
from .materialise import objects_to_grid_input


class PenaltyCalculator:
    _instance: ClassVar[Optional["PenaltyCalculator"]] = None
    _initialized: bool = False

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
        return cls._instance

    def __init__(self):
        if not self._initialized:
            self._initialized = True
            self._engine: MaskGridEngine | None = None
            self._np_grids = np.empty((0, 0), dtype=np.uint8)
            self._grids = None
            self._map = None
            self._index = None

    def bind_engine(self, engine: MaskGridEngine) -> None:
        self._engine = engine

    def _get_engine(self, H: int, W: int) -> MaskGridEngine:
        e = self._engine
        if e is None or e.gs != config.grid_size or e.layout.rmax * e.gs < H or e.layout.cmax * e.gs < W:
            Hh, Ww = max(H, 4 * config.grid_size), max(W, 4 * config.grid_size)
            e = MaskGridEngine(H=Hh, W=Ww, mh=max(2, Hh // 4), mw=max(4, (Ww // 4) // 4 * 4), max_n=1,
                               gs=config.grid_size, max_batch=1)
            self._engine = e
        return e

    def _pre_compute_easy_segments(self, np_grids: np.ndarray, grids) -> None:
        """PenaltyCalculator.py:26-55: remember the grid; the segments are found on the device."""
        self._np_grids = np.asarray(np_grids)
        self._grids = grids
        self._map = None

    def _ensure_map(self, grid_lookup) -> None:
        if self._map is not None:
            return
        gs = config.grid_size
        use_easy = self._np_grids.ndim == 2 and self._np_grids.size > 0
        gi = objects_to_grid_input(self._grids, grid_lookup, gs, use_easy)
        H = int(max([gi["rows_y"].max(initial=0)] + ([gi["plane_y"].max(initial=0)] if "plane_y" in gi else []))) + gs
        W = int(gi["x0"] + gi["occ"].shape[1] * gs)
        eng = self._get_engine(H, W)
        rec = eng.decode(eng.grids_to_records([gi]))[0]
        self._map = rec.penalty
        self._index = {id(g): (k, c) for k, row in enumerate(self._grids) for c, g in enumerate(row)}

    def calculate_penalty(self, grid, grid_lookup):
        """PenaltyCalculator.py:112-142."""
        if grid.empty:
            return 0
        if self._grids is None:
            raise RuntimeError("_pre_compute_easy_segments must be called first (FrameProcessor.py:175)")
        self._ensure_map(grid_lookup)
        k, c = self._index[id(grid)]
        return float(self._map[k, c])

    def get_penalty_colour(self, penalty: float) -> tuple[int, int, int]:
        """PenaltyCalculator.py:144-152."""
        key = min(config.penalty_colour_gradient.keys(), key=lambda x: abs(x - penalty))
        return config.penalty_colour_gradient[key]


penalty_calculator = PenaltyCalculator()
