"""Drop-in for the reference's ProtrusionDetector live path (ProtrusionDetector.py:10-57, :59-158,
:419-442, :535): top-most occupied pixel row of the 21x21-px cell raster, split into runs, run
centres returned as Coordinates.  The raster is never materialised: the device kernel works on
the occupancy bits (closed form).  `.binary` is therefore None unless debug is on."""
from __future__ import annotations

from typing import ClassVar, Optional

import numpy as np

from . import config, models
from .engine import MaskGridEngine
from .materialise import objects_to_grid_input


class ProtrusionDetector:
    _instance: ClassVar[Optional["ProtrusionDetector"]] = None
    _initialized: bool = False

    def __new__(cls, debug: bool = False, imshow: bool = False):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
            cls._instance.debug = debug
            cls._instance.imshow = imshow
        return cls._instance

    def __init__(self, debug: bool = False, imshow: bool = False):
        if not self._initialized:
            self._initialized = True
            self.frame = None
            self._grids = None                   # list[list[Grid]], or a callable that builds it (lazy object view)
            self.height = 0
            self.width = 0
            self.binary = None
            self.frames_processed = 0
            self._engine: MaskGridEngine | None = None
            self._precomputed = None

    def bind_engine(self, engine: MaskGridEngine) -> None:
        self._engine = engine

    @property
    def grids(self):
        if callable(self._grids):
            self._grids = self._grids()
        return self._grids

    @grids.setter
    def grids(self, value) -> None:
        self._grids = value

    def from_record(self, frame: np.ndarray, peaks: np.ndarray, grids_source) -> list:
        """__call__ for a frame whose peaks the fused kernel already produced; `.grids` is built from `grids_source()`
        only if somebody reads it."""
        Coordinate, _, _ = models.classes()
        self.frame = frame
        self._grids = grids_source
        self.height, self.width = frame.shape[:2]
        self.frames_processed += 1
        self._precomputed = None
        return [Coordinate(x=int(x), y=int(y)) for x, y in peaks]

    def set_precomputed(self, grids, peaks: np.ndarray) -> None:
        """FrameProcessor hands over the peaks the fused kernel already produced for `grids`."""
        self._precomputed = (id(grids), peaks)

    def __call__(self, frame: np.ndarray, grids, grid_lookup) -> list:
        Coordinate, _, _ = models.classes()
        self.frame = frame
        self.grids = grids
        self.height, self.width = frame.shape[:2]
        self.frames_processed += 1
        if self._precomputed is not None and self._precomputed[0] == id(grids):
            peaks = self._precomputed[1]
        else:
            gs = config.grid_size
            e = self._engine
            if e is None or e.gs != gs or e.H != self.height or e.W != self.width:
                e = MaskGridEngine(H=self.height, W=self.width, mh=max(2, self.height // 4),
                                   mw=max(4, (self.width // 4) // 4 * 4), max_n=1, gs=gs, max_batch=1)
                self._engine = e
            gi = objects_to_grid_input(grids, None, gs)
            peaks = e.decode(e.grids_to_records([gi]))[0].peaks
        self._precomputed = None
        return [Coordinate(x=int(x), y=int(y)) for x, y in peaks]
