"""Boundary record types the host side consumes (reference models.py:17-42).

Same field names and semantics as the reference's pydantic models so that A* / PathAnalyser code
written against `vision_assist.models` works on these objects unchanged.  When this package is
used inside the reference repository call `bind_models(vision_assist.models)` so that the very
same classes are produced.
"""
from __future__ import annotations

from typing import Literal

from pydantic import BaseModel, computed_field

from . import config


class Coordinate(BaseModel):
    x: int
    y: int

    @computed_field
    @property
    def midpoint(self) -> tuple[int, int]:
        return (self.x + (config.grid_size // 2), self.y + (config.grid_size // 2))

    def to_tuple(self) -> tuple[int, int]:
        return (self.x, self.y)


class Grid(BaseModel):
    coords: Coordinate
    centre: Coordinate
    penalty: float | None
    row: int
    col: int
    empty: bool
    artificial: bool


class Peak(BaseModel):
    centre: Coordinate
    left: Coordinate | None = None
    right: Coordinate | None = None
    orientation: Literal["left", "right", "up"]


_bound = {"Coordinate": Coordinate, "Grid": Grid, "Peak": Peak}


def bind_models(module) -> None:
    """Produce `module.Coordinate/Grid/Peak` (e.g. the reference's vision_assist.models) instead."""
    for k in _bound:
        _bound[k] = getattr(module, k)


def classes():
    return _bound["Coordinate"], _bound["Grid"], _bound["Peak"]
