"""Test-only harness that imports the UNMODIFIED reference from /root/reference.

The reference is a package called `vision_assist` whose directory is the repo root
(`from vision_assist.X import ...`, FrameProcessor.py:7-14) and it imports `ultralytics`
only for a type annotation (FrameProcessor.py:5,24).  We expose it through a symlink in a
temp dir and stub `ultralytics`.  /root/reference only exists in the build container, so
everything that uses this module is skipped on the GPU box; what travels there are the
golden vectors under tests/golden/ generated with it (tests/golden/make_golden.py).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import tempfile
import types

REFERENCE_ROOT = "/root/reference"
_state: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "FrameProcessor.py"))


def load():
    """Return a namespace with the reference modules (cached)."""
    if _state:
        return types.SimpleNamespace(**_state)
    if not available():
        raise RuntimeError("reference not present at /root/reference")
    d = tempfile.mkdtemp(prefix="va_ref_")
    os.symlink(REFERENCE_ROOT, os.path.join(d, "vision_assist"))
    sys.path.insert(0, d)
    if "ultralytics" not in sys.modules:
        stub = types.ModuleType("ultralytics")
        stub.YOLO = object
        sys.modules["ultralytics"] = stub
    for name in ("config", "models", "utils", "PenaltyCalculator", "ProtrusionDetector",
                 "PathFinder", "PathAnalyser", "PathVisualiser", "FrameProcessor"):
        _state[name] = importlib.import_module(f"vision_assist.{name}")
    spec = importlib.util.spec_from_file_location(
        "va_ref_ops", os.path.join(REFERENCE_ROOT, "testing/old/segmenting_using_tflite/ops.py"))
    ops = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(ops)
    except Exception:
        # the vendored file imports ultralytics.utils helpers at module scope; provide stubs
        ops = _load_ops_with_stubs()
    _state["ops"] = ops
    return types.SimpleNamespace(**_state)


def _load_ops_with_stubs():
    import logging
    u = sys.modules["ultralytics"]
    utils = types.ModuleType("ultralytics.utils")
    utils.LOGGER = logging.getLogger("ultralytics")
    metrics = types.ModuleType("ultralytics.utils.metrics")
    metrics.batch_probiou = lambda *a, **k: None
    u.utils = utils
    utils.metrics = metrics
    sys.modules["ultralytics.utils"] = utils
    sys.modules["ultralytics.utils.metrics"] = metrics
    spec = importlib.util.spec_from_file_location(
        "va_ref_ops", os.path.join(REFERENCE_ROOT, "testing/old/segmenting_using_tflite/ops.py"))
    ops = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ops)
    return ops


class FakeMasks:
    def __init__(self, xy):
        self.xy = xy


class FakeResult:
    def __init__(self, xy):
        self.masks = FakeMasks(xy) if xy is not None else None


class FakeModel:
    """Duck-typed model: predict() -> [result(.masks.xy)] (FrameProcessor.py:67-73, 322)."""

    def __init__(self, xy):
        self.xy = xy

    def predict(self, frame, conf=0.5, verbose=False):
        return [FakeResult(self.xy)]


def new_frame_processor(ref, model=None):
    """A fresh (non-singleton) reference FrameProcessor."""
    FP = ref.FrameProcessor.FrameProcessor
    FP._instance = None
    FP._initialized = False
    PD = ref.ProtrusionDetector.ProtrusionDetector
    PD._instance = None
    PD._initialized = False
    return FP(model, False, False, False)
