"""vision_assist_b200 - B200-native (sm_100a) implementation of Vision Assist's per-frame
data-parallel stage: YOLOv8-seg mask assembly -> occupancy grid -> penalty map -> protrusion peaks.

Host code is Python/PyTorch (device memory, streams, torch.distributed) above a C-ABI CUDA library
(`libva_sm100.so`, include/vision_assist_b200.h).  There is no CPU fallback.
"""
from . import config, models  # noqa: F401
from .engine import FrameRecord, MaskGridEngine  # noqa: F401

__all__ = ["MaskGridEngine", "FrameRecord", "config", "models"]
