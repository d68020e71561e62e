"""pybatchrender_b200 -- B200-native drop-in for the pixel path of PyBatchRender.

Public surface mirrors ``pybatchrender/__init__.py:8-70`` of the reference: ``PBRConfig``,
``PBRRenderer``, ``PBRNode``, ``PBRCam``, ``PBRLight``, ``PBRShaderContext``, ``PBREnv``, ``envs``.
Rendering goes through hand-written sm_100a CUDA kernels in ``csrc/libpbr_b200.so`` (C ABI in
``include/pbr_b200.h``); there is no CPU, OpenGL or PyTorch fallback for the pixel path.
"""
from __future__ import annotations

__version__ = "0.1.0"

from .config import PBRConfig
from .renderer.shader_context import PBRShaderContext
from .renderer.node import PBRNode
from .renderer.camera import PBRCam
from .renderer.light import PBRLight
from .renderer.renderer import PBRRenderer
from .renderer.frame_grabber import CPUFrameGrabber, GPUFrameGrabber
from . import dist


def native_available() -> bool:
    """True when libpbr_b200.so is built and a CUDA device is visible."""
    import os
    import torch
    from . import _native
    return os.path.exists(_native.LIB_PATH) and torch.cuda.is_available()


GPU_AVAILABLE = None  # resolved lazily by native_available(); kept for reference API parity
FrameGrabber = GPUFrameGrabber   # reference __init__.py:43-50: "prefer GPU if available"; rendering here is GPU-only

try:  # env layer is optional (mirrors the guarded imports of the reference package)
    from .env import PBREnv
    from . import envs
except Exception:  # pragma: no cover
    PBREnv = None
    envs = None

__all__ = ["PBRConfig", "PBRShaderContext", "PBRNode", "PBRCam", "PBRLight", "PBRRenderer", "PBREnv",
           "envs", "dist", "native_available", "GPUFrameGrabber", "CPUFrameGrabber", "FrameGrabber", "__version__"]
