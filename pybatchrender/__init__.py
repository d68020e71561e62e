"""Drop-in import name: ``import pybatchrender as pbr`` resolves to :mod:`pybatchrender_b200`.

Code written against the reference imports ``pybatchrender`` absolutely (its Steering env does
``from pybatchrender import PBRConfig, PBRRenderer, PBREnv`` and ``pybatchrender.renderer.renderer``
is referenced by dotted path).  This package only aliases modules; all code
lives in ``pybatchrender_b200``.
"""
import importlib
import sys

import pybatchrender_b200 as _impl
from pybatchrender_b200 import *  # noqa: F401,F403
from pybatchrender_b200 import (CPUFrameGrabber, FrameGrabber, GPUFrameGrabber, PBRCam, PBRConfig, PBREnv, PBRLight,
                                PBRNode, PBRRenderer, PBRShaderContext, __version__, envs)

for _name in ("config", "env", "envs", "envs.cartpole", "envs.cartpole.config", "envs.cartpole.env",
              "envs.cartpole.renderer", "renderer", "renderer.renderer", "renderer.node", "renderer.camera",
              "renderer.light", "renderer.shader_context", "renderer.frame_grabber", "meshes", "dist"):
    try:
        sys.modules[f"{__name__}.{_name}"] = importlib.import_module(f"pybatchrender_b200.{_name}")
    except Exception:  # pragma: no cover - optional pieces
        pass

GPU_AVAILABLE = _impl.native_available()
