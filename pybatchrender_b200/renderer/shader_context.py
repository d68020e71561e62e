"""Base class shared by nodes, camera and light.

Reference: ``pybatchrender/renderer/shader_context.py:11-116``.  In the reference this class owns
GLSL program creation, RGBA32F buffer textures and ``setShaderInput`` plumbing.  Here the "shader
inputs" are plain device tensors that ``PBRRenderer`` hands to ``libpbr_b200.so`` each frame, so
what remains is the shared math:

* ``_pack_columns``        -- (B,4,4) -> texel j = column j              (shader_context.py:42-45)
* ``_rotation_mats_from_hpr`` -- R = Rz(H) @ Ry(P) @ Rx(R), radians      (shader_context.py:47-84)
  (note: P about Y and R about X -- the reference's own convention, not Panda3D's)
* uniform broadcast: objects without geometry push a value to every registered node
  (shader_context.py:87-97); nodes keep the values in ``shader_inputs``.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Literal

import torch


class PBRShaderContext(ABC):
    def __init__(self, showbase, backend: Literal["instanced", "loop"] = "instanced") -> None:
        self.base = showbase
        if backend == "loop":
            raise NotImplementedError("Loop backend is not implemented yet. Use instanced backend instead.")
        self.backend = backend

    @property
    def device(self) -> torch.device:
        return getattr(self.base, "device", torch.device("cpu"))

    @staticmethod
    def _pack_columns(mat_batch: torch.Tensor) -> torch.Tensor:
        return mat_batch.transpose(1, 2).to(torch.float32)

    @staticmethod
    def _rotation_mats_from_hpr(hpr_b3: torch.Tensor) -> torch.Tensor:
        """Euler (H, P, R) in radians -> [B,3,3] with R = Rz(H) @ Ry(P) @ Rx(R)."""
        hpr = hpr_b3.to(torch.float32)
        c, s = torch.cos(hpr), torch.sin(hpr)
        ch, cp, cr = c.unbind(-1)
        sh, sp, sr = s.unbind(-1)
        # closed form of the triple product (the zero/one entries of the factors make every
        # term below exact with respect to the reference's two batched matmuls up to one rounding)
        rows = [
            ch * cp, ch * sp * sr - sh * cr, ch * sp * cr + sh * sr,
            sh * cp, sh * sp * sr + ch * cr, sh * sp * cr - ch * sr,
            -sp, cp * sr, cp * cr,
        ]
        return torch.stack(rows, dim=-1).reshape(-1, 3, 3)

    def _set_shader_input(self, input_name: str, value) -> None:
        """Nodes store the value; camera/light (no geometry of their own) broadcast it to all nodes."""
        own = getattr(self, "shader_inputs", None)
        if own is not None:
            own[input_name] = value
            return
        for node in getattr(self.base, "_pbr_nodes", []):
            inputs = getattr(node, "shader_inputs", None)
            if inputs is not None:
                inputs[input_name] = value

    def _auto_screen_size_input(self) -> None:
        win = self.base.win
        self._set_shader_input("screenSize", (float(win.getXSize()), float(win.getYSize())))

    @abstractmethod
    def _register_self(self) -> None:
        raise NotImplementedError
