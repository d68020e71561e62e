"""``PBRLight`` -- one global ambient + one directional light + a strength blend.

Reference: ``pybatchrender/renderer/light.py:8-48`` (defaults lines 11-14) and the fragment shader
that consumes it, ``pybatchrender/shaders/basic.frag:33-37``:

    ndl   = max(dot(n, normalize(dirLightDir)), 0)      # the vector is used un-negated
    light = ambientCol + ndl * dirLightCol
    rgb   = color.rgb * mix(1, light, clamp(lightingStrength, 0, 1))

The values are uniforms of the frame (``pbr_frame_desc.ambient / dir_dir / dir_col / strength``);
they are also broadcast to every node's ``shader_inputs`` like the reference does.
"""
from __future__ import annotations

from typing import Literal

from .shader_context import PBRShaderContext


class PBRLight(PBRShaderContext):
    def __init__(self, showbase,
                 ambient: tuple[float, float, float] = (0.2, 0.2, 0.25),
                 dir_dir: tuple[float, float, float] = (0.4, -0.6, -0.7),
                 dir_col: tuple[float, float, float] = (1.0, 1.0, 1.0),
                 strength: float = 1.0,
                 backend: Literal["loop", "instanced"] = "instanced") -> None:
        super().__init__(showbase, backend=backend)
        self.set_strength(strength)
        self.set_directional(dir_dir, dir_col)
        self.set_ambient(ambient)
        self._register_self()

    def set_strength(self, strength: float) -> None:
        self.strength = float(strength)
        self._set_shader_input("lightingStrength", self.strength)

    def set_directional(self, dir_dir, dir_col) -> None:
        self.dir_dir = tuple(float(x) for x in dir_dir)
        self.dir_col = tuple(float(x) for x in dir_col)
        self._set_shader_input("dirLightDir", self.dir_dir)
        self._set_shader_input("dirLightCol", self.dir_col)

    def set_ambient(self, amb_col) -> None:
        self.ambient = tuple(float(x) for x in amb_col)
        self._set_shader_input("ambientCol", self.ambient)

    def attach(self, node) -> None:
        node._set_shader_input("lightingStrength", self.strength)
        node._set_shader_input("dirLightDir", self.dir_dir)
        node._set_shader_input("dirLightCol", self.dir_col)
        node._set_shader_input("ambientCol", self.ambient)

    def _register_self(self) -> None:
        self.base._pbr_light = self
