"""``PBRLight``: the frame's single ambient term + single directional light + strength blend.

Mirrors the public surface of the reference's light object (``pybatchrender/renderer/light.py:8-48``;
defaults from lines 11-14) and feeds the fragment-shader maths of ``shaders/basic.frag:33-37``::

    ndl   = max(dot(n, normalize(dirLightDir)), 0)      # the vector is used un-negated
    light = ambientCol + ndl * dirLightCol
    rgb   = color.rgb * mix(1, light, clamp(lightingStrength, 0, 1))

Here the four values travel as plain uniforms of the frame (``pbr_frame_desc.ambient``, ``dir_dir``,
``dir_col``, ``strength``).  They are also mirrored into every node's ``shader_inputs`` dict, which is
what the reference's uniform broadcast amounts to.
"""
from __future__ import annotations

from typing import Literal

from .shader_context import PBRShaderContext

_UNIFORM_OF = {"strength": "lightingStrength", "dir_dir": "dirLightDir", "dir_col": "dirLightCol",
               "ambient": "ambientCol"}


def _vec3(v) -> tuple[float, float, float]:
    x, y, z = (float(c) for c in v)
    return (x, y, z)


class PBRLight(PBRShaderContext):
    def __init__(self, showbase, ambient=(0.2, 0.2, 0.25), dir_dir=(0.4, -0.6, -0.7), dir_col=(1.0, 1.0, 1.0),
                 strength: float = 1.0, backend: Literal["loop", "instanced"] = "instanced") -> None:
        super().__init__(showbase, backend=backend)
        self.strength = float(strength)
        self.dir_dir, self.dir_col, self.ambient = _vec3(dir_dir), _vec3(dir_col), _vec3(ambient)
        self._publish(*_UNIFORM_OF)
        self._register_self()

    def _publish(self, *attrs: str) -> None:
        """Broadcast the named attributes to all registered nodes (light has no geometry of its own)."""
        for attr in attrs:
            self._set_shader_input(_UNIFORM_OF[attr], getattr(self, attr))

    def set_strength(self, strength: float) -> None:
        self.strength = float(strength)
        self._publish("strength")

    def set_directional(self, dir_dir, dir_col) -> None:
        self.dir_dir, self.dir_col = _vec3(dir_dir), _vec3(dir_col)
        self._publish("dir_dir", "dir_col")

    def set_ambient(self, amb_col) -> None:
        self.ambient = _vec3(amb_col)
        self._publish("ambient")

    def attach(self, node) -> None:
        """Give one (late-created) node the current values."""
        for attr, uniform in _UNIFORM_OF.items():
            node._set_shader_input(uniform, getattr(self, attr))

    def _register_self(self) -> None:
        self.base._pbr_light = self
