"""The CartPole scene: one shared rail, a cart and a pole per scene, a fixed camera, default light.

Scene constants are the reference's (``pybatchrender/envs/cartpole/renderer.py:18-96``, SURVEY.md 8
row a18): all three parts are ``models/box``; rail 6 x 0.05 x 0.05 and cart 1.2 x 0.8 x 0.5 pivot at
their centre, the pole 0.1 x 0.1 x 2 pivots 5 % above its lower end and stands (0.8 + 0.1) / 2 in
front of the cart; cart colours ramp over the scenes; camera at (5, 5, 2) looking at the origin.
``_step(state[B, 4])`` maps ``x = state[:, 0]`` to the cart / pole position and ``theta = state[:, 2]``
to the pole's P angle (a rotation about +Y in the reference's own HPR convention,
``shader_context.py:47-84``).

What differs from the reference: the state is consumed where it lives.  The reference moves it to
the CPU and re-uploads three whole buffers per step (``renderer.py:112,130-138``); here the two
nodes are *bound* to ``state[:, 0]`` / ``state[:, 2]`` (``PBRNode.set_pose``, strided views) and the
raster kernel computes their model matrices itself while it transforms the vertices: a step is one
kernel launch.  Without a GPU (``cfg.device == 'cpu'``) the generic
torch setters run instead so that the host logic stays testable -- rendering itself needs CUDA.
"""
from __future__ import annotations

import math

import torch

from ...config import PBRConfig
from ...renderer.renderer import PBRRenderer

_PARTS = {
    #        size (x, y, z)        pivot (bounds-relative)   shared
    "rail": ((6.0, 0.05, 0.05), (0.5, 0.5, 0.5), True),
    "cart": ((1.2, 0.8, 0.5), (0.5, 0.5, 0.5), False),
    "pole": ((0.1, 0.1, 2.0), (0.5, 0.5, 0.05), False),
}
_RAIL_RGBA = (0.2, 0.2, 0.2, 1.0)
_POLE_RGBA = (1.0, 0.7, 0.2, 1.0)
_CART_RGBA_RAMP = ((0.6, 0.8, 1.0, 1.0), (1.0, 0.6, 0.8, 1.0))
# pose channel (0..2 position, 3..5 H/P/R) <- state column: x drives both nodes, theta the pole's P angle
_CART_COLUMNS = ((0, 0),)
_POLE_COLUMNS = ((0, 0), (4, 2))


class CartPoleRenderer(PBRRenderer):
    def __init__(self, cfg: PBRConfig | dict | None = None, **cfg_overrides):
        super().__init__(cfg, **cfg_overrides)
        dev, N, I = self.device, int(self.cfg.num_scenes), 1
        f32 = dict(dtype=torch.float32, device=dev)

        for name, (size, pivot, shared) in _PARTS.items():
            setattr(self, f"{name}_size", size)
            setattr(self, name, self.add_node("models/box", instances_per_scene=I, model_scale=size,
                                              model_pivot_relative_point=pivot, shared_across_scenes=shared))
        self.rail_pos_color, self.pole_pos_color = _RAIL_RGBA, _POLE_RGBA
        self.cart_pos_color_range = _CART_RGBA_RAMP

        # positions: everything on the rail axis, the pole in front of the cart
        self.pole_y = 0.5 * (self.cart_size[1] + self.pole_size[1])
        self.rail_pos = torch.zeros((1, I, 3), **f32)
        self.cart_pos = torch.zeros((N, I, 3), **f32)
        self.pole_pos = torch.zeros((N, I, 3), **f32)
        self.pole_pos[..., 1] = self.pole_y
        for node, pos in ((self.rail, self.rail_pos), (self.cart, self.cart_pos), (self.pole, self.pole_pos)):
            node.set_positions(pos)

        # colours: the cart ramps linearly over the *global* batch (a shard keeps its slice, see dist.py)
        first = int(getattr(self.cfg, "scene_offset", 0) or 0)
        n_global = int(getattr(self.cfg, "global_num_scenes", 0) or 0) or N
        lo, hi = (torch.tensor(c, dtype=torch.float32) for c in _CART_RGBA_RAMP)
        ramp = torch.linspace(0.0, 1.0, steps=n_global, dtype=torch.float32)[first:first + N]
        self.rail_base_color = torch.tensor(_RAIL_RGBA, **f32).expand(1, I, 4).contiguous()
        self.cart_base_color = (lo.unsqueeze(0) + (hi - lo).unsqueeze(0) * ramp.unsqueeze(1)).to(dev)
        self.pole_base_color = torch.tensor(_POLE_RGBA, **f32).expand(N, I, 4).contiguous()
        self.rail.set_colors(self.rail_base_color)
        self.cart.set_colors(self.cart_base_color)
        self.pole.set_colors(self.pole_base_color)

        # the pole starts lying along +x (P = 90 degrees) until the first state arrives
        self.pole_hpr = torch.zeros((N, I, 3), **f32)
        self.pole_hpr[..., 1] = 0.5 * math.pi
        self.pole.set_hprs(self.pole_hpr)
        self.cart_x_pos = torch.zeros((N, I, 1), **f32)
        self.pole_theta = torch.zeros((N, I, 1), **f32)

        cam = self.add_camera()
        cam.set_positions(torch.tensor([5.0, 5.0, 2.0]))
        cam.look_at(torch.tensor([0.0, 0.0, 0.0]))
        self.add_light()
        self.setup_environment()

    def _fit_batch(self, state: torch.Tensor) -> torch.Tensor:
        """Repeat the last row / drop rows so that there is one state per scene (reference lines 117-122)."""
        have, want = int(state.shape[0]), int(self.cfg.num_scenes)
        if have > want:
            return state[:want]
        if have < want:
            return torch.cat([state, state[-1:].expand(want - have, -1)], dim=0)
        return state

    def _step(self, state_batch: torch.Tensor | None = None):
        if state_batch is None:
            return
        state = state_batch
        if not (isinstance(state, torch.Tensor) and state.dtype == torch.float32 and not state.requires_grad):
            state = torch.as_tensor(state_batch, dtype=torch.float32).detach()
        if state.device != self.device:
            state = state.to(self.device, non_blocking=True)
        if state.shape[0] != self.num_scenes:
            state = self._fit_batch(state)

        if self._native is not None:
            # cart = T(x, 0, 0); pole = T(x, pole_y, 0) . Ry(theta): bound to the state columns, evaluated by
            # the raster kernel itself -- no launch, no matrix buffer written or read
            if self.cart._pose is None or self.pole._pose is None:
                x, theta = state[:, 0], state[:, 2]
                self.cart.set_pose(pos=(x, 0.0, 0.0))
                self.pole.set_pose(pos=(x, self.pole_y, 0.0), hpr=(0.0, theta, 0.0))
            else:                      # every later step: re-point the two state columns, nothing else
                self.cart.bind_pose_columns(state, _CART_COLUMNS)
                self.pole.bind_pose_columns(state, _POLE_COLUMNS)
            return

        # generic path (CPU): same sequence of setter calls as the reference
        x, theta = state[:, 0], state[:, 2]
        self.cart_x_pos[:, :, 0] = x.unsqueeze(1)
        self.cart_pos[:, :, 0:1] = self.cart_x_pos
        self.cart.set_positions(self.cart_pos)
        self.pole_pos[:, :, 0:1] = self.cart_x_pos
        self.pole.set_positions(self.pole_pos, lazy=True)
        self.pole_theta[:, :, 0] = theta.unsqueeze(1)
        self.pole_hpr[:, :, 1:2] = self.pole_theta
        self.pole.set_hprs(self.pole_hpr)
