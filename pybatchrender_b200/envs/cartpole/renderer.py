"""CartPole scene: rail (shared), cart and pole (per scene), camera at (5,5,2) looking at the origin.

Reference: ``pybatchrender/envs/cartpole/renderer.py:18-138``.  Scene constants
(sizes, colours, pole offset, camera, default light) are the reference's.  ``_step(state[B,4])`` maps
``x = state[:,0]`` to the cart and pole x position and ``theta = state[:,2]`` to the pole's P angle
(rotation about +Y in the reference's HPR convention, shader_context.py:47-84).

Difference: the state stays on the device.  The reference does ``state.detach().cpu()`` and three
full buffer re-uploads per step (renderer.py:112,130-138); here one fused pose kernel
(``pbr_compose_transforms``) reads ``state[:,0]`` / ``state[:,2]`` in place through strided channel
views and writes both nodes' matrix buffers.  ``cfg.device == 'cpu'`` (no GPU) falls back to the
generic torch setters so that the host logic stays testable; it cannot render.
"""
from __future__ import annotations

import math

import torch

from ...config import PBRConfig
from ...renderer.renderer import PBRRenderer


class CartPoleRenderer(PBRRenderer):
    def __init__(self, cfg: PBRConfig | dict | None = None, **cfg_overrides):
        super().__init__(cfg, **cfg_overrides)
        instances_per_scene = 1
        dev = self.device

        self.rail_size = (6.0, 0.05, 0.05)
        self.cart_size = (1.2, 0.8, 0.5)
        self.pole_size = (0.1, 0.1, 2.0)

        self.rail_pos_color = (0.2, 0.2, 0.2, 1.0)
        self.cart_pos_color_range = ((0.6, 0.8, 1.0, 1.0), (1.0, 0.6, 0.8, 1.0))
        self.pole_pos_color = (1.0, 0.7, 0.2, 1.0)

        self.rail = self.add_node("models/box", model_pivot_relative_point=(0.5, 0.5, 0.5),
                                  model_scale=self.rail_size, instances_per_scene=instances_per_scene,
                                  shared_across_scenes=True)
        self.cart = self.add_node("models/box", model_pivot_relative_point=(0.5, 0.5, 0.5),
                                  model_scale=self.cart_size, instances_per_scene=instances_per_scene,
                                  shared_across_scenes=False)
        self.pole = self.add_node("models/box", model_pivot_relative_point=(0.5, 0.5, 0.05),
                                  model_scale=self.pole_size, instances_per_scene=instances_per_scene,
                                  shared_across_scenes=False)

        N = int(self.cfg.num_scenes)
        # scenes of a shard keep their global colour ramp position (see pybatchrender_b200.dist)
        g0 = int(getattr(self.cfg, "scene_offset", 0) or 0)
        gN = int(getattr(self.cfg, "global_num_scenes", 0) or 0) or N

        self.pole_y = (self.cart_size[1] + self.pole_size[1]) * 0.5
        self.rail_pos = torch.zeros((1, instances_per_scene, 3), dtype=torch.float32, device=dev)
        self.cart_pos = torch.zeros((N, instances_per_scene, 3), dtype=torch.float32, device=dev)
        self.pole_pos = self.cart_pos.clone()
        self.pole_pos[:, :, 1] = self.pole_y
        self.rail.set_positions(self.rail_pos)
        self.cart.set_positions(self.cart_pos)
        self.pole.set_positions(self.pole_pos)

        self.rail_base_color = torch.tensor(self.rail_pos_color, dtype=torch.float32,
                                            device=dev).repeat(1, instances_per_scene, 1)
        start = torch.tensor(self.cart_pos_color_range[0], dtype=torch.float32)
        end = torch.tensor(self.cart_pos_color_range[1], dtype=torch.float32)
        t = torch.linspace(0.0, 1.0, steps=gN, dtype=torch.float32)[g0:g0 + N]
        self.cart_base_color = (start.unsqueeze(0) + (end - start).unsqueeze(0) * t.unsqueeze(1)).to(dev)
        self.pole_base_color = torch.tensor(self.pole_pos_color, dtype=torch.float32,
                                            device=dev).repeat(N, instances_per_scene, 1)
        self.rail.set_colors(self.rail_base_color)
        self.cart.set_colors(self.cart_base_color)
        self.pole.set_colors(self.pole_base_color)

        self.pole_hpr = torch.zeros((N, instances_per_scene, 3), dtype=torch.float32, device=dev)
        self.pole_hpr[:, :, 1] = math.pi * 0.5
        self.pole.set_hprs(self.pole_hpr)

        self.cart_x_pos = self.cart_pos[:, :, 0:1].clone()
        self.pole_theta = torch.zeros_like(self.cart_x_pos)

        self.add_camera()
        self._pbr_cam.set_positions(torch.tensor([5, 5, 2], dtype=torch.float32))
        self._pbr_cam.look_at(torch.tensor([0, 0, 0], dtype=torch.float32))
        self.add_light()

        self.setup_environment()

    def _fit_batch(self, state: torch.Tensor) -> torch.Tensor:
        """Pad with the last row / truncate to ``num_scenes`` (reference renderer.py:117-122)."""
        B, N = int(state.shape[0]), int(self.cfg.num_scenes)
        if B < N:
            state = torch.cat([state, state[-1:].repeat(N - B, 1)], dim=0)
        elif B > N:
            state = state[:N]
        return state

    def _step(self, state_batch: torch.Tensor | None = None):
        if state_batch is None:
            return
        state = torch.as_tensor(state_batch, dtype=torch.float32).detach()
        if state.device != self.device:
            state = state.to(self.device, non_blocking=True)
        state = self._fit_batch(state)

        if self._native is not None:
            # fused pose kernel: cart = T(x,0,0); pole = T(x, pole_y, 0) * Ry(theta)
            x, theta = state[:, 0], state[:, 2]
            self._native.compose([
                dict(pos=(x, 0.0, 0.0), hpr=(0.0, 0.0, 0.0), scale=1.0, out=self.cart.matbuf),
                dict(pos=(x, self.pole_y, 0.0), hpr=(0.0, theta, 0.0), scale=1.0, out=self.pole.matbuf),
            ], self.device)
            self._last_state = state       # keep the views alive until the stream work is enqueued
            return

        x = state[:, 0:1]
        theta = state[:, 2:3]
        self.cart_x_pos[:, :, 0] = x
        self.cart_pos[:, :, 0:1] = self.cart_x_pos
        self.cart.set_positions(self.cart_pos)
        self.pole_pos[:, :, 0:1] = self.cart_x_pos
        self.pole.set_positions(self.pole_pos, lazy=True)
        self.pole_theta[:, :, 0] = theta
        self.pole_hpr[:, :, 1:2] = self.pole_theta
        self.pole.set_hprs(self.pole_hpr)
