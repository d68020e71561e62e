"""The steering scene: a player box, two lane rails that travel with it, a field of obstacle
spheres per scene and a camera that follows the player every step.

Scene semantics follow the reference (``pybatchrender/envs/steering/renderer.py:33-271``, SURVEY.md
8 row f3): the player is ``player_model`` scaled to ``player_dimensions``; both rails are
``border_model`` scaled to ``rail_dimensions`` at x = -/+ lane_width/2, z = rail_offset.z, following the
player's y; every obstacle (cone flag or not -- the reference only ever instantiates
``obstacle_sphere_model``) is a sphere at (x, y, 0), gold or red by its flag; the player turns
``crash_player_color`` while the grace flag ``state[:, 3]`` is set; the camera eye is the player
position plus ``camera_eye_offset`` with the default forward direction.  ``setup_environment`` runs
on the first ``build_obstacles`` call.

What is different: nothing is staged through the host.  The reference copies the state to the CPU
and re-serialises four buffers per step (renderer.py:226-271); here the player / rail matrices are
written by one pose-kernel launch from strided views of the device state, the colour switch is a
``torch.where`` into the colour buffer, and the obstacle node is reused across resets when its
instance count is unchanged (the reference destroys and reloads the model, renderer.py:176-185).
"""
from __future__ import annotations

import os

import torch

from ...meshes import MODELS_DIR
from ...renderer.renderer import PBRRenderer


class SteeringRenderer(PBRRenderer):
    def __init__(self, cfg=None, **cfg_overrides) -> None:
        super().__init__(cfg, **cfg_overrides)
        c, dev, n = self.cfg, self.device, int(self.cfg.num_scenes)
        self._num_scenes = n
        self._setup_called = False
        f32 = dict(dtype=torch.float32, device=dev)

        self._color_player = torch.tensor(c.player_color, **f32)
        self._color_crash = torch.tensor(getattr(c, "crash_player_color", c.player_color), **f32)
        self._color_red = torch.tensor(c.red_obstacle_color, **f32)
        self._color_gold = torch.tensor(c.gold_obstacle_color, **f32)
        self._color_border = torch.tensor(c.edge_color, **f32)
        self.set_background_color(*(float(v) for v in c.background_color))

        self.sphere_node = None
        self.cone_node = None          # never populated (the reference does not instantiate cones either)

        # ---- player
        self._player_pos = torch.zeros((n, 1, 3), **f32)
        self._player_color_buf = self._color_player.expand(n, 1, 4).clone()
        self.player_node = self.add_node(
            self._resolve_model_path(c.player_model), instances_per_scene=1, model_scale=c.player_dimensions,
            model_hpr=c.player_model_hpr, model_scale_units="absolute",
            model_pivot_relative_point=c.player_pivot_relative_point, shared_across_scenes=False)
        self.player_node.set_positions(self._player_pos)
        self.player_node.set_colors(self._player_color_buf)

        # ---- rails
        half, z_off = 0.5 * float(c.lane_width), float(c.rail_offset[2])
        self._rail_x = (-half, half)
        self._rail_z = z_off
        rails = []
        for x in self._rail_x:
            node = self.add_node(self._resolve_model_path(c.border_model), instances_per_scene=1,
                                 model_scale=c.rail_dimensions, model_hpr=c.border_model_hpr,
                                 model_scale_units="absolute", shared_across_scenes=False)
            pos = torch.zeros((n, 1, 3), **f32)
            pos[:, 0, 0], pos[:, 0, 2] = x, z_off
            node.set_positions(pos)
            node.set_colors(self._color_border.expand(n, 1, 4).clone())
            rails.append((node, pos))
        (self.left_border, self._left_border_pos), (self.right_border, self._right_border_pos) = rails

        self._camera_eye_offset = torch.tensor(c.camera_eye_offset, **f32).reshape(1, 3)
        self.add_camera(z_far=float(c.z_far), z_near=float(c.z_near), fov_y_deg=float(c.fov_y_deg))
        self.add_light(ambient=c.ambient_light, dir_dir=c.directional_light_dir)
        # setup_environment() is deferred to build_obstacles(), like the reference

    @staticmethod
    def _resolve_model_path(model: str) -> str:
        """'models/<x>' -> the packaged file when it exists, else the built-in of that name."""
        if model.startswith("models/"):
            packaged = os.path.join(MODELS_DIR, model[len("models/"):])
            return packaged if os.path.exists(packaged) else model
        return model if os.path.isabs(model) else os.path.abspath(os.path.join(os.path.dirname(MODELS_DIR), model))

    def _finish_setup(self) -> None:
        if not self._setup_called:
            self.setup_environment()
            self._setup_called = True

    def build_obstacles(self, obstacles: torch.Tensor | None) -> None:
        """``obstacles[B, N, 4]`` = (x, y, gold flag, cone flag) -> the obstacle node's instances."""
        if obstacles is None or obstacles.numel() == 0:
            if self.sphere_node is not None:
                self.sphere_node.np.removeNode()
                self.sphere_node = None
            self._finish_setup()
            return
        obs = obstacles.detach().to(self.device, torch.float32)
        n_scenes, n_obs = int(obs.shape[0]), int(obs.shape[1])
        if self.sphere_node is not None and self.sphere_node.instances_per_scene != n_obs:
            self.sphere_node.np.removeNode()
            self.sphere_node = None
        if self.sphere_node is None:
            c = self.cfg
            self.sphere_node = self.add_node(
                self._resolve_model_path(c.obstacle_sphere_model), instances_per_scene=n_obs,
                model_scale=c.obstacle_dimensions, model_hpr=c.obstacle_model_hpr, model_scale_units="absolute",
                model_pivot_relative_point=c.obstacle_pivot_relative_point, shared_across_scenes=False)
        where = torch.zeros((n_scenes, n_obs, 3), dtype=torch.float32, device=self.device)
        where[..., 0:2] = obs[..., 0:2]
        gold = (obs[..., 2] >= 0.5).unsqueeze(-1)
        self.sphere_node.set_positions(where)
        self.sphere_node.set_colors(torch.where(gold, self._color_gold, self._color_red))
        self._finish_setup()

    def _fit_batch(self, state: torch.Tensor) -> torch.Tensor:
        have, want = int(state.shape[0]), self._num_scenes
        if have > want:
            return state[:want]
        if have < want:
            return torch.cat([state, state[-1:].expand(want - have, -1)], dim=0)
        return state

    def _step(self, state_batch: torch.Tensor | None = None) -> None:
        """state[:, 0:2] = player x, y; state[:, 3] = grace flag (optional)."""
        if state_batch is None:
            return
        state = torch.as_tensor(state_batch, dtype=torch.float32).detach()
        if state.device != self.device:
            state = state.to(self.device, non_blocking=True)
        state = self._fit_batch(state)
        x, y = state[:, 0], state[:, 1]

        self._player_pos[:, 0, 0], self._player_pos[:, 0, 1] = x, y
        self._left_border_pos[:, 0, 1] = y
        self._right_border_pos[:, 0, 1] = y
        if self._native is not None:
            self.player_node.set_pose(pos=(x, y, 0.0))
            self.left_border.set_pose(pos=(self._rail_x[0], y, self._rail_z))
            self.right_border.set_pose(pos=(self._rail_x[1], y, self._rail_z))
        else:
            self.player_node.set_positions(self._player_pos)
            self.left_border.set_positions(self._left_border_pos)
            self.right_border.set_positions(self._right_border_pos)

        if state.shape[1] > 3:
            in_grace = (state[:, 3] > 0.5).reshape(-1, 1, 1)
            self._player_color_buf = torch.where(in_grace, self._color_crash, self._color_player)
        else:
            self._player_color_buf = self._color_player.expand(self._num_scenes, 1, 4)
        self.player_node.set_colors(self._player_color_buf)

        if self._pbr_cam is not None:
            self._pbr_cam.set_positions(self._player_pos[:, 0, :] + self._camera_eye_offset)
