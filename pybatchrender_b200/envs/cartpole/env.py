"""CartPole TorchRL-style environment (reference ``pybatchrender/envs/cartpole/env.py:39-218``).

Physics, reset sampling, termination and auto-reset follow the reference (Euler integration of the
Gym equations).  The state never leaves ``cfg.device``: ``render_pixels(next_obs)`` hands the CUDA
tensor to the renderer, whose pose kernel reads ``x`` and ``theta`` in place -- the reference's
per-step ``state.cpu()`` synchronisation (envs/cartpole/renderer.py:112) does not exist here, and a
rendering failure raises instead of silently dropping ``pixels`` (env.py:162-167, 192-196 there).
"""
import math

import torch

from ..._rl_compat import TensorDict
from ...env import PBREnv
from .config import CartPoleConfig
from .renderer import CartPoleRenderer


class CartPoleEnv(PBREnv):
    def __init__(self, renderer: CartPoleRenderer, cfg: CartPoleConfig | None = None, **cfg_overrides):
        if cfg is None:
            cfg = renderer.cfg
        super().__init__(renderer=renderer, cfg=cfg, device=torch.device(cfg.device),
                         batch_size=torch.Size([cfg.num_scenes]))
        g = lambda name, default: getattr(cfg, name, default)   # noqa: E731
        self.gravity = float(g("gravity", 9.8))
        self.masscart = float(g("masscart", 1.0))
        self.masspole = float(g("masspole", 0.1))
        self.total_mass = self.masscart + self.masspole
        self.length = float(g("length", 0.5))
        self.polemass_length = self.masspole * self.length
        self.force_mag = float(g("force_mag", 10.0))
        self.tau = float(g("tau", 0.02))
        self.theta_threshold = float(g("theta_threshold_deg", 12.0)) * 2 * math.pi / 360.0
        self.x_threshold = float(g("x_threshold", 2.4))
        self.max_steps = int(g("max_steps", 500))
        self.auto_reset = bool(g("auto_reset", True))
        self.seed = int(g("seed", 0))
        self.render = bool(g("render", True))
        th = g("init_theta_range_deg", (-30.0, 30.0))
        thd = g("init_theta_dot_range_deg", (-15.0, 15.0))
        self._init_x_range = tuple(g("init_x_range", (-2.0, 2.0)))
        self._init_x_dot_range = tuple(g("init_x_dot_range", (-1.0, 1.0)))
        self._init_theta_range_rad = (math.radians(min(th)), math.radians(max(th)))
        self._init_theta_dot_range_rad = (math.radians(min(thd)), math.radians(max(thd)))
        self.set_default_specs(direct_obs_dim=4, actions=2, with_pixels=self.render, pixels_only=False,
                               discrete_actions=True)
        if self.seed is not None:
            self.set_seed(self.seed)

    def _sample_initial_state(self, batch_shape) -> torch.Tensor:
        def u(lo, hi):
            return torch.empty(*batch_shape, 1, dtype=torch.float32, device=self.device).uniform_(lo, hi)
        return torch.cat([u(*self._init_x_range), u(*self._init_x_dot_range), u(*self._init_theta_range_rad),
                          u(*self._init_theta_dot_range_rad)], dim=-1)

    def _dynamics(self, obs: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        x, x_dot, theta, theta_dot = obs.unbind(-1)
        force = torch.where(action == 1, self.force_mag, -self.force_mag).to(torch.float32)
        costheta, sintheta = torch.cos(theta), torch.sin(theta)
        temp = (force + self.polemass_length * theta_dot.pow(2) * sintheta) / self.total_mass
        thetaacc = (self.gravity * sintheta - costheta * temp) / (
            self.length * (4.0 / 3.0 - self.masspole * costheta.pow(2) / self.total_mass))
        xacc = temp - self.polemass_length * thetaacc * costheta / self.total_mass
        return torch.stack([x + self.tau * x_dot, x_dot + self.tau * xacc, theta + self.tau * theta_dot,
                            theta_dot + self.tau * thetaacc], dim=-1)

    def _termination(self, obs: torch.Tensor, step_count: torch.Tensor) -> torch.Tensor:
        x, _, theta, _ = obs.unbind(-1)
        return ((x.abs() > self.x_threshold) | (theta.abs() > self.theta_threshold)
                | (step_count >= (self.max_steps - 1))).unsqueeze(-1)

    def _reward(self, obs: torch.Tensor, done: torch.Tensor) -> torch.Tensor:
        return torch.ones_like(done, dtype=torch.float32, device=self.device)

    def _set_seed(self, seed: int) -> None:
        torch.manual_seed(int(seed))

    def _reset(self, tensordict: TensorDict | None = None) -> TensorDict:
        bs = self.batch_size if self.batch_size != torch.Size([]) else torch.Size([1])
        state = self._sample_initial_state(bs)
        fields = {
            "observation": state,
            "step_count": torch.zeros(*bs, dtype=torch.long, device=self.device),
            "done": torch.zeros(*bs, 1, dtype=torch.bool, device=self.device),
        }
        if self.render:
            fields["pixels"] = self.render_pixels(state)
        return TensorDict(fields, batch_size=self.batch_size)

    @torch.no_grad()
    def _step(self, tensordict: TensorDict) -> TensorDict:
        obs = tensordict.get("observation", None)
        if obs is None:
            td0 = self._reset()
            obs, step_count = td0["observation"], td0["step_count"]
        else:
            step_count = tensordict.get("step_count", None)
            if step_count is None:
                step_count = torch.zeros_like(obs[..., 0], dtype=torch.long, device=self.device)
        action = tensordict["action"].to(self.device)
        next_obs_raw = self._dynamics(obs, action)
        done = self._termination(next_obs_raw, step_count)
        reward = self._reward(next_obs_raw, done)
        pixels = self.render_pixels(next_obs_raw) if self.render else None
        if self.auto_reset:
            reset_state = self._sample_initial_state(next_obs_raw.shape[:-1])
            next_obs = torch.where(done, reset_state, next_obs_raw)
            next_step_count = torch.where(done.squeeze(-1), torch.zeros_like(step_count), step_count + 1)
        else:
            next_obs, next_step_count = next_obs_raw, step_count + 1
        out = {"observation": next_obs, "reward": reward, "done": done, "step_count": next_step_count}
        if pixels is not None:
            out["pixels"] = pixels
        return TensorDict(out, batch_size=self.batch_size)
