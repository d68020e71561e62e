"""CartPole as a batched TorchRL-style environment whose pixels come from the B200 rasteriser.

Behaviour follows the reference (``pybatchrender/envs/cartpole/env.py:39-218``): explicit-Euler
integration of the classic cart-pole equations, uniform reset ranges, termination on |x|, |theta|
or the step budget, +1 reward per step, optional auto-reset.  Two deliberate differences: the state
never leaves ``cfg.device`` (the renderer's pose kernel reads ``x`` / ``theta`` in place instead of
the reference's per-step ``state.cpu()``, envs/cartpole/renderer.py:112), and a rendering failure
raises instead of silently dropping ``pixels`` (reference env.py:162-167, 192-196).
"""
import math

import torch

from ..._rl_compat import TensorDict
from ...env import PBREnv
from .config import CartPoleConfig
from .renderer import CartPoleRenderer

_DEFAULTS = CartPoleConfig.__dataclass_fields__


class CartPoleEnv(PBREnv):
    def __init__(self, renderer: CartPoleRenderer, cfg: CartPoleConfig | None = None, **cfg_overrides):
        cfg = renderer.cfg if cfg is None else cfg
        super().__init__(renderer=renderer, cfg=cfg, device=torch.device(cfg.device),
                         batch_size=torch.Size([cfg.num_scenes]))

        def opt(name):
            return getattr(cfg, name, _DEFAULTS[name].default)

        for name in ("gravity", "masscart", "masspole", "length", "force_mag", "tau", "x_threshold"):
            setattr(self, name, float(opt(name)))
        self.total_mass = self.masspole + self.masscart
        self.polemass_length = self.masspole * self.length
        self.theta_threshold = math.radians(float(opt("theta_threshold_deg")))
        self.max_steps, self.seed = int(opt("max_steps")), int(opt("seed"))
        self.auto_reset, self.render = bool(opt("auto_reset")), bool(opt("render"))
        # (low, high) per state component: x, x_dot, theta, theta_dot
        deg = [sorted(opt("init_theta_range_deg")), sorted(opt("init_theta_dot_range_deg"))]
        self._reset_ranges = (tuple(opt("init_x_range")), tuple(opt("init_x_dot_range")),
                              tuple(math.radians(a) for a in deg[0]), tuple(math.radians(a) for a in deg[1]))
        self._init_theta_range_rad, self._init_theta_dot_range_rad = self._reset_ranges[2], self._reset_ranges[3]
        self.set_default_specs(direct_obs_dim=4, actions=2, with_pixels=self.render, pixels_only=False,
                               discrete_actions=True)
        self.set_seed(self.seed)

    # ------------------------------------------------------------------ pieces of a transition
    def _sample_initial_state(self, batch_shape) -> torch.Tensor:
        cols = [torch.empty(*batch_shape, 1, dtype=torch.float32, device=self.device).uniform_(lo, hi)
                for lo, hi in self._reset_ranges]
        return torch.cat(cols, dim=-1)

    def _dynamics(self, obs: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        x, v, th, w = obs.unbind(-1)
        push = torch.where(action == 1, self.force_mag, -self.force_mag).to(torch.float32)
        c, s = torch.cos(th), torch.sin(th)
        tmp = (push + self.polemass_length * w.pow(2) * s) / self.total_mass
        th_acc = (self.gravity * s - c * tmp) / (self.length * (4.0 / 3.0 - self.masspole * c.pow(2) / self.total_mass))
        x_acc = tmp - self.polemass_length * th_acc * c / self.total_mass
        dt = self.tau
        return torch.stack([x + dt * v, v + dt * x_acc, th + dt * w, w + dt * th_acc], dim=-1)

    def _termination(self, obs: torch.Tensor, step_count: torch.Tensor) -> torch.Tensor:
        out_of_track = obs[..., 0].abs() > self.x_threshold
        fallen = obs[..., 2].abs() > self.theta_threshold
        timed_out = step_count >= self.max_steps - 1
        return (out_of_track | fallen | timed_out).unsqueeze(-1)

    def _reward(self, obs: torch.Tensor, done: torch.Tensor) -> torch.Tensor:
        return torch.ones_like(done, dtype=torch.float32, device=self.device)

    def _set_seed(self, seed: int) -> None:
        torch.manual_seed(int(seed))

    # ------------------------------------------------------------------ TorchRL hooks
    def _reset(self, tensordict: TensorDict | None = None) -> TensorDict:
        bs = self.batch_size if len(self.batch_size) else torch.Size([1])
        state = self._sample_initial_state(bs)
        td = {"observation": state,
              "step_count": torch.zeros(*bs, dtype=torch.long, device=self.device),
              "done": torch.zeros(*bs, 1, dtype=torch.bool, device=self.device)}
        if self.render:
            td["pixels"] = self.render_pixels(state)
        return TensorDict(td, batch_size=self.batch_size)

    @torch.no_grad()
    def _step(self, tensordict: TensorDict) -> TensorDict:
        obs = tensordict.get("observation", None)
        steps = tensordict.get("step_count", None) if obs is not None else None
        if obs is None:
            fresh = self._reset()
            obs, steps = fresh["observation"], fresh["step_count"]
        elif steps is None:
            steps = torch.zeros_like(obs[..., 0], dtype=torch.long, device=self.device)

        moved = self._dynamics(obs, tensordict["action"].to(self.device))
        done = self._termination(moved, steps)
        out = {"reward": self._reward(moved, done), "done": done}
        if self.render:
            out["pixels"] = self.render_pixels(moved)       # frames show the pre-reset state, like the reference
        if self.auto_reset:
            out["observation"] = torch.where(done, self._sample_initial_state(moved.shape[:-1]), moved)
            out["step_count"] = torch.where(done.squeeze(-1), torch.zeros_like(steps), steps + 1)
        else:
            out["observation"], out["step_count"] = moved, steps + 1
        return TensorDict(out, batch_size=self.batch_size)
