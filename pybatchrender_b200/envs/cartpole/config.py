"""``CartPoleConfig``: the CartPole environment's knobs on top of :class:`PBRConfig`.

Field names and defaults follow the reference (``pybatchrender/envs/cartpole/config.py:8-71``) so
that ``pbr.envs.make("CartPole-v0", **overrides)`` accepts the same overrides; angles are given in
degrees here and converted to radians inside the environment.
"""
from dataclasses import dataclass

from ...config import PBRConfig

Range = tuple[float, float]


@dataclass
class CartPoleConfig(PBRConfig):
    # --- rendering defaults of this env
    tile_resolution: tuple[int, int] | None = (64, 64)
    num_channels: int = 3
    offscreen: bool = True
    report_fps: bool = False
    render: bool = True

    # --- episode handling / spaces
    direct_obs_dim: int | None = 4
    action_n: int | None = 2
    action_type: str = "discrete"
    max_steps: int = 500
    auto_reset: bool = True
    seed: int = 0

    # --- physics (classic-control cart-pole; `length` is half the pole, `tau` the Euler step in s)
    gravity: float = 9.8
    masscart: float = 1.0
    masspole: float = 0.1
    length: float = 0.5
    force_mag: float = 10.0
    tau: float = 0.02
    x_threshold: float = 2.4
    theta_threshold_deg: float = 90.0

    # --- reset distribution (uniform)
    init_x_range: Range = (-2.0, 2.0)
    init_x_dot_range: Range = (-1.0, 1.0)
    init_theta_range_deg: Range = (-30.0, 30.0)
    init_theta_dot_range_deg: Range = (-15.0, 15.0)

    # --- periodic example dumps (< 0 never, 0 every step, n every n steps)
    save_every_steps: int = 50
    save_examples_num: int = 16
    save_out_dir: str | None = None

    # --- TorchRL ParallelEnv bookkeeping
    worker_index: int = 0
    num_workers: int = 1
