"""CartPole environment configuration (reference ``pybatchrender/envs/cartpole/config.py:8-71``)."""
from dataclasses import dataclass

from ...config import PBRConfig


@dataclass
class CartPoleConfig(PBRConfig):
    # TorchRL env defaults
    direct_obs_dim: int | None = 4
    action_n: int | None = 2
    action_type: str = "discrete"
    max_steps: int = 500
    auto_reset: bool = True

    # Rendering defaults
    num_channels: int = 3
    tile_resolution: tuple[int, int] | None = (64, 64)
    offscreen: bool = True
    report_fps: bool = False

    # Physics (Gym's classic-control equations)
    gravity: float = 9.8
    masscart: float = 1.0
    masspole: float = 0.1
    length: float = 0.5          # half-pole length
    force_mag: float = 10.0
    tau: float = 0.02            # seconds between updates
    theta_threshold_deg: float = 90.0
    x_threshold: float = 2.4

    seed: int = 0
    render: bool = True

    # Reset ranges (angles in degrees here, radians inside the env)
    init_x_range: tuple[float, float] = (-2.0, 2.0)
    init_theta_range_deg: tuple[float, float] = (-30.0, 30.0)
    init_x_dot_range: tuple[float, float] = (-1.0, 1.0)
    init_theta_dot_range_deg: tuple[float, float] = (-15.0, 15.0)

    # Saving controls
    save_every_steps: int = 50
    save_examples_num: int = 16
    save_out_dir: str | None = None

    # Parallel env controls
    worker_index: int = 0
    num_workers: int = 1
