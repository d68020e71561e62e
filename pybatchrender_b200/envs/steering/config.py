"""``SteeringConfig``: knobs of the lane-steering task on top of :class:`PBRConfig`.

Names and defaults are the reference's (``pybatchrender/envs/steering/config.py:11-111``) so that
``pbr.envs.make("Steering-v0", **overrides)`` takes the same overrides.  Units: lengths in world
units, speeds in units / s, ``tau`` and ``grace_period`` in seconds, angles in degrees.
"""
from dataclasses import dataclass, field

from ...config import PBRConfig

Vec3 = tuple[float, float, float]
RGBA = tuple[float, float, float, float]
_CENTRE: Vec3 = (0.5, 0.5, 0.5)
_NO_TURN: Vec3 = (0.0, 0.0, 0.0)


def _lane_slots() -> list[int]:
    # odd x positions across the 25-wide lane: -11, -9, ..., 11
    return [x for x in range(-11, 12, 2)]


@dataclass
class SteeringConfig(PBRConfig):
    # ---- what the agent sees / does.  observation = (x, y, speed, grace flag, next obstacle x, y)
    direct_obs_dim: int = 6
    action_n: int = 1
    action_type: str = "continuous"
    action_low: float = -327.68          # steering-wheel range / 1000
    action_high: float = 327.68
    max_steps: int = 10000
    auto_reset: bool = True
    seed: int = 0
    render: bool = True

    # ---- frame
    num_channels: int = 3
    tile_resolution: tuple[int, int] = (64, 64)
    offscreen: bool = True
    report_fps: bool = False
    background_color: RGBA = (0.0, 0.0, 0.0, 1.0)

    # ---- follow camera: eye = player + offset, looking along camera_forward_vector
    camera_eye_offset: Vec3 = (0.0, -16.3, 4.0)
    camera_forward_vector: Vec3 = (0.0, 1.0, 0.0)
    fov_y_deg: float = 40.0
    z_near: float = 0.5
    z_far: float = 1000.0

    # ---- light
    ambient_light: Vec3 = (0.2, 0.2, 0.2)
    directional_light_dir: Vec3 = (0.22, 0.44, 0.88)
    directional_light_color: Vec3 = (0.8, 0.8, 0.8)

    # ---- the cast: model, size, pivot, pre-rotation, colour
    player_model: str = "models/box"
    player_dimensions: Vec3 = (2.0, 2.0, 2.0)
    player_pivot_relative_point: Vec3 = _CENTRE
    player_model_hpr: Vec3 = _NO_TURN
    player_color: RGBA = (0.0, 0.0, 1.0, 1.0)
    crash_player_color: RGBA = (0.4, 0.4, 0.4, 1.0)     # shown while the grace period runs

    obstacle_sphere_model: str = "models/smiley"
    obstacle_cone_model: str = "models/cone.egg"
    obstacle_dimensions: Vec3 = (2.0, 2.0, 2.0)
    obstacle_pivot_relative_point: Vec3 = _CENTRE
    obstacle_model_hpr: Vec3 = _NO_TURN
    red_obstacle_color: RGBA = (1.0, 0.0, 0.0, 1.0)
    gold_obstacle_color: RGBA = (1.0, 0.5, 0.0, 1.0)

    border_model: str = "models/cylinder/scene.gltf"
    rail_dimensions: Vec3 = (0.2, 1012.0, 0.2)
    rail_offset: Vec3 = (0.0, -12.0, 0.5)
    rail_pivot_relative_point: Vec3 = (0.5, 0.5, 0.0)
    border_model_hpr: Vec3 = _NO_TURN
    edge_color: RGBA = (0.2, 0.2, 0.2, 1.0)

    # ---- track
    lane_width: float = 25.0
    track_length: float = 1000.0
    number_of_obstacles: int = 100
    distance_to_first_obstacle: float = 6
    obstacle_y_spacing: float = 12.0
    obstacle_x_position: list[int] = field(default_factory=_lane_slots)
    obstacle_seed: int | None = None
    prob_gold: float = 0.2
    prob_cone: float = 0.5

    # ---- motion
    tau: float = 1.0 / 60.0
    speed_initial: float = 96.0
    minimal_speed: float = 18.0
    speed_increment: float = 0.072        # per obstacle passed
    speed_decrement: float = 6.12         # per crash
    grace_period: float = 100.0 / 60.0
    steering_speed: float = 20.0
    wheel_sensitivity: float = 800.0      # larger = less lateral motion per unit of wheel input

    # ---- reward / episode end
    collision_penalty: float = -1.0
    terminate_on_collision: bool = False
    terminate_on_track_end: bool = True

    # ---- debugging output, worker bookkeeping
    save_every_steps: int = -1
    save_examples_num: int = 16
    save_out_dir: str | None = None
    worker_index: int = 0
    num_workers: int = 1
