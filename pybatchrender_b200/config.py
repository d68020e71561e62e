"""``PBRConfig`` -- field-for-field mirror of the reference's config dataclass.

Reference: ``pybatchrender/config.py:16-199``.  Semantics kept:

* ``tiles`` given as an int (or derived from ``num_scenes``) becomes ``(cols, rows)`` with
  ``cols = ceil(sqrt(n))``, ``rows = ceil(n / cols)``                      (config.py:63-69)
* whichever of ``tiles`` / ``tile_resolution`` / ``window_resolution`` are missing are filled
  in, default tile 64x64                                                   (config.py:71-112)
* ``window == tiles * tile`` is checked when all three are given            (config.py:73-82)
* ``num_scenes <= cols * rows``                                            (config.py:114-120)
* ``device`` is one of cpu / cuda / mps, auto = cuda when available        (config.py:128-145)

The tile grid no longer describes a real window (every scene is rendered straight into its own
slice of the output tensor) but it still fixes the projection aspect (quirk Q1: aspect is the
*window* aspect ``cols*W / rows*H``, reference ``camera.py:154``), so it is computed identically.
Fields that only made sense for Panda3D (``panda3d_backend``, ``extra_prc_file_data``,
``cuda_gl_interop`` ...) are kept so existing configs still construct; ``build_prc`` remains a
harmless string builder.
"""
from __future__ import annotations

import logging
import math
from dataclasses import asdict, dataclass
from typing import TypeVar

import torch

T = TypeVar("T", bound="PBRConfig")

_DEFAULT_TILE = (64, 64)


def grid_for(n: int) -> tuple[int, int]:
    """(cols, rows) of the near-square grid that holds ``n`` tiles (reference config.py:66-69)."""
    cols = math.ceil(math.sqrt(n))
    rows = math.ceil(n / cols)
    return cols, rows


_PANDA_PRC_EXTRAS = "audio-library-name null\ntextures-power-2 none\nsync-video 0\n"


@dataclass
class PBRConfig:
    # ---- batch geometry: how many scenes, how big each tile is.  `tiles` / `window_resolution`
    # describe the reference's single tiled window; they survive because the *window* aspect enters
    # the projection (quirk Q1).
    num_scenes: int | None = None
    tile_resolution: tuple[int, int] | None = None
    tiles: tuple[int, int] | int | None = None
    window_resolution: tuple[int, int] | None = None
    batch_inner_dim: int | None = None
    num_channels: int = 3
    device: str | None = None

    # ---- scene sharding (not in the reference; see pybatchrender_b200/dist.py): this process renders
    # scenes [scene_offset, scene_offset + num_scenes) of a batch of global_num_scenes, and `tiles`
    # is then the *global* grid so that the projection aspect matches the unsharded render
    scene_offset: int = 0
    global_num_scenes: int | None = None

    # ---- env layer (TorchRL)
    direct_obs_dim: int | None = None
    action_n: int | None = None
    action_type: str = "discrete"
    max_steps: int = 500
    auto_reset: bool = True
    num_workers: int = 1

    # ---- accepted for compatibility with existing configs; no effect on the CUDA rasteriser
    offscreen: bool = True
    interactive: bool = False
    manual_camera_control: bool = False
    cuda_gl_interop: bool = True
    panda3d_backend: str | None = "arm"
    extra_prc_file_data: str = _PANDA_PRC_EXTRAS
    render_mode: str = "rgb_array"
    warmup_steps: int = 5
    dt: float = 1 / 60
    report_fps: bool = True
    report_fps_interval: float = 1.0
    log_level: int = logging.DEBUG
    clip_camera: tuple[float, float] = (3.0, 500.0)
    min_objects: int = 50
    max_objects: int = 100

    # ------------------------------------------------------------------ resolution
    def process_resolution(self) -> None:
        if self.tiles is None and self.num_scenes is not None:
            self.tiles = self.num_scenes
        if isinstance(self.tiles, int):
            self.tiles = grid_for(self.tiles)

        have = [self.tiles is not None, self.tile_resolution is not None, self.window_resolution is not None]
        missing = 3 - sum(have)

        if missing == 0:
            tw = (self.tiles[0] * self.tile_resolution[0], self.tiles[1] * self.tile_resolution[1])
            if tuple(self.window_resolution) != tw:
                raise ValueError(
                    "window_resolution must equal tiles * tile_resolution. "
                    f"Got window_resolution={self.window_resolution}, tiles={self.tiles}, "
                    f"tile_resolution={self.tile_resolution}."
                )
        elif missing == 1:
            if self.tiles is None:
                self.tiles = (self.window_resolution[0] // self.tile_resolution[0],
                              self.window_resolution[1] // self.tile_resolution[1])
            elif self.tile_resolution is None:
                self.tile_resolution = (self.window_resolution[0] // self.tiles[0],
                                        self.window_resolution[1] // self.tiles[1])
            else:
                self.window_resolution = (self.tiles[0] * self.tile_resolution[0],
                                          self.tiles[1] * self.tile_resolution[1])
        else:
            # Two or three unknowns: the single known quantity (if any) wins, the rest default
            # to one 64x64 tile.
            if self.tiles is not None:
                self.tile_resolution = _DEFAULT_TILE
                self.window_resolution = (_DEFAULT_TILE[0] * self.tiles[0], _DEFAULT_TILE[1] * self.tiles[1])
            elif self.tile_resolution is not None:
                self.tiles = (1, 1)
                self.window_resolution = (self.tile_resolution[0], self.tile_resolution[1])
            elif self.window_resolution is not None:
                self.tiles = (1, 1)
                self.tile_resolution = (self.window_resolution[0], self.window_resolution[1])
            else:
                self.tiles = (1, 1)
                self.tile_resolution = _DEFAULT_TILE
                self.window_resolution = _DEFAULT_TILE

        capacity = self.tiles[0] * self.tiles[1]
        if self.num_scenes is None:
            self.num_scenes = capacity
        elif self.num_scenes > capacity:
            raise ValueError(f"{self.num_scenes=} can't fit into {self.tiles=}")

        if self.batch_inner_dim is None:
            self.batch_inner_dim = capacity
        elif self.batch_inner_dim != capacity:
            raise ValueError("batch_inner_dim must equal tiles[0] * tiles[1]")

    # ------------------------------------------------------------------ device
    def process_device(self) -> None:
        if self.device is not None and self.device not in ("cpu", "cuda", "mps"):
            raise ValueError(f"Invalid device: {self.device}")
        if self.device is None:
            # the reference maps an available-but-unselected MPS to 'cpu' too (config.py:138-141)
            self.device = "cuda" if torch.cuda.is_available() else "cpu"
        if self.device == "cuda" and not torch.cuda.is_available():
            raise RuntimeError("device is set to CUDA but CUDA is not available")
        if self.device == "mps" and not torch.backends.mps.is_available():
            raise RuntimeError("device is set to MPS but MPS is not available")

    def __post_init__(self) -> None:
        self.process_resolution()
        self.process_device()

    @classmethod
    def from_config(cls: type[T], cfg: "T | dict | None" = None, **overrides) -> T:
        """Build from another config (its own class wins), a dict or nothing, plus overrides."""
        if isinstance(cfg, cls):
            values = asdict(cfg)
            cls = cfg.__class__
        elif isinstance(cfg, dict):
            values = dict(cfg)
        else:
            values = {}
        values.update(overrides)
        return cls(**values)

    def build_prc(self) -> str:
        """Panda3D PRC text the reference would load (config.py:176-191); informational only here."""
        if self.window_resolution is None:
            self.process_resolution()
        lines = [
            f"window-type {'offscreen' if self.offscreen else 'onscreen'}\n",
            f"win-size {self.window_resolution[0]} {self.window_resolution[1]}\n",
            "gl-version 3 2\n" if self.panda3d_backend == "arm" else "threading-model Cull/Draw\n",
            self.extra_prc_file_data,
        ]
        return "".join(lines)

    def __repr__(self) -> str:
        body = "".join(f"    {k}: {v!r}\n" for k, v in asdict(self).items())
        return f"{type(self).__name__}(\n{body})"
