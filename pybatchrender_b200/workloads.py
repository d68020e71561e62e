"""Synthetic scenes of the BASELINE configurations that are not environments.

* ``many_cubes`` -- config 3, the shape of the reference's ``pybatchrender/demo_many_cubes.py:34-54``:
  ``instances`` per-scene boxes around the default camera, positions U(-15,15)^3, HPR U(-pi,pi)^3,
  uniform scale U(0.5,1.8), colours U(0,1)^3, from ``numpy.random.default_rng(seed)`` (the demo's
  unseeded ``np.random.rand`` replaced by a seeded generator).
* ``mixed_meshes`` -- config 5: four per-scene nodes x ``per_node`` instances -- procedural box,
  ``models/cone.egg``, ``models/cylinder/scene.gltf`` and a UV sphere (stand-in for Panda3D's un-vendored
  ``models/smiley``) -- camera at (0,-40,10) looking at the origin, instances drawn as in config 3.

Both take the *global* batch description plus a shard (``rank`` / ``world_size``): a rank builds only its
own scenes' rows but keeps the global tile grid, so the concatenated shards equal the single-process
frame (``dist.shard_config``).  ``cartpole_state`` is the state distribution of the CartPole configs.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .config import PBRConfig
from .dist import shard_config
from .renderer.renderer import PBRRenderer


def cartpole_state(n: int, seed: int) -> torch.Tensor:
    """x ~ U(-2,2), theta ~ U(-30deg,30deg): the reference's reset ranges (envs/cartpole/config.py:57-62);
    x_dot / theta_dot do not reach the pixels."""
    g = torch.Generator().manual_seed(seed)
    s = torch.zeros(n, 4)
    s[:, 0] = torch.rand(n, generator=g) * 4.0 - 2.0
    s[:, 1] = torch.rand(n, generator=g) * 2.0 - 1.0
    s[:, 2] = (torch.rand(n, generator=g) * 60.0 - 30.0) * (math.pi / 180.0)
    s[:, 3] = (torch.rand(n, generator=g) * 30.0 - 15.0) * (math.pi / 180.0)
    return s


def _sharded(num_scenes, tile, device, rank, world_size, channels=3):
    cfg = PBRConfig(num_scenes=int(num_scenes), tile_resolution=tuple(tile), num_channels=channels,
                    **({} if device is None else {"device": device}))
    if world_size > 1:
        cfg = shard_config(cfg, rank, world_size)
    first = int(getattr(cfg, "scene_offset", 0) or 0)
    return PBRRenderer(cfg), first, int(cfg.num_scenes)


def many_cubes(num_scenes: int = 1024, instances: int = 256, tile=(128, 128), seed: int = 123, device=None,
               rank: int = 0, world_size: int = 1, spread: float = 15.0) -> PBRRenderer:
    r, first, n = _sharded(num_scenes, tile, device, rank, world_size)
    node = r.add_node("models/box", instances_per_scene=instances, model_pivot_relative_point=(0.5, 0.5, 0.5))
    rng = np.random.default_rng(seed)
    B = num_scenes * instances
    rows = slice(first * instances, (first + n) * instances)
    pos = rng.uniform(-spread, spread, (B, 3))
    hpr = rng.uniform(-np.pi, np.pi, (B, 3))
    sc = rng.uniform(0.5, 1.8, (B, 1))
    col = np.concatenate([rng.uniform(0, 1, (B, 3)), np.ones((B, 1))], axis=1)
    node.set_positions(torch.tensor(pos[rows], dtype=torch.float32), lazy=True)
    node.set_hprs(torch.tensor(hpr[rows], dtype=torch.float32), lazy=True)
    node.set_scales(torch.tensor(sc[rows], dtype=torch.float32))
    node.set_colors(torch.tensor(col[rows], dtype=torch.float32))
    cam = r.add_camera()
    cam.set_positions(torch.tensor([0.0, -12.0, 0.0]))
    r.add_light()
    r.setup_environment()
    return r


def mixed_meshes(num_scenes: int = 16384, per_node: int = 16, tile=(256, 256), seed: int = 123, device=None,
                 rank: int = 0, world_size: int = 1, spread: float = 15.0) -> PBRRenderer:
    r, first, n = _sharded(num_scenes, tile, device, rank, world_size)
    nodes = [
        r.add_node("models/box", instances_per_scene=per_node, model_pivot_relative_point=(0.5, 0.5, 0.5)),
        r.add_node("models/cone.egg", instances_per_scene=per_node, model_pivot_relative_point=(0.5, 0.5, 0.5)),
        r.add_node("models/cylinder/scene.gltf", instances_per_scene=per_node, model_scale=2.0,
                   model_scale_units="absolute", model_pivot_relative_point=(0.5, 0.5, 0.5)),
        r.add_node("models/smiley", instances_per_scene=per_node),
    ]
    g = torch.Generator().manual_seed(seed)
    B = num_scenes * per_node
    rows = slice(first * per_node, (first + n) * per_node)
    for node in nodes:
        node.set_positions(((torch.rand(B, 3, generator=g) * 2 - 1) * spread)[rows], lazy=True)
        node.set_hprs(((torch.rand(B, 3, generator=g) * 2 - 1) * math.pi)[rows], lazy=True)
        node.set_scales((torch.rand(B, 1, generator=g) * 1.3 + 0.5)[rows])
        node.set_colors(torch.cat([torch.rand(B, 3, generator=g), torch.ones(B, 1)], 1)[rows])
    cam = r.add_camera()
    cam.set_positions(torch.tensor([0.0, -40.0, 10.0]))
    cam.look_at(torch.tensor([0.0, 0.0, 0.0]))
    r.add_light()
    r.setup_environment()
    return r
