"""Scene sharding across the GPUs of one box.

Scenes are independent: each reads its own VP / instance rows plus shared read-only data and writes
``out[scene]``.  So the multi-GPU path is a contiguous block partition of ``num_scenes`` with **no
collective on the hot path**: rank r owns scenes ``[r*N/G, (r+1)*N/G)``, one process per GPU.  The
reference's analogue is TorchRL ``ParallelEnv`` with one Panda3D process per worker
(``pybatchrender/env.py:360-416``).

Only two things depend on the *global* batch: the tile grid that fixes the projection aspect
(quirk Q1, reference ``camera.py:154``) and per-scene constants that are functions of the global
scene index (e.g. CartPole's cart colour ramp, ``envs/cartpole/renderer.py:72-76``).  Both are
carried by ``PBRConfig.tiles`` / ``scene_offset`` / ``global_num_scenes``.

``gather_frames`` is the optional collective (frames of all ranks onto one policy rank, NCCL over
NVLink); it is never part of ``renderer.step`` and is timed separately by ``bench.py --gather``.
"""
from __future__ import annotations

import os
from dataclasses import asdict

import torch

from .config import PBRConfig, grid_for


def shard_range(num_scenes: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous block partition: (begin, count) of rank's scenes; the first ``N % G`` ranks get one more."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, rem = divmod(int(num_scenes), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def shard_config(cfg: PBRConfig | dict, rank: int | None = None, world_size: int | None = None):
    """Config of one rank's shard of the global batch described by ``cfg``.

    ``num_scenes`` becomes the local count, ``tiles`` stays the global grid (same aspect, hence
    bit-identical images to the single-process render), ``scene_offset`` / ``global_num_scenes``
    record where the shard sits."""
    if rank is None:
        rank = int(os.environ.get("RANK", "0"))
    if world_size is None:
        world_size = int(os.environ.get("WORLD_SIZE", "1"))
    cls = PBRConfig
    if isinstance(cfg, PBRConfig):
        cls = type(cfg)
        values = asdict(cfg)
    else:
        values = dict(cfg)
    full = cls(**values)
    n_global = int(full.num_scenes)
    begin, count = shard_range(n_global, rank, world_size)
    if count == 0:
        raise ValueError(f"rank {rank} of {world_size} gets no scenes out of {n_global}")
    values = asdict(full)
    values.update(num_scenes=count, tiles=tuple(full.tiles), tile_resolution=tuple(full.tile_resolution),
                  window_resolution=tuple(full.window_resolution), batch_inner_dim=None,
                  scene_offset=int(full.scene_offset) + begin, global_num_scenes=n_global)
    return cls(**values)


def global_grid(num_scenes: int) -> tuple[int, int]:
    return grid_for(int(num_scenes))


def gather_frames(local: torch.Tensor, dst: int = 0, group=None) -> torch.Tensor | None:
    """Collect every rank's ``[N_r, C, H, W]`` uint8 frames on rank ``dst`` (``[sum N_r, C, H, W]``).

    Equal shard sizes use one ``gather``; ragged shards fall back to ``all_gather_object``-free
    padding.  Returns None on the other ranks."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(sizes)
    if local.shape[0] < n_max:
        pad = torch.zeros((n_max - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    local = local.contiguous()
    if rank == dst:
        bufs = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, gather_list=bufs, dst=dst, group=group)
        return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
    dist.gather(local, gather_list=None, dst=dst, group=group)
    return None
