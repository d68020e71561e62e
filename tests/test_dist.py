"""Scene sharding: partition logic, the shard configs, and a world_size-2 gloo run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from hypothesis import given, settings, strategies as st

from pybatchrender_b200 import PBRConfig
from pybatchrender_b200.dist import gather_frames, shard_config, shard_range
from pybatchrender_b200.envs.cartpole import CartPoleConfig, CartPoleRenderer


@given(st.integers(1, 100000), st.integers(1, 16))
@settings(max_examples=200, deadline=None)
def test_shard_ranges_partition_the_batch(n, world):
    nxt = 0
    for r in range(world):
        b, c = shard_range(n, r, world)
        assert b == nxt and c in (n // world, n // world + 1)
        nxt = b + c
    assert nxt == n


def test_shard_config_keeps_the_global_grid():
    g = PBRConfig(num_scenes=4098, tile_resolution=(64, 64), device="cpu")
    shards = [shard_config(g, r, 4) for r in range(4)]
    assert sum(s.num_scenes for s in shards) == 4098
    for s in shards:
        assert s.tiles == g.tiles == (65, 64) and s.window_resolution == g.window_resolution
        assert s.global_num_scenes == 4098
    assert [s.scene_offset for s in shards] == [0, 1025, 2050, 3074]
    with pytest.raises(ValueError):
        shard_config(PBRConfig(num_scenes=2, device="cpu"), 3, 4)


def test_sharded_cartpole_state_equals_the_global_one():
    """Per-scene buffers of the shards, concatenated, equal the single-process buffers (same aspect,
    same cart colour ramp), so the sharded render is the same image set."""
    n, world = 37, 3
    g = CartPoleConfig(num_scenes=n, device="cpu")
    full = CartPoleRenderer(g)
    state = torch.rand(n, 4) - 0.5
    full._step(state)
    parts = []
    for r in range(world):
        cfg = shard_config(g, r, world)
        sh = CartPoleRenderer(cfg)
        sh._step(state[cfg.scene_offset:cfg.scene_offset + cfg.num_scenes])
        parts.append(sh)
        assert torch.equal(sh._pbr_cam.P_k44, full._pbr_cam.P_k44)
        assert torch.equal(sh.rail.matbuf, full.rail.matbuf)
    for name in ("cart", "pole"):
        cat_m = torch.cat([getattr(p, name).matbuf for p in parts])
        cat_c = torch.cat([getattr(p, name).colbuf for p in parts])
        assert torch.equal(cat_m, getattr(full, name).matbuf)
        assert torch.equal(cat_c, getattr(full, name).colbuf)
    assert torch.equal(torch.cat([p._pbr_cam.viewbuf for p in parts]), full._pbr_cam.viewbuf)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = shard_config(CartPoleConfig(num_scenes=n, device="cpu", tile_resolution=(8, 8)))
        # stand-in frames: scene index in every byte, so the gather order is checkable
        idx = torch.arange(cfg.scene_offset, cfg.scene_offset + cfg.num_scenes, dtype=torch.uint8)
        local = idx.view(-1, 1, 1, 1).expand(-1, 3, 8, 8).contiguous()
        out = gather_frames(local, dst=0)
        t = torch.tensor([float(cfg.num_scenes)])
        dist.all_reduce(t)
        if rank == 0:
            q.put((out[:, 0, 0, 0].tolist(), float(t.item())))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather():
    n, world = 11, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    order, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert order == list(range(n)) and total == n
