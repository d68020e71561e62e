"""Steering-v0: host logic on CPU, pixels on the GPU against the oracle."""
import dataclasses
import importlib.util
import os

import numpy as np
import pytest
import torch

import pybatchrender_b200 as pbr
from pybatchrender_b200.envs.steering import SteeringConfig, SteeringEnv, SteeringRenderer

REF_CFG = "/root/reference/pybatchrender/envs/steering/config.py"


def _env(n=8, **kw):
    kw.setdefault("device", "cpu")
    kw.setdefault("render", False)
    return pbr.envs.make("Steering-v0", num_scenes=n, **kw)


def test_registered_and_specs():
    assert "Steering-v0" in pbr.envs.list_envs()
    env = _env(4)
    assert isinstance(env, SteeringEnv) and isinstance(env._renderer, SteeringRenderer)
    a = env.action_spec.rand()
    assert a.shape == (4, 1) and float(a.min()) >= env.action_low and float(a.max()) <= env.action_high
    td = env.reset()
    assert td["observation"].shape == (4, 6) and "pixels" not in td.keys()
    # x uniformly inside the drivable strip, first obstacle ahead
    assert float(td["observation"][:, 0].abs().max()) <= env.x_max == 12.5 - 1.0 - 0.1
    assert torch.equal(td["observation"][:, 5], torch.full((4,), 6.0))


@pytest.mark.skipif(not os.path.isfile(REF_CFG), reason="reference checkout not present")
def test_config_fields_and_defaults_match_the_reference():
    spec = importlib.util.spec_from_file_location("_ref_steering_config", REF_CFG)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)             # its `from pybatchrender import PBRConfig` resolves to the alias package
    ref = {f.name: f for f in dataclasses.fields(mod.SteeringConfig)}
    mine = {f.name: f for f in dataclasses.fields(SteeringConfig)}
    assert set(ref) <= set(mine)
    a, b = mod.SteeringConfig(num_scenes=4, device="cpu"), SteeringConfig(num_scenes=4, device="cpu")
    for name in ref:
        assert getattr(a, name) == getattr(b, name), name


def _pop_loop(obstacles, idx, px, py, ppx, ppy, thr):
    """Per-scene restatement of the reference's pop-while-passed loop (env.py:259-301)."""
    hits, passes = 0, 0
    n = obstacles.shape[0]
    while idx < n and obstacles[idx, 1] <= py:
        oy, ox = float(obstacles[idx, 1]), float(obstacles[idx, 0])
        den = py - ppy
        den = 1.0 if abs(den) < 1e-6 else den
        t = np.float32((np.float32(oy) - np.float32(ppy)) / np.float32(den))
        xa = np.float32(ppx) + np.float32(np.float32(px) - np.float32(ppx)) * t
        hits += int(abs(np.float32(xa) - np.float32(ox)) <= thr)
        idx += 1
        passes += 1
    return hits, passes, idx


def test_collision_pop_equals_the_sequential_loop():
    torch.manual_seed(3)
    env = _env(64, number_of_obstacles=30, obstacle_y_spacing=3.0)
    env.reset()
    rng = np.random.default_rng(0)
    for _ in range(40):
        env._prev_player_x, env._prev_player_y = env._player_x.clone(), env._player_y.clone()
        env._player_y = env._player_y + torch.tensor(rng.uniform(0.0, 11.0, 64), dtype=torch.float32)   # 0..3 obstacles per step
        env._player_x = (env._player_x + torch.tensor(rng.uniform(-2, 2, 64), dtype=torch.float32)).clamp(env.x_min, env.x_max)
        before = env._next_obstacle_idx.clone()
        hits, passes = env._check_collisions()
        for s in range(64):
            h, p, i = _pop_loop(env._obstacles[s].numpy(), int(before[s]), float(env._player_x[s]), float(env._player_y[s]),
                                float(env._prev_player_x[s]), float(env._prev_player_y[s]), env.collision_threshold)
            assert (int(hits[s]), int(passes[s]), int(env._next_obstacle_idx[s])) == (h, p, i), s
    assert int(env._next_obstacle_idx.max()) == 30        # some scenes ran off the end of the track


def test_transition_rules():
    env = _env(2, number_of_obstacles=3, distance_to_first_obstacle=1.0, obstacle_y_spacing=50.0, max_steps=10_000)
    td = env.reset()
    # put obstacle 0 right in front of scene 0 and far to the side of scene 1
    env._obstacles[:, 0, 0] = torch.tensor([0.0, 11.0])
    env._player_x = torch.tensor([0.5, -11.0])
    td["action"] = torch.zeros(2, 1)
    nxt = env.step(td)["next"]
    assert nxt["reward"].flatten().tolist() == [-1.0, 0.0]
    speeds = nxt["observation"][:, 2]
    assert torch.allclose(speeds, torch.tensor([96.0 + 0.072 - 6.12, 96.0 + 0.072]))
    assert nxt["observation"][:, 3].tolist() == [1.0, 0.0]                 # grace flag
    assert nxt["observation"][:, 5].tolist() == [51.0, 51.0]               # next obstacle
    assert not bool(nxt["done"].any())
    # steering: x moves by action * speed / sensitivity * tau and is clamped to the strip
    nxt["action"] = torch.tensor([[327.68], [-327.68]])
    x0, v = nxt["observation"][:, 0].clone(), nxt["observation"][:, 2].clone()
    n2 = env.step(nxt)["next"]
    want = (x0 + torch.tensor([327.68, -327.68]) * v / 800.0 / 60.0).clamp(env.x_min, env.x_max)
    assert torch.allclose(n2["observation"][:, 0], want, atol=1e-5)
    # running off the last obstacle ends the episode; the step counter restarts (auto_reset)
    env._player_y = torch.full((2,), 200.0)
    n2["action"] = torch.zeros(2, 1)
    n3 = env.step(n2)["next"]
    assert bool(n3["done"].all()) and n3["step_count"].tolist() == [0, 0]
    assert torch.allclose(n3["observation"][:, 5], n3["observation"][:, 1] + 1000.0)


def test_renderer_host_state_follows_the_player():
    r = SteeringRenderer(SteeringConfig(num_scenes=3, device="cpu"))
    obs = torch.zeros(3, 7, 4)
    obs[..., 0] = torch.arange(7) - 3.0
    obs[..., 1] = torch.arange(7) * 12.0 + 6
    obs[:, ::2, 2] = 1.0
    r.build_obstacles(obs)
    assert r._setup_called and r.sphere_node.instances_per_scene == 7
    first = r.sphere_node
    r.build_obstacles(obs)                                   # same count: the node is reused
    assert r.sphere_node is first and len(r._drawable_nodes()) == 4
    r.build_obstacles(obs[:, :5])                            # new count: replaced
    assert r.sphere_node is not first and len(r._drawable_nodes()) == 4
    cols = r.sphere_node.colbuf.reshape(3, 5, 4)
    assert cols[0, 0].tolist() == [1.0, 0.5, 0.0, 1.0] and cols[0, 1].tolist() == [1.0, 0.0, 0.0, 1.0]

    state = torch.tensor([[1.0, 10.0, 96.0, 0.0], [-2.0, 20.0, 96.0, 1.0], [3.0, 30.0, 96.0, 0.0]])
    r._step(state)
    mats = r.player_node.matbuf.reshape(3, 4, 4)             # texel j = column j: translation in row 3
    assert torch.allclose(mats[:, 3, :3], torch.tensor([[1.0, 10.0, 0.0], [-2.0, 20.0, 0.0], [3.0, 30.0, 0.0]]))
    left = r.left_border.matbuf.reshape(3, 4, 4)[:, 3, :3]
    assert torch.allclose(left, torch.tensor([[-12.5, 10.0, 0.5], [-12.5, 20.0, 0.5], [-12.5, 30.0, 0.5]]))
    assert r.player_node.colbuf.reshape(3, 4)[1].tolist() == pytest.approx([0.4, 0.4, 0.4, 1.0])
    assert r.player_node.colbuf.reshape(3, 4)[0].tolist() == [0.0, 0.0, 1.0, 1.0]
    # camera: eye = player + (0, -16.3, 4), per scene
    assert not r._pbr_cam.uniform
    r.build_obstacles(None)
    assert r.sphere_node is None and len(r._drawable_nodes()) == 3


# ---------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_steering_frames_match_the_oracle():
    from util import oracle_render
    torch.manual_seed(11)
    env = pbr.envs.make("Steering-v0", num_scenes=6, device="cuda", number_of_obstacles=40, tile_resolution=(96, 64))
    td = env.reset()
    assert td["pixels"].shape == (6, 3, 64, 96) and td["pixels"].is_cuda
    assert np.array_equal(td["pixels"].cpu().numpy(), oracle_render(env._renderer))
    for _ in range(25):
        td["action"] = env.action_spec.rand()
        td = env.step(td)["next"]
    px = td["pixels"]
    assert np.array_equal(px.cpu().numpy(), oracle_render(env._renderer))
    assert int((px != 0).any(1).sum()) > 200                 # player, rails and obstacles are in view
    assert env._renderer._native.device_status(torch.cuda.current_device()) == 0


@pytest.mark.gpu
def test_steering_pose_kernel_equals_generic_setters():
    cfg = SteeringConfig(num_scenes=5, device="cuda")
    a, b = SteeringRenderer(cfg), SteeringRenderer(cfg)
    b._native_keep, b._native = b._native, None              # force the torch setter path for the poses
    state = torch.tensor(np.random.default_rng(2).uniform(-10, 300, (5, 4)), dtype=torch.float32, device="cuda")
    a._step(state)
    b._step(state)
    for name in ("player_node", "left_border", "right_border"):
        assert torch.equal(getattr(a, name).matbuf, getattr(b, name).matbuf), name
    assert torch.equal(a._pbr_cam.viewbuf, b._pbr_cam.viewbuf)
