"""SURVEY.md 8 row f1 on the GPU: the TorchRL-style CartPole environment produces, at every step, exactly the
oracle's frames of the state it rendered (reference loop: examples/scripts/cartpole_benchmark.py:143-164 over
pybatchrender/envs/cartpole/env.py:149-218)."""
import numpy as np
import pytest
import torch

import pybatchrender as pbr
from util import host_threads, oracle_render

pytestmark = pytest.mark.gpu


def test_cartpole_env_pixels_match_the_oracle_every_step():
    torch.manual_seed(5)
    n = 4096
    env = pbr.envs.make("CartPole-v0", num_scenes=n, device="cuda", max_steps=8)     # episodes end inside the loop
    td = env.reset()
    px = td["pixels"]
    assert px.shape == (n, 3, 64, 64) and px.dtype == torch.uint8 and px.is_cuda
    assert np.array_equal(px.cpu().numpy(), oracle_render(env._renderer, n_threads=host_threads())), "reset"
    seen_done = 0
    for step in range(20):
        td["action"] = env.action_spec.rand()
        out = env.step(td)
        nxt = out["next"]
        got = nxt["pixels"].cpu().numpy()
        # frames show the state the physics moved to (before an auto-reset replaces finished episodes)
        ref = oracle_render(env._renderer, n_threads=host_threads())
        assert np.array_equal(got, ref), f"step {step}: {(got != ref).any((1, 2, 3)).sum()} scenes differ"
        seen_done += int(nxt["done"].sum())
        td = nxt
    assert (got != 0).any()
    assert seen_done > 0, "no episode ended in 20 steps: the auto-reset branch was not exercised"


def test_cartpole_env_small_batches_and_other_tiles():
    torch.manual_seed(6)
    for n, tile in ((5, (64, 64)), (130, (84, 84)), (33, (32, 48))):
        env = pbr.envs.make("CartPole-v0", num_scenes=n, device="cuda", tile_resolution=tile)
        td = env.reset()
        for _ in range(6):
            td["action"] = env.action_spec.rand()
            td = env.step(td)["next"]
        assert td["pixels"].shape == (n, 3, tile[1], tile[0])
        assert np.array_equal(td["pixels"].cpu().numpy(), oracle_render(env._renderer)), (n, tile)
