"""Env registry / CartPole env (host logic, no rendering) and mesh bake rules."""
import math

import numpy as np
import pytest
import torch

import pybatchrender_b200 as pbr
from pybatchrender_b200 import meshes


def test_registry_surface():
    assert "CartPole-v0" in pbr.envs.list_envs() and pbr.envs.is_registered("CartPole-v0")
    env_cls, rend_cls, cfg_cls = pbr.envs.get_env_classes("CartPole-v0")
    assert cfg_cls.__name__ == "CartPoleConfig"
    with pytest.raises(ValueError):
        pbr.envs.make("Nope-v0")
    pbr.envs.register("Tmp-v0", env_cls, rend_cls, cfg_cls)
    with pytest.raises(ValueError):
        pbr.envs.register("Tmp-v0", env_cls, rend_cls, cfg_cls)
    pbr.envs.unregister("Tmp-v0")
    assert not pbr.envs.is_registered("Tmp-v0")


def test_cartpole_env_loop_like_the_reference_benchmark():
    env = pbr.envs.make("CartPole-v0", num_scenes=16, device="cpu", render=False, seed=3)
    td = env.reset()
    assert td["observation"].shape == (16, 4) and "pixels" not in td.keys()
    x0 = td["observation"][:, 0]
    assert float(x0.min()) >= -2.0 and float(x0.max()) <= 2.0
    total = 0.0
    for _ in range(5):
        td["action"] = env.action_spec.rand()
        td = env.step(td)
        total += td["next", "reward"].sum().item()
        assert td["next", "done"].dtype == torch.bool and td["next", "done"].shape == (16, 1)
        td = td["next"]
    assert total == 5 * 16
    assert int(td["step_count"].max()) <= 5


def test_cartpole_dynamics_match_gym_equations():
    env = pbr.envs.make("CartPole-v0", num_scenes=1, device="cpu", render=False)
    obs = torch.tensor([[0.1, -0.2, 0.05, 0.3]])
    nxt = env._dynamics(obs, torch.tensor([1]))[0].numpy()
    x, xd, th, thd = 0.1, -0.2, 0.05, 0.3
    force, mp_, mc, l, g, tau = 10.0, 0.1, 1.0, 0.5, 9.8, 0.02
    temp = (force + mp_ * l * thd ** 2 * math.sin(th)) / (mp_ + mc)
    tha = (g * math.sin(th) - math.cos(th) * temp) / (l * (4 / 3 - mp_ * math.cos(th) ** 2 / (mp_ + mc)))
    xa = temp - mp_ * l * tha * math.cos(th) / (mp_ + mc)
    np.testing.assert_allclose(nxt, [x + tau * xd, xd + tau * xa, th + tau * thd, thd + tau * tha], rtol=1e-5)


def test_auto_reset_and_termination():
    env = pbr.envs.make("CartPole-v0", num_scenes=4, device="cpu", render=False, x_threshold=0.0)
    td = env.reset()
    td["action"] = env.action_spec.rand()
    td = env.step(td)
    assert bool(td["next", "done"].all())                  # |x| > 0 everywhere -> done
    assert int(td["next", "step_count"].max()) == 0         # auto-reset zeroes the counters


def test_save_batch_examples_grid_layout(tmp_path):
    env = pbr.envs.make("CartPole-v0", num_scenes=16, device="cpu", render=False)
    px = torch.zeros(16, 3, 4, 4, dtype=torch.uint8)
    for i in range(16):
        px[i] = i * 10
    path, data = env.save_batch_examples(pixels=px, num=16, scale=3, out_dir=str(tmp_path), return_bytes=True)
    from PIL import Image
    import io
    a = np.array(Image.open(io.BytesIO(data)))
    assert a.shape == (48, 48, 3)
    tiles = a[::3, ::3].reshape(4, 4, 4, 4, 3).transpose(0, 2, 1, 3, 4).reshape(16, 4, 4, 3)
    assert [int(t[0, 0, 0]) for t in tiles] == [i * 10 for i in range(16)]   # tile n at row n//4, col n%4


def test_box_mesh_is_the_unit_cube_with_face_normals():
    m = meshes.box()
    assert m.pos.shape == (24, 3) and m.idx.shape == (12, 3)
    lo, hi = m.tight_bounds()
    assert lo.tolist() == [0, 0, 0] and hi.tolist() == [1, 1, 1]
    for t in m.idx:
        p0, p1, p2 = m.pos[t]
        n = np.cross(p1 - p0, p2 - p0)
        n /= np.linalg.norm(n)
        assert np.allclose(n, m.nrm[t[0]]) and np.allclose(m.nrm[t[0]], m.nrm[t[1]])   # CCW outward, flat


def test_bake_scale_pivot_like_the_cartpole_nodes():
    pole = meshes.bake(meshes.box(), model_scale=(0.1, 0.1, 2.0), pivot_rel=(0.5, 0.5, 0.05))
    lo, hi = pole.tight_bounds()
    np.testing.assert_allclose(lo, [-0.05, -0.05, -0.1], atol=1e-7)
    np.testing.assert_allclose(hi, [0.05, 0.05, 1.9], atol=1e-7)
    rail = meshes.bake(meshes.box(), model_scale=(6.0, 0.05, 0.05), pivot_rel=(0.5, 0.5, 0.5))
    lo, hi = rail.tight_bounds()
    np.testing.assert_allclose(lo, [-3, -0.025, -0.025], atol=1e-7)
    # normals stay axis aligned under the non-uniform scale
    assert set(map(tuple, np.abs(rail.nrm).round(6))) == {(1, 0, 0), (0, 1, 0), (0, 0, 1)}


def test_bake_absolute_units_and_hpr():
    m = meshes.bake(meshes.box(), model_scale=4.0, model_scale_units="absolute")
    lo, hi = m.tight_bounds()
    assert np.allclose(hi - lo, 4.0)
    m = meshes.bake(meshes.box(), model_scale=(2.0, 3.0, 5.0), model_scale_units="absolute")
    lo, hi = m.tight_bounds()
    assert np.allclose(hi - lo, [2, 3, 5])
    m = meshes.bake(meshes.box(), model_hpr=(90.0, 0.0, 0.0))       # heading: +x -> +y
    assert np.allclose(m.pos.min(0), [-1, 0, 0], atol=1e-6) and np.allclose(m.pos.max(0), [0, 1, 1], atol=1e-6)
    with pytest.raises(ValueError):
        meshes.bake(meshes.box(), model_scale=(1, 2))
    with pytest.raises(FileNotFoundError):
        meshes.load_mesh("models/does_not_exist")


def test_sphere_is_closed_and_outward():
    s = meshes.uv_sphere(1.0, 12, 8)
    for t in s.idx:
        p0, p1, p2 = s.pos[t]
        n = np.cross(p1 - p0, p2 - p0)
        assert np.dot(n, (p0 + p1 + p2) / 3) > 0


def test_drop_in_import_name():
    """``import pybatchrender`` (the reference's package name) resolves to this implementation."""
    import pybatchrender as ref_name
    from pybatchrender.renderer.renderer import PBRRenderer as R1
    from pybatchrender.envs.cartpole.renderer import CartPoleRenderer as C1
    from pybatchrender import PBRConfig, PBRRenderer, PBREnv  # noqa: F401  (Steering's import line)
    assert R1 is pbr.PBRRenderer and ref_name.PBRConfig is pbr.PBRConfig
    assert "CartPole-v0" in ref_name.envs.list_envs()
    assert C1(dict(num_scenes=2, device="cpu")).num_scenes == 2
