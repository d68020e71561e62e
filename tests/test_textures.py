"""Textured nodes (reference node.py:277-287, basic.frag:31-37)."""
import numpy as np
import pytest
import torch

from pybatchrender_b200 import PBRRenderer
from util import oracle_render


def _pattern(h, w, seed=0):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[::2, ::2] = (255, 255, 255)
    return img


def _scene(device, *, num_scenes=4, inst=5, tile=(64, 64), tex_shape=(7, 13), eye=(0.0, -9.0, 1.0), spread=3.0,
           mixed=True, two_sided=False, seed=3, channels=3, sphere=True, use=1.0):
    from pybatchrender_b200 import meshes
    r = PBRRenderer(dict(num_scenes=num_scenes, tile_resolution=tile, device=device, num_channels=channels))
    rng = np.random.default_rng(seed)
    nodes = [r.add_node("models/box", instances_per_scene=inst, texture=_pattern(*tex_shape, seed=seed),
                        model_pivot_relative_point=(0.5, 0.5, 0.5))]
    if sphere:
        m = meshes.uv_sphere(1.0, 10, 6)
        m.two_sided = two_sided
        if two_sided:
            m.idx = m.idx[: (2 * m.idx.shape[0]) // 3].copy()
        meshes.register_mesh(f"test/tex_sphere_{int(two_sided)}", m)
        nodes.append(r.add_node(f"test/tex_sphere_{int(two_sided)}", instances_per_scene=2,
                                texture=_pattern(16, 16, seed=seed + 1), shared_across_scenes=True))
    if mixed:
        nodes.append(r.add_node("models/box", instances_per_scene=2, model_pivot_relative_point=(0.5, 0.5, 0.5)))
    for node in nodes:
        B = node.buf_instances
        node.set_positions(torch.tensor(rng.uniform(-spread, spread, (B, 3)), dtype=torch.float32), lazy=True)
        node.set_hprs(torch.tensor(rng.uniform(-np.pi, np.pi, (B, 3)), dtype=torch.float32), lazy=True)
        node.set_scales(torch.tensor(rng.uniform(1.0, 2.5, (B, 1)), dtype=torch.float32))
        node.set_colors(torch.tensor(np.concatenate([rng.uniform(0.3, 1, (B, 3)), np.ones((B, 1))], 1), dtype=torch.float32))
    if use != 1.0:
        nodes[0]._set_shader_input("useTexture", use)
    cam = r.add_camera()
    cam.set_positions(torch.tensor(eye, dtype=torch.float32))
    cam.look_at(torch.tensor([0.0, 0.0, 0.0]))
    r.add_light()
    r.setup_environment()
    return r


def test_oracle_texture_orientation_and_filtering():
    """One unlit box face filling the view with a 2x2 image: the picture appears upright (row 0 of
    the array at the top), quadrant centres are the pure texel colours, the centre is their blend."""
    r = PBRRenderer(dict(num_scenes=1, tile_resolution=(64, 64), device="cpu"))
    tex = np.array([[(255, 0, 0), (0, 255, 0)], [(0, 0, 255), (255, 255, 0)]], np.uint8)
    n = r.add_node("models/box", instances_per_scene=1, texture=tex, model_pivot_relative_point=(0.5, 0.5, 0.5))
    n.set_scales(torch.tensor([[4.0]]))
    r.add_camera().set_positions(torch.tensor([0.0, -6.0, 0.0]))
    r.setup_environment()
    assert n.texture_image.shape == (2, 2, 4) and n.texture_image[0, 0].tolist() == [0, 0, 255, 255]   # row 0 = bottom
    img = oracle_render(r)[0]

    def near(px, want):
        return np.abs(px.astype(int) - np.array(want)).max() <= 24
    assert near(img[:, 16, 16], (255, 0, 0)) and near(img[:, 16, 48], (0, 255, 0))
    assert near(img[:, 48, 16], (0, 0, 255)) and near(img[:, 48, 48], (255, 255, 0))
    assert near(img[:, 32, 32], (128, 128, 64))
    # texture=None / False switch the sampler off; True alone leaves the node white
    n.set_texture(None)
    assert n.shader_inputs["useTexture"] == 0.0 and n.texture_image is None
    plain = oracle_render(r)[0]
    n.set_texture(True)
    assert n.shader_inputs["useTexture"] == 1.0
    assert np.array_equal(oracle_render(r)[0], plain)
    with pytest.raises(ValueError):
        n.set_texture(np.zeros((4, 4), np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [
    dict(),                                                    # staged path, shared textured spheres + plain boxes
    dict(num_scenes=3, inst=2, sphere=False, mixed=False),     # small frame: fused general kernel
    dict(tile=(160, 96), channels=4, tex_shape=(1, 1)),
    dict(eye=(0.3, -1.2, 0.2), spread=1.5, two_sided=True),    # camera among the instances: clipped, two-sided
    dict(use=0.35, tex_shape=(32, 5)),
])
@pytest.mark.parametrize("fused", [0, 2, 8, 16], ids=["auto", "fused", "staged", "binned"])
def test_textured_frames_bit_exact(kw, fused):
    r = _scene("cuda", **kw)
    r.render_flags = fused        # PBR_FRAME_FORCE_FUSED / _STAGED / _BINNED
    got = r.render().cpu().numpy()
    assert np.array_equal(got, oracle_render(r)), f"textured {kw} fused={fused}"
    assert r._native.device_status(torch.cuda.current_device()) == 0


@pytest.mark.gpu
def test_texture_with_static_layer_and_cartpole_kernel():
    """A shared textured node goes into the static layer (general kernel), the per-scene boxes stay
    on the one-warp-per-scene kernel; the composite equals the oracle."""
    r = PBRRenderer(dict(num_scenes=8, tile_resolution=(64, 64), device="cuda"))
    floor = r.add_node("models/box", instances_per_scene=1, texture=_pattern(8, 8), model_scale=(6.0, 6.0, 0.2),
                       model_pivot_relative_point=(0.5, 0.5, 0.5), shared_across_scenes=True)
    floor.set_positions(torch.tensor([[0.0, 0.0, -1.0]]))
    box = r.add_node("models/box", instances_per_scene=1, model_pivot_relative_point=(0.5, 0.5, 0.5))
    box.set_positions(torch.tensor(np.random.default_rng(0).uniform(-2, 2, (8, 3)), dtype=torch.float32))
    cam = r.add_camera()
    cam.set_positions(torch.tensor([4.0, -6.0, 3.0]))
    cam.look_at(torch.tensor([0.0, 0.0, 0.0]))
    r.add_light()
    r.setup_environment()
    assert np.array_equal(r.render().cpu().numpy(), oracle_render(r))
    assert r._base_sig is not None                        # the static layer was used
