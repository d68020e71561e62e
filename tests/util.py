"""Shared helpers for the parity tests: renderer state -> oracle frame, synthetic scenes."""
from __future__ import annotations

import numpy as np
import torch

import oracle


def oracle_frame(renderer) -> oracle.OracleFrame:
    fa = renderer.frame_arrays()
    return oracle.OracleFrame(
        num_scenes=fa["num_scenes"], tile_w=fa["tile_w"], tile_h=fa["tile_h"], channels=fa["channels"],
        vp=fa["vp"], bg=fa["bg"], ambient=fa["ambient"], dir_dir=fa["dir_dir"], dir_col=fa["dir_col"],
        strength=fa["strength"], nodes=[oracle.OracleNode(**n) for n in fa["nodes"]])


def oracle_render(renderer, n_threads: int = 8, **kw) -> np.ndarray:
    return oracle.render(oracle_frame(renderer), n_threads=n_threads, **kw)


def oracle_subset(renderer, scenes, n_threads: int = 8) -> np.ndarray:
    """Oracle frames of the listed scenes only, as a compact [len(scenes), C, H, W] array: the per-scene rows
    (VP, instance matrices and colours) of those scenes are gathered into a small frame, so sampling a huge
    batch costs neither a full-size output nor one oracle call per scene."""
    fa = renderer.frame_arrays()
    scenes = np.asarray(list(scenes), dtype=np.int64)
    nodes = []
    for n in fa["nodes"]:
        n = dict(n)
        if not n["shared"]:
            I = n["instances_per_scene"]
            rows = (scenes[:, None] * I + np.arange(I)[None, :]).reshape(-1)
            n["mats"], n["cols"] = n["mats"][rows], n["cols"][rows]
        nodes.append(oracle.OracleNode(**n))
    fr = oracle.OracleFrame(
        num_scenes=len(scenes), tile_w=fa["tile_w"], tile_h=fa["tile_h"], channels=fa["channels"],
        vp=fa["vp"][scenes], bg=fa["bg"], ambient=fa["ambient"], dir_dir=fa["dir_dir"], dir_col=fa["dir_col"],
        strength=fa["strength"], nodes=nodes)
    return oracle.render(fr, n_threads=n_threads)


def host_threads() -> int:
    import os
    return max(1, min(64, os.cpu_count() or 1))


def cartpole_states(n: int, seed: int = 0, device="cpu") -> torch.Tensor:
    """x ~ U(-2,2), theta ~ U(-30deg,30deg) (reference envs/cartpole/config.py:57-62)."""
    g = torch.Generator().manual_seed(seed)
    s = torch.zeros(n, 4)
    s[:, 0] = torch.rand(n, generator=g) * 4.0 - 2.0
    s[:, 1] = torch.rand(n, generator=g) * 2.0 - 1.0
    s[:, 2] = (torch.rand(n, generator=g) * 60.0 - 30.0) * (np.pi / 180.0)
    s[:, 3] = (torch.rand(n, generator=g) * 30.0 - 15.0) * (np.pi / 180.0)
    return s.to(device)


def many_cubes_renderer(num_scenes=8, instances=16, tile=(64, 64), seed=123, device=None, channels=3,
                        spread=15.0, eye=(0.0, -12.0, 0.0), shared=False, two_sided=False, light=True):
    """Scene shaped like reference demo_many_cubes.py:34-54: random boxes around the default camera
    (positions U(-spread,spread)^3, HPR U(-pi,pi)^3, scale U(0.5,1.8), colours U(0,1)^3)."""
    from pybatchrender_b200 import PBRRenderer
    from pybatchrender_b200 import meshes
    cfg = dict(num_scenes=num_scenes, tile_resolution=tile, num_channels=channels)
    if device is not None:
        cfg["device"] = device
    r = PBRRenderer(cfg)
    model = "models/box"
    if two_sided:
        m = meshes.box()
        m.two_sided = True
        # open box: drop the +z face so that inside faces become visible
        m.idx = m.idx[:8].copy()
        meshes.register_mesh("test/open_box", m)
        model = "test/open_box"
    node = r.add_node(model, instances_per_scene=instances, model_pivot_relative_point=(0.5, 0.5, 0.5),
                      shared_across_scenes=shared)
    rng = np.random.default_rng(seed)
    B = node.buf_instances
    node.set_positions(torch.tensor(rng.uniform(-spread, spread, (B, 3)), dtype=torch.float32), lazy=True)
    node.set_hprs(torch.tensor(rng.uniform(-np.pi, np.pi, (B, 3)), dtype=torch.float32), lazy=True)
    node.set_scales(torch.tensor(rng.uniform(0.5, 1.8, (B, 1)), dtype=torch.float32))
    col = np.concatenate([rng.uniform(0, 1, (B, 3)), np.ones((B, 1))], axis=1)
    node.set_colors(torch.tensor(col, dtype=torch.float32))
    cam = r.add_camera()
    cam.set_positions(torch.tensor(eye, dtype=torch.float32))
    if light:
        r.add_light()
    r.setup_environment()
    return r


def mixed_mesh_renderer(num_scenes=8, boxes=6, spheres=5, tile=(64, 64), seed=7, device=None, channels=3,
                        spread=6.0, eye=(0.0, -12.0, 0.0), segments=10, rings=6, two_sided=False,
                        shared_spheres=False):
    """Flat boxes + smooth-shaded UV spheres (per-vertex normals, non-uniform scales) in one frame:
    the shape of reference demo_mixed_meshes.py (BASELINE config 5) without its model files."""
    from pybatchrender_b200 import PBRRenderer
    from pybatchrender_b200 import meshes
    cfg = dict(num_scenes=num_scenes, tile_resolution=tile, num_channels=channels)
    if device is not None:
        cfg["device"] = device
    r = PBRRenderer(cfg)
    sph = meshes.uv_sphere(1.0, segments, rings)
    if two_sided:
        sph.two_sided = True
        sph.idx = sph.idx[: (2 * sph.idx.shape[0]) // 3].copy()       # open bowl: inside faces show
    name = f"test/sphere_{segments}_{rings}_{int(two_sided)}"
    meshes.register_mesh(name, sph)
    rng = np.random.default_rng(seed)
    nodes = []
    if boxes:
        nodes.append((r.add_node("models/box", instances_per_scene=boxes, model_pivot_relative_point=(0.5, 0.5, 0.5)), 1))
    if spheres:
        nodes.append((r.add_node(name, instances_per_scene=spheres, shared_across_scenes=shared_spheres,
                                 model_scale=(1.0, 0.7, 1.3)), 1))
    for node, k in nodes:
        B = node.buf_instances
        node.set_positions(torch.tensor(rng.uniform(-spread, spread, (B, 3)), dtype=torch.float32), lazy=True)
        node.set_hprs(torch.tensor(rng.uniform(-np.pi, np.pi, (B, 3)), dtype=torch.float32), lazy=True)
        node.set_scales(torch.tensor(rng.uniform(0.6, 2.2, (B, k)), dtype=torch.float32))
        col = np.concatenate([rng.uniform(0, 1, (B, 3)), np.ones((B, 1))], axis=1)
        node.set_colors(torch.tensor(col, dtype=torch.float32))
    cam = r.add_camera()
    cam.set_positions(torch.tensor(eye, dtype=torch.float32))
    r.add_light()
    r.setup_environment()
    return r


def config5_renderer(num_scenes=16384, per_node=16, tile=(256, 256), seed=123, device=None, spread=15.0):
    """BASELINE config 5 ("mixed-mesh"): four per-scene nodes x ``per_node`` instances --
    procedural box, models/cone.egg, models/cylinder/scene.gltf, UV sphere (stand-in for the
    un-vendored models/smiley) -- camera eye (0,-40,10) looking at the origin, instances drawn like
    config 3 (positions U(-spread,spread)^3, HPR U(-pi,pi)^3, scales U(0.5,1.8), colours U(0,1)^3)."""
    from pybatchrender_b200 import PBRRenderer
    cfg = dict(num_scenes=num_scenes, tile_resolution=tile)
    if device is not None:
        cfg["device"] = device
    r = PBRRenderer(cfg)
    nodes = [
        r.add_node("models/box", instances_per_scene=per_node, model_pivot_relative_point=(0.5, 0.5, 0.5)),
        r.add_node("models/cone.egg", instances_per_scene=per_node, model_pivot_relative_point=(0.5, 0.5, 0.5)),
        r.add_node("models/cylinder/scene.gltf", instances_per_scene=per_node, model_scale=2.0,
                   model_scale_units="absolute", model_pivot_relative_point=(0.5, 0.5, 0.5)),
        r.add_node("models/smiley", instances_per_scene=per_node),
    ]
    g = torch.Generator().manual_seed(seed)
    for node in nodes:
        B = node.buf_instances
        node.set_positions((torch.rand(B, 3, generator=g) * 2 - 1) * spread, lazy=True)
        node.set_hprs((torch.rand(B, 3, generator=g) * 2 - 1) * float(np.pi), lazy=True)
        node.set_scales(torch.rand(B, 1, generator=g) * 1.3 + 0.5)
        node.set_colors(torch.cat([torch.rand(B, 3, generator=g), torch.ones(B, 1)], 1))
    cam = r.add_camera()
    cam.set_positions(torch.tensor([0.0, -40.0, 10.0]))
    cam.look_at(torch.tensor([0.0, 0.0, 0.0]))
    r.add_light()
    r.setup_environment()
    return r
