#!/usr/bin/env python
"""Generate tests/golden/host_math.npz by running the REFERENCE's own host classes under stubs.

Authoring container only (needs /root/reference; see tests/ref_stubs.py).  What is captured is the
exact byte content the reference uploads into its buffer textures -- the input contract of the
shaders -- for the CartPole scene, plus a few camera / rotation known answers:

    python tests/golden/make_host_golden.py
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_stubs  # noqa: E402


def f32(tex):
    return np.frombuffer(tex.ram, dtype=np.float32).copy()


def cartpole_buffers(R, num_scenes, state):
    cfg = R["PBRConfig"](num_scenes=num_scenes, tile_resolution=(64, 64), device="cpu")
    base = ref_stubs.FakeBase(*cfg.window_resolution)
    mk = lambda shared: R["PBRNode"](base, "models/box", num_scenes=num_scenes, instances_per_scene=1,  # noqa: E731
                                     shared_across_scenes=shared)
    rail, cart, pole = mk(True), mk(False), mk(False)
    N = num_scenes
    cart_pos = torch.zeros((N, 1, 3))
    pole_pos = cart_pos.clone() + torch.tensor([0, (0.8 + 0.1) * 0.5, 0])
    rail.set_positions(torch.zeros((1, 1, 3)))
    cart.set_positions(cart_pos)
    pole.set_positions(pole_pos)
    start, end = torch.tensor([0.6, 0.8, 1.0, 1.0]), torch.tensor([1.0, 0.6, 0.8, 1.0])
    t = torch.linspace(0.0, 1.0, steps=N)
    rail.set_colors(torch.ones((1, 1, 4)) * torch.tensor([0.2, 0.2, 0.2, 1.0]))
    cart.set_colors(start.unsqueeze(0) + (end - start).unsqueeze(0) * t.unsqueeze(1))
    pole.set_colors(torch.ones((N, 1, 4)) * torch.tensor([1.0, 0.7, 0.2, 1.0]))
    hpr = torch.zeros((N, 1, 3))
    hpr[:, :, 1] = math.pi * 0.5
    pole.set_hprs(hpr)
    cam = R["PBRCam"](base, num_scenes=N, cols=cfg.tiles[0], rows=cfg.tiles[1])
    cam.set_positions(torch.tensor([5, 5, 2], dtype=torch.float32))
    cam.look_at(torch.tensor([0, 0, 0], dtype=torch.float32))
    # one _step as in reference envs/cartpole/renderer.py:98-138
    x, theta = state[:, 0:1], state[:, 2:3]
    cart_pos[:, :, 0:1] = x.unsqueeze(-1)[:, :, 0:1] if x.ndim == 2 else x
    cart_pos[:, :, 0] = x
    cart.set_positions(cart_pos)
    pole_pos[:, :, 0] = x
    pole.set_positions(pole_pos, lazy=True)
    hpr[:, :, 1] = theta
    pole.set_hprs(hpr)
    return dict(tiles=np.array(cfg.tiles), window=np.array(cfg.window_resolution),
                viewbuf=f32(cam.viewbuf).reshape(N, 16), tilebuf=f32(cam.tilebuf).reshape(N, 4),
                rail_mat=f32(rail.matbuf).reshape(-1, 16), cart_mat=f32(cart.matbuf).reshape(-1, 16),
                pole_mat=f32(pole.matbuf).reshape(-1, 16), rail_col=f32(rail.colbuf).reshape(-1, 4),
                cart_col=f32(cart.colbuf).reshape(-1, 4), pole_col=f32(pole.colbuf).reshape(-1, 4),
                P=cam.P_k44.numpy().copy(), V0=cam.V_k44[0].numpy().copy(), VP0=cam.VP_k44[0].numpy().copy())


def main():
    R = ref_stubs.install()
    out = {}
    g = torch.Generator().manual_seed(7)
    for n in (4, 37, 4098):
        st = torch.zeros(n, 4)
        st[:, 0] = torch.rand(n, generator=g) * 4 - 2
        st[:, 2] = torch.rand(n, generator=g) - 0.5
        bufs = cartpole_buffers(R, n, st)
        out[f"state_{n}"] = st.numpy()
        keep = slice(None) if n <= 64 else slice(0, 64)
        for k, v in bufs.items():
            out[f"{k}_{n}"] = v[keep] if (v.ndim == 2 and v.shape[0] == n) else v
    # rotation matrices from HPR (shader_context.py:47-84)
    hpr = torch.tensor(np.random.default_rng(3).uniform(-math.pi, math.pi, (32, 3)), dtype=torch.float32)
    out["hpr"] = hpr.numpy()
    out["rot_from_hpr"] = R["PBRShaderContext"]._rotation_mats_from_hpr(hpr).numpy()
    # per-scene camera paths (camera.py:268-331)
    cfg = R["PBRConfig"](num_scenes=6, tile_resolution=(32, 48), device="cpu")
    base = ref_stubs.FakeBase(*cfg.window_resolution)
    cam = R["PBRCam"](base, num_scenes=6, cols=cfg.tiles[0], rows=cfg.tiles[1], fov_y_deg=40.0, z_near=0.1, z_far=50.0)
    rng = np.random.default_rng(5)
    eye = torch.tensor(rng.uniform(-5, 5, (6, 3)), dtype=torch.float32)
    tgt = torch.tensor(rng.uniform(-1, 1, (6, 3)), dtype=torch.float32)
    cam.set_positions_and_lookat(eye, tgt)
    out["cam6_eye"], out["cam6_tgt"] = eye.numpy(), tgt.numpy()
    out["cam6_viewbuf_lookat"] = f32(cam.viewbuf).reshape(6, 16)
    hp = torch.tensor(rng.uniform(-1, 1, (6, 3)), dtype=torch.float32)
    cam.set_hprs(hp)
    out["cam6_hpr"] = hp.numpy()
    out["cam6_viewbuf_hpr"] = f32(cam.viewbuf).reshape(6, 16)
    out["cam6_tiles"] = np.array(cfg.tiles)
    # _rearrange_img (renderer.py:352-363)
    cfg2 = R["PBRConfig"](num_scenes=5, tile_resolution=(4, 2), device="cpu")
    fake = type("S", (), {"cfg": cfg2})()
    img = torch.arange(cfg2.window_resolution[1] * cfg2.window_resolution[0] * 3, dtype=torch.int64)
    img = (img % 251).to(torch.uint8).reshape(cfg2.window_resolution[1], cfg2.window_resolution[0], 3)
    out["rearr_in"] = img.numpy()
    out["rearr_out"] = R["rearrange"](fake, img).contiguous().numpy()
    out["rearr_tiles"] = np.array(cfg2.tiles)
    # config arithmetic (config.py:61-126)
    rows = []
    for n in (1, 2, 3, 4, 5, 16, 17, 64, 100, 4096, 4098, 65536):
        c = R["PBRConfig"](num_scenes=n, device="cpu")
        rows.append([n, c.tiles[0], c.tiles[1], c.window_resolution[0], c.window_resolution[1], c.batch_inner_dim])
    out["config_rows"] = np.array(rows)
    path = os.path.join(HERE, "host_math.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
