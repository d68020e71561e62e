"""Device-time of the other BASELINE.json configs (parity cases, not bench lines) -- one GPU.

python profiles/other_configs.py   (prints one line per config: ms per frame, scene-frames/s, HBM roofline fraction)
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pybatchrender_b200.envs.cartpole import CartPoleRenderer  # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from util import config5_renderer, many_cubes_renderer  # noqa: E402

PEAK = bench.measured_peak()[0]


def time_render(r, reps):
    out = r.render()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r.render(out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, n, ms, bytes_per_scene):
    fps = n / (ms * 1e-3)
    print(json.dumps({"config": name, "scenes": n, "ms_per_frame": round(ms, 4), "scene_frames_per_s": round(fps),
                      "algorithmic_bytes_per_scene": bytes_per_scene,
                      "hbm_roofline_frac": round(fps * bytes_per_scene / (PEAK * 1e9), 4)}))


def main():
    n = 65536
    r = CartPoleRenderer(dict(num_scenes=n, tile_resolution=(84, 84), device="cuda"))
    r._step(bench.cartpole_state(n, 0, torch).cuda())
    report("config4: CartPole 65536 x 84^2 (1 GPU)", n, time_render(r, 20), 3 * 84 * 84 + 64 + 160)
    del r
    torch.cuda.empty_cache()
    n = 1024
    r = many_cubes_renderer(num_scenes=n, instances=256, tile=(128, 128), device="cuda")
    report("config3: many cubes 1024 x 256 boxes, 128^2", n, time_render(r, 5), 3 * 128 * 128 + 64 + 256 * 80)
    del r
    torch.cuda.empty_cache()
    n = 16384
    r = config5_renderer(num_scenes=n, device="cuda")
    report("config5: mixed meshes 16384 x 64 instances, 256^2 (1 GPU)", n, time_render(r, 2), 3 * 256 * 256 + 64 + 64 * 80)


if __name__ == "__main__":
    main()
