import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
from pybatchrender_b200 import PBRRenderer
from util import oracle_render
def scene(seed, scale):
    r = PBRRenderer(dict(num_scenes=6, tile_resolution=(64, 48), device="cuda"))
    rng = np.random.default_rng(seed)
    node = r.add_node("models/smiley", instances_per_scene=150)
    B = node.buf_instances
    node.set_positions(torch.tensor(rng.uniform(-9, 9, (B, 3)), dtype=torch.float32), lazy=True)
    node.set_hprs(torch.tensor(rng.uniform(-np.pi, np.pi, (B, 3)), dtype=torch.float32), lazy=True)
    node.set_scales(torch.tensor(rng.uniform(0.3, 2.0, (B, 1)) * scale, dtype=torch.float32))
    node.set_colors(torch.tensor(np.concatenate([rng.uniform(0.2, 1, (B, 3)), np.ones((B, 1))], 1), dtype=torch.float32))
    cam = r.add_camera(); cam.set_positions(torch.tensor([0.0, -12.0, 0.0]))
    r.add_light(); r.setup_environment()
    return r
for seed, scale in [(0,0.02),(1,0.05),(2,0.15)]:
    r = scene(seed, scale)
    ref = oracle_render(r)
    for flags in (0, 2):
        r.render_flags = flags
        got = r.render().cpu().numpy()
        d = (got != ref).any(1)
        print(seed, scale, "flags", flags, "mismatch px", int(d.sum()), "nonbg ref", int((ref!=0).any(1).sum()), "nonbg got", int((got!=0).any(1).sum()))
        if d.sum():
            s,y,x = np.argwhere(d)[0]; print("  first", s,y,x, got[s,:,y,x], ref[s,:,y,x])
