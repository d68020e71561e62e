import os, sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
which = sys.argv[1]
if which == "steering":
    import pybatchrender_b200 as pbr
    env = pbr.envs.make("Steering-v0", num_scenes=1024, device="cuda")
    env.reset(); r = env._renderer
elif which == "config5":
    from util import config5_renderer
    r = config5_renderer(num_scenes=1024, device="cuda")
elif which == "config3":
    from util import many_cubes_renderer
    r = many_cubes_renderer(num_scenes=1024, instances=256, tile=(128, 128), device="cuda")
for _ in range(3): r.render()
torch.cuda.synchronize()
