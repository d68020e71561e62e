"""Large-scene path on BASELINE config 3 (many-cubes 1024 x 256 boxes, 128^2) or 5 (mixed-mesh x 64 instances, 256^2;
N scenes): a few frames, for ncu launch lists / full captures of cull_kernel, geom_kernel and raster_staged_kernel.

    python profiles/staged_workloads.py 3            # config 3, 1024 scenes
    python profiles/staged_workloads.py 5 2048       # config 5, 2048 scenes
"""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from pybatchrender_b200 import workloads

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else (1024 if cfg == 3 else 2048)
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
r = workloads.many_cubes(num_scenes=n, device='cuda') if cfg == 3 else workloads.mixed_meshes(num_scenes=n, device='cuda')
out = r.render()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(frames):
    r.render(out=out)
e1.record()
torch.cuda.synchronize()
print(f"config {cfg}: {n} scenes, {e0.elapsed_time(e1) / frames:.3f} ms per frame")
