"""DRAM traffic of the small-scene kernel over a RANGE of consecutive launches (not one isolated launch, whose
50 MB of pixel writes are still dirty in the 126 MB L2 when it ends): 16 steps cycling the 4-buffer output ring.

    ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        python profiles/traffic_range.py

Expected from the algorithm: 16 x 50.3 MB of writes (minus what is still in L2 at the end of the range), reads ~ 0."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from pybatchrender_b200 import workloads
from pybatchrender_b200.envs.cartpole import CartPoleRenderer

N, STEPS, RING = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 16, 4
r = CartPoleRenderer(dict(num_scenes=N, tile_resolution=(64, 64), device='cuda'))
st = [workloads.cartpole_state(N, i).cuda() for i in range(16)]
outs = [torch.empty((N, 3, 64, 64), dtype=torch.uint8, device='cuda') for _ in range(RING)]
for i in range(8):
    r.step(st[i % 16], out=outs[i % RING])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for i in range(STEPS):
    r.step(st[i % 16], out=outs[i % RING])
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("range done:", STEPS, "launches,", STEPS * N * 12512, "algorithmic bytes")
