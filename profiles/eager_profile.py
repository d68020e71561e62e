"""Where the host time of an eager `renderer.step` goes (cProfile over 3000 CartPole steps at 4096 scenes)."""
import cProfile, pstats, sys, time
sys.path.insert(0, '/root/repo')
import torch
from pybatchrender_b200 import workloads
from pybatchrender_b200.envs.cartpole import CartPoleRenderer
N = 4096
r = CartPoleRenderer(dict(num_scenes=N, tile_resolution=(64, 64), device='cuda'))
st = [workloads.cartpole_state(N, i).cuda() for i in range(16)]
outs = [torch.empty((N, 3, 64, 64), dtype=torch.uint8, device='cuda') for _ in range(4)]
for i in range(50): r.step(st[i % 16], out=outs[i % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(3000): r.step(st[i % 16], out=outs[i % 4])
dt = time.perf_counter() - t0
torch.cuda.synchronize()
print(f"host time per step: {dt / 3000 * 1e6:.2f} us")
pr = cProfile.Profile(); pr.enable()
for i in range(3000): r.step(st[i % 16], out=outs[i % 4])
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
