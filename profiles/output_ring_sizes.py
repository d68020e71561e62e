import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
from pybatchrender_b200.envs.cartpole import CartPoleRenderer
import bench
N = 4096
r = CartPoleRenderer(dict(num_scenes=N, tile_resolution=(64, 64), device='cuda'))
st = [bench.cartpole_state(N, i, torch).cuda() for i in range(4)]
for nout in (1, 2, 3, 4, 8):
    outs = [torch.empty((N, 3, 64, 64), dtype=torch.uint8, device='cuda') for _ in range(nout)]
    for i in range(5): r.step(st[i % 4], out=outs[i % nout])
    torch.cuda.synchronize()
    res = []
    for mode in ('raster', 'step'):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(16):
                if mode == 'raster': r.render(out=outs[i % nout])
                else: r.step(st[i % 4], out=outs[i % nout])
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(40): g.replay()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 640 * 1000)
        res.append(best)
    print(f"RING output ring of {nout} x 50.3 MB: raster {res[0]:.2f} us, step {res[1]:.2f} us", flush=True)
