import torch, sys
sys.path.insert(0,'/root/repo')
from pybatchrender_b200.envs.cartpole import CartPoleRenderer
import bench
N=4096
r=CartPoleRenderer(dict(num_scenes=N,tile_resolution=(64,64),device='cuda'))
st=[bench.cartpole_state(N,i,torch).cuda() for i in range(16)]
outs=[torch.empty((N,3,64,64),dtype=torch.uint8,device='cuda') for _ in range(4)]
for dbg in (0,1,2,3):
    r.render_flags = dbg<<8
    for i in range(5): r.step(st[i%16],out=outs[i%4])
    torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16): r.render(out=outs[i%4])
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    print('debug',dbg,'us/frame',e0.elapsed_time(e1)/800*1000)
# static layer off
r.render_flags=0; r.static_layer=False
for i in range(5): r.step(st[i%16],out=outs[i%4])
torch.cuda.synchronize()
g=torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(16): r.render(out=outs[i%4])
for _ in range(3): g.replay()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): g.replay()
e1.record(); torch.cuda.synchronize()
print('no static layer us/frame',e0.elapsed_time(e1)/800*1000)
for dbg in (1, 2):
    r.render_flags = dbg << 8
    for i in range(5): r.step(st[i%16],out=outs[i%4])
    torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16): r.render(out=outs[i%4])
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    print('no static layer, debug',dbg,'us/frame',e0.elapsed_time(e1)/800*1000)
# pure device fill for reference
x = outs[0]
for _ in range(3): x.fill_(7)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200): outs[i%4].fill_(7)
e1.record(); torch.cuda.synchronize()
print('torch fill_ of one 50 MB frame us', e0.elapsed_time(e1)/200*1000)
