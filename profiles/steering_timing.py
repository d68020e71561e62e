"""Steering-v0 on one GPU: device time of a frame and wall-clock env.step throughput.

python profiles/steering_timing.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time, json
import pybatchrender_b200 as pbr
for n in (1024, 4096):
    env = pbr.envs.make("Steering-v0", num_scenes=n, device="cuda")
    td = env.reset()
    r = env._renderer
    for _ in range(3): r.render()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): r.render()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    t=time.time()
    for _ in range(20):
        td["action"] = env.action_spec.rand(); td = env.step(td)["next"]
    torch.cuda.synchronize()
    print(json.dumps({"steering_scenes": n, "render_ms": round(ms,3), "render_fps": round(n/ms*1e3), "env_step_fps": round(20*n/(time.time()-t))}))
