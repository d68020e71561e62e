import os, sys, json
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from util import config5_renderer, many_cubes_renderer
def t(r, reps=3):
    out = r.render(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r.render(out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
r = config5_renderer(num_scenes=1024, device="cuda"); a = t(r); del r
r = many_cubes_renderer(num_scenes=1024, instances=256, tile=(128,128), device="cuda"); b = t(r, 10)
print(json.dumps({"band_kb": os.environ.get("PBR_B200_BAND_KB","56"), "config5_1024_ms": round(a,3), "config3_ms": round(b,3)}))
