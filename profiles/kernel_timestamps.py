"""Time stamps inside raster_warp_kernel (%globaltimer per warp at phase boundaries).

Build a timing library first (it dumps the stamps into the head of the overflow pool and exports pbr_debug_pool):
    cd pybatchrender_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -Xcompiler -fPIC \
        -DPBR_W_TIMING -shared -o /tmp/timing.so pbr_b200.cu -lcudart
    python profiles/kernel_timestamps.py /tmp/timing.so
"""
import os, sys, ctypes
os.environ['PBR_B200_LIB'] = os.path.abspath(sys.argv[1])      # a build with -DPBR_W_TIMING
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from pybatchrender_b200.envs.cartpole import CartPoleRenderer
from pybatchrender_b200 import _native
from pybatchrender_b200 import workloads
N = 4096
r = CartPoleRenderer(dict(num_scenes=N, tile_resolution=(64, 64), device='cuda'))
st = [workloads.cartpole_state(N, i).cuda() for i in range(4)]
outs = [torch.empty((N, 3, 64, 64), dtype=torch.uint8, device='cuda') for _ in range(5)]
for i in range(20): r.step(st[i % 4], out=outs[i % 5])
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ['PBR_B200_LIB'])
nc = (N + 13) // 14
buf = np.zeros((nc, 16, 16), dtype=np.uint64)
rc = lib.pbr_debug_pool(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(buf.nbytes))
assert rc == 0
t = buf[:, :, :7].astype(np.int64)
tf = buf[:, :, 8:14].astype(np.int64)
t0 = t[:, :, 0].min()
rel = (t - t0) / 1000.0
names = ['entry', 'after A (vertices)', 'after setup', 'before tma wait/queue', 'after tma wait', 'after barrier', 'sweep done']
scene_w = slice(0, 14)
for k, nm in enumerate(names):
    w = scene_w if k in (1, 2) else slice(0, 16)
    v = rel[:, w, k]
    print(f"{nm:26s} mean {v.mean():6.2f}  p10 {np.percentile(v,10):6.2f}  p50 {np.percentile(v,50):6.2f}  p90 {np.percentile(v,90):6.2f}  max {v.max():6.2f} us")
# per-SM end time
sm = (buf[:, 0, 7] & 0xffffffff).astype(int)
wslot = (buf[:, :, 7] >> 32).astype(int)
end = rel[:, :, 6].max(1)
per_sm = {}
for s_, e in zip(sm, end): per_sm[s_] = max(per_sm.get(s_, 0), e)
e = np.array(list(per_sm.values()))
print('per-SM end: mean %.2f min %.2f max %.2f' % (e.mean(), e.min(), e.max()), 'SMs', len(e))
# barrier skew within CTA: after barrier - own arrival
skew = rel[:, :14, 5] - rel[:, :14, 3]
print('barrier wait per scene warp: mean %.2f p90 %.2f max %.2f' % (skew.mean(), np.percentile(skew, 90), skew.max()))
sw = rel[:, :, 6] - rel[:, :, 5]
print('sweep duration per warp: mean %.2f p10 %.2f p90 %.2f' % (sw.mean(), np.percentile(sw, 10), np.percentile(sw, 90)))
cta_end = rel[:, :, 6].max(1); cta_first = rel[:, :, 6].min(1)
print('sweep end skew within CTA: mean %.2f' % (cta_end - cta_first).mean())
bar = rel[:, :, 5].max(1)
order = np.argsort(-bar)
print('CTAs with barrier release > 8 us:', int((bar > 8).sum()), 'of', nc, '; > 7 us:', int((bar > 7).sum()))
for c in order[:8]:
    a = rel[c, :14, 1]; s2 = rel[c, :14, 2]; q = rel[c, :14, 3]
    print(f" cta {c} sm {sm[c]} barrier {bar[c]:.2f} end {cta_end[c]:.2f}  slowest warp: afterA {a.max():.2f} setup {s2.max():.2f} queue {q.max():.2f} (median queue {np.median(q):.2f})")
late = np.argsort(-cta_end)[:8]
for c in late:
    print(f" late cta {c} sm {sm[c]} barrier {bar[c]:.2f} end {cta_end[c]:.2f}")
print('corr(barrier, end) = %.2f' % np.corrcoef(bar, cta_end)[0, 1])
# how long after the barrier does the CTA run
print('sweep span per CTA (end - barrier): mean %.2f p10 %.2f p90 %.2f max %.2f' % ((cta_end - bar).mean(), np.percentile(cta_end - bar, 10), np.percentile(cta_end - bar, 90), (cta_end - bar).max()))
print('per warp index (mean over CTAs): afterA, setup, queue, sweepdone')
for w in range(16):
    v = rel[:-1, w, :]
    print(f"  warp {w:2d}: A {v[:,1].mean():6.2f} setup {v[:,2].mean():6.2f} queue {v[:,3].mean():6.2f} barrier {v[:,5].mean():6.2f} done {v[:,6].mean():6.2f}")
d = (rel[:-1, :14, 2] - rel[:-1, :14, 1]).ravel()
print('setup - afterA histogram (us):', np.histogram(d, bins=[0,1,2,3,4,5,6,7,8,12])[0].tolist())
d = (rel[:-1, :14, 3] - rel[:-1, :14, 2]).ravel()
print('queue - setup histogram (us):', np.histogram(d, bins=[0,0.5,1,1.5,2,3,4,5,8])[0].tolist())
c = 100
print('cta 100 per warp setup-A:', np.round(rel[c, :14, 2] - rel[c, :14, 1], 2).tolist())
print('cta 100 per warp queue-setup:', np.round(rel[c, :14, 3] - rel[c, :14, 2], 2).tolist())

# ---- do CTAs in high hardware warp slots run their geometry faster than those in low slots?
hi = wslot[:, 0] >= 16
geo = rel[:, :14, 5].max(1) - rel[:, :, 0].min(1)          # entry -> barrier release
swp = rel[:, :, 6].max(1) - rel[:, :14, 5].max(1)
life = rel[:, :, 6].max(1) - rel[:, :, 0].min(1)
for name, sel in (("warp slots 0-15 ", ~hi), ("warp slots 16-31", hi)):
    if sel.any():
        print(f"CTAs in {name}: {int(sel.sum()):4d}  geometry {geo[sel].mean():5.2f} us  sweep {swp[sel].mean():5.2f} us  lifetime {life[sel].mean():5.2f} us")
print("warp slot of warp 0, first CTAs:", wslot[:12, 0].tolist())

# ---- fine-grained phases (stamps 8..13 of each warp), relative to the CTA's own entry
ent = t[:, :, 0].min(1, keepdims=True)
names2 = ['entry barrier', 'M barrier', 'A barrier', 'B barrier', 'S barrier', 'tail: colours + clipped done']
prev = np.zeros(nc - 1)
for k, nm in enumerate(names2):
    v = (tf[:-1, :14, k] - ent[:-1]) / 1000.0             # (the last CTA may have warps without a scene: no stamp)
    m = v.mean(1)
    print(f"  {nm:30s} mean {m.mean():6.2f} us after CTA entry  (+{(m - prev).mean():5.2f})   slowest warp {v.max(1).mean():6.2f}")
    prev = m
q = (t[:-1, :14, 3] - ent[:-1]) / 1000.0
print(f"  {'block list built (stamp 3)':30s} mean {q.mean():6.2f} us after CTA entry  (+{(q.mean(1) - prev).mean():5.2f})   slowest warp {q.max(1).mean():6.2f}")
bq = (t[:, :14, 5] - ent) / 1000.0
print(f"  {'pre-sweep barrier released':30s} mean {bq.mean():6.2f} us after CTA entry")

print("per warp index, mean us after CTA entry: entry-barrier, M, A, B, S barriers")
for w in range(15):
    v = (tf[:, w, :5] - ent) / 1000.0
    print(f"  warp {w:2d}: " + "  ".join(f"{v[:, k].mean():6.2f}" for k in range(5)))

stg = (buf[:, :14, 14].astype(np.int64) - t[:, :, 0].min(1, keepdims=True)) / 1000.0
print(f"stage-in barrier released {stg.mean():.2f} us after CTA entry (p10 {np.percentile(stg,10):.2f}, p90 {np.percentile(stg,90):.2f})")

# ---- are the two CTAs of an SM in step (both in geometry, then both in the sweep) or staggered?
ent1 = rel[:, :, 0].min(1); bar1 = rel[:, :14, 5].max(1); end1 = rel[:, :, 6].max(1)
by_sm = {}
for c in range(nc): by_sm.setdefault(int(sm[c]), []).append(c)
d_ent, ovl = [], []
for s_, cs in by_sm.items():
    if len(cs) != 2: continue
    a, b = cs
    d_ent.append(abs(ent1[a] - ent1[b]))
    # share of a's sweep during which b is in its geometry (and the reverse)
    for x, y in ((a, b), (b, a)):
        lo, hi2 = max(bar1[x], ent1[y]), min(end1[x], bar1[y])
        ovl.append(max(0.0, hi2 - lo) / max(1e-9, end1[x] - bar1[x]))
d_ent = np.array(d_ent)
print(f"entry offset between the two CTAs of an SM: mean {d_ent.mean():.2f} us p10 {np.percentile(d_ent,10):.2f} p50 {np.percentile(d_ent,50):.2f} p90 {np.percentile(d_ent,90):.2f}; "
      f"share of a sweep spent beside the other CTA's geometry (same frame only): {np.mean(ovl):.2f}")
print("entry times of the last frame's CTAs (us after the first): p10 %.2f p50 %.2f p90 %.2f max %.2f" % tuple(np.percentile(ent1, [10, 50, 90, 100])))

# ---- what releases the pre-sweep barrier: the slowest scene warp, or the TMA thread waiting for the CTA's bulk stores?
tw = buf.shape[1] - 1                                    # the TMA thread's warp is the last one
t_tma = (t[:, tw, 4] - ent[:, 0]) / 1000.0               # its bulk stores have completed (stamp 4)
t_tma0 = (t[:, tw, 3] - ent[:, 0]) / 1000.0              # it starts to wait for them (stamp 3)
t_scn = (t[:, :14, 4].max(1) - ent[:, 0]) / 1000.0       # the slowest scene warp has its block list
print(f"TMA thread: starts to wait for the stores {t_tma0.mean():.2f} us after CTA entry, stores complete {t_tma.mean():.2f} "
      f"(p10 {np.percentile(t_tma,10):.2f}, p90 {np.percentile(t_tma,90):.2f}); slowest scene warp ready {t_scn.mean():.2f}; "
      f"CTAs whose barrier waits for the stores: {int((t_tma > t_scn).sum())} of {nc}")
