#!/usr/bin/env python
"""bench.py -- scene-frames/s of the CartPole 4096 x 64^2 render-to-tensor hot path.

Contract (see DESIGN.md section "Measurement"):

    python bench.py --gpus N --steps K --warmup W [--impl reference]

* a *step* = ``renderer.step(state)``: pose kernel + raster kernel for one batch of 4096 CartPole
  scenes (BASELINE.json configs[1]); per-GPU work is fixed, scenes are sharded with no collective
  (``scaling: weak``); under torchrun every rank renders its own 4096 scenes.
* ``value``: whole-job scene-frames/s with the state ring already resident in HBM, K steps replayed
  from CUDA graphs, timed with CUDA events on the launching stream, max over ranks.
* ``e2e``: same metric through the public API with HOST buffers: pinned state -> H2D -> step ->
  D2H of the uint8 frames into pinned memory, every step, copies inside the timed region.  Two pinned
  frame buffers: the D2H of frame i runs beside the H2D + render of frame i+1 and the host waits for
  every frame (``value_sync_every_step``: the same loop with a stream synchronize after every frame).
* ``roofline``: the raster kernel alone (same inputs), algorithmic bytes / average launch time
  against MEASURED_PEAKS.json's HBM copy bandwidth.
* ``cpu_baseline``: the CPU oracle (``oracle/``, a port of the reference pipeline) timed on the
  host cores on a bounded sample of the same workload (rank 0, N=1 only).
* ``--impl reference``: the reference's own pipeline cannot run (Panda3D/OpenGL absent), so this arm
  times the oracle port on all host cores, on the same config / metric.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SCENES_PER_GPU = 4096
TILE = (64, 64)
ALGO_BYTES_PER_SCENE = 3 * 64 * 64 + 64 + 2 * (64 + 16)      # 12,512 B
STATE_RING = 16
OUT_RING = 4            # 4 x 50.3 MB of output > 126 MB L2, so pixel writes cannot stay cached
NCU_DRAM_BYTES_PER_LAUNCH = 1003008 + 1779712      # ncu --set full, profiles/r01q_raster_warp_ncu.txt


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer loop (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather", action="store_true",
                    help="N>1: also time the optional NCCL gather of all frames onto rank 0 (reported separately)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs while the timed region executes
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def cartpole_state(n, seed, torch):
    """x ~ U(-2,2), theta ~ U(-30deg,30deg): reference envs/cartpole/config.py:57-62 reset ranges."""
    g = torch.Generator().manual_seed(seed)
    s = torch.zeros(n, 4)
    s[:, 0] = torch.rand(n, generator=g) * 4.0 - 2.0
    s[:, 1] = torch.rand(n, generator=g) * 2.0 - 1.0
    s[:, 2] = (torch.rand(n, generator=g) * 60.0 - 30.0) * (3.141592653589793 / 180.0)
    s[:, 3] = (torch.rand(n, generator=g) * 30.0 - 15.0) * (3.141592653589793 / 180.0)
    return s


def oracle_frame_of(renderer):
    import oracle
    fa = renderer.frame_arrays()
    return oracle.OracleFrame(num_scenes=fa["num_scenes"], tile_w=fa["tile_w"], tile_h=fa["tile_h"],
                              channels=fa["channels"], vp=fa["vp"], bg=fa["bg"], ambient=fa["ambient"],
                              dir_dir=fa["dir_dir"], dir_col=fa["dir_col"], strength=fa["strength"],
                              nodes=[oracle.OracleNode(**nd) for nd in fa["nodes"]])


def time_cpu_oracle(target_s: float, cores: int, seed0: int = 1000):
    """Times the oracle port: whole CartPole 4096x64^2 steps (host pose update + raster) until about
    ``target_s`` seconds have elapsed.  Returns (scene_frames_per_s, steps, seconds)."""
    import numpy as np
    import torch

    import oracle
    from pybatchrender_b200.envs.cartpole import CartPoleRenderer
    r = CartPoleRenderer(dict(num_scenes=SCENES_PER_GPU, tile_resolution=TILE, device="cpu"))
    out = np.zeros((SCENES_PER_GPU, 3, TILE[1], TILE[0]), np.uint8)
    states = [cartpole_state(SCENES_PER_GPU, seed0 + i, torch) for i in range(4)]
    r._step(states[0])
    oracle.render(oracle_frame_of(r), n_threads=cores, out=out)     # warm-up (page faults, lib load)
    t0 = time.perf_counter()
    steps = 0
    while True:
        r._step(states[steps % 4])
        oracle.render(oracle_frame_of(r), n_threads=cores, out=out)
        steps += 1
        dt = time.perf_counter() - t0
        if dt >= target_s or steps >= 4096:
            break
    return SCENES_PER_GPU * steps / dt, steps, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    K, W = max(1, args.steps), max(0, args.warmup)
    # bounded sample: each "step" here is one full 4096-scene frame on the CPU; cap the wall time
    import numpy as np
    import torch

    import oracle
    from pybatchrender_b200.envs.cartpole import CartPoleRenderer
    r = CartPoleRenderer(dict(num_scenes=SCENES_PER_GPU, tile_resolution=TILE, device="cpu"))
    out = np.zeros((SCENES_PER_GPU, 3, TILE[1], TILE[0]), np.uint8)
    states = [cartpole_state(SCENES_PER_GPU, 1000 + i, torch) for i in range(4)]
    budget_s = 120.0
    t_w = time.perf_counter()
    for i in range(min(W, 3)):
        r._step(states[i % 4])
        oracle.render(oracle_frame_of(r), n_threads=cores, out=out)
    per = (time.perf_counter() - t_w) / max(1, min(W, 3)) if W else 0.2
    k_eff = max(1, min(K, int(budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    for i in range(k_eff):
        r._step(states[i % 4])
        oracle.render(oracle_frame_of(r), n_threads=cores, out=out)
    dt = time.perf_counter() - t0
    value = SCENES_PER_GPU * k_eff / dt
    line = {
        "impl": "reference",
        "metric": "scene-frames/sec to torch tensor (CartPole 4096x64^2)",
        "value": value, "unit": "scene-frames/s", "n_gpus": args.gpus, "steps": k_eff, "warmup": min(W, 3),
        "ms_per_step": 1e3 * dt / k_eff, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+i64 (u8 out)", "data": "synthetic",
        "config": {"workload": "CartPole-v0 num_scenes=4096 tile 64x64 (BASELINE configs[1])",
                   "note": "reference pipeline (Panda3D + OpenGL) cannot run in this image; this arm times "
                           "the CPU oracle port of it on the host cores"},
        "cpu_baseline": {"value": value, "unit": "scene-frames/s", "cores": cores, "kind": "port",
                         "sample": f"{k_eff} full frames of 4096 scenes ({dt:.1f} s)"},
        "e2e": {"value": value, "unit": "scene-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 renderer has no CPU fallback "
                         "(use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from pybatchrender_b200.config import grid_for
    from pybatchrender_b200.envs.cartpole import CartPoleRenderer

    K, W = max(1, args.steps), max(3, args.warmup)
    N = SCENES_PER_GPU
    r = CartPoleRenderer(dict(num_scenes=N, tile_resolution=TILE, device="cuda"))
    states = [cartpole_state(N, 1000 * rank + i, torch).to(dev) for i in range(STATE_RING)]
    outs = [torch.empty((N, 3, TILE[1], TILE[0]), dtype=torch.uint8, device=dev) for _ in range(OUT_RING)]
    launches_per_step = 2          # pose kernel + raster kernel

    def step(i):
        return r.step(states[i % STATE_RING], out=outs[i % OUT_RING])

    # ---- graphs: STATE_RING steps per replay (+ a tail graph so that exactly K steps are timed)
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            step(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()

    def capture(n_steps, fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n_steps):
                fn(i)
        return g

    g_full = capture(STATE_RING, step)
    tail = K % STATE_RING
    g_tail = capture(tail, step) if tail else None

    def run_steps(k):
        for _ in range(k // STATE_RING):
            g_full.replay()
        if k % STATE_RING:
            if k % STATE_RING == tail and g_tail is not None:
                g_tail.replay()
            else:
                for i in range(k % STATE_RING):
                    step(i)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    run_steps(W)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)

    # ---- roofline leg: the raster kernel alone
    def raster_only(i):
        r.render(out=outs[i % OUT_RING])
    g_r = capture(STATE_RING, raster_only)
    for _ in range(3):
        g_r.replay()
    torch.cuda.synchronize()
    reps = max(1, K // STATE_RING)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for _ in range(reps):
        g_r.replay()
    r1.record()
    torch.cuda.synchronize()
    raster_ms = r0.elapsed_time(r1) / (reps * STATE_RING)

    # ---- e2e leg: host buffers, copies inside the timed region, result on the host every step
    e2e_steps = args.e2e_steps or max(8, min(K, 200))
    h_states = [cartpole_state(N, 5000 + 1000 * rank + i, torch).pin_memory() for i in range(4)]
    h_out = torch.empty((N, 3, TILE[1], TILE[0]), dtype=torch.uint8).pin_memory()
    d_state = torch.empty((N, 4), dtype=torch.float32, device=dev)

    def e2e_step(i):
        d_state.copy_(h_states[i % 4], non_blocking=True)
        px = r.step(d_state, out=outs[i % OUT_RING])
        h_out.copy_(px, non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller owns the frame on the host now

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_sync_s = time.perf_counter() - t0

    # the same loop the way a consumer of frames would write it: two pinned host buffers, the D2H copy of
    # frame i runs on a copy stream beside the H2D + render of frame i+1; the host owns frame i when its
    # copy event has completed (waited for before that buffer is reused, and for all frames at the end).
    # Every step still moves its own state in and its own frame out inside the timed region.
    h_outs = [h_out, torch.empty_like(h_out).pin_memory()]
    copy_stream = torch.cuda.Stream(dev)
    landed = [None, None]

    def e2e_step_pipelined(i):
        d_state.copy_(h_states[i % 4], non_blocking=True)
        px = r.step(d_state, out=outs[i % OUT_RING])
        rendered = torch.cuda.Event()
        rendered.record()
        if landed[i % 2] is not None:
            landed[i % 2].synchronize()                 # frame i-2 is on the host: its buffer is free again
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(rendered)
            h_outs[i % 2].copy_(px, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        landed[i % 2] = ev

    def e2e_drain():
        for ev in landed:
            if ev is not None:
                ev.synchronize()

    for i in range(4):
        e2e_step_pipelined(i)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step_pipelined(i)
    e2e_drain()
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- optional: frames of all ranks onto rank 0 (NCCL gather over NVLink); never part of step()
    gather_ms = None
    if distributed and args.gather:
        from pybatchrender_b200.dist import gather_frames
        for _ in range(2):
            gather_frames(outs[0], dst=0)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(10):
            gather_frames(outs[0], dst=0)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1) / 10

    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    t_ms = torch.tensor([ms, e2e_s * 1e3, raster_ms, e2e_sync_s * 1e3], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, raster_ms_max, e2e_sync_ms_max = (float(x) for x in t_ms.tolist())

    if rank == 0:
        peak, peak_src = measured_peak()
        total_scenes = N * world
        value = total_scenes * K / (ms_max * 1e-3)
        achieved = ALGO_BYTES_PER_SCENE * N / (raster_ms_max * 1e-3) / 1e9
        line = {
            "metric": "scene-frames/sec to torch tensor (CartPole 4096x64^2)",
            "value": value, "unit": "scene-frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+i32 (u8 out)", "data": "synthetic",
            "config": {"workload": "CartPole-v0 num_scenes=4096 per GPU, tile 64x64 (BASELINE configs[1])",
                       "scenes_per_gpu": N, "tile": list(TILE), "channels": 3,
                       "parallelism": f"scene-sharded x{world}, no collective",
                       "l2": f"output ring of {OUT_RING} x {N * 3 * TILE[0] * TILE[1] / 1e6:.1f} MB (> 126 MB L2); "
                             f"state ring of {STATE_RING}",
                       "launch": f"CUDA graphs of {STATE_RING} steps; raster kernel as a programmatic dependent of the pose kernel"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_note": "dram__bytes_read+write of one isolated launch under ncu (profiles/r01q_*): "
                                         "the 50 MB of pixel writes are still dirty in the 126 MB L2 when the kernel "
                                         "ends, so DRAM sees them later; no re-reads",
                         "kernel": "raster_warp_kernel<14, true>",
                         "kernel_ms": raster_ms_max, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SCENE * N,
                         "peak_source": peak_src},
            "e2e": {"value": total_scenes * e2e_steps / (e2e_ms_max * 1e-3), "unit": "scene-frames/s",
                    "h2d_bytes_per_step": N * 4 * 4, "d2h_bytes_per_step": N * 3 * TILE[0] * TILE[1],
                    "steps": e2e_steps,
                    "how": "pinned host state -> H2D -> step -> D2H of the frames every step; two pinned frame buffers, "
                           "the D2H of frame i overlaps the H2D + render of frame i+1, the host waits for every frame",
                    "value_sync_every_step": total_scenes * e2e_steps / (e2e_sync_ms_max * 1e-3)},
            "gpu_launches": launches_per_step * K,
            "clocks": sampler.summary(),
            "gather": None if gather_ms is None else {
                "ms": gather_ms, "bytes_into_rank0": (world - 1) * N * 3 * TILE[0] * TILE[1],
                "GBps_into_rank0": (world - 1) * N * 3 * TILE[0] * TILE[1] / (gather_ms * 1e-3) / 1e9,
                "note": "optional collective, timed separately, not part of value / e2e"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, steps, dt = time_cpu_oracle(10.0, cores)
            line["cpu_baseline"] = {"value": v, "unit": "scene-frames/s", "cores": cores, "kind": "port",
                                    "sample": f"{steps} full frames of 4096 scenes in {dt:.1f} s"}
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
