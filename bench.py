#!/usr/bin/env python
"""bench.py -- scene-frames/s of the render-to-tensor hot path on the BASELINE configurations.

Contract (see DESIGN.md section "Measurement"):

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config {2,3,4,5}]

Headline = ``--config 2`` (default): CartPole 4096 x 64^2 per GPU (BASELINE.json configs[1]); scenes are sharded
with no collective (``scaling: weak``), under torchrun every rank renders its own 4096 scenes.  The other BASELINE
configurations -- 3: many-cubes 1024 x 256 boxes at 128^2 (weak), 4: CartPole 65,536 x 84^2 *strong*-sharded over
the ranks, 5: mixed-mesh 16,384 x 64 instances at 256^2 (strong) -- are measured in the same run with a short
device-timed loop and reported under ``extra`` (``--no-extras`` skips them; ``--config C`` makes C the headline).

* a *step* = ``renderer.step(state)``: ONE raster kernel for CartPole (the pose of cart and pole is computed in the
  kernel from the state tensor), geometry pre-pass + staged raster for the large scenes.
* ``value``: whole-job scene-frames/s with the inputs already resident in HBM, K steps (CartPole: replayed from
  CUDA graphs that were replayed before the timed region as well), CUDA events on the launching stream, max over ranks.
  ``value_eager``: the same K steps issued eagerly through ``renderer.step`` (host-side launch cost included).
* ``verified``: after the timed region every output-ring buffer is compared with the CPU oracle's frame of the state
  that was rendered into it last (all scenes; config 5: 64 scenes spread over the batch).  A mismatch aborts the line.
* ``e2e``: same metric through the public API with HOST buffers: pinned inputs -> H2D -> step -> D2H of the uint8
  frames into pinned memory, every step, copies inside the timed region (two pinned frame buffers; the host waits for
  every frame).  ``d2h_ceiling_GBps``: a plain pinned D2H copy of one frame on all ranks at once, the bound of e2e.
* ``roofline``: the dominant raster kernel alone (same inputs), algorithmic bytes / average launch time against
  MEASURED_PEAKS.json's HBM copy bandwidth; ``traffic`` from the ncu range capture committed under profiles/.
* ``env_step``: full ``env.step`` of the CartPole environment (physics + render, no per-step sync; 5 warm-up + 100
  steps: the protocol of the reference's examples/scripts/cartpole_benchmark.py:135-168).
* ``gather`` (N > 1): NCCL gather of all ranks' frames onto rank 0, timed separately, never part of ``value``.
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle (``oracle/``, a port of the reference pipeline -- Panda3D +
  OpenGL cannot run here) on all host cores, persistent thread pool, frame marshalled once, best of 5 (3) blocks of >= 1 s (3 s).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STATE_RING = 16
GRAPH_STEPS = 128       # steps per CUDA graph of runs longer than 256 steps (multiple of STATE_RING and of the output ring)

# name, scenes (global for strong scaling, per GPU for weak), tile, algorithmic bytes per scene-frame
CONFIGS = {
    2: dict(kind="cartpole", scenes=4096, tile=(64, 64), scaling="weak", algo=3 * 64 * 64 + 64 + 2 * 80,
            metric="scene-frames/sec to torch tensor (CartPole 4096x64^2)",
            workload="CartPole-v0 num_scenes=4096 per GPU, tile 64x64 (BASELINE configs[1])"),
    3: dict(kind="cubes", scenes=1024, tile=(128, 128), scaling="weak", algo=3 * 128 * 128 + 64 + 256 * 80,
            metric="scene-frames/sec to torch tensor (many-cubes 1024x256 boxes, 128^2)",
            workload="demo_many_cubes: 1024 scenes x 256 box instances per GPU, tile 128x128 (BASELINE configs[2])"),
    4: dict(kind="cartpole", scenes=65536, tile=(84, 84), scaling="strong", algo=3 * 84 * 84 + 64 + 2 * 80,
            metric="scene-frames/sec to torch tensor (CartPole 65536x84^2, scene-sharded)",
            workload="CartPole-v0 num_scenes=65536 in total, tile 84x84, scene-sharded over the GPUs (BASELINE configs[3])"),
    5: dict(kind="mixed", scenes=16384, tile=(256, 256), scaling="strong", algo=3 * 256 * 256 + 64 + 64 * 80,
            metric="scene-frames/sec to torch tensor (mixed-mesh 16384x64 instances, 256^2, scene-sharded)",
            workload="mixed-mesh: 16384 scenes x 64 instances (box, cone.egg, cylinder glTF, sphere) in total, tile "
                     "256x256, scene-sharded over the GPUs (BASELINE configs[4])"),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of raster_warp_kernel<14, true> from an ncu range over
# consecutive launches cycling the 4-buffer output ring (profiles/r02*_traffic_range.txt); None until captured
NCU_TRAFFIC = {2: None}
try:
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as _f:
        NCU_TRAFFIC.update({int(k): v for k, v in json.load(_f).items()})
except Exception:
    pass


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer loop (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configs and env_step")
    ap.add_argument("--no-gather", action="store_true", help="N>1: skip the NCCL gather timing")
    ap.add_argument("--gather", action="store_true", help="(default at N>1; kept for compatibility)")
    ap.add_argument("--no-verify", action="store_true", help="profiling runs only: skip the oracle comparison")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs while the timed regions execute
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
# oracle helpers (the checker of `verified`, and the CPU legs)
# ------------------------------------------------------------------------------------------------
def oracle_frame_of(renderer, scenes=None):
    """Renderer state -> oracle frame; ``scenes``: gather only those scenes' rows into a compact frame."""
    import numpy as np

    import oracle
    fa = renderer.frame_arrays()
    nodes = []
    vp = fa["vp"]
    K = fa["num_scenes"]
    if scenes is not None:
        scenes = np.asarray(list(scenes), dtype=np.int64)
        vp, K = vp[scenes], len(scenes)
    for n in fa["nodes"]:
        n = dict(n)
        if scenes is not None and not n["shared"]:
            I = n["instances_per_scene"]
            rows = (scenes[:, None] * I + np.arange(I)[None, :]).reshape(-1)
            n["mats"], n["cols"] = n["mats"][rows], n["cols"][rows]
        nodes.append(oracle.OracleNode(**n))
    return oracle.OracleFrame(num_scenes=K, tile_w=fa["tile_w"], tile_h=fa["tile_h"], channels=fa["channels"], vp=vp,
                              bg=fa["bg"], ambient=fa["ambient"], dir_dir=fa["dir_dir"], dir_col=fa["dir_col"],
                              strength=fa["strength"], nodes=nodes)


class CpuArm:
    """The CPU port on the same workload: a renderer on the CPU device (host pose update, the reference's
    envs/cartpole/renderer.py:98-138 on the host) + the oracle rasteriser, frame marshalled once."""

    def __init__(self, config: int, cores: int, sample_scenes: int | None = None):
        import numpy as np
        import torch

        import oracle
        from pybatchrender_b200 import workloads
        from pybatchrender_b200.envs.cartpole import CartPoleRenderer
        c = CONFIGS[config]
        self.c, self.cores = c, cores
        W, H = c["tile"]
        # the host pose update is a few small torch ops; with torch's own thread pool left on, its spinning
        # OpenMP workers fight the rasteriser's threads for the cores (measured: 20 ms per step instead of 8)
        torch.set_num_threads(1)
        if c["kind"] == "cartpole":
            n = c["scenes"]
            self.r = CartPoleRenderer(dict(num_scenes=n, tile_resolution=c["tile"], device="cpu"))
            self.states = [workloads.cartpole_state(n, 1000 + i) for i in range(4)]
            self.r._step(self.states[0])
            scenes = None
        else:
            build = workloads.many_cubes if c["kind"] == "cubes" else workloads.mixed_meshes
            self.r = build(device="cpu")
            self.states = None
            n = c["scenes"]
            scenes = None
        # bounded sample: a contiguous prefix of the scenes (the scenes are i.i.d.)
        self.n_total = n
        self.n = n if sample_scenes is None else min(n, sample_scenes)
        fa_frame = oracle_frame_of(self.r)
        if self.states is not None:
            # alias the live CPU matrix buffers so that the per-step host pose update is seen without re-marshalling
            live = [nd for nd in self.r._drawable_nodes()]
            for on, nd in zip(fa_frame.nodes, live):
                on.mats = nd._matbuf.numpy()
                on.cols = nd.colbuf.numpy()
        self.out = np.zeros((n, 3, H, W), np.uint8)
        self.packed = oracle.PackedFrame(fa_frame, self.out, n_threads=cores, scene_begin=0, scene_count=self.n)

    def step(self, i: int) -> None:
        if self.states is not None:
            self.r._step(self.states[i % 4])
        self.packed.render()

    def time_blocks(self, steps_per_block: int, blocks: int = 3, min_block_s: float = 1.0, warmup: int = 3):
        for i in range(max(1, warmup)):
            self.step(i)
        per = None
        for i in range(3):                       # fastest of three single steps sizes the blocks
            t0 = time.perf_counter()
            self.step(i)
            dt = max(time.perf_counter() - t0, 1e-4)
            per = dt if per is None or dt < per else per
        # a block lasts >= min_block_s and at most ~4 s however many steps were asked for
        target_s = min(max(min_block_s, steps_per_block * per), max(min_block_s, 4.0))
        n_steps = max(1, int(math.ceil(target_s / per)))
        best = None
        self.blocks = blocks
        done = 0
        while done < blocks:
            t0 = time.perf_counter()
            for i in range(n_steps):
                self.step(i)
            dt = time.perf_counter() - t0
            if dt < 0.8 * min_block_s:
                # the sizing steps were slower than the steady state: lengthen the block, do not count this one
                n_steps = int(math.ceil(n_steps * 1.1 * min_block_s / dt))
                best = None
                done = 0
                continue
            best = dt if best is None or dt < best else best
            done += 1
        return self.n * n_steps / best, n_steps, best

    def sample_text(self, n_steps, dt):
        part = "full frames" if self.n == self.n_total else f"frames of the first {self.n} of {self.n_total} scenes"
        return f"best of {self.blocks} blocks of {n_steps} {part} ({dt:.2f} s per block), {self.cores} threads, persistent pool"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    c = CONFIGS[args.config]
    K, W = max(1, args.steps), max(0, args.warmup)
    sample = {2: None, 3: 256, 4: 16384, 5: 64}[args.config]
    arm = CpuArm(args.config, cores, sample)
    value, n_steps, dt = arm.time_blocks(K, blocks=5, min_block_s=1.0, warmup=min(W, 50))
    line = {
        "impl": "reference", "metric": c["metric"], "value": value, "unit": "scene-frames/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": 1e3 * dt / n_steps * (arm.n_total / arm.n), "higher_is_better": True,
        "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32+i64 (u8 out)", "data": "synthetic",
        "config": {"workload": c["workload"],
                   "note": "reference pipeline (Panda3D + OpenGL) cannot run in this image; this arm times the CPU "
                           "oracle port of it on the host cores; ms_per_step is scaled to the full batch"},
        "cpu_baseline": {"value": value, "unit": "scene-frames/s", "cores": cores, "kind": "port",
                         "sample": arm.sample_text(n_steps, dt)},
        "e2e": {"value": value, "unit": "scene-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU workloads
# ------------------------------------------------------------------------------------------------
class Workload:
    """One BASELINE configuration on this rank: renderer, input ring, output ring, step functions."""

    def __init__(self, config: int, dev, rank: int, world: int):
        import torch

        from pybatchrender_b200 import workloads
        from pybatchrender_b200.dist import shard_config
        from pybatchrender_b200.envs.cartpole import CartPoleConfig, CartPoleRenderer
        c = CONFIGS[config]
        self.c, self.config, self.dev, self.rank, self.world = c, config, dev, rank, world
        W, H = c["tile"]
        strong = c["scaling"] == "strong"
        self.total_scenes = c["scenes"] if strong else c["scenes"] * world
        if c["kind"] == "cartpole":
            cfg = CartPoleConfig(num_scenes=c["scenes"], tile_resolution=c["tile"], device="cuda")
            if strong and world > 1:
                cfg = shard_config(cfg, rank, world)
            self.r = CartPoleRenderer(cfg)
            n = int(cfg.num_scenes)
            self.states = [workloads.cartpole_state(n, 1000 * rank + i).to(dev) for i in range(STATE_RING)]
            self.kernel = "raster_warp_kernel<14, true>"
        else:
            build = workloads.many_cubes if c["kind"] == "cubes" else workloads.mixed_meshes
            if strong:
                self.r = build(device="cuda", rank=rank, world_size=world)
            else:
                self.r = build(device="cuda", seed=123 + rank)
            n = int(self.r.num_scenes)
            self.states = None
            self.kernel = ("cull_kernel + bin_xform_kernel + bin_tri_kernel + bin_blocks_kernel<0|1> + bin_scan_kernel + "
                           "raster_binned_kernel<%s> (the whole large-scene pipeline, per launch chunk)" % (
                               "true" if c["kind"] == "mixed" else "false"))
        self.n = n
        self.frame_bytes = n * 3 * H * W
        # output ring: more than the 126 MB L2 in flight, and >= 3 buffers where it is cheap so that consecutive
        # frames (which overlap through the programmatic launch chain) never write the same memory
        self.out_ring = 4 if self.frame_bytes <= (256 << 20) else (3 if self.frame_bytes <= (2 << 30) else 1)
        self.outs = [torch.empty((n, 3, H, W), dtype=torch.uint8, device=dev) for _ in range(self.out_ring)]
        self.use_graphs = c["kind"] == "cartpole"

    def step(self, i: int):
        out = self.outs[i % self.out_ring]
        if self.states is not None:
            return self.r.step(self.states[i % STATE_RING], out=out)
        return self.r.render(out=out)

    def raster_only(self, i: int):
        # small scenes: the step IS the raster kernel (pose folded in), so the roofline leg launches exactly what the
        # timed steps launch, state ring included (with one state re-used its rows stay in L2: 0.3 us per frame)
        if self.states is not None:
            return self.step(i)
        return self.r.render(out=self.outs[i % self.out_ring])

    # e2e: host inputs of one step
    def make_host_inputs(self):
        import torch

        from pybatchrender_b200 import workloads
        if self.states is not None:
            self.h_in = [[workloads.cartpole_state(self.n, 5000 + 1000 * self.rank + i).pin_memory()] for i in range(4)]
            self.d_in = [torch.empty((self.n, 4), dtype=torch.float32, device=self.dev)]
        else:
            nodes = [nd for nd in self.r._drawable_nodes() if not nd.shared_across]
            self.d_in = [nd._matbuf for nd in nodes]
            self.h_in = [[t.cpu().pin_memory() for t in self.d_in] for _ in range(2)]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.d_in)

    def e2e_render(self, i: int):
        for h, d in zip(self.h_in[i % len(self.h_in)], self.d_in):
            d.copy_(h, non_blocking=True)
        out = self.outs[i % self.out_ring]
        if self.states is not None:
            return self.r.step(self.d_in[0], out=out)
        return self.r.render(out=out)

    def verify(self, last_step_of_buffer: dict, cores: int):
        """Every ring buffer against the oracle's frame of the state rendered into it last."""
        import numpy as np
        import torch

        import oracle
        checked = 0
        for b, i in sorted(last_step_of_buffer.items()):
            got = self.outs[b]
            if self.states is not None:
                self.r._step(self.states[i % STATE_RING])
            if self.c["kind"] == "mixed":
                sample = sorted(set(range(0, self.n, max(1, self.n // 64))) | {self.n - 1})
                ref = oracle.render(oracle_frame_of(self.r, sample), n_threads=cores)
                got = got[torch.tensor(sample, device=got.device)]
            elif self.states is None and checked > 0:
                ref = None                      # static scene: the other ring buffers must equal the verified one
                if not torch.equal(got, self.outs[sorted(last_step_of_buffer)[0]]):
                    return False, f"ring buffer {b} differs from buffer 0"
            else:
                ref = oracle.render(oracle_frame_of(self.r), n_threads=cores)
            if ref is not None:
                same = np.array_equal(got.cpu().numpy(), ref)
                if not same:
                    return False, f"ring buffer {b} (step {i}) differs from the oracle"
            checked += 1
        what = "64 sampled scenes" if self.c["kind"] == "mixed" else "all scenes"
        return True, f"{checked} ring buffer(s), {what} each, bit-exact vs the CPU oracle"


def timed_run(wl: Workload, K: int, W: int, barrier, native):
    """K steps of ``wl.step`` timed with CUDA events (graphs for CartPole); returns (ms, launches, last writer of
    every ring buffer)."""
    import gc

    import torch
    last = {}
    gc.collect()            # renderers of earlier legs are destroyed now, not in the middle of a capture
    if wl.use_graphs:
        side = torch.cuda.Stream(wl.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                wl.step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

        def capture(n_steps):
            g = torch.cuda.CUDAGraph()
            l0 = native.kernel_launches()
            with torch.cuda.graph(g):
                for i in range(n_steps):
                    wl.step(i)
            return g, native.kernel_launches() - l0

        if K <= 256:
            # short runs (the driver's --steps 20): ONE graph of exactly K steps, so that no host-side graph launch
            # falls inside the timed region (with 8 ranks on one host a late second launch showed as 19.3 us per step)
            g_all, l_all = capture(K)

            def run_steps(k):
                g_all.replay()
            for _ in range(max(2, (W + K - 1) // K)):
                g_all.replay()
            launches = l_all
            for i in range(K):
                last[i % wl.out_ring] = i
        else:
            # long runs: graphs of GRAPH_STEPS steps (a multiple of the state and output rings) plus a tail graph; the
            # first launch of every replay has no frame ahead of it to overlap with (+ ~7 us per replay boundary)
            g_full, l_full = capture(GRAPH_STEPS)
            tail = K % GRAPH_STEPS
            g_tail, l_tail = capture(tail) if tail else (None, 0)

            def run_steps(k):
                # exactly k steps: whole graphs, then the tail graph (k % GRAPH_STEPS == tail by construction)
                for _ in range(k // GRAPH_STEPS):
                    g_full.replay()
                if k % GRAPH_STEPS:
                    g_tail.replay()
            # warm-up: >= W steps and, whatever W is, at least two replays of every graph that is timed
            for _ in range(max(2, (W + GRAPH_STEPS - 1) // GRAPH_STEPS)):
                g_full.replay()
            if g_tail is not None:
                for _ in range(2):
                    g_tail.replay()
            launches = (K // GRAPH_STEPS) * l_full + (l_tail if tail else 0)
            # what each ring buffer holds at the end: the full graph ran (in the warm-up at least), then the tail graph
            for i in list(range(GRAPH_STEPS)) + list(range(tail)):
                last[i % wl.out_ring] = i
    else:
        for i in range(max(3, W)):
            wl.step(i)

        def run_steps(k):
            for i in range(k):
                wl.step(i)
        launches = None
        for i in range(K):
            last[i % wl.out_ring] = i
    barrier()
    l0 = native.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if wl.use_graphs:
        # ~0.2 ms of device-side delay ahead of the start event: the host enqueues the event and the graph while it
        # runs, so the events bracket the K steps on the device and not the host's launch latency (with 8 ranks on one
        # host that latency showed as +1 us per step at K = 20); torch's own spin kernel, not one of ours
        torch.cuda._sleep(400_000)
    e0.record()
    run_steps(K)
    e1.record()
    barrier()
    if launches is None:
        launches = native.kernel_launches() - l0
    return e0.elapsed_time(e1), launches, last


ROOFLINE_LAUNCHES = 128


def roofline_leg(wl: Workload, K: int):
    """The raster kernel(s) alone on resident inputs: average duration of a launch, CUDA events on the launching stream.

    Small-scene configurations (one kernel per frame, consecutive launches overlap through programmatic dependent
    launch): ONE graph of 128 ... 256 launches (K clamped to that range), replayed twice untimed, then timed -- elapsed / launches.  The first
    launch of a chain has no frame ahead of it to overlap with (it alone takes ~23 us), so a chain of 16-20 launches
    reads 3-5 % slower than the kernel runs in steady state; `roofline.launches_timed` says how many were averaged."""
    import torch
    if wl.use_graphs:
        n = min(max(K, ROOFLINE_LAUNCHES), 2 * ROOFLINE_LAUNCHES)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n):
                wl.raster_only(i)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        g.replay()
        r1.record()
        torch.cuda.synchronize()
        return r0.elapsed_time(r1) / n
    n = max(2, min(K, 20))
    for i in range(2):
        wl.raster_only(i)
    torch.cuda.synchronize()
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record()
    for i in range(n):
        wl.raster_only(i)
    r1.record()
    torch.cuda.synchronize()
    return r0.elapsed_time(r1) / n


def max_over_ranks(values, dev, distributed):
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def measure_config(config, K, W, dev, rank, world, distributed, barrier, native, cores, verify=True, full=False,
                   e2e_steps=0):
    """Device-timed value + roofline + verification of one configuration; ``full`` adds eager and e2e legs."""
    import torch
    wl = Workload(config, dev, rank, world)
    c = wl.c
    ms, launches, last = timed_run(wl, K, W, barrier, native)
    res = {}
    if verify:
        ok, how = wl.verify(last, cores)
        flag = max_over_ranks([0.0 if ok else 1.0], dev, distributed)[0]
        if flag != 0.0:
            raise SystemExit(f"bench.py: config {config}: frames differ from the oracle on some rank ({how}); no line printed")
        res["verified"], res["verified_how"] = True, how
    else:
        res["verified"], res["verified_how"] = None, "skipped (--no-verify)"
    raster_ms = roofline_leg(wl, K)
    legs = [ms, raster_ms]

    if full:
        # eager: the same K steps issued one by one through renderer.step (bounded so that it stays short)
        k_e = min(K, 2000)
        for i in range(16):
            wl.step(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(k_e):
            wl.step(i)
        host_s = time.perf_counter() - t0
        e1.record()
        barrier()
        eager_ms = e0.elapsed_time(e1)
        legs += [eager_ms / k_e, host_s * 1e3 / k_e]

        # e2e, host buffers: sync after every frame, and pipelined over two pinned frame buffers
        wl.make_host_inputs()
        H, Wd = c["tile"][1], c["tile"][0]
        h_outs = [torch.empty((wl.n, 3, H, Wd), dtype=torch.uint8).pin_memory() for _ in range(2)]
        n_e2e = e2e_steps or max(8, min(K, 200 if wl.frame_bytes < (256 << 20) else 10))

        def e2e_sync(i):
            px = wl.e2e_render(i)
            h_outs[0].copy_(px, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for i in range(3):
            e2e_sync(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_sync(i)
        barrier()
        e2e_sync_s = time.perf_counter() - t0

        copy_stream = torch.cuda.Stream(dev)
        landed = [None, None]

        def e2e_pipe(i):
            px = wl.e2e_render(i)
            rendered = torch.cuda.Event()
            rendered.record()
            if landed[i % 2] is not None:
                landed[i % 2].synchronize()             # frame i-2 is on the host: its buffer is free again
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(rendered)
                h_outs[i % 2].copy_(px, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            landed[i % 2] = ev

        def drain():
            for ev in landed:
                if ev is not None:
                    ev.synchronize()

        for i in range(4):
            e2e_pipe(i)
        drain()
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_pipe(i)
        drain()
        barrier()
        e2e_s = time.perf_counter() - t0

        # the bound of e2e: plain pinned D2H copies of one frame, all ranks at once
        for _ in range(2):
            h_outs[0].copy_(wl.outs[0], non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        n_copy = 10
        for k in range(n_copy):
            h_outs[k % 2].copy_(wl.outs[0], non_blocking=True)
        barrier()
        d2h_s = (time.perf_counter() - t0) / n_copy
        legs += [e2e_s * 1e3, e2e_sync_s * 1e3, d2h_s * 1e3]

    legs = max_over_ranks(legs, dev, distributed)
    ms, raster_ms = legs[0], legs[1]
    peak, peak_src = measured_peak()
    per_rank = wl.n
    value = wl.total_scenes * K / (ms * 1e-3)
    achieved = c["algo"] * per_rank / (raster_ms * 1e-3) / 1e9
    traffic = NCU_TRAFFIC.get(config)
    res.update({
        "metric": c["metric"], "value": value, "unit": "scene-frames/s", "ms_per_step": ms / K, "steps": K,
        "scaling": c["scaling"],
        "config": {"workload": c["workload"], "scenes_per_gpu": per_rank, "scenes_total": wl.total_scenes,
                   "tile": list(c["tile"]), "channels": 3,
                   "parallelism": f"scene-sharded x{world}, no collective",
                   "l2": f"output ring of {wl.out_ring} x {wl.frame_bytes / 1e6:.1f} MB (> 126 MB L2 in flight)"
                         + (f"; state ring of {STATE_RING}" if wl.states is not None else ""),
                   "launch": ((f"one CUDA graph of the {K} timed steps" if K <= 256 else f"CUDA graphs of {GRAPH_STEPS} steps")
                              + ", replayed before the timed region; one kernel per step, consecutive frames overlap "
                                "through programmatic dependent launch; a 0.2 ms device-side delay ahead of the start event "
                                "keeps the host's launch latency out of the timed region")
                   if wl.use_graphs else "eager: instance cull + geometry pre-pass + staged raster per launch chunk"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic,
                     "traffic_over_algorithmic": None if traffic is None else traffic / (c["algo"] * per_rank),
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu range over consecutive "
                                     "launches cycling the output ring (profiles/)" if traffic is not None
                     else "not captured for this configuration",
                     "kernel": wl.kernel, "kernel_ms": raster_ms,
                     "launches_timed": min(max(K, ROOFLINE_LAUNCHES), 2 * ROOFLINE_LAUNCHES) if wl.use_graphs else max(2, min(K, 20)),
                     "frac_over_timed_steps": (c["algo"] * per_rank / (ms / K * 1e-3) / 1e9 / peak) if launches == K else None,
                     "algorithmic_bytes_per_launch": c["algo"] * per_rank, "peak_source": peak_src},
        "gpu_launches": launches,
    })
    if full:
        eager_ms_step, host_ms_step, e2e_ms, e2e_sync_ms, d2h_ms = legs[2:7]
        res["value_eager"] = {"value": wl.total_scenes / (eager_ms_step * 1e-3), "unit": "scene-frames/s",
                              "ms_per_step": eager_ms_step, "host_ms_per_step": host_ms_step,
                              "how": "the same steps issued eagerly through renderer.step (no CUDA graph), CUDA events; "
                                     "host_ms_per_step = wall time the Python loop spends per call"}
        ceiling = wl.frame_bytes * world / (d2h_ms * 1e-3) / 1e9
        e2e_val = wl.total_scenes * n_e2e / (e2e_ms * 1e-3)
        res["e2e"] = {"value": e2e_val, "unit": "scene-frames/s", "h2d_bytes_per_step": wl.h2d_bytes,
                      "d2h_bytes_per_step": wl.frame_bytes, "steps": n_e2e,
                      "how": "pinned host inputs -> H2D -> step -> D2H of the frames every step; two pinned frame buffers, "
                             "the D2H of frame i overlaps the H2D + render of frame i+1, the host waits for every frame",
                      "value_sync_every_step": wl.total_scenes * n_e2e / (e2e_sync_ms * 1e-3),
                      "d2h_ceiling_GBps": ceiling,
                      "d2h_GBps": e2e_val * 3 * c["tile"][0] * c["tile"][1] / 1e9,
                      "frac_of_d2h_ceiling": (e2e_val * 3 * c["tile"][0] * c["tile"][1] / 1e9) / ceiling,
                      "ceiling_note": "plain pinned device->host copies of one frame issued on all ranks at once (aggregate GB/s)"}
    return res, wl


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

    from pybatchrender_b200 import _native
    native = _native.Native()
    cores = os.cpu_count() or 1
    K, W = max(1, args.steps), max(3, args.warmup)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    head, wl = measure_config(args.config, K, W, dev, rank, world, distributed, barrier, native, cores,
                              verify=not args.no_verify, full=True, e2e_steps=args.e2e_steps)

    # ---- frames of all ranks onto rank 0 (NCCL gather over NVLink); never part of step()
    gather = None
    if distributed and not args.no_gather:
        from pybatchrender_b200.dist import gather_frames
        for _ in range(2):
            gather_frames(wl.outs[0], dst=0)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(10):
            gather_frames(wl.outs[0], dst=0)
        g1.record()
        barrier()
        g_ms = max_over_ranks([g0.elapsed_time(g1) / 10], dev, distributed)[0]
        nbytes = (world - 1) * wl.frame_bytes
        gather = {"ms": g_ms, "bytes_into_rank0": nbytes, "GBps_into_rank0": nbytes / (g_ms * 1e-3) / 1e9,
                  "note": "optional collective (torch.distributed gather, NCCL), timed separately, not part of value / e2e"}

    # ---- env.step of the CartPole environment (reference examples/scripts/cartpole_benchmark.py:135-168)
    env_step = None
    if CONFIGS[args.config]["kind"] == "cartpole" and not args.no_extras:
        import pybatchrender_b200 as pbr
        n_env = wl.n
        env = pbr.envs.make("CartPole-v0", num_scenes=n_env, tile_resolution=CONFIGS[args.config]["tile"], device="cuda")
        td = env.reset()
        for _ in range(5):
            td["action"] = env.action_spec.rand()
            td = env.step(td)["next"]
        barrier()
        t0 = time.perf_counter()
        for _ in range(100):
            td["action"] = env.action_spec.rand()
            td = env.step(td)["next"]
        barrier()
        env_s = max_over_ranks([time.perf_counter() - t0], dev, distributed)[0]
        env_step = {"value": n_env * world * 100 / env_s, "unit": "scene-frames/s", "ms_per_step": env_s * 10.0,
                    "steps": 100, "warmup": 5,
                    "how": "pbr.envs.make('CartPole-v0').step: physics + render + auto-reset, random actions, no "
                           "per-step sync, wall clock (reference on an NVIDIA L4: 980,562 scene-frames/s at 4096 "
                           "scenes, examples/notebooks/cartpole_benchmark.ipynb raw line 886)"}
        del env, td

    del wl
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, short device-timed loops
    extra = {}
    if not args.no_extras:
        short = {2: (min(K, 200), 16), 3: (min(K, 50), 3), 4: (min(K, 48), 16), 5: (min(K, 3), 1)}
        for cfg_id in sorted(CONFIGS):
            if cfg_id == args.config:
                continue
            k_c, w_c = short[cfg_id]
            try:
                res, w2 = measure_config(cfg_id, k_c, w_c, dev, rank, world, distributed, barrier, native, cores,
                                         verify=not args.no_verify, full=False)
                del w2
                extra[f"config{cfg_id}"] = {k: res[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "scaling",
                                                                "verified", "verified_how", "gpu_launches")}
                extra[f"config{cfg_id}"]["workload"] = res["config"]["workload"]
                extra[f"config{cfg_id}"]["scenes_per_gpu"] = res["config"]["scenes_per_gpu"]
                extra[f"config{cfg_id}"]["roofline_frac"] = res["roofline"]["frac"]
                extra[f"config{cfg_id}"]["kernel_ms"] = res["roofline"]["kernel_ms"]
                extra[f"config{cfg_id}"]["kernel"] = res["roofline"]["kernel"]
            except SystemExit:
                raise
            except Exception as e:      # an extra must not take the headline down; say what happened
                extra[f"config{cfg_id}"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()

    sampler.stop_flag = True
    sampler.join(timeout=1.0)

    if rank == 0:
        c = CONFIGS[args.config]
        line = {
            "metric": head["metric"], "value": head["value"], "unit": "scene-frames/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": c["scaling"],
            "vs_baseline": None, "dtype": "f32+i32 (u8 out)", "data": "synthetic", "config": head["config"],
            "verified": head["verified"], "verified_how": head["verified_how"],
            "roofline": head["roofline"], "e2e": head["e2e"], "value_eager": head["value_eager"],
            "gpu_launches": head["gpu_launches"], "clocks": sampler.summary(), "gather": gather,
            "env_step": env_step, "extra": extra,
        }
        if world == 1 and not args.no_cpu_baseline:
            sample = {2: None, 3: 256, 4: 16384, 5: 64}[args.config]
            arm = CpuArm(args.config, cores, sample)
            v, n_steps, dt = arm.time_blocks(8, blocks=3, min_block_s=3.0)
            line["cpu_baseline"] = {"value": v, "unit": "scene-frames/s", "cores": cores, "kind": "port",
                                    "sample": arm.sample_text(n_steps, dt)}
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
