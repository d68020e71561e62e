#!/usr/bin/env python
"""CartPole throughput demo with the command line of the reference's
``examples/scripts/cartpole_benchmark.py`` (same flags, same printed quantities), on the B200
rasteriser.

    python examples/scripts/cartpole_benchmark.py --num-scenes 4096 --steps 500
    python examples/scripts/cartpole_benchmark.py --gif --gif-steps 60 --save-dir ./outputs
    torchrun --nproc-per-node 8 examples/scripts/cartpole_benchmark.py --parallel --num-scenes 32768

``--parallel`` in the reference spawns TorchRL ``ParallelEnv`` workers because one Panda3D window
lives in one process; here it means "shard the scenes over the ranks of this torchrun job" (one
process per GPU, ``pybatchrender_b200.dist.shard_config``) and ``--num-workers`` is ignored.
``--window`` is accepted and ignored: there is no window.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import pybatchrender_b200 as pbr  # noqa: E402


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="CartPole environment benchmark.",
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--num-scenes", type=int, default=1024, help="Scenes rendered per step (whole job).")
    p.add_argument("--tile-resolution", type=int, nargs=2, metavar=("W", "H"), default=(64, 64))
    p.add_argument("--steps", type=int, default=500, help="Environment steps to time.")
    p.add_argument("--parallel", action="store_true", help="Shard scenes over the torchrun ranks (one per GPU).")
    p.add_argument("--num-workers", type=int, default=4, help="Ignored (reference: ParallelEnv workers).")
    p.add_argument("--save-every", type=int, default=-1, help="Save an image grid every N steps (-1: never).")
    p.add_argument("--save-num", type=int, default=16)
    p.add_argument("--save-dir", type=str, default="./outputs")
    p.add_argument("--no-render", action="store_true", help="State only, no pixels.")
    p.add_argument("--window", action="store_false", dest="offscreen", default=True, help="Ignored.")
    p.add_argument("--device", type=str, default=None, help="cuda (default when available) or cpu (needs --no-render).")
    p.add_argument("--gif", action="store_true", help="Write an animated GIF instead of timing.")
    p.add_argument("--gif-steps", type=int, default=100)
    p.add_argument("--gif-interval", type=int, default=2)
    p.add_argument("--gif-scale", type=int, default=3)
    p.add_argument("--gif-duration", type=int, default=100)
    return p.parse_args(argv)


def make_env(args):
    overrides = dict(num_scenes=args.num_scenes, tile_resolution=tuple(args.tile_resolution),
                     render=not args.no_render, offscreen=args.offscreen)
    if args.device:
        overrides["device"] = args.device
    rank, world = 0, 1
    if args.parallel and "RANK" in os.environ:
        import torch.distributed as dist
        from pybatchrender_b200.dist import shard_config
        from pybatchrender_b200.envs.cartpole import CartPoleConfig, CartPoleEnv, CartPoleRenderer
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        if not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
        cfg = shard_config(CartPoleConfig(**overrides), rank=rank, world_size=world)
        return CartPoleEnv(renderer=CartPoleRenderer(cfg), cfg=cfg), rank, world
    return pbr.envs.make("CartPole-v0", **overrides), rank, world


def run(args) -> dict:
    env, rank, world = make_env(args)
    local = int(env.cfg.num_scenes)
    td = env.reset()
    if rank == 0:
        print("=" * 60)
        print(f"CartPole benchmark: {args.num_scenes} scenes at {tuple(args.tile_resolution)}, "
              f"{world} process(es), device {env.device}, render={not args.no_render}")
        print("=" * 60)
    cuda = env.device.type == "cuda"

    if args.gif:
        frames = []
        for t in range(args.gif_steps):
            td["action"] = env.action_spec.rand()
            td = env.step(td)["next"]
            if t % max(1, args.gif_interval) == 0:
                frames.append(td["pixels"].clone())
        path = env.save_batch_gif(frames, num=args.save_num, scale=args.gif_scale, out_dir=args.save_dir,
                                  filename_prefix="cartpole", duration_ms=args.gif_duration)
        print(f"wrote {path} ({len(frames)} frames)")
        return {"gif": path}

    total_reward = 0.0
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(args.steps):
        td["action"] = env.action_spec.rand()
        td = env.step(td)
        total_reward = total_reward + td["next", "reward"].sum()
        td = td["next"]
        if args.save_every >= 0 and not args.no_render and (args.save_every == 0 or t % args.save_every == 0):
            env.save_batch_examples(pixels=td["pixels"], num=args.save_num, out_dir=args.save_dir,
                                    filename_prefix=f"cartpole_step{t:05d}_rank{rank}")
    if cuda:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    fps = args.steps * local / dt
    print(f"[rank {rank}] {args.steps} steps x {local} scenes in {dt:.3f} s -> {fps:,.0f} scene-frames/s "
          f"(mean reward per step {float(total_reward) / max(1, args.steps * local):.3f})")
    return {"fps": fps, "seconds": dt, "scenes": local}


if __name__ == "__main__":
    run(parse_args())
