"""``PBREnv`` -- batched RL environment base class whose pixels come from a ``PBRRenderer``.

Reference: ``pybatchrender/env.py:18-440``.  The pixel path is one line there and here:
``render_pixels(obs) = renderer.step(obs)`` (env.py:92-95) -- but here ``obs`` stays on the GPU and
the returned ``[B,C,H,W]`` uint8 tensor was written by the CUDA rasteriser directly.

Kept API: ctor ``(renderer, cfg=None, device="cpu", batch_size=[128])``, ``set_default_specs``,
``render_pixels``, ``save_batch_examples`` (PNG grid, nearest-neighbour ``scale``; the notebook
goldens were produced by it), classmethod ``make_parallel_env`` (TorchRL ``ParallelEnv``; needs
TorchRL -- on B200 boxes prefer one process per GPU with ``pybatchrender_b200.dist.shard_config``).
"""
from __future__ import annotations

import io
import math
import os
from abc import ABC, abstractmethod

import torch

from ._rl_compat import Categorical, Composite, EnvBase, HAVE_TORCHRL, ParallelEnv, TensorDict, Unbounded
from .config import PBRConfig
from .renderer.renderer import PBRRenderer


class PBREnv(EnvBase, ABC):
    def __init__(self, renderer: PBRRenderer, cfg: PBRConfig | None = None,
                 device: str | torch.device = "cpu", batch_size: torch.Size = torch.Size([128]), **kwargs) -> None:
        super().__init__(device=torch.device(device), batch_size=batch_size, **kwargs)
        self._renderer = renderer
        self.cfg = cfg if cfg is not None else renderer.cfg

    def set_default_specs(self, *, direct_obs_dim: int | None = None, actions: int | None = None,
                          with_pixels: bool = False, pixels_only: bool = False,
                          discrete_actions: bool = True) -> None:
        assert with_pixels or not pixels_only, "Cannot set both with_pixels and pixels_only to True."
        bs = self.batch_size if self.batch_size != torch.Size([]) else torch.Size([1])
        fields = {}
        if with_pixels:
            C = int(self.cfg.num_channels)
            W, H = int(self.cfg.tile_resolution[0]), int(self.cfg.tile_resolution[1])
            fields["pixels"] = Unbounded(shape=bs + torch.Size([C, H, W]), dtype=torch.uint8, device=self.device)
        if not pixels_only:
            if direct_obs_dim is None:
                raise ValueError("direct_obs_dim must be provided when pixels_only=False.")
            fields["observation"] = Unbounded(shape=bs + torch.Size([int(direct_obs_dim)]), dtype=torch.float32,
                                              device=self.device)
        self.observation_spec = Composite(**fields, shape=bs)
        if actions is None:
            raise ValueError("actions must be provided.")
        if discrete_actions:
            self.action_spec = Categorical(n=int(actions), shape=bs, dtype=torch.long, device=self.device)
        else:
            self.action_spec = Unbounded(shape=bs + torch.Size([int(actions)]), dtype=torch.float32,
                                         device=self.device)
        self.reward_spec = Unbounded(shape=bs + torch.Size([1]), dtype=torch.float32, device=self.device)
        self.done_spec = Unbounded(shape=bs + torch.Size([1]), dtype=torch.bool, device=self.device)

    @abstractmethod
    def _step(self, tensordict: TensorDict) -> TensorDict:
        ...

    @abstractmethod
    def _reset(self, tensordict: TensorDict | None = None) -> TensorDict:
        ...

    def render_pixels(self, obs: torch.Tensor | None = None) -> torch.Tensor:
        if self._renderer is None:
            raise RuntimeError("Renderer is not initialized. Construct env with a renderer and pass it to PBREnv.")
        return self._renderer.step(obs)

    # ------------------------------------------------------------------ image export
    @staticmethod
    def _make_grid_frame(pixels: torch.Tensor, indices, cols: int, scale: int = 1):
        """[B,C,H,W] uint8 -> numpy [rows*H*scale, cols*W*scale, C]; tile n at row n//cols, col n%cols."""
        import numpy as np
        sel = pixels[list(indices)].detach().to("cpu")
        n, C, H, W = sel.shape
        rows = math.ceil(n / cols)
        grid = torch.zeros((rows * H, cols * W, C), dtype=torch.uint8)
        for k in range(n):
            r, c = divmod(k, cols)
            grid[r * H:(r + 1) * H, c * W:(c + 1) * W] = sel[k].permute(1, 2, 0)
        arr = grid.numpy()
        if scale > 1:
            arr = np.repeat(np.repeat(arr, scale, axis=0), scale, axis=1)
        return arr

    def save_batch_examples(self, *, pixels: torch.Tensor | None = None, indices=None, num: int = 16,
                            scale: int = 1, out_dir: str | None = None, filename_prefix: str = "batch",
                            return_bytes: bool = False):
        """PNG grid of ``num`` scenes (reference env.py:97-187); returns the path (and the bytes)."""
        from PIL import Image
        if pixels is None:
            pixels = self.render_pixels(None)
        B = int(pixels.shape[0])
        if indices is None:
            indices = list(range(min(int(num), B)))
        indices = [i for i in indices if 0 <= i < B][: int(num)]
        cols = max(1, math.ceil(math.sqrt(len(indices))))
        arr = self._make_grid_frame(pixels, indices, cols, int(scale))
        img = Image.fromarray(arr[..., :3] if arr.shape[-1] == 4 else arr)
        out_dir = out_dir or getattr(self.cfg, "save_out_dir", None) or "./outputs"
        os.makedirs(out_dir, exist_ok=True)
        fpath = os.path.join(out_dir, f"{filename_prefix}.png")
        img.save(fpath)
        if return_bytes:
            buf = io.BytesIO()
            img.save(buf, format="PNG")
            return str(fpath), buf.getvalue()
        return str(fpath)

    def save_batch_gif(self, frames: list[torch.Tensor], *, indices=None, num: int = 16, scale: int = 1,
                       out_dir: str | None = None, filename_prefix: str = "batch", duration_ms: int = 50):
        """Animated GIF of a list of ``[B,C,H,W]`` frames (reference env.py:246-358)."""
        from PIL import Image
        B = int(frames[0].shape[0])
        if indices is None:
            indices = list(range(min(int(num), B)))
        cols = max(1, math.ceil(math.sqrt(len(indices))))
        imgs = [Image.fromarray(self._make_grid_frame(f, indices, cols, int(scale))[..., :3]) for f in frames]
        out_dir = out_dir or "./outputs"
        os.makedirs(out_dir, exist_ok=True)
        fpath = os.path.join(out_dir, f"{filename_prefix}.gif")
        imgs[0].save(fpath, save_all=True, append_images=imgs[1:], duration=duration_ms, loop=0)
        return str(fpath)

    # ------------------------------------------------------------------ parallel envs
    @classmethod
    def make_parallel_env(cls, *, config: PBRConfig, renderer_cls: type[PBRRenderer], num_workers: int,
                          mp_start_method: str = "spawn", shared_memory: bool = True):
        if not HAVE_TORCHRL:
            raise RuntimeError(
                "make_parallel_env needs TorchRL's ParallelEnv, which is not installed; on a multi-GPU box use "
                "one process per GPU with pybatchrender_b200.dist.shard_config instead")
        import multiprocessing as mp
        try:
            mp.set_start_method(mp_start_method, force=True)
        except RuntimeError:
            pass
        kwargs = [dict(env_cls=cls, renderer_cls=renderer_cls, config=config, worker_index=i,
                       num_workers=num_workers) for i in range(int(num_workers))]
        return ParallelEnv(int(num_workers), create_env_fn=cls._make_env_worker, shared_memory=shared_memory,
                           mp_start_method=mp_start_method, create_env_kwargs=kwargs)

    @staticmethod
    def _make_env_worker(env_cls, renderer_cls, config, worker_index: int, num_workers: int):
        cfg = type(config).from_config(config, worker_index=worker_index, num_workers=num_workers)
        if getattr(cfg, "seed", None) is not None:
            cfg.seed = int(cfg.seed) + int(worker_index)
        renderer = renderer_cls(cfg)
        return env_cls(renderer=renderer, cfg=cfg)
