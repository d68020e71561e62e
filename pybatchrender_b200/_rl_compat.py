"""TorchRL / tensordict when installed, otherwise a minimal stand-in with the same call pattern.

The reference's env layer subclasses ``torchrl.envs.EnvBase`` and returns ``tensordict.TensorDict``
(``pybatchrender/env.py:5-7``).  Neither package is in this image, so ``PBREnv`` is written against
this module: with TorchRL present it *is* a TorchRL env; without it the shim below supports exactly
what the reference's benchmark loop uses (``examples/scripts/cartpole_benchmark.py:135-164``):

    td = env.reset(); td["action"] = env.action_spec.rand(); td = env.step(td)
    td["next", "reward"]; td.get(("next", "pixels"), None); td = td["next"]; td.keys()
"""
from __future__ import annotations

import torch

try:  # pragma: no cover - not available in the build image
    from tensordict import TensorDict
    from torchrl.data.tensor_specs import Bounded, Categorical, Composite, Unbounded
    from torchrl.envs import EnvBase, ParallelEnv
    HAVE_TORCHRL = True
except Exception:
    HAVE_TORCHRL = False
    ParallelEnv = None

    class TensorDict(dict):
        """dict of tensors with a batch size and tuple keys for nesting (``td["next", "reward"]``)."""

        def __init__(self, source=None, batch_size=None, device=None):
            super().__init__()
            self.batch_size = torch.Size(batch_size) if batch_size is not None else torch.Size([])
            self.device = device
            for k, v in (source or {}).items():
                self[k] = v

        def __getitem__(self, key):
            if isinstance(key, tuple):
                cur = self
                for k in key:
                    cur = dict.__getitem__(cur, k)
                return cur
            return dict.__getitem__(self, key)

        def __setitem__(self, key, value):
            if isinstance(key, tuple):
                cur = self
                for k in key[:-1]:
                    if k not in cur:
                        dict.__setitem__(cur, k, TensorDict({}, batch_size=self.batch_size))
                    cur = dict.__getitem__(cur, k)
                dict.__setitem__(cur, key[-1], value)
            else:
                dict.__setitem__(self, key, value)

        def get(self, key, default=None):
            try:
                return self[key]
            except KeyError:
                return default

        def set(self, key, value):
            self[key] = value
            return self

        def to(self, device):
            out = TensorDict({}, batch_size=self.batch_size, device=device)
            for k, v in self.items():
                dict.__setitem__(out, k, v.to(device) if hasattr(v, "to") else v)
            return out

        def clone(self):
            out = TensorDict({}, batch_size=self.batch_size, device=self.device)
            for k, v in self.items():
                dict.__setitem__(out, k, v.clone() if hasattr(v, "clone") else v)
            return out

    class _Spec:
        def __init__(self, shape, dtype, device=None):
            self.shape = torch.Size(shape)
            self.dtype = dtype
            self.device = device

        def rand(self):
            if self.dtype == torch.bool:
                return torch.zeros(self.shape, dtype=torch.bool, device=self.device)
            if self.dtype.is_floating_point:
                return torch.randn(self.shape, dtype=self.dtype, device=self.device)
            return torch.zeros(self.shape, dtype=self.dtype, device=self.device)

        def zero(self):
            return torch.zeros(self.shape, dtype=self.dtype, device=self.device)

    class Unbounded(_Spec):
        def __init__(self, shape, dtype=torch.float32, device=None):
            super().__init__(shape, dtype, device)

    class Bounded(_Spec):
        def __init__(self, low, high, shape, dtype=torch.float32, device=None):
            super().__init__(shape, dtype, device)
            self.low, self.high = float(low), float(high)

        def rand(self):
            u = torch.rand(self.shape, dtype=self.dtype, device=self.device)
            return u * (self.high - self.low) + self.low

    class Categorical(_Spec):
        def __init__(self, n, shape, dtype=torch.long, device=None):
            super().__init__(shape, dtype, device)
            self.n = int(n)

        def rand(self):
            return torch.randint(0, self.n, tuple(self.shape), dtype=self.dtype, device=self.device)

    class Composite(dict):
        def __init__(self, shape=None, **fields):
            super().__init__(**fields)
            self.shape = torch.Size(shape) if shape is not None else torch.Size([])

        def rand(self):
            return TensorDict({k: v.rand() for k, v in self.items()}, batch_size=self.shape)

    class EnvBase:
        """The slice of ``torchrl.envs.EnvBase`` the reference relies on."""

        def __init__(self, device=None, batch_size=None, **_kw):
            self.device = torch.device(device) if device is not None else torch.device("cpu")
            self.batch_size = torch.Size(batch_size) if batch_size is not None else torch.Size([])
            self.observation_spec = None
            self.action_spec = None
            self.reward_spec = None
            self.done_spec = None

        def set_seed(self, seed: int):
            self._set_seed(int(seed))
            return int(seed)

        def _set_seed(self, seed: int) -> None:
            torch.manual_seed(seed)

        def reset(self, tensordict=None):
            return self._reset(tensordict)

        def step(self, tensordict):
            nxt = self._step(tensordict)
            tensordict["next"] = nxt
            return tensordict

        def rand_step(self, tensordict=None):
            if tensordict is None:
                tensordict = TensorDict({}, batch_size=self.batch_size)
            tensordict["action"] = self.action_spec.rand()
            return self.step(tensordict)

        def rollout(self, max_steps: int, tensordict=None):
            td = self.reset() if tensordict is None else tensordict
            out = []
            for _ in range(int(max_steps)):
                td = self.rand_step(td)
                out.append(td)
                td = td["next"]
            return out

        def close(self):
            return None
