"""Gym-style environment registry (reference ``pybatchrender/envs/__init__.py:39-220``).

``register(name, env_cls, renderer_cls, config_cls)``, ``make(name, **overrides)``, ``list_envs``,
``is_registered``, ``get_env_classes``, ``unregister``, ``make_parallel`` and discovery of third-party
envs through the ``pybatchrender.envs`` entry-point group are kept.  Unlike the reference (one
Panda3D ``ShowBase`` per process) any number of envs / renderers can live in one process.
"""
from __future__ import annotations

_ENV_REGISTRY: dict[str, tuple] = {}


def register(name: str, env_cls, renderer_cls, config_cls, override: bool = False) -> None:
    if name in _ENV_REGISTRY and not override:
        raise ValueError(f"Environment '{name}' is already registered")
    _ENV_REGISTRY[name] = (env_cls, renderer_cls, config_cls)


def unregister(name: str) -> None:
    _ENV_REGISTRY.pop(name, None)


def list_envs() -> list[str]:
    return sorted(_ENV_REGISTRY)


def is_registered(name: str) -> bool:
    return name in _ENV_REGISTRY


def get_env_classes(name: str):
    if name not in _ENV_REGISTRY:
        raise ValueError(f"Unknown environment: '{name}'. Available: {list_envs()}")
    return _ENV_REGISTRY[name]


def make(name: str, **config_overrides):
    """``env = make("CartPole-v0", num_scenes=1024, tile_resolution=(64, 64))``"""
    env_cls, renderer_cls, config_cls = get_env_classes(name)
    cfg = config_cls(**config_overrides)
    renderer = renderer_cls(cfg)
    return env_cls(renderer=renderer, cfg=cfg)


def make_parallel(name: str, num_workers: int, shared_memory: bool = True, **config_overrides):
    env_cls, renderer_cls, config_cls = get_env_classes(name)
    cfg = config_cls(num_workers=int(num_workers), **config_overrides)
    return env_cls.make_parallel_env(config=cfg, renderer_cls=renderer_cls, num_workers=int(num_workers),
                                     shared_memory=shared_memory)


def _register_builtin_envs() -> None:
    from .cartpole import CartPoleConfig, CartPoleEnv, CartPoleRenderer
    if CartPoleEnv is not None and "CartPole-v0" not in _ENV_REGISTRY:
        register("CartPole-v0", CartPoleEnv, CartPoleRenderer, CartPoleConfig)
    from .steering import SteeringConfig, SteeringEnv, SteeringRenderer
    if SteeringEnv is not None and "Steering-v0" not in _ENV_REGISTRY:
        register("Steering-v0", SteeringEnv, SteeringRenderer, SteeringConfig)


def _discover_entry_points() -> None:
    try:
        from importlib.metadata import entry_points
        for ep in entry_points(group="pybatchrender.envs"):
            try:
                env_cls, renderer_cls, config_cls = ep.load()
                if ep.name not in _ENV_REGISTRY:
                    register(ep.name, env_cls, renderer_cls, config_cls)
            except Exception:
                continue
    except Exception:
        pass


_register_builtin_envs()
_discover_entry_points()
