from .config import CartPoleConfig
from .renderer import CartPoleRenderer

try:
    from .env import CartPoleEnv
except Exception:  # pragma: no cover
    CartPoleEnv = None

__all__ = ["CartPoleConfig", "CartPoleRenderer", "CartPoleEnv"]
