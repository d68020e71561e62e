from .config import SteeringConfig
from .renderer import SteeringRenderer

try:
    from .env import SteeringEnv
except Exception:  # pragma: no cover
    SteeringEnv = None

__all__ = ["SteeringConfig", "SteeringRenderer", "SteeringEnv"]
