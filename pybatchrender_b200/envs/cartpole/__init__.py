from .renderer import CartPoleRenderer

__all__ = ["CartPoleRenderer"]
