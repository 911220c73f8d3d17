"""B200-native RenderFormer inference kernels and host engine."""
from .config import RenderFormerConfig  # noqa: F401

__all__ = ["RenderFormerConfig"]
