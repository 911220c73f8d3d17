"""renderformer.models.config -> renderformer_b200.config.RenderFormerConfig."""
from renderformer_b200.config import RenderFormerConfig

__all__ = ["RenderFormerConfig"]
