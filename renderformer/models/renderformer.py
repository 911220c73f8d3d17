"""renderformer.models.renderformer -> renderformer_b200.model.RenderFormer."""
from renderformer_b200.model import RenderFormer

__all__ = ["RenderFormer"]
